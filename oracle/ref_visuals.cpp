// oracle/ref_visuals.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C entry points over the reference's OWN, UNMODIFIED visual processors: FFTDataDistributor, SpectrumVisualProcessor (both
// branches: full band and zoomed view) and ScopeVisualProcessor.  oracle/build_ref.py compiles
// /root/reference/src/process/{FFTDataDistributor,SpectrumVisualProcessor,ScopeVisualProcessor}.cpp from where they lie
// (nothing is copied into this repository) into oracle/_ref/libvisuals_ref.so together with this file; the two headers they
// cannot have here (the wx application, the OpenGL panel) are shadowed by oracle/shadow/.  tests/test_oracle_visuals.py pins
// the restated glue of oracle/ref_chain.cpp (what the GPU parity tests compare against) to these classes: same frames in,
// bit-identical frames / points / ceiling / floor out.
//
// The classes are driven the way their threads drive them (FFTVisualDataThread.cpp:40-84, SpectrumVisualDataThread.cpp):
// push the input queue, run(), drain the output queue.
#include "CubicSDR.h"                     // oracle/shadow
#include "FFTDataDistributor.h"
#include "SpectrumVisualProcessor.h"
#include "ScopeVisualProcessor.h"

#include <cstring>

static OracleAppStub g_app;
OracleAppStub &wxGetApp() { return g_app; }

namespace {
template <class Q>
std::shared_ptr<Q> big_queue() {
    auto q = std::make_shared<Q>();
    q->set_max_num_items(1u << 20);
    return q;
}
DemodulatorThreadIQDataPtr make_iq(const float *iq, unsigned n, long long rate, long long freq) {
    auto p = std::make_shared<DemodulatorThreadIQData>();
    p->frequency = freq;
    p->sampleRate = rate;
    p->data.resize(n);
    if (n) std::memcpy(reinterpret_cast<void *>(p->data.data()), iq, (size_t)n * 2 * sizeof(float));
    return p;
}

struct Distrib {
    FFTDataDistributor d;
    FFTDataDistributor::VisualInputQueueTypePtr in = big_queue<FFTDataDistributor::VisualInputQueueType>();
    FFTDataDistributor::VisualOutputQueueTypePtr out = big_queue<FFTDataDistributor::VisualOutputQueueType>();
};
struct Spectrum {
    SpectrumVisualProcessor p;
    SpectrumVisualProcessor::VisualInputQueueTypePtr in = big_queue<SpectrumVisualProcessor::VisualInputQueueType>();
    SpectrumVisualProcessor::VisualOutputQueueTypePtr out = big_queue<SpectrumVisualProcessor::VisualOutputQueueType>();
};
struct Scope {
    ScopeVisualProcessor p;
    ScopeVisualProcessor::VisualInputQueueTypePtr in = big_queue<ScopeVisualProcessor::VisualInputQueueType>();
    ScopeVisualProcessor::VisualOutputQueueTypePtr out = big_queue<ScopeVisualProcessor::VisualOutputQueueType>();
};
}  // namespace

extern "C" {

// ---- FFTDataDistributor ------------------------------------------------------------------------------------------------------
void *refsrc_distrib_create(unsigned fft_size, unsigned lines_per_second) {
    Distrib *s = new Distrib;
    s->d.setInput(s->in);
    s->d.attachOutput(s->out);
    s->d.setFFTSize(fft_size);
    s->d.setLinesPerSecond(lines_per_second);
    return s;
}
void refsrc_distrib_set_fft_size(void *h, unsigned n) { static_cast<Distrib *>(h)->d.setFFTSize(n); }
void refsrc_distrib_set_lines_per_second(void *h, unsigned n) { static_cast<Distrib *>(h)->d.setLinesPerSecond(n); }
// one block in; the frames it released are written back to back into frames (cap_samples complex samples of room);
// returns the number of frames, *n_samples = their total length
unsigned refsrc_distrib_push(void *h, const float *iq, unsigned n, long long rate, long long freq, float *frames, unsigned cap_samples,
                             unsigned *n_samples) {
    Distrib *s = static_cast<Distrib *>(h);
    s->in->push(make_iq(iq, n, rate, freq));
    s->d.run();
    unsigned nf = 0, used = 0;
    DemodulatorThreadIQDataPtr f;
    while (s->out->try_pop(f)) {
        const unsigned m = (unsigned)f->data.size();
        if (used + m <= cap_samples) std::memcpy(frames + 2 * (size_t)used, f->data.data(), (size_t)m * 2 * sizeof(float));
        used += m;
        ++nf;
    }
    *n_samples = used;
    return nf;
}
void refsrc_distrib_destroy(void *h) { delete static_cast<Distrib *>(h); }

// ---- SpectrumVisualProcessor ---------------------------------------------------------------------------------------------------
void *refsrc_spectrum_create(unsigned fft_size, long long device_sample_rate) {
    g_app.sampleRate = device_sample_rate;
    Spectrum *s = new Spectrum;
    s->p.setInput(s->in);
    s->p.attachOutput(s->out);
    s->p.setup(fft_size);
    return s;
}
void refsrc_spectrum_set_view(void *h, int view, long long center, long bandwidth) {
    static_cast<Spectrum *>(h)->p.setView(view != 0, center, bandwidth);
}
// what the application does for the full-band display: canvas->setBandwidth(sampleRate), canvas->setCenterFrequency(frequency)
// (src/AppFrame.cpp:1396-1399), which the canvases hand to their processor
void refsrc_spectrum_set_band(void *h, long long center, long bandwidth) {
    static_cast<Spectrum *>(h)->p.setCenterFrequency(center);
    static_cast<Spectrum *>(h)->p.setBandwidth(bandwidth);
}
void refsrc_spectrum_set_average_rate(void *h, float r) { static_cast<Spectrum *>(h)->p.setFFTAverageRate(r); }
void refsrc_spectrum_set_hide_dc(void *h, int on) { static_cast<Spectrum *>(h)->p.setHideDC(on != 0); }
int refsrc_spectrum_desired_input(void *h) { return static_cast<Spectrum *>(h)->p.getDesiredInputSize(); }
// one process() call on one frame; returns the number of display points produced (0: nothing was distributed);
// points_y[i] = spectrum_points[2 i + 1]; ceil_floor = {fft_ceiling, fft_floor}
int refsrc_spectrum_frame(void *h, const float *iq, unsigned n, long long freq, long long rate, float *points_y, unsigned cap,
                          double *ceil_floor, long long *center_out, int *bandwidth_out) {
    Spectrum *s = static_cast<Spectrum *>(h);
    s->in->push(make_iq(iq, n, rate, freq));
    s->p.run();
    SpectrumVisualDataPtr o;
    if (!s->out->try_pop(o) || !o) return 0;
    const unsigned m = (unsigned)(o->spectrum_points.size() / 2);
    for (unsigned i = 0; i < m && i < cap; ++i) points_y[i] = o->spectrum_points[2 * (size_t)i + 1];
    ceil_floor[0] = o->fft_ceiling;
    ceil_floor[1] = o->fft_floor;
    if (center_out) *center_out = o->centerFreq;
    if (bandwidth_out) *bandwidth_out = o->bandwidth;
    return (int)m;
}
void refsrc_spectrum_destroy(void *h) { delete static_cast<Spectrum *>(h); }

// ---- ScopeVisualProcessor -------------------------------------------------------------------------------------------------------
void *refsrc_scope_create(int fft_size) {
    Scope *s = new Scope;
    s->p.setInput(s->in);
    s->p.attachOutput(s->out);
    s->p.setup(fft_size);
    return s;
}
// one AudioThreadInput in; wave / spec receive waveform_points of the scope / spectrum render data (as is: interleaved pairs);
// returns the scope's float count, *n_spec the spectrum's; floor_ceil = {fft_floor, fft_ceil}; *mode = ScopeMode of the scope data
unsigned refsrc_scope_process(void *h, const float *data, unsigned n, int channels, int sample_rate, int input_rate, int type,
                              float *wave, unsigned wave_cap, float *spec, unsigned spec_cap, unsigned *n_spec, double *floor_ceil,
                              int *mode) {
    Scope *s = static_cast<Scope *>(h);
    auto a = std::make_shared<AudioThreadInput>();
    a->channels = channels;
    a->sampleRate = sample_rate;
    a->inputRate = input_rate;
    a->type = type;
    a->data.assign(data, data + n);
    s->in->push(a);
    s->p.run();
    unsigned nw = 0;
    *n_spec = 0;
    ScopeRenderDataPtr r;
    while (s->out->try_pop(r)) {
        const unsigned m = (unsigned)r->waveform_points.size();
        if (r->spectrum) {
            std::memcpy(spec, r->waveform_points.data(), (size_t)(m < spec_cap ? m : spec_cap) * sizeof(float));
            *n_spec = m;
            floor_ceil[0] = r->fft_floor;
            floor_ceil[1] = r->fft_ceil;
        } else {
            std::memcpy(wave, r->waveform_points.data(), (size_t)(m < wave_cap ? m : wave_cap) * sizeof(float));
            nw = m;
            if (mode) *mode = (int)r->mode;
        }
    }
    return nw;
}
void refsrc_scope_destroy(void *h) { delete static_cast<Scope *>(h); }

}  // extern "C"
