// oracle/ref_chain.cpp -- TEST INFRASTRUCTURE ONLY (checker + CPU baseline; never on the product path).
//
// CubicSDR's hot-path GLUE restated on top of the reference's own liquid-dsp 1.5.0 binary (transplanted by
// oracle/build_ref.py).  The arithmetic therefore IS the reference's; only the ~100 lines of C++ glue between
// liquid calls are restated, each block citing the reference lines it follows (paths relative to
// /root/reference/).  Exposed as a plain C API (SysV) for ctypes / bench.py.
//
// PINNED against the reference itself, run here: oracle/build_ref.py (steps 6-8) compiles the reference's own UNMODIFIED sources
// -- SDRPostThread, DemodulatorInstance with its pre-demod / worker / demodulator threads, the nine analog modems,
// FFTDataDistributor, SpectrumVisualProcessor, ScopeVisualProcessor, AudioThread's RtAudio callback, AudioFileWAV -- over the
// same liquid binary (oracle/ref_modems.cpp, ref_visuals.cpp, ref_demod.cpp, ref_audio.cpp), and tests/test_oracle_modems.py,
// test_oracle_visuals.py and test_oracle_demod.py assert that everything restated below returns the SAME BITS as those
// objects do on the same inputs.  (Kept as a restatement rather than replaced by them because the reference's objects need
// their threads, queues and ~20 ms of set-up per demodulator: too slow and too asynchronous for 512-channel parity runs and for
// the timed CPU baseline.)
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "liquid_ms.h"

#define API extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------------------------------------
// thin SysV -> ms_abi pass-throughs used by unit tests to pin the restatement (oracle/liquid_port.c) and the
// product's host-side filter design against the reference binary
// ---------------------------------------------------------------------------------------------------------
API const char *ref_libversion() { return liquid_libversion(); }
API unsigned ref_estimate_req_filter_len(float df, float as) { return estimate_req_filter_len(df, as); }
API float ref_estimate_req_filter_df(float as, unsigned n) { return estimate_req_filter_df(as, n); }
API float ref_kaiser_beta_As(float as) { return kaiser_beta_As(as); }
API int ref_firdespm_halfband_as(unsigned m, float as, float *h) { return liquid_firdespm_halfband_as(m, as, h); }
API int ref_firdes_kaiser(unsigned n, float fc, float as, float mu, float *h) { return liquid_firdes_kaiser(n, fc, as, mu, h); }
API float ref_besseli0f(float z) { return liquid_besseli0f(z); }
API float ref_kaiser_window(unsigned i, unsigned n, float beta) { return liquid_kaiser(i, n, beta); }

API void *ref_nco_create() { return nco_crcf_create(0 /*LIQUID_NCO*/); }
API void ref_nco_destroy(void *q) { nco_crcf_destroy((nco_crcf)q); }
API void ref_nco_set_frequency(void *q, float f) { nco_crcf_set_frequency((nco_crcf)q, f); }
API float ref_nco_get_frequency(void *q) { return nco_crcf_get_frequency((nco_crcf)q); }
API void ref_nco_set_phase(void *q, float p) { nco_crcf_set_phase((nco_crcf)q, p); }
API float ref_nco_get_phase(void *q) { return nco_crcf_get_phase((nco_crcf)q); }
// object layout {int type; float sintab[1024]; uint32 theta; uint32 d_theta; ...}
API uint32_t ref_nco_raw_theta(void *q) { uint32_t v; memcpy(&v, (char *)q + 4 + 4096, 4); return v; }
API uint32_t ref_nco_raw_dtheta(void *q) { uint32_t v; memcpy(&v, (char *)q + 4 + 4096 + 4, 4); return v; }
API void ref_nco_sintab(void *q, float *tab) { memcpy(tab, (char *)q + 4, 4096); }
API void ref_nco_mix_block(void *q, int up, lq_cf *x, lq_cf *y, unsigned n) {
    if (up) nco_crcf_mix_block_up((nco_crcf)q, x, y, n); else nco_crcf_mix_block_down((nco_crcf)q, x, y, n);
}

API void *ref_resamp2_crcf_create(unsigned m, float f0, float as) { return resamp2_crcf_create(m, f0, as); }
API void ref_resamp2_crcf_destroy(void *q) { resamp2_crcf_destroy((resamp2_crcf)q); }
API void ref_resamp2_crcf_decim(void *q, lq_cf *x, unsigned n_pairs, lq_cf *y) {
    for (unsigned i = 0; i < n_pairs; i++) resamp2_crcf_decim_execute((resamp2_crcf)q, x + 2 * i, y + i);
}
API void *ref_resamp_crcf_create(float r, unsigned m, float fc, float as, unsigned npfb) { return resamp_crcf_create(r, m, fc, as, npfb); }
API void ref_resamp_crcf_destroy(void *q) { resamp_crcf_destroy((resamp_crcf)q); }
API unsigned ref_resamp_crcf_execute_block(void *q, lq_cf *x, unsigned n, lq_cf *y) {
    unsigned tot = 0, nw;
    for (unsigned i = 0; i < n; i++) { resamp_crcf_execute((resamp_crcf)q, x[i], y + tot, &nw); tot += nw; }
    return tot;
}
API void *ref_msresamp_crcf_create(float r, float as) { return msresamp_crcf_create(r, as); }
API void ref_msresamp_crcf_destroy(void *q) { msresamp_crcf_destroy((msresamp_crcf)q); }
API float ref_msresamp_crcf_get_delay(void *q) { return msresamp_crcf_get_delay((msresamp_crcf)q); }
API unsigned ref_msresamp_crcf_execute(void *q, lq_cf *x, unsigned n, lq_cf *y) {
    unsigned ny = 0; msresamp_crcf_execute((msresamp_crcf)q, x, n, y, &ny); return ny;
}
API void *ref_msresamp_rrrf_create(float r, float as) { return msresamp_rrrf_create(r, as); }
API void ref_msresamp_rrrf_destroy(void *q) { msresamp_rrrf_destroy((msresamp_rrrf)q); }
API float ref_msresamp_rrrf_get_delay(void *q) { return msresamp_rrrf_get_delay((msresamp_rrrf)q); }
API unsigned ref_msresamp_rrrf_execute(void *q, float *x, unsigned n, float *y) {
    unsigned ny = 0; msresamp_rrrf_execute((msresamp_rrrf)q, x, n, y, &ny); return ny;
}
API void *ref_freqdem_create(float kf) { return freqdem_create(kf); }
API void ref_freqdem_destroy(void *q) { freqdem_destroy((freqdem)q); }
API void ref_freqdem_block(void *q, lq_cf *x, unsigned n, float *y) { freqdem_demodulate_block((freqdem)q, x, n, y); }
API void *ref_firfilt_dc_blocker_create(unsigned m, float as) { return firfilt_rrrf_create_dc_blocker(m, as); }
API void ref_firfilt_destroy(void *q) { firfilt_rrrf_destroy((firfilt_rrrf)q); }
API void ref_firfilt_block(void *q, float *x, unsigned n, float *y) {
    for (unsigned i = 0; i < n; i++) { firfilt_rrrf_push((firfilt_rrrf)q, x[i]); firfilt_rrrf_execute((firfilt_rrrf)q, y + i); }
}
API void *ref_iirfilt_lowpass_create(unsigned order, float fc) { return iirfilt_crcf_create_lowpass(order, fc); }
API void *ref_iirfilt_dc_blocker_create(float alpha) { return iirfilt_crcf_create_dc_blocker(alpha); }
API void ref_iirfilt_destroy(void *q) { iirfilt_crcf_destroy((iirfilt_crcf)q); }
API void ref_iirfilt_block(void *q, lq_cf *x, unsigned n, lq_cf *y) { iirfilt_crcf_execute_block((iirfilt_crcf)q, x, n, y); }
API void *ref_firhilb_create(unsigned m, float as) { return firhilbf_create(m, as); }
API void ref_firhilb_destroy(void *q) { firhilbf_destroy((firhilbf)q); }
API void ref_firhilb_c2r_block(void *q, lq_cf *x, unsigned n, float *lsb, float *usb) {
    for (unsigned i = 0; i < n; i++) firhilbf_c2r_execute((firhilbf)q, x[i], lsb + i, usb + i);
}
API void ref_firhilb_r2c_block(void *q, float *x, unsigned n, lq_cf *y) {
    for (unsigned i = 0; i < n; i++) firhilbf_r2c_execute((firhilbf)q, x[i], y + i);
}
API int ref_iirdes(int ftype, int btype, int format, unsigned n, float fc, float f0, float ap, float as, float *b, float *a) {
    return liquid_iirdes(ftype, btype, format, n, fc, f0, ap, as, b, a);
}
// the stages of liquid_iirdes, one by one (pins for the product's host-side restatement, csrc/design.cpp)
API void ref_cheby2_azpk(unsigned n, float es, lq_cf *z, lq_cf *p, lq_cf *k) { cheby2_azpkf(n, es, z, p, k); }
API float ref_iirdes_freqprewarp(int btype, float fc, float f0) { return iirdes_freqprewarp(btype, fc, f0); }
API void ref_bilinear_zpk(lq_cf *za, unsigned nza, lq_cf *pa, unsigned npa, const lq_cf *ka, float m, lq_cf *zd, lq_cf *pd, lq_cf *kd) {
    bilinear_zpkf(za, nza, pa, npa, *ka, m, zd, pd, kd);
}
API void ref_dzpk_lp2bp(lq_cf *zd, lq_cf *pd, unsigned n, float f0, lq_cf *zdt, lq_cf *pdt) { iirdes_dzpk_lp2bp(zd, pd, n, f0, zdt, pdt); }
API void ref_dzpk2sos(lq_cf *zd, lq_cf *pd, unsigned n, const lq_cf *kd, float *b, float *a) { iirdes_dzpk2sosf(zd, pd, n, *kd, b, a); }
API void *ref_iirfilt_prototype_create(int ftype, int btype, int format, unsigned order, float fc, float f0, float ap, float as) {
    return iirfilt_crcf_create_prototype(ftype, btype, format, order, fc, f0, ap, as);
}
API void *ref_firfilt_create(float *h, unsigned n) { return firfilt_rrrf_create(h, n); }
API void *ref_iirfilt_rrrf_create(float *b, unsigned nb, float *a, unsigned na) { return iirfilt_rrrf_create(b, nb, a, na); }
API void ref_iirfilt_rrrf_destroy(void *q) { iirfilt_rrrf_destroy((iirfilt_rrrf)q); }
API void ref_iirfilt_rrrf_block(void *q, float *x, unsigned n, float *y) {
    for (unsigned i = 0; i < n; i++) iirfilt_rrrf_execute((iirfilt_rrrf)q, x[i], y + i);
}
// nco_crcf phase-locked loop: bandwidth, then one pll_step per phase error; raw phase / frequency words after each
API void ref_nco_pll_run(float bw, const float *dphi, unsigned n, uint32_t *theta, uint32_t *dtheta) {
    nco_crcf q = nco_crcf_create(1 /* LIQUID_VCO */);
    nco_crcf_reset(q);
    nco_crcf_pll_set_bandwidth(q, bw);
    for (unsigned i = 0; i < n; i++) {
        nco_crcf_pll_step(q, dphi[i]);
        nco_crcf_step(q);
        theta[i] = ref_nco_raw_theta(q); dtheta[i] = ref_nco_raw_dtheta(q);
    }
    nco_crcf_destroy(q);
}
API void ref_fft_forward(lq_cf *x, lq_cf *y, unsigned n) {
    fftplan p = fft_create_plan(n, x, y, +1 /*LIQUID_FFT_FORWARD*/, 0);
    fft_execute(p);
    fft_destroy_plan(p);
}

// ---------------------------------------------------------------------------------------------------------
// SDRPostThread::runSingleCH DC blocker     src/sdr/SDRPostThread.cpp:29 (create), :284 (execute_block)
// ---------------------------------------------------------------------------------------------------------
struct RefIngress { iirfilt_crcf dc; };
API void *ref_ingress_create() { auto *g = new RefIngress; g->dc = iirfilt_crcf_create_dc_blocker(0.0005f); return g; }
API void ref_ingress_destroy(void *h) { auto *g = (RefIngress *)h; iirfilt_crcf_destroy(g->dc); delete g; }
API void ref_ingress_block(void *h, lq_cf *in, unsigned n, lq_cf *out) {
    iirfilt_crcf_execute_block(((RefIngress *)h)->dc, in, n, out);
}

// ---------------------------------------------------------------------------------------------------------
// SDRPostThread::runPFBCH channelizer       src/sdr/SDRPostThread.cpp:401-455
//   initPFBCH :401-414  firpfbch_crcf_create_kaiser(LIQUID_ANALYZER, numChannels, 4, 60)
//   runPFBCH  :447-451  one analyzer_execute per numChannels input samples -> numChannels interleaved outputs
// ---------------------------------------------------------------------------------------------------------
struct RefPfb { firpfbch_crcf q; unsigned M; };
API void *ref_pfb_create(unsigned num_channels, unsigned m, float as_db) {
    auto *p = new RefPfb; p->M = num_channels;
    p->q = firpfbch_crcf_create_kaiser(0 /* LIQUID_ANALYZER */, num_channels, m, as_db);
    return p;
}
API void ref_pfb_destroy(void *h) { auto *p = (RefPfb *)h; firpfbch_crcf_destroy(p->q); delete p; }
// n must be a multiple of num_channels (SDRThread::getOptimalElementCount guarantees it, SoapySDRThread.cpp:668-674)
API void ref_pfb_execute(void *h, lq_cf *in, unsigned n, lq_cf *out) {
    auto *p = (RefPfb *)h;
    for (unsigned i = 0; i + p->M <= n; i += p->M) firpfbch_crcf_analyzer_execute(p->q, in + i, out + i);
}
// SDRPostThread::runPFBCH2 channelizer      src/sdr/SDRPostThread.cpp:457-511
//   initPFBCH2 :457-470  firpfbch2_crcf_create_kaiser(LIQUID_ANALYZER, numChannels, 4, 60)
//   runPFBCH2  :503-506  one execute per numChannels/2 input samples -> numChannels interleaved outputs
struct RefPfb2 { firpfbch2_crcf q; unsigned M; };
API void *ref_pfb2_create(unsigned num_channels, unsigned m, float as_db) {
    auto *p = new RefPfb2; p->M = num_channels;
    p->q = firpfbch2_crcf_create_kaiser(0 /* LIQUID_ANALYZER */, num_channels, m, as_db);
    return p;
}
API void ref_pfb2_destroy(void *h) { auto *p = (RefPfb2 *)h; firpfbch2_crcf_destroy(p->q); delete p; }
// n must be a multiple of num_channels / 2; out receives 2 n samples
API void ref_pfb2_execute(void *h, lq_cf *in, unsigned n, lq_cf *out) {
    auto *p = (RefPfb2 *)h;
    const unsigned M2 = p->M / 2;
    for (unsigned i = 0; i + M2 <= n; i += M2) firpfbch2_crcf_execute(p->q, in + i, out + 2 * i);
}
// SDRThread::getOptimalChannelCount          src/sdr/SoapySDRThread.cpp:676-693  (CHANNELIZER_RATE_MAX = 500000, CubicSDRDefs.h:63)
API int ref_optimal_channel_count(long long sampleRate_in) {
    if (sampleRate_in <= 500000) return 1;
    int optimal_rate = 500000;
    int optimal_count = int(ceil(double(sampleRate_in) / double(optimal_rate)));
    if (optimal_count % 2 == 1) optimal_count--;
    if (optimal_count < 2) optimal_count = 2;
    return optimal_count;
}

// ---------------------------------------------------------------------------------------------------------
// one demodulator channel: DemodulatorPreThread + Modem + DemodulatorThread metering
// ---------------------------------------------------------------------------------------------------------
enum { M_FM = 0, M_NBFM = 1, M_AM = 2, M_USB = 3, M_LSB = 4, M_IQ = 5, M_DSB = 7, M_CW = 8, M_FMS = 9 };

struct RefChan {
    int modem;
    long long sampleRate;      // wideband rate
    int bandwidth;             // modem rate (after checkSampleRate)
    int requestedBandwidth;    // what setBandwidth was given: DemodulatorThreadPostIQData::sampleRate (DemodulatorPreThread.cpp:222, currentBandwidth)
    int audioRate;
    // DemodulatorPreThread members (src/demod/DemodulatorPreThread.cpp:22, DemodulatorWorkerThread.cpp:95-101)
    nco_crcf freqShifter;
    msresamp_crcf iqResampler;
    double iqResampleRatio;
    long long shiftFrequency;
    // ModemAnalog (src/modules/modem/ModemAnalog.cpp:6, :21-33)
    msresamp_rrrf audioResampler;
    double audioResampleRatio;
    float aOutputCeil, aOutputCeilMA, aOutputCeilMAA;
    // modem objects (ModemFM.cpp:7, ModemAM.cpp:9, ModemUSB.cpp:8-11)
    freqdem demodFM;
    firfilt_rrrf mDCBlock;
    iirfilt_crcf ssbFilt;
    nco_crcf ssbShift;
    firhilbf c2rFilt;
    ampmodem demodAM_DSB = nullptr;                                            // ModemDSB.cpp:7
    // ModemCW members (ModemCW.cpp:15-25) and its kit's msresamp_cccf (:117-127)
    msresamp_cccf cwInputResampler = nullptr; nco_crcf cwLO = nullptr; firhilbf cwToReal = nullptr;
    float cwBeepFrequency = 650.0f, cwGain = 15.0f;
    std::vector<lq_cf> cwInput;
    // ModemKitFMStereo (ModemFMStereo.h:7-32) and ModemFMStereo's buffers (:55-58)
    msresamp_rrrf stereoResampler = nullptr;
    firfilt_rrrf firStereoLeft = nullptr, firStereoRight = nullptr;
    iirfilt_crcf iirStereoPilot = nullptr;
    int demph = 75;
    iirfilt_rrrf iirDemphR = nullptr, iirDemphL = nullptr;
    firhilbf firStereoR2C = nullptr, firStereoC2R = nullptr;
    nco_crcf stereoPilot = nullptr;
    std::vector<float> demodStereoData, resampledStereoData;
    std::vector<lq_cf> dbg_x, dbg_v, dbg_x2;          // test instrumentation: the loop's intermediates of the last block
    std::vector<lq_cf> in_buf, out_buf, resampled;
    std::vector<float> demodOut, audioOut;
    // DemodulatorThread members (src/demod/DemodulatorThread.cpp:25-26): squelchLevel(-100), signalLevel(-100),
    // signalFloor(-30), signalCeil(30), squelchEnabled(false)
    float squelchLevel = -100, signalLevel = -100, signalFloor = -30, signalCeil = 30;
    bool squelchEnabled = false, squelched = false;
};

static int modem_from_name(const char *n) {
    std::string s(n);
    if (s == "FM") return M_FM;
    if (s == "NBFM") return M_NBFM;
    if (s == "AM") return M_AM;
    if (s == "USB") return M_USB;
    if (s == "LSB") return M_LSB;
    if (s == "I/Q") return M_IQ;
    if (s == "DSB") return M_DSB;
    if (s == "CW") return M_CW;
    if (s == "FMS") return M_FMS;
    return -1;
}

// Modem*::checkSampleRate: ModemAnalog.cpp:14-19 (MIN_BANDWIDTH = 500, src/CubicSDRDefs.h), ModemUSB.cpp:29-37
API int ref_modem_check_sample_rate(const char *name, long long bw) {
    int m = modem_from_name(name);
    if (m < 0) return -1;
    if (m == M_FMS) return bw < 100000 ? 100000 : (int)bw;                  // ModemFMStereo.cpp:27-35
    if (bw < 500) return 500;
    if ((m == M_USB || m == M_LSB) && (bw % 2)) return (int)(bw + 1);
    return (int)bw;
}

API void *ref_chan_create(const char *modem, long long sampleRate, long long bandwidth, int audioRate) {
    int m = modem_from_name(modem);
    if (m < 0) return nullptr;
    auto *c = new RefChan();
    c->modem = m;
    c->sampleRate = sampleRate;
    c->bandwidth = ref_modem_check_sample_rate(modem, bandwidth);
    c->requestedBandwidth = (int)bandwidth;
    if (m == M_IQ) c->bandwidth = audioRate;                               // ModemIQ::checkSampleRate (ModemIQ.cpp:31-33)
    c->audioRate = audioRate;
    c->freqShifter = nco_crcf_create(0);                                   // DemodulatorPreThread.cpp:22
    c->shiftFrequency = 0;
    // DemodulatorWorkerThread.cpp:95-101
    float As = 60.0f;
    c->iqResampleRatio = (double)c->bandwidth / (double)sampleRate;
    c->iqResampler = msresamp_crcf_create((float)c->iqResampleRatio, As);
    // ModemAnalog::buildKit ModemAnalog.cpp:21-33
    c->audioResampleRatio = double(audioRate) / double(c->bandwidth);
    c->audioResampler = msresamp_rrrf_create((float)c->audioResampleRatio, As);
    c->aOutputCeil = c->aOutputCeilMA = c->aOutputCeilMAA = 1;             // ModemAnalog.cpp:6
    c->demodFM = nullptr; c->mDCBlock = nullptr; c->ssbFilt = nullptr; c->ssbShift = nullptr; c->c2rFilt = nullptr;
    if (m == M_FM || m == M_NBFM) c->demodFM = freqdem_create(0.5f);       // ModemFM.cpp:7
    if (m == M_AM) c->mDCBlock = firfilt_rrrf_create_dc_blocker(25, 30.0f); // ModemAM.cpp:9
    if (m == M_DSB) c->demodAM_DSB = ampmodem_create(0.5f, 0 /* LIQUID_AMPMODEM_DSB */, 1);   // ModemDSB.cpp:7
    if (m == M_CW) {                                                       // ModemCW.cpp:21-23, 117-127
        c->cwLO = nco_crcf_create(0);
        c->cwToReal = firhilbf_create(5, 60.0f);
        c->cwInputResampler = msresamp_cccf_create((float)c->audioResampleRatio, 60.0f);
    }
    if (m == M_FMS) {                                                      // ModemFMStereo.cpp:6-9, 89-167 (buildKit)
        c->demodFM = freqdem_create(0.5f);
        c->stereoResampler = msresamp_rrrf_create((float)c->audioResampleRatio, As);
        float firStereoCutoff = 16000.0f / float(audioRate);
        float ft = 1000.0f / float(audioRate);
        float mu = 0.0f;
        if (firStereoCutoff < 0) firStereoCutoff = 0;
        if (firStereoCutoff > 0.5) firStereoCutoff = 0.5;
        unsigned int h_len = estimate_req_filter_len(ft, As);
        std::vector<float> h(h_len);
        liquid_firdes_kaiser(h_len, firStereoCutoff, As, mu, h.data());
        c->firStereoLeft = firfilt_rrrf_create(h.data(), h_len);
        c->firStereoRight = firfilt_rrrf_create(h.data(), h_len);
        float bw = float(c->bandwidth);
        if (bw < 100000.0) bw = 100000.0;
        unsigned int order = 5;
        float f0 = ((float)19000 / bw);
        float fc = ((float)19500 / bw);
        float Ap = 1.0f;
        c->iirStereoPilot = iirfilt_crcf_create_prototype(2 /* CHEBY2 */, 2 /* BANDPASS */, 0 /* SOS */, order, fc, f0, Ap, As);
        c->firStereoR2C = firhilbf_create(5, 60.0f);
        c->firStereoC2R = firhilbf_create(5, 60.0f);
        c->stereoPilot = nco_crcf_create(1 /* LIQUID_VCO */);
        nco_crcf_reset(c->stereoPilot);
        nco_crcf_pll_set_bandwidth(c->stereoPilot, 0.25f);
        if (c->demph) {                                                     // :149-166, _demph = 75 (:8)
            double f = (1.0 / (2.0 * M_PI * double(c->demph) * 1e-6));
            double t = 1.0 / (2.0 * M_PI * f);
            t = 1.0 / (2.0 * double(audioRate) * tan(1.0 / (2.0 * double(audioRate) * t)));
            double tb = (1.0 + 2.0 * t * double(audioRate));
            float b_demph[2] = {(float)(1.0 / tb), (float)(1.0 / tb)};
            float a_demph[2] = {1.0, (float)((1.0 - 2.0 * t * double(audioRate)) / tb)};
            c->iirDemphL = iirfilt_rrrf_create(b_demph, 2, a_demph, 2);
            c->iirDemphR = iirfilt_rrrf_create(b_demph, 2, a_demph, 2);
        }
    }
    if (m == M_USB || m == M_LSB) {                                        // ModemUSB.cpp:8-11
        c->ssbFilt = iirfilt_crcf_create_lowpass(6, 0.25);
        c->ssbShift = nco_crcf_create(0);
        nco_crcf_set_frequency(c->ssbShift, (float)((2.0 * M_PI) * 0.25));
        c->c2rFilt = firhilbf_create(5, 90.0);
    }
    return c;
}

API void ref_chan_destroy(void *h) {
    auto *c = (RefChan *)h;
    nco_crcf_destroy(c->freqShifter);
    msresamp_crcf_destroy(c->iqResampler);
    msresamp_rrrf_destroy(c->audioResampler);
    if (c->demodFM) freqdem_destroy(c->demodFM);
    if (c->mDCBlock) firfilt_rrrf_destroy(c->mDCBlock);
    if (c->ssbFilt) iirfilt_crcf_destroy(c->ssbFilt);
    if (c->ssbShift) nco_crcf_destroy(c->ssbShift);
    if (c->c2rFilt) firhilbf_destroy(c->c2rFilt);
    if (c->demodAM_DSB) ampmodem_destroy(c->demodAM_DSB);
    if (c->cwLO) nco_crcf_destroy(c->cwLO);
    if (c->cwToReal) firhilbf_destroy(c->cwToReal);
    if (c->cwInputResampler) msresamp_cccf_destroy(c->cwInputResampler);
    if (c->stereoResampler) msresamp_rrrf_destroy(c->stereoResampler);
    if (c->firStereoLeft) firfilt_rrrf_destroy(c->firStereoLeft);
    if (c->firStereoRight) firfilt_rrrf_destroy(c->firStereoRight);
    if (c->iirStereoPilot) iirfilt_crcf_destroy(c->iirStereoPilot);
    if (c->iirDemphL) iirfilt_rrrf_destroy(c->iirDemphL);
    if (c->iirDemphR) iirfilt_rrrf_destroy(c->iirDemphR);
    if (c->firStereoR2C) firhilbf_destroy(c->firStereoR2C);
    if (c->firStereoC2R) firhilbf_destroy(c->firStereoC2R);
    if (c->stereoPilot) nco_crcf_destroy(c->stereoPilot);
    delete c;
}

API int ref_chan_bandwidth(void *h) { return ((RefChan *)h)->bandwidth; }

// --- DemodulatorThread.cpp:142-163 (level) and :223-233 (peak); abMagnitude :49-58, linearToDb :60-67; then :165-198:
// floor / ceiling trackers, smoothed level, squelch decision.  The members are std::atomic<float>; the arithmetic is the
// reference's mixed float / double expressions, statement by statement.  c->audioOut holds the block's audio, c->resampled
// the decimated IQ the modem was given.
static int ref_chan_meter_block(RefChan *c, unsigned numAudioWritten, size_t bufSize, bool useSignalOutput, double *level_db, float *peak) {
    const lq_cf *d = c->resampled.data();
    double currentSignalLevel = 0;
    double sampleTime = double(bufSize) / double(c->requestedBandwidth);                     // :143 (inp->sampleRate = the bandwidth that was ASKED for, not the one checkSampleRate made of it)
    if (numAudioWritten) {                                                                    // :145 `!ati->data.empty()`
        double accum = 0;
        if (useSignalOutput) {
            for (unsigned i = 0; i < numAudioWritten; i++) { double v = (double)c->audioOut[i]; accum += sqrt(v * v + 0.0); }
            accum /= double(numAudioWritten);
        } else {
            for (size_t i = 0; i < bufSize; i++) { double a = d[i].real, b = d[i].imag; accum += sqrt(a * a + b * b); }
            accum /= double(bufSize);
        }
        if (accum <= 1e-20) accum = 1e-20;
        currentSignalLevel = 20.0 * log10(accum);
        float sf = c->signalFloor, sc = c->signalCeil, sl = c->squelchLevel;                 // :165
        if (currentSignalLevel > sc) sc = currentSignalLevel;
        if (currentSignalLevel < sf) sf = currentSignalLevel;
        if (sl + 1.0f > sc) sc = sl + 1.0f;
        if ((sf + 2.0f) > sc) sc = sf + 2.0f;
        sc -= (sc - (currentSignalLevel + 2.0f)) * sampleTime * 0.05f;                        // :185
        sf += ((currentSignalLevel - 5.0f) - sf) * sampleTime * 0.15f;                        // :186
        c->signalFloor = sf; c->signalCeil = sc;
    }
    if (level_db) *level_db = currentSignalLevel;
    if (currentSignalLevel > c->signalLevel) c->signalLevel = c->signalLevel + (currentSignalLevel - c->signalLevel) * 0.5;     // :192-196
    else c->signalLevel = c->signalLevel + (currentSignalLevel - c->signalLevel) * 0.05 * sampleTime * 30.0;
    c->squelched = c->squelchEnabled && (c->signalLevel < c->squelchLevel);                   // :198
    if (peak) {                                                                               // :223-233
        float p = 0;
        for (unsigned i = 0; i < numAudioWritten; i++) { float a = fabs(c->audioOut[i]); if (a > p) p = a; }
        *peak = p;
    }
    return 0;
}

// One block through the channel.  shift = demod frequency - block centre frequency.
// Returns 0, or 1 if the block was skipped by the |shift| > 0.75 fs rule.
// Outputs (any may be null): decimated IQ, demodulated (pre-audio-resample, post-autogain) samples, audio,
// level_db (DemodulatorThread.cpp:142-163), peak (:223-233).
API int ref_chan_process(void *h, const lq_cf *block, unsigned n, long long shift,
                         lq_cf *iq_out, unsigned *n_iq, float *demod_out, float *audio_out, unsigned *n_audio,
                         double *level_db, float *peak) {
    auto *c = (RefChan *)h;
    if (n_iq) *n_iq = 0;
    if (n_audio) *n_audio = 0;
    // --- DemodulatorPreThread.cpp:154-163
    if (shift != c->shiftFrequency) {
        c->shiftFrequency = shift;
        if (std::llabs(c->shiftFrequency) <= (int)((double)(c->sampleRate / 2) * 1.5)) {
            nco_crcf_set_frequency(c->freqShifter, (2.0 * M_PI) * (((double)std::llabs(c->shiftFrequency)) / ((double)c->sampleRate)));
        }
    }
    if (std::llabs(c->shiftFrequency) > (int)((double)(c->sampleRate / 2) * 1.5)) return 1;
    if (n == 0) return 0;
    // --- :166-195
    c->in_buf.assign(block, block + n);
    c->out_buf.resize(n);
    lq_cf *in_buf = c->in_buf.data(), *out_buf = c->out_buf.data();
    if (c->shiftFrequency != 0) {
        if (c->shiftFrequency < 0) nco_crcf_mix_block_up(c->freqShifter, in_buf, out_buf, n);
        else nco_crcf_mix_block_down(c->freqShifter, in_buf, out_buf, n);
        std::swap(in_buf, out_buf);
    }
    // --- :199-211
    size_t out_size = (size_t)ceil((double)n * c->iqResampleRatio) + 512;
    c->resampled.resize(out_size);
    unsigned numWritten = 0;
    msresamp_crcf_execute(c->iqResampler, in_buf, n, c->resampled.data(), &numWritten);
    if (iq_out) memcpy(iq_out, c->resampled.data(), numWritten * sizeof(lq_cf));
    if (n_iq) *n_iq = numWritten;

    // --- DemodulatorThread.cpp:120-140 -> Modem::demodulate
    size_t bufSize = numWritten;                                            // ModemAnalog::initOutputBuffers :42-65
    if (!bufSize) return 0;
    if (c->modem == M_CW) {                                                 // ModemCW::demodulate (ModemCW.cpp:155-209)
        size_t audio_out_size = (size_t)ceil((double)(bufSize)*c->audioResampleRatio) + 512;
        c->cwInput.resize(audio_out_size);
        unsigned int outSize = 0;
        msresamp_cccf_execute(c->cwInputResampler, c->resampled.data(), (unsigned int)bufSize, c->cwInput.data(), &outSize);
        c->demodOut.resize(outSize);
        nco_crcf_set_frequency(c->cwLO, 2.0f * (float)M_PI * c->cwBeepFrequency / c->audioRate);
        for (unsigned int i = 0; i < outSize; i++) {
            lq_cf sig; float lsb;
            nco_crcf_mix_up(c->cwLO, c->cwInput[i], &sig);
            nco_crcf_step(c->cwLO);
            firhilbf_c2r_execute(c->cwToReal, sig, &lsb, &c->demodOut[i]);
        }
        c->aOutputCeilMA = c->aOutputCeilMA + (c->aOutputCeil - c->aOutputCeilMA) * 0.025f;    // mAutoGain (:187-199)
        c->aOutputCeilMAA = c->aOutputCeilMAA + (c->aOutputCeilMA - c->aOutputCeilMAA) * 0.025f;
        c->aOutputCeil = 0;
        for (size_t i = 0; i < outSize; i++) if (c->demodOut[i] > c->aOutputCeil) c->aOutputCeil = c->demodOut[i];
        c->cwGain = 10.0f * std::log10(0.5f / c->aOutputCeilMAA);
        for (size_t i = 0; i < outSize; i++) c->demodOut[i] *= std::pow(10.0f, c->cwGain / 10.0f);
        c->audioOut.assign(c->demodOut.begin(), c->demodOut.end());
        if (audio_out) memcpy(audio_out, c->audioOut.data(), outSize * sizeof(float));
        if (n_audio) *n_audio = outSize;
        return ref_chan_meter_block(c, outSize, bufSize, true, level_db, peak);
    }
    if (c->modem == M_FMS) {                                                // ModemFMStereo::demodulate (ModemFMStereo.cpp:186-287)
        size_t audio_out_size = (size_t)ceil((double)(bufSize)*c->audioResampleRatio) + 512;
        c->demodOut.resize(bufSize);
        freqdem_demodulate_block(c->demodFM, c->resampled.data(), (int)bufSize, c->demodOut.data());
        std::vector<float> resampledOutputData(audio_out_size);
        unsigned int numAudioWritten = 0;
        msresamp_rrrf_execute(c->audioResampler, c->demodOut.data(), (int)bufSize, resampledOutputData.data(), &numAudioWritten);
        c->demodStereoData.resize(bufSize);
        float phase_error = 0;
        for (size_t i = 0; i < bufSize; i++) {
            lq_cf u, v, w, x, y;
            if (i == 0) { c->dbg_x.resize(bufSize); c->dbg_v.resize(bufSize); c->dbg_x2.resize(bufSize); }
            firhilbf_r2c_execute(c->firStereoR2C, c->demodOut[i], &x);                  // real -> complex
            c->dbg_x[i] = x;
            iirfilt_crcf_execute(c->iirStereoPilot, x, &v);                              // 19 kHz pilot band-pass
            c->dbg_v[i] = v;
            nco_crcf_cexpf(c->stereoPilot, &w);
            w.imag = -w.imag;
            u.real = v.real * w.real - v.imag * w.imag;
            u.imag = v.real * w.imag + v.imag * w.real;
            phase_error = atan2f(u.imag, u.real);
            nco_crcf_pll_step(c->stereoPilot, phase_error);
            nco_crcf_step(c->stereoPilot);
            nco_crcf_mix_down(c->stereoPilot, x, &y);                                   // 38 kHz down-mix
            nco_crcf_mix_down(c->stereoPilot, y, &x);
            c->dbg_x2[i] = x;
            float usb_discard;
            firhilbf_c2r_execute(c->firStereoC2R, x, &c->demodStereoData[i], &usb_discard);
        }
        c->resampledStereoData.resize(audio_out_size);
        msresamp_rrrf_execute(c->stereoResampler, c->demodStereoData.data(), (int)bufSize, c->resampledStereoData.data(), &numAudioWritten);
        c->audioOut.resize(numAudioWritten * 2);
        for (size_t i = 0; i < numAudioWritten; i++) {
            float l, r, ld, rd;
            if (c->demph) {
                iirfilt_rrrf_execute(c->iirDemphL, 0.568f * (resampledOutputData[i] - (c->resampledStereoData[i])), &ld);
                iirfilt_rrrf_execute(c->iirDemphR, 0.568f * (resampledOutputData[i] + (c->resampledStereoData[i])), &rd);
                firfilt_rrrf_push(c->firStereoLeft, ld);
                firfilt_rrrf_execute(c->firStereoLeft, &l);
                firfilt_rrrf_push(c->firStereoRight, rd);
                firfilt_rrrf_execute(c->firStereoRight, &r);
            } else {
                firfilt_rrrf_push(c->firStereoLeft, 0.568f * (resampledOutputData[i] - (c->resampledStereoData[i])));
                firfilt_rrrf_execute(c->firStereoLeft, &l);
                firfilt_rrrf_push(c->firStereoRight, 0.568f * (resampledOutputData[i] + (c->resampledStereoData[i])));
                firfilt_rrrf_execute(c->firStereoRight, &r);
            }
            c->audioOut[i * 2] = l;
            c->audioOut[i * 2 + 1] = r;
        }
        if (demod_out) memcpy(demod_out, c->demodStereoData.data(), bufSize * sizeof(float));   // the 38 kHz difference signal, for localisation
        if (audio_out) memcpy(audio_out, c->audioOut.data(), numAudioWritten * 2 * sizeof(float));
        if (n_audio) *n_audio = numAudioWritten * 2;
        return ref_chan_meter_block(c, numAudioWritten * 2, bufSize, false, level_db, peak);
    }
    if (c->modem == M_IQ) {                                                 // ModemIQ::demodulate (ModemIQ.cpp:39-56): channels = 2
        c->audioOut.resize(bufSize * 2);
        for (size_t i = 0; i < bufSize; i++) { c->audioOut[i * 2] = c->resampled[i].imag; c->audioOut[i * 2 + 1] = c->resampled[i].real; }
        if (audio_out) memcpy(audio_out, c->audioOut.data(), bufSize * 2 * sizeof(float));
        if (n_audio) *n_audio = (unsigned)(bufSize * 2);
        return ref_chan_meter_block(c, (unsigned)(bufSize * 2), bufSize, false, level_db, peak);
    }
    size_t audio_out_size = (size_t)ceil((double)bufSize * c->audioResampleRatio) + 512;
    c->demodOut.resize(bufSize);
    c->audioOut.resize(audio_out_size);
    lq_cf *d = c->resampled.data();
    bool autoGain = false, useSignalOutput = false;
    switch (c->modem) {
    case M_FM: case M_NBFM:                                                 // ModemFM.cpp:26-39
        freqdem_demodulate_block(c->demodFM, d, (unsigned)bufSize, c->demodOut.data());
        break;
    case M_AM:                                                              // ModemAM.cpp:29-50
        for (size_t i = 0; i < bufSize; i++) {
            float I = d[i].real, Q = d[i].imag;
            firfilt_rrrf_push(c->mDCBlock, sqrt(I * I + Q * Q));
            firfilt_rrrf_execute(c->mDCBlock, &c->demodOut[i]);
        }
        autoGain = true; useSignalOutput = true;
        break;
    case M_USB:                                                             // ModemUSB.cpp:43-64
        for (size_t i = 0; i < bufSize; i++) {
            lq_cf x, y; float discard;
            nco_crcf_step(c->ssbShift);
            nco_crcf_mix_down(c->ssbShift, d[i], &x);
            iirfilt_crcf_execute(c->ssbFilt, x, &y);
            nco_crcf_mix_up(c->ssbShift, y, &x);
            firhilbf_c2r_execute(c->c2rFilt, x, &discard, &c->demodOut[i]);
        }
        autoGain = true; useSignalOutput = true;
        break;
    case M_DSB:                                                             // ModemDSB.cpp:27-41
        for (size_t i = 0; i < bufSize; i++) ampmodem_demodulate(c->demodAM_DSB, d[i], &c->demodOut[i]);
        autoGain = true; useSignalOutput = true;
        break;
    case M_LSB:                                                             // ModemLSB.cpp:43-64
        for (size_t i = 0; i < bufSize; i++) {
            lq_cf x, y; float discard;
            nco_crcf_step(c->ssbShift);
            nco_crcf_mix_up(c->ssbShift, d[i], &x);
            iirfilt_crcf_execute(c->ssbFilt, x, &y);
            nco_crcf_mix_down(c->ssbShift, y, &x);
            firhilbf_c2r_execute(c->c2rFilt, x, &c->demodOut[i], &discard);
        }
        autoGain = true; useSignalOutput = true;
        break;
    }
    // --- ModemAnalog::buildAudioOutput ModemAnalog.cpp:67-93
    if (autoGain) {
        c->aOutputCeilMA = c->aOutputCeilMA + (c->aOutputCeil - c->aOutputCeilMA) * 0.025f;
        c->aOutputCeilMAA = c->aOutputCeilMAA + (c->aOutputCeilMA - c->aOutputCeilMAA) * 0.025f;
        c->aOutputCeil = 0;
        for (size_t i = 0; i < bufSize; i++) if (c->demodOut[i] > c->aOutputCeil) c->aOutputCeil = c->demodOut[i];
        float gain = 0.5f / c->aOutputCeilMAA;
        for (size_t i = 0; i < bufSize; i++) c->demodOut[i] *= gain;
    }
    unsigned numAudioWritten = 0;
    msresamp_rrrf_execute(c->audioResampler, c->demodOut.data(), (unsigned)bufSize, c->audioOut.data(), &numAudioWritten);
    if (demod_out) memcpy(demod_out, c->demodOut.data(), bufSize * sizeof(float));
    if (audio_out) memcpy(audio_out, c->audioOut.data(), numAudioWritten * sizeof(float));
    if (n_audio) *n_audio = numAudioWritten;

    return ref_chan_meter_block(c, numAudioWritten, bufSize, useSignalOutput, level_db, peak);
}

// test instrumentation: intermediates of the FM stereo loop of the last block (which: 0 real->complex, 1 pilot filter, 2 down-mix)
API unsigned ref_chan_fms_debug(void *h, int which, lq_cf *out, unsigned cap) {
    auto *c = (RefChan *)h;
    const std::vector<lq_cf> &v = which == 0 ? c->dbg_x : (which == 1 ? c->dbg_v : c->dbg_x2);
    unsigned n = (unsigned)std::min<size_t>(cap, v.size());
    memcpy(out, v.data(), n * sizeof(lq_cf));
    return n;
}
// DemodulatorThread::setSquelchEnabled / setSquelchLevel (:384-397: setting a level switches the squelch on)
API void ref_chan_set_squelch(void *h, int enabled, float level) {
    auto *c = (RefChan *)h;
    c->squelchEnabled = enabled != 0;
    c->squelchLevel = level;
}
// getSignalLevel / getSignalFloor / getSignalCeil (:372-382) and AudioThreadInput::is_squelch_active of the last block (:235-238)
API void ref_chan_meter(void *h, float *level_floor_ceil3, int *squelched) {
    auto *c = (RefChan *)h;
    if (level_floor_ceil3) { level_floor_ceil3[0] = c->signalLevel; level_floor_ceil3[1] = c->signalFloor; level_floor_ceil3[2] = c->signalCeil; }
    if (squelched) *squelched = c->squelched ? 1 : 0;
}


// ---------------------------------------------------------------------------------------------------------
// Audio-side consumers
//   audioCallback                         src/audio/AudioThread.cpp:88-243   (the mix of every bound AudioThread)
//   AudioFileWAV::writePayloadToFileStream src/audio/AudioFileWAV.cpp:143-165 (16-bit PCM)
//   ScopeVisualProcessor::process         src/process/ScopeVisualProcessor.cpp:60-217 (waveform + audio spectrum)
// ---------------------------------------------------------------------------------------------------------
struct RefAudioInput { int sampleRate, channels; float peak; std::vector<float> data; };
struct RefMixSource {                                   // the members of a bound AudioThread the callback touches
    std::vector<RefAudioInput *> queue;                 // inputQueue (front = oldest)
    RefAudioInput *currentInput = nullptr;
    size_t audioQueuePtr = 0;
    float gain = 1.0f;
    bool active = true;
    bool try_pop() { if (queue.empty()) { return false; } delete currentInput; currentInput = queue.front(); queue.erase(queue.begin()); return true; }
};
struct RefMixer {
    int sampleRate;
    std::vector<RefMixSource *> all;                    // by source index (what the caller holds)
    std::vector<RefMixSource *> boundThreads;           // the controller's list, in binding order (the callback's order)
};
API void *ref_mixer_create(int sampleRate) { auto *m = new RefMixer; m->sampleRate = sampleRate; return m; }
API void ref_mixer_destroy(void *h) {
    auto *m = (RefMixer *)h;
    for (auto *s : m->all) { for (auto *q : s->queue) delete q; delete s->currentInput; delete s; }
    delete m;
}
// AudioThread::setGain (AudioThread.cpp:595-604): clamped to [0, 2]
static float ref_mixer_clamp_gain(float gain_in) {
    if (gain_in < 0.0) gain_in = 0.0;
    if (gain_in > 2.0) gain_in = 2.0;
    return gain_in;
}
// AudioThread::setActive (AudioThread.cpp:557-589): becoming active binds the thread to the controller (appended: a source
// that was inactive for a while mixes LAST afterwards), becoming inactive removes it; every call flushes the inputQueue
// (currentInput and audioQueuePtr stay)
API void ref_mixer_set_active(void *h, int idx, int state_in) {
    auto *m = (RefMixer *)h; RefMixSource *s = m->all[(size_t)idx];
    const bool state = state_in != 0;
    if (state && !s->active) {
        if (std::find(m->boundThreads.begin(), m->boundThreads.end(), s) == m->boundThreads.end()) m->boundThreads.push_back(s);   // bindThread :56-63
    } else if (!state && s->active) {
        auto i = std::find(m->boundThreads.begin(), m->boundThreads.end(), s);                                                     // removeThread :65-74
        if (i != m->boundThreads.end()) m->boundThreads.erase(i);
    }
    for (auto *q : s->queue) delete q;
    s->queue.clear();
    s->active = state;
}
// a demodulator's AudioThread: constructed inactive with gain 1 (:20-27), given its gain, then activated
API int ref_mixer_add_source(void *h, float gain) {
    auto *m = (RefMixer *)h; auto *s = new RefMixSource;
    s->active = false; s->gain = ref_mixer_clamp_gain(gain);
    m->all.push_back(s);
    ref_mixer_set_active(h, (int)m->all.size() - 1, 1);
    return (int)m->all.size() - 1;
}
API void ref_mixer_set_gain(void *h, int idx, float gain) { ((RefMixer *)h)->all[(size_t)idx]->gain = ref_mixer_clamp_gain(gain); }
API void ref_mixer_push(void *h, int idx, const float *data, unsigned n, int channels, float peak, int sampleRate) {
    auto *in = new RefAudioInput; in->sampleRate = sampleRate; in->channels = channels; in->peak = peak; in->data.assign(data, data + n);
    ((RefMixer *)h)->all[(size_t)idx]->queue.push_back(in);
}
// audioCallback (AudioThread.cpp:88-243), statement by statement; `src` = the controller, `srcmix` = a bound thread
API void ref_mixer_callback(void *h, float *out, unsigned nBufferFrames) {
    auto *src = (RefMixer *)h;
    memset(out, 0, nBufferFrames * 2 * sizeof(float));
    double peak = 0.0;
    for (size_t j = 0; j < src->boundThreads.size(); j++) {
        RefMixSource *srcmix = src->boundThreads[j];
        if (srcmix->queue.empty() || !srcmix->active) continue;                       // :122 (an empty inputQueue skips the thread, whatever currentInput still holds)
        if (!srcmix->currentInput) {                                                  // :126-134
            srcmix->audioQueuePtr = 0;
            srcmix->try_pop();
            continue;
        }
        if (srcmix->currentInput->sampleRate != src->sampleRate) {                    // :136-155
            while (srcmix->try_pop()) {
                if (srcmix->currentInput && srcmix->currentInput->sampleRate == src->sampleRate) break;
                delete srcmix->currentInput; srcmix->currentInput = nullptr;
            }
            srcmix->audioQueuePtr = 0;
            if (!srcmix->currentInput) continue;
        }
        if (srcmix->currentInput->channels == 0 || srcmix->currentInput->data.empty()) {   // :158-171
            if (!srcmix->queue.empty()) {
                srcmix->audioQueuePtr = 0;
                delete srcmix->currentInput; srcmix->currentInput = nullptr;
                if (!srcmix->try_pop()) continue;
            }
            continue;
        }
        double mixPeak = srcmix->currentInput->peak * srcmix->gain;                   // :173
        if (srcmix->currentInput->channels == 1) {                                    // :175-202
            for (unsigned int i = 0; i < nBufferFrames; i++) {
                if (srcmix->audioQueuePtr >= srcmix->currentInput->data.size()) {
                    srcmix->audioQueuePtr = 0;
                    delete srcmix->currentInput; srcmix->currentInput = nullptr;
                    if (!srcmix->try_pop()) break;
                    double srcPeak = srcmix->currentInput->peak * srcmix->gain;
                    if (mixPeak < srcPeak) mixPeak = srcPeak;
                }
                if (srcmix->currentInput && !srcmix->currentInput->data.empty()) {
                    float v = srcmix->currentInput->data[srcmix->audioQueuePtr] * srcmix->gain;
                    out[i * 2] += v;
                    out[i * 2 + 1] += v;
                }
                srcmix->audioQueuePtr++;
            }
        } else {                                                                      // :203-226
            for (unsigned int i = 0, iMax = srcmix->currentInput->channels * nBufferFrames; i < iMax; i++) {
                if (srcmix->audioQueuePtr >= srcmix->currentInput->data.size()) {
                    srcmix->audioQueuePtr = 0;
                    delete srcmix->currentInput; srcmix->currentInput = nullptr;
                    if (!srcmix->try_pop()) break;
                    double srcPeak = srcmix->currentInput->peak * srcmix->gain;
                    if (mixPeak < srcPeak) mixPeak = srcPeak;
                }
                if (srcmix->currentInput && !srcmix->currentInput->data.empty())
                    out[i] = out[i] + srcmix->currentInput->data[srcmix->audioQueuePtr] * srcmix->gain;
                srcmix->audioQueuePtr++;
            }
        }
        peak += mixPeak;                                                              // :228
    }
    if (peak > 1.0) {                                                                 // :231-238
        float invPeak = (float)(1.0 / peak);
        for (unsigned int i = 0; i < nBufferFrames * 2; i++) out[i] *= invPeak;
    }
}
// AudioFileWAV::writePayloadToFileStream (:143-165): the 16-bit little-endian words it writes for data[start, end)
API unsigned ref_wav_payload(const float *data, unsigned start, unsigned end, int channels, float peak, short *out) {
    float intScale = (peak < 1.0) ? 32767.0f : (32767.0f / peak);
    unsigned k = 0;
    if (channels == 1) {
        for (size_t i = start, iMax = end; i < iMax; i++) out[k++] = (short)int(data[i] * intScale);
    } else if (channels == 2) {
        for (size_t i = start, iMax = end / 2; i < iMax; i++) { out[k++] = (short)int(data[i * 2] * intScale); out[k++] = (short)int(data[i * 2 + 1] * intScale); }
    }
    return k;
}
// ScopeVisualProcessor (:14-217): waveform (type 0, SCOPE_MODE_Y) and the audio spectrum with its running averages
struct RefScope {
    unsigned fftSize = 0, maxScopeSamples = 1024;
    double fft_average_rate = 0.65f, fft_ceil_ma = 0, fft_ceil_maa = 0, fft_floor_ma = 0, fft_floor_maa = 0;
    std::vector<lq_cf> fftInData, fftOutput;
    std::vector<double> fft_result, fft_result_ma, fft_result_maa;
    fftplan fftPlan = nullptr;
};
API void *ref_scope_create(unsigned fftSize, unsigned maxScopeSamples) {
    auto *s = new RefScope; s->fftSize = fftSize; s->maxScopeSamples = maxScopeSamples;
    s->fftInData.resize(fftSize); s->fftOutput.resize(fftSize);
    s->fftPlan = fft_create_plan(fftSize, s->fftInData.data(), s->fftOutput.data(), +1, 0);      // :23-34
    return s;
}
API void ref_scope_destroy(void *h) { auto *s = (RefScope *)h; if (s->fftPlan) fft_destroy_plan(s->fftPlan); delete s; }
// one AudioThreadInput through process(): waveform points (x, y interleaved; returns their count / 2) and spectrum points
API unsigned ref_scope_process(void *h, const float *data, unsigned n, int channels, int sampleRate, int inputRate,
                               float *waveform_xy, float *spectrum_xy, unsigned *n_spec, double *floor_ceil2) {
    auto *s = (RefScope *)h;
    size_t i, iMax = n;
    if (iMax > s->maxScopeSamples) iMax = s->maxScopeSamples;                                    // :73-76
    float peak = 1.0f;
    for (i = 0; i < iMax; i++) { float p = fabs(data[i]); if (p > peak) peak = p; }              // :87-94
    for (i = 0; i < iMax; i++) {                                                                  // :119-122 (type 0)
        waveform_xy[i * 2] = (float)((((double)i / (double)iMax) - 0.5) * 2.0);
        waveform_xy[i * 2 + 1] = data[i] / peak;
    }
    const unsigned n_wave = (unsigned)iMax;
    const unsigned fftSize = s->fftSize;
    iMax = n;
    if (channels == 1) {                                                                          // :133-142
        for (i = 0; i < fftSize; i++) { s->fftInData[i].real = i < iMax ? data[i] : 0; s->fftInData[i].imag = 0; }
    } else {                                                                                      // :143-153
        iMax = iMax / 2;
        for (i = 0; i < fftSize; i++) { s->fftInData[i].real = i < iMax ? data[i] + data[iMax + i] : 0; s->fftInData[i].imag = 0; }
    }
    double fft_ceil = 0, fft_floor = 1;
    if (s->fft_result.size() < (fftSize / 2)) { s->fft_result.resize(fftSize / 2); s->fft_result_ma.resize(fftSize / 2); s->fft_result_maa.resize(fftSize / 2); }
    fft_execute(s->fftPlan);
    for (i = 0; i < (fftSize / 2); i++) { double a = (double)s->fftOutput[i].real, b = (double)s->fftOutput[i].imag; s->fft_result[i] = sqrt(a * a + b * b); }
    for (i = 0; i < (fftSize / 2); i++) {                                                         // :182-192
        s->fft_result_ma[i] += (s->fft_result[i] - s->fft_result_ma[i]) * s->fft_average_rate;
        s->fft_result_maa[i] += (s->fft_result_ma[i] - s->fft_result_maa[i]) * s->fft_average_rate;
        if (s->fft_result_maa[i] > fft_ceil) fft_ceil = s->fft_result_maa[i];
        if (s->fft_result_maa[i] < fft_floor) fft_floor = s->fft_result_maa[i];
    }
    s->fft_ceil_ma = s->fft_ceil_ma + (fft_ceil - s->fft_ceil_ma) * 0.05;
    s->fft_ceil_maa = s->fft_ceil_maa + (s->fft_ceil_ma - s->fft_ceil_maa) * 0.05;
    s->fft_floor_ma = s->fft_floor_ma + (fft_floor - s->fft_floor_ma) * 0.05;
    s->fft_floor_maa = s->fft_floor_maa + (s->fft_floor_ma - s->fft_floor_maa) * 0.05;
    unsigned int outSize = fftSize / 2;
    if (sampleRate != inputRate) outSize = (int)floor((float)outSize * ((float)sampleRate / (float)inputRate));
    for (i = 0; i < outSize; i++) {                                                               // :207-211
        float v = (log10(s->fft_result_maa[i] + 0.25 - (s->fft_floor_maa - 0.75)) / log10((s->fft_ceil_maa + 0.25) - (s->fft_floor_maa - 0.75)));
        spectrum_xy[i * 2] = (float)((double)i / (double)(outSize));
        spectrum_xy[i * 2 + 1] = v;
    }
    if (n_spec) *n_spec = outSize;
    if (floor_ceil2) { floor_ceil2[0] = s->fft_floor_maa; floor_ceil2[1] = s->fft_ceil_maa; }
    return n_wave;
}

// ---------------------------------------------------------------------------------------------------------
// FFTDataDistributor framing             src/process/FFTDataDistributor.cpp:28-144
// ---------------------------------------------------------------------------------------------------------
struct RefDistrib {
    unsigned fftSize; unsigned linesPerSecond; double lineRateAccum;
    std::vector<lq_cf> buf; size_t bufferMax, bufferOffset, bufferedItems; long long sampleRate; long long frequency = 0;
};
API void *ref_distrib_create(unsigned fftSize, unsigned lps) {
    auto *d = new RefDistrib(); d->fftSize = fftSize; d->linesPerSecond = lps; d->lineRateAccum = 0;
    d->bufferMax = d->bufferOffset = d->bufferedItems = 0; d->sampleRate = 0; return d;
}
API void ref_distrib_destroy(void *h) { delete (RefDistrib *)h; }
API void ref_distrib_set_fft_size(void *h, unsigned fftSize) { ((RefDistrib *)h)->fftSize = fftSize; }   // FFTDataDistributor.cpp:15-18
// returns number of frames written to frames_out (each fftSize samples)
API unsigned ref_distrib_push(void *h, const lq_cf *in, unsigned n, long long sampleRate, long long frequency, lq_cf *frames_out, unsigned cap_frames) {
    auto *d = (RefDistrib *)h;
    unsigned nf = 0;
    if (d->sampleRate != sampleRate || d->frequency != frequency) {         // :41-53
        d->frequency = frequency;
        d->bufferMax = std::max((size_t)(sampleRate * 0.250), (size_t)(1.2 * d->fftSize));
        d->bufferOffset = 0; d->bufferedItems = 0; d->sampleRate = sampleRate; d->buf.resize(d->bufferMax);
    }
    if (d->bufferMax < (size_t)(1.2 * d->fftSize)) { d->bufferMax = (size_t)(1.2 * d->fftSize); d->buf.resize(d->bufferMax); }
    size_t add = n;
    if (d->bufferOffset + d->bufferedItems + n > d->bufferMax) {            // :66-76
        memmove(&d->buf[0], &d->buf[d->bufferOffset], d->bufferedItems * sizeof(lq_cf));
        d->bufferOffset = 0;
        if (d->bufferedItems + n > d->bufferMax) add = d->bufferMax - d->bufferedItems;
    }
    memcpy(&d->buf[d->bufferOffset + d->bufferedItems], in, add * sizeof(lq_cf));
    d->bufferedItems += add;
    double inputTime = (double)d->bufferedItems / (double)d->sampleRate;    // :89-97
    double inputLines = (double)d->bufferedItems / (double)d->fftSize;
    double lineRateStep = ((double)d->linesPerSecond * inputTime) / inputLines;
    if (d->bufferedItems >= d->fftSize) {                                   // :100-142
        size_t numProcessed = 0;
        if (d->lineRateAccum + (lineRateStep * ((double)d->bufferedItems / (double)d->fftSize)) < 1.0) {
            d->lineRateAccum += (lineRateStep * ((double)d->bufferedItems / (double)d->fftSize));
            numProcessed = d->bufferedItems;
        } else {
            for (size_t i = 0, iMax = d->bufferedItems; i < iMax; i += d->fftSize) {
                if ((i + d->fftSize) > iMax) break;
                d->lineRateAccum += lineRateStep;
                if (d->lineRateAccum >= 1.0) {
                    if (nf < cap_frames) memcpy(frames_out + (size_t)nf * d->fftSize, &d->buf[d->bufferOffset + i], d->fftSize * sizeof(lq_cf));
                    nf++;
                    while (d->lineRateAccum >= 1.0) d->lineRateAccum -= 1.0;
                }
                numProcessed += d->fftSize;
            }
        }
        if (numProcessed) { d->bufferedItems -= numProcessed; d->bufferOffset += numProcessed; }
        if (d->bufferedItems <= 0) { d->bufferedItems = 0; d->bufferOffset = 0; }
    }
    return nf;
}

// ---------------------------------------------------------------------------------------------------------
// SpectrumVisualProcessor, non-view branch   src/process/SpectrumVisualProcessor.cpp:10-45,140-178,388-452,494-576
// + WaterfallPanel::step u8 rule             src/panel/WaterfallPanel.cpp:51-83
// ---------------------------------------------------------------------------------------------------------
struct RefSpectrum {
    unsigned fftSize, fftSizeInternal;
    std::vector<lq_cf> fftInput, fftOutput;
    fftplan plan;
    double fft_ceil_ma, fft_ceil_maa, fft_floor_ma, fft_floor_maa;
    float fft_average_rate, scaleFactor;
    std::vector<double> fft_result, fft_result_ma, fft_result_maa, fft_result_temp;
    // view (zoom) state: SpectrumVisualProcessor.cpp:10-41 constructor values
    bool is_view = false, lastView = false;
    long long centerFreq = 0; long bandwidth = 0;
    nco_crcf freqShifter = nullptr; long shiftFrequency = 0;
    long lastInputBandwidth = 0, lastBandwidth = 0;
    msresamp_crcf resampler = nullptr; double resamplerRatio = 0;
    unsigned lastDataSize = 0, desiredInputSize = 0;
    std::vector<lq_cf> fftInData, fftLastData, shiftBuffer, resampleBuffer;
};
API void *ref_spectrum_create(unsigned fftSize) {
    auto *s = new RefSpectrum();
    s->fftSize = fftSize; s->fftSizeInternal = fftSize * 2;                 // :145 (SPECTRUM_VZM = 2)
    s->fftInput.assign(s->fftSizeInternal, lq_cf{0, 0}); s->fftOutput.assign(s->fftSizeInternal, lq_cf{0, 0});
    s->plan = fft_create_plan(s->fftSizeInternal, s->fftInput.data(), s->fftOutput.data(), +1, 0);  // :177
    s->fft_ceil_ma = s->fft_ceil_maa = 100.0; s->fft_floor_ma = s->fft_floor_maa = 0.0;           // :32-33
    s->fft_average_rate = 0.65f; s->scaleFactor = 1.0f;                                           // :36-37
    s->fft_result.assign(s->fftSizeInternal, 0.0); s->fft_result_ma.assign(s->fftSizeInternal, 0.0);
    s->fft_result_maa.assign(s->fftSizeInternal, 0.0);                                            // :252-264
    s->fft_result_temp.assign(s->fftSizeInternal, 0.0);
    s->freqShifter = nco_crcf_create(0);                                                          // :28
    s->fftInData.assign(s->fftSizeInternal, lq_cf{0, 0}); s->fftLastData.assign(s->fftSizeInternal, lq_cf{0, 0});   // setup() :166-175
    return s;
}
API void ref_spectrum_destroy(void *h) {
    auto *s = (RefSpectrum *)h;
    fft_destroy_plan(s->plan); nco_crcf_destroy(s->freqShifter);
    if (s->resampler) msresamp_crcf_destroy(s->resampler);
    delete s;
}
// setView(bView, centerFreq, bandwidth) :64-71
API void ref_spectrum_set_view(void *h, int is_view, long long centerFreq, long bandwidth) {
    auto *s = (RefSpectrum *)h;
    s->is_view = is_view != 0; s->centerFreq = centerFreq; s->bandwidth = bandwidth;
}
// the frame length FFTVisualDataThread gives the distributor for this view (:283-297, FFTVisualDataThread.cpp:55-61)
API unsigned ref_spectrum_desired_input(void *h, long long sampleRate) {
    auto *s = (RefSpectrum *)h;
    if (!s->is_view) return s->fftSizeInternal;
    long resampleBw = sampleRate;
    while (resampleBw / 2 >= (long)s->bandwidth) resampleBw /= 2;
    double ratio = (double)(resampleBw) / (double)sampleRate;
    return (unsigned)(size_t)(s->fftSizeInternal / ratio);
}
API void ref_spectrum_set_average_rate(void *h, float r) { ((RefSpectrum *)h)->fft_average_rate = r; }

// FFT + magnitudes + averaging + points for the frame in s->fftInput (SpectrumVisualProcessor.cpp:437-576).  `rebin` != 0:
// a new resampler was made in view mode (zoom step, :454-492), with bwDiff's sign in it.  visualRatio = bandwidth / resampleBw
// (:532; 1 when not in view mode).
static void ref_spectrum_execute(RefSpectrum *s, int rebin, double visualRatio, float *points_y, unsigned char *wf_line,
                                 double *fft_result_out, double *fft_maa_out, double *ceil_floor_out) {
    const unsigned N = s->fftSizeInternal;
    float fft_ceil = 0, fft_floor = 1;                                      // :437
    fft_execute(s->plan);                                                   // :439
    for (int i = 0, iMax = N / 2; i < iMax; i++) {                          // :441-452
        float a = s->fftOutput[i].real, b = s->fftOutput[i].imag;
        float c = sqrt(a * a + b * b);
        float x = s->fftOutput[N / 2 + i].real, y = s->fftOutput[N / 2 + i].imag;
        float z = sqrt(x * x + y * y);
        s->fft_result[i] = (z);
        s->fft_result[N / 2 + i] = (c);
    }
    if (rebin) {                                                            // :454-492 (newResampler && lastView)
        std::vector<double> &ma = s->fft_result_ma, &maa = s->fft_result_maa, &tmp = s->fft_result_temp;
        if (rebin < 0) {                                                    // zoomed in: every old bin of the middle half becomes two
            for (unsigned i = 0; i < N; i++) tmp[i] = ma[(N / 4) + (i / 2)];
            for (unsigned i = 0; i < N; i++) { ma[i] = tmp[i]; tmp[i] = maa[(N / 4) + (i / 2)]; }
            for (unsigned i = 0; i < N; i++) maa[i] = tmp[i];
        } else {                                                            // zoomed out: every second old bin, zeros outside
            for (size_t i = 0; i < N; i++) {
                if (i < N / 4) tmp[i] = 0; else if (i >= N - N / 4) tmp[i] = 0; else tmp[i] = ma[(i - N / 4) * 2];
            }
            for (unsigned i = 0; i < N; i++) {
                ma[i] = tmp[i];
                if (i < N / 4) tmp[i] = 0; else if (i >= N - N / 4) tmp[i] = 0; else tmp[i] = maa[(i - N / 4) * 2];
            }
            for (unsigned i = 0; i < N; i++) maa[i] = tmp[i];
        }
    }
    for (int i = 0, iMax = N; i < iMax; i++) {                              // :494-511
        if (s->fft_result_maa[i] != s->fft_result_maa[i]) s->fft_result_maa[i] = s->fft_result[i];
        s->fft_result_maa[i] += (s->fft_result_ma[i] - s->fft_result_maa[i]) * s->fft_average_rate;
        if (s->fft_result_ma[i] != s->fft_result_ma[i]) s->fft_result_ma[i] = s->fft_result[i];
        s->fft_result_ma[i] += (s->fft_result[i] - s->fft_result_ma[i]) * s->fft_average_rate;
        if (s->fft_result_maa[i] > fft_ceil || fft_ceil != fft_ceil) fft_ceil = s->fft_result_maa[i];
        if (s->fft_result_maa[i] < fft_floor || fft_floor != fft_floor) fft_floor = s->fft_result_maa[i];
    }
    if (s->fft_ceil_ma != s->fft_ceil_ma) s->fft_ceil_ma = fft_ceil;         // :513-521
    s->fft_ceil_ma = s->fft_ceil_ma + (fft_ceil - s->fft_ceil_ma) * 0.05;
    if (s->fft_ceil_maa != s->fft_ceil_maa) s->fft_ceil_maa = fft_ceil;
    s->fft_ceil_maa = s->fft_ceil_maa + (s->fft_ceil_ma - s->fft_ceil_maa) * 0.05;
    if (s->fft_floor_ma != s->fft_floor_ma) s->fft_floor_ma = fft_floor;
    s->fft_floor_ma = s->fft_floor_ma + (fft_floor - s->fft_floor_ma) * 0.05;
    if (s->fft_floor_maa != s->fft_floor_maa) s->fft_floor_maa = fft_floor;
    s->fft_floor_maa = s->fft_floor_maa + (s->fft_floor_ma - s->fft_floor_maa) * 0.05;

    float sf = s->scaleFactor;                                              // :532-576
    double visualStart = (double(N) / 2.0) - (double(N) * (visualRatio / 2.0));
    double visualAccum = 0, acc = 0, accCount = 0, i = 0;
    double point_ceil = s->fft_ceil_maa, point_floor = s->fft_floor_maa;
    for (int x = 0, xMax = s->fftSize; x < xMax; x++) {
        visualAccum += visualRatio * double(2);
        while (visualAccum >= 1.0) {
            unsigned int idx = round(visualStart + i);
            if (idx > 0 && idx < N) acc += s->fft_result_maa[idx]; else acc += s->fft_floor_maa;
            accCount += 1.0; visualAccum -= 1.0; i++;
        }
        if (accCount) {
            float yv = ((log10((acc / accCount) + 0.25 - (point_floor - 0.75)) / log10((point_ceil + 0.25) - (point_floor - 0.75)))) * sf;
            if (points_y) points_y[x] = yv;
            if (wf_line) {                                                  // WaterfallPanel.cpp:64-70
                float v = yv;
                float wv = v < 0 ? 0 : (v > 0.99 ? 0.99 : v);
                wf_line[x] = (unsigned char)floor(wv * 255.0);
            }
            acc = 0.0; accCount = 0.0;
        }
    }
    if (fft_result_out) memcpy(fft_result_out, s->fft_result.data(), N * sizeof(double));
    if (fft_maa_out) memcpy(fft_maa_out, s->fft_result_maa.data(), N * sizeof(double));
    if (ceil_floor_out) { ceil_floor_out[0] = point_ceil; ceil_floor_out[1] = point_floor; }
}

// Non-view frame: >= fftSizeInternal samples (n given); outputs: points_y[fftSize], u8 line[fftSize], optional copies of
// fft_result (linear magnitudes, fftshifted) and fft_result_maa, and the ceil/floor pair used for this frame.
API void ref_spectrum_frame(void *h, const lq_cf *frame, unsigned n, float *points_y, unsigned char *wf_line,
                            double *fft_result_out, double *fft_maa_out, double *ceil_floor_out) {
    auto *s = (RefSpectrum *)h;
    const unsigned N = s->fftSizeInternal;
    // :388-403 (non-view): copy first N samples, zero-pad a short frame (priming path not modelled: n >= N expected)
    if (n < N) { memcpy(s->fftInput.data(), frame, n * sizeof(lq_cf)); memset(&s->fftInput[n], 0, (N - n) * sizeof(lq_cf)); }
    else memcpy(s->fftInput.data(), frame, N * sizeof(lq_cf));
    ref_spectrum_execute(s, 0, 1.0, points_y, wf_line, fft_result_out, fft_maa_out, ceil_floor_out);
    s->lastView = false;                                                    // :636
}

// One SpectrumVisualProcessor::process() call in view mode (:283-386, :400-420, :437-636): `frame` is what the distributor
// handed over (n samples taken at centre `frequency`, rate `sampleRate`).  Returns 1 if a transform was executed.
API int ref_spectrum_view_frame(void *h, const lq_cf *frame, unsigned n, long long frequency, long long sampleRate, float *points_y,
                                unsigned char *wf_line, double *fft_result_out, double *fft_maa_out, double *ceil_floor_out) {
    auto *s = (RefSpectrum *)h;
    const unsigned N = s->fftSizeInternal;
    if (!s->is_view || !sampleRate || !n) return 0;
    unsigned int num_written;
    long resampleBw = sampleRate;
    bool newResampler = false;
    int bwDiff = 0;
    while (resampleBw / 2 >= (long)s->bandwidth) resampleBw /= 2;            // :289-291 (SPECTRUM_VZM = 2)
    s->resamplerRatio = (double)(resampleBw) / (double)sampleRate;
    size_t desired_input_size = N / s->resamplerRatio;
    s->desiredInputSize = desired_input_size;
    if (n < desired_input_size) desired_input_size = n;                     // :299-302
    if (s->centerFreq != frequency) {                                       // :304-347
        if ((s->centerFreq - frequency) != s->shiftFrequency || s->lastInputBandwidth != sampleRate) {
            if (std::llabs(frequency - s->centerFreq) < (sampleRate / 2)) {  // wxGetApp().getSampleRate(): the device rate
                long lastShiftFrequency = s->shiftFrequency;
                s->shiftFrequency = s->centerFreq - frequency;
                nco_crcf_set_frequency(s->freqShifter, (2.0 * M_PI) * (((double)labs(s->shiftFrequency)) / ((double)sampleRate)));
                long freqDiff = s->shiftFrequency - lastShiftFrequency;
                if (s->lastBandwidth != 0) {
                    double binPerHz = double(s->lastBandwidth) / double(N);
                    unsigned int numShift = floor(double(labs(freqDiff)) / binPerHz);
                    if (numShift < N / 2 && numShift) {
                        if (freqDiff > 0) {
                            memmove(&s->fft_result_ma[0], &s->fft_result_ma[numShift], (N - numShift) * sizeof(double));
                            memmove(&s->fft_result_maa[0], &s->fft_result_maa[numShift], (N - numShift) * sizeof(double));
                        } else {
                            memmove(&s->fft_result_ma[numShift], &s->fft_result_ma[0], (N - numShift) * sizeof(double));
                            memmove(&s->fft_result_maa[numShift], &s->fft_result_maa[0], (N - numShift) * sizeof(double));
                        }
                    }
                }
            }
        }
        s->shiftBuffer.resize(desired_input_size);
        std::vector<lq_cf> in(frame, frame + desired_input_size);
        if (s->shiftFrequency < 0) nco_crcf_mix_block_up(s->freqShifter, in.data(), s->shiftBuffer.data(), (unsigned)desired_input_size);
        else nco_crcf_mix_block_down(s->freqShifter, in.data(), s->shiftBuffer.data(), (unsigned)desired_input_size);
    } else {
        s->shiftBuffer.assign(frame, frame + desired_input_size);
    }
    if (!s->resampler || resampleBw != s->lastBandwidth || s->lastInputBandwidth != sampleRate) {   // :354-368
        if (s->resampler) msresamp_crcf_destroy(s->resampler);
        s->resampler = msresamp_crcf_create((float)s->resamplerRatio, 60.0f);
        bwDiff = resampleBw - s->lastBandwidth;
        s->lastBandwidth = resampleBw;
        s->lastInputBandwidth = sampleRate;
        newResampler = true;
    }
    unsigned int out_size = ceil((double)(desired_input_size)*s->resamplerRatio) + 512;
    s->resampleBuffer.resize(out_size);
    msresamp_crcf_execute(s->resampler, s->shiftBuffer.data(), (unsigned)desired_input_size, s->resampleBuffer.data(), &num_written);
    lq_cf *fftInData = s->fftInData.data(), *fftLastData = s->fftLastData.data(), *fftInput = s->fftInput.data();
    if (num_written < N) {                                                  // :380-385
        memcpy(fftInData, s->resampleBuffer.data(), num_written * sizeof(lq_cf));
        memset(&(fftInData[num_written]), 0, (N - num_written) * sizeof(lq_cf));
    } else memcpy(fftInData, s->resampleBuffer.data(), N * sizeof(lq_cf));
    bool execute = false;                                                   // :398-420
    if (num_written >= N) {
        execute = true;
        memcpy(fftInput, fftInData, N * sizeof(lq_cf));
        memcpy(fftLastData, fftInput, N * sizeof(lq_cf));
    } else {
        if (s->lastDataSize + num_written < N) {                            // priming
            unsigned int num_copy = N - s->lastDataSize;
            if (num_written > num_copy) num_copy = num_written;
            memcpy(fftLastData, fftInData, num_copy * sizeof(lq_cf));
            s->lastDataSize += num_copy;
        } else {
            unsigned int num_last = (N - num_written);
            memcpy(fftInput, fftLastData + (s->lastDataSize - num_last), num_last * sizeof(lq_cf));
            memcpy(fftInput + num_last, fftInData, num_written * sizeof(lq_cf));
            memcpy(fftLastData, fftInput, N * sizeof(lq_cf));
            execute = true;
        }
    }
    if (execute) ref_spectrum_execute(s, (newResampler && s->lastView) ? (bwDiff < 0 ? -1 : 1) : 0, (double(s->bandwidth) / double(resampleBw)), points_y, wf_line,
                                      fft_result_out, fft_maa_out, ceil_floor_out);
    s->lastView = s->is_view;                                               // :636
    return execute ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------------
// CPU baseline: whole path, threaded the way the reference is (one ingress, one per channel, one spectrum;
// here the channel work is spread over `threads` workers).  Returns seconds for `blocks` blocks of n samples
// after `warm_blocks` untimed ones.
// ---------------------------------------------------------------------------------------------------------
API double ref_bench_path(const lq_cf *iq, unsigned n, unsigned warm_blocks, unsigned blocks, long long sampleRate,
                          unsigned n_chan, const char *const *modem_names, const long long *bandwidths, const long long *shifts,
                          int audioRate, unsigned fftSize /*0 = no spectrum*/, unsigned threads, double *audio_checksum) {
    std::vector<void *> chans(n_chan);
    for (unsigned c = 0; c < n_chan; c++) chans[c] = ref_chan_create(modem_names[c], sampleRate, bandwidths[c], audioRate);
    void *ing = ref_ingress_create();
    void *spec = fftSize ? ref_spectrum_create(fftSize) : nullptr;
    void *dist = fftSize ? ref_distrib_create(fftSize * 2, 1000000) : nullptr;
    std::vector<lq_cf> dc(n);
    std::vector<lq_cf> frames;
    unsigned cap_frames = fftSize ? (n / (fftSize * 2) + 2) : 0;
    if (fftSize) frames.resize((size_t)cap_frames * fftSize * 2);
    std::vector<float> pts(fftSize ? fftSize : 1);
    std::vector<unsigned char> line(fftSize ? fftSize : 1);
    if (threads < 1) threads = 1;
    std::vector<double> sums(threads, 0.0);
    auto t0 = std::chrono::steady_clock::now();
    for (unsigned b = 0; b < warm_blocks + blocks; b++) {
        if (b == warm_blocks) t0 = std::chrono::steady_clock::now();   // first-touch page faults and cold filters stay outside
        ref_ingress_block(ing, const_cast<lq_cf *>(iq), n, dc.data());
        std::atomic<unsigned> next(0);
        auto work = [&](unsigned tid) {
            std::vector<float> audio(8192);
            for (;;) {
                unsigned c = next.fetch_add(1);
                if (c >= n_chan) break;
                auto *rc = (RefChan *)chans[c];
                size_t need = (size_t)ceil((double)n * rc->iqResampleRatio * rc->audioResampleRatio) + 2048;
                if (audio.size() < need) audio.resize(need);
                unsigned na = 0;
                ref_chan_process(chans[c], dc.data(), n, shifts[c], nullptr, nullptr, nullptr, audio.data(), &na, nullptr, nullptr);
                double s = 0; for (unsigned i = 0; i < na; i++) s += audio[i];
                sums[tid] += s;
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < threads; t++) pool.emplace_back(work, t);
        if (spec) {   // the spectrum thread's work rides on the calling thread, then it joins the channel pool
            unsigned nf = ref_distrib_push(dist, dc.data(), n, sampleRate, 0, frames.data(), cap_frames);
            for (unsigned f = 0; f < nf && f < cap_frames; f++)
                ref_spectrum_frame(spec, frames.data() + (size_t)f * fftSize * 2, fftSize * 2, pts.data(), line.data(), nullptr, nullptr, nullptr);
        }
        work(0);
        for (auto &t : pool) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    double tot = 0; for (double s : sums) tot += s;
    if (audio_checksum) *audio_checksum = tot;
    for (auto c : chans) ref_chan_destroy(c);
    ref_ingress_destroy(ing);
    if (spec) { ref_spectrum_destroy(spec); ref_distrib_destroy(dist); }
    return std::chrono::duration<double>(t1 - t0).count();
}

// The same benchmark through the reference's channelizer path (SDRPostThread::runPFBCH + runDemodChannels, mode 1, or
// runPFBCH2, mode 2): the filterbank runs on the calling thread like SDRPostThread, the spectrum sees the raw block, every
// demodulator (created at the channel rate) reads its de-interleaved channel; channel 0 goes through the DC blocker.
// freqs are absolute demodulator frequencies, `center` the tuner frequency.
API double ref_bench_pfb_path(const lq_cf *iq, unsigned n, unsigned warm_blocks, unsigned blocks, long long sampleRate, long long center,
                              int mode, unsigned n_chan, const char *const *modem_names, const long long *bandwidths,
                              const long long *freqs, int audioRate, unsigned fftSize, unsigned threads, double *audio_checksum) {
    const int M = ref_optimal_channel_count(sampleRate);
    if (M < 2 || n % (unsigned)M) return -1.0;
    const long long chanBw = sampleRate / M;
    const int osf = mode == 2 ? 2 : 1;
    const long long demodRate = chanBw * osf;
    std::vector<void *> chans(n_chan);
    for (unsigned c = 0; c < n_chan; c++) chans[c] = ref_chan_create(modem_names[c], demodRate, bandwidths[c], audioRate);
    void *pfb = osf == 2 ? ref_pfb2_create((unsigned)M, 4, 60.0f) : ref_pfb_create((unsigned)M, 4, 60.0f);
    void *dcf = ref_ingress_create();
    void *spec = fftSize ? ref_spectrum_create(fftSize) : nullptr;
    void *dist = fftSize ? ref_distrib_create(fftSize * 2, 1000000) : nullptr;
    // updateChannels (:115-124) and getChannelAt (:127-139)
    std::vector<long long> centers((size_t)M + 1);
    for (int i = 0; i < M / 2; i++) { int ofs = (int)(chanBw * i); centers[i] = center + ofs; centers[i + M / 2] = center - (sampleRate / 2) + ofs; }
    centers[M] = center + (sampleRate / 2);
    std::vector<int> pick(n_chan, -1);
    for (unsigned c = 0; c < n_chan; c++) {
        long long minDelta = sampleRate;
        for (int i = 0; i < M + 1; i++) { long long fd = std::llabs(freqs[c] - centers[i]); if (fd < minDelta) { minDelta = fd; pick[c] = i; } }
    }
    const size_t chanDataSize = (size_t)osf * n / M;
    std::vector<lq_cf> dataOut((size_t)osf * n);
    std::vector<std::vector<lq_cf>> chanData((size_t)M + 1);
    std::vector<char> active((size_t)M + 1, 0);
    for (unsigned c = 0; c < n_chan; c++) if (pick[c] >= 0) active[pick[c]] = 1;
    std::vector<lq_cf> dcBuf(chanDataSize), frames;
    unsigned cap_frames = fftSize ? (n / (fftSize * 2) + 2) : 0;
    if (fftSize) frames.resize((size_t)cap_frames * fftSize * 2);
    std::vector<float> pts(fftSize ? fftSize : 1);
    std::vector<unsigned char> line(fftSize ? fftSize : 1);
    if (threads < 1) threads = 1;
    std::vector<double> sums(threads, 0.0);
    auto t0 = std::chrono::steady_clock::now();
    for (unsigned b = 0; b < warm_blocks + blocks; b++) {
        if (b == warm_blocks) t0 = std::chrono::steady_clock::now();
        if (osf == 2) ref_pfb2_execute(pfb, const_cast<lq_cf *>(iq), n, dataOut.data());
        else ref_pfb_execute(pfb, const_cast<lq_cf *>(iq), n, dataOut.data());
        for (int i = 0; i < M + 1; i++) {                                   // runDemodChannels :333-397
            if (!active[i]) continue;
            chanData[i].resize(chanDataSize);
            int idx = (i == M) ? M / 2 : i;
            if (i == 0) {
                for (size_t j = 0; j < chanDataSize; j++) { dcBuf[j] = dataOut[idx]; idx += M; }
                ref_ingress_block(dcf, dcBuf.data(), (unsigned)chanDataSize, chanData[i].data());
            } else {
                for (size_t j = 0; j < chanDataSize; j++) { chanData[i][j] = dataOut[idx]; idx += M; }
            }
        }
        std::atomic<unsigned> next(0);
        auto work = [&](unsigned tid) {
            std::vector<float> audio(8192);
            for (;;) {
                unsigned c = next.fetch_add(1);
                if (c >= n_chan) break;
                if (pick[c] < 0) continue;
                auto *rc = (RefChan *)chans[c];
                size_t need = (size_t)ceil((double)chanDataSize * rc->iqResampleRatio * rc->audioResampleRatio) + 2048;
                if (audio.size() < need) audio.resize(need);
                unsigned na = 0;
                ref_chan_process(chans[c], chanData[pick[c]].data(), (unsigned)chanDataSize, freqs[c] - centers[pick[c]], nullptr, nullptr,
                                 nullptr, audio.data(), &na, nullptr, nullptr);
                double s = 0; for (unsigned i = 0; i < na; i++) s += audio[i];
                sums[tid] += s;
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < threads; t++) pool.emplace_back(work, t);
        if (spec) {
            unsigned nf = ref_distrib_push(dist, const_cast<lq_cf *>(iq), n, sampleRate, center, frames.data(), cap_frames);
            for (unsigned f = 0; f < nf && f < cap_frames; f++)
                ref_spectrum_frame(spec, frames.data() + (size_t)f * fftSize * 2, fftSize * 2, pts.data(), line.data(), nullptr, nullptr, nullptr);
        }
        work(0);
        for (auto &t : pool) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    double tot = 0; for (double s : sums) tot += s;
    if (audio_checksum) *audio_checksum = tot;
    for (auto c : chans) ref_chan_destroy(c);
    if (osf == 2) ref_pfb2_destroy(pfb); else ref_pfb_destroy(pfb);
    ref_ingress_destroy(dcf);
    if (spec) { ref_spectrum_destroy(spec); ref_distrib_destroy(dist); }
    return std::chrono::duration<double>(t1 - t0).count();
}

