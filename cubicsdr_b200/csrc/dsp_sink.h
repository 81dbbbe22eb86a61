// cubicsdr_b200/csrc/dsp_sink.h -- audio-side consumers on the device: what the reference does with the
// demodulators' audio AFTER DemodulatorThread, kept on the GPU so that hundreds of channels need no host fan-in.
//
//   mixer   audioCallback                          src/audio/AudioThread.cpp:88-243   the RtAudio callback's mix of every bound thread
//   pcm16   AudioFileWAV::writePayloadToFileStream src/audio/AudioFileWAV.cpp:143-165 16-bit PCM words of a recording
//   scope   ScopeVisualProcessor::process          src/process/ScopeVisualProcessor.cpp:60-217 waveform points + audio spectrum
//
// A mixer source is one bound AudioThread: its inputQueue (entries = the AudioThreadInput of one block each), currentInput and
// audioQueuePtr live in HBM; the samples sit in a per-source ring the feed kernel fills from the channel's result slot after every
// block.  Rendering n frames is the callback, statement by statement: a serial walk per source that decides which sample every
// output slot takes (phase A, one thread per source), the sum over the sources IN BINDING ORDER per output slot (phase B: float
// additions are not associative, the order is the reference's: MixArgs::order is the controller's boundThreads, which
// AudioThread::setActive edits), the peak sum and the normalisation (phases C, D).
#pragma once
#include "dsp_core.h"

namespace sdrb200 {

constexpr int kMixEntries = 100;           // AudioThreadInput blocks queued per source: the reference's inputQueue holds 100 (pipeAudioData->
                                           // set_max_num_items(100), DemodulatorInstance.cpp:72-73); a full queue drops the new block, like try_push
constexpr int kMixMaxSources = 1024;

struct MixEntry { int start, n, channels, rate; float peak; };
struct MixSource {
    int chan;                      // engine channel slot whose result slot feeds this source; -1 = fed by sdr_mixer_push only
    int feed_rate;                 // that channel's audio rate (AudioThreadInput::sampleRate of what it produces)
    float gain;
    int active;
    int ring_off, ring_cap;        // this source's sample ring inside MixArgs::rings (floats)
    int w;                         // ring write position
    int q_head, q_count;           // inputQueue: ent[(q_head + k) % kMixEntries], k < q_count
    MixEntry ent[kMixEntries];
    int has_cur;                   // currentInput != nullptr
    MixEntry cur;
    int ptr;                       // audioQueuePtr
};
struct MixArgs {
    MixSource *src;
    int n_src;
    const int *order;              // the controller's boundThreads: source indices in binding order (inactive sources are not in it)
    int n_order;
    float *rings;
    int rate;                      // the output device's sample rate (src->getSampleRate())
    // feed
    const ChanStatic *cs;
    const unsigned char *results;
    // render
    int n_frames;
    int *map;                      // [n_src][2 n_frames] ring index feeding each output float, -1 = none
    double *mix_peak;              // [n_src] this source's mixPeak, < 0 = the source did not reach `peak += mixPeak`
    float *out;                    // [2 n_frames]
};

SDR_HD SDR_FORCEINLINE bool mix_try_pop(MixSource &s) {
    if (s.q_count == 0) return false;
    s.cur = s.ent[s.q_head]; s.has_cur = 1;
    s.q_head = (s.q_head + 1) % kMixEntries; s.q_count -= 1;
    return true;
}
// Room for an entry of n samples in a source's ring: entries are contiguous (one never wraps), live data runs from the oldest
// entry still referenced (currentInput, else the head of the queue) to the write position.  Returns the start, or -1 when the
// new entry would run into live data (the block is dropped, as DemodulatorThread's try_push on a full queue drops it).
SDR_HD inline int mix_ring_take(const MixSource &s, int n) {
    if (n > s.ring_cap || s.q_count == kMixEntries) return -1;
    int oldest = -1;
    if (s.has_cur && s.cur.n > 0) oldest = s.cur.start;
    else if (s.q_count > 0) { for (int k = 0; k < s.q_count && oldest < 0; ++k) { const MixEntry &e = s.ent[(s.q_head + k) % kMixEntries]; if (e.n > 0) oldest = e.start; } }
    const int w = s.w;
    if (oldest < 0) return w + n <= s.ring_cap ? w : 0;                       // nothing live
    if (w > oldest) { if (w + n <= s.ring_cap) return w; return n < oldest ? 0 : -1; }
    if (w < oldest) return w + n < oldest ? w : -1;
    return -1;                                                                // w == oldest with live data: the ring is full
}
SDR_HD inline void mix_queue_push(MixSource &s, int start, int n, int channels, int rate, float peak) {
    MixEntry e;
    e.start = start; e.n = n; e.channels = channels; e.rate = rate; e.peak = peak;
    s.ent[(s.q_head + s.q_count) % kMixEntries] = e;
    s.q_count += 1;
    s.w = start + n;
}
// after a block: the channel's audio becomes one queued AudioThreadInput (thread j = source j)
SDR_HD inline void mix_feed_thread(const MixArgs &a, int j) {
    if (j >= a.n_src) return;
    MixSource &s = a.src[j];
    if (s.chan < 0) return;
    const ChanStatic &cs = a.cs[s.chan];
    const ChanResultHeader *hdr = reinterpret_cast<const ChanResultHeader *>(a.results + cs.res_off);
    const float *audio = reinterpret_cast<const float *>(a.results + cs.res_off + kResHdrBytes);
    int n = (int)hdr->n_audio;
    if (n > cs.res_audio_cap) n = cs.res_audio_cap;
    if (hdr->skipped || n <= 0) return;
    const int w = mix_ring_take(s, n);
    if (w < 0) return;
    float *ring = a.rings + s.ring_off;
    for (int i = 0; i < n; ++i) ring[w + i] = audio[i];
    mix_queue_push(s, w, n, (cs.modem == MODEM_IQ || cs.modem == MODEM_FMS) ? 2 : 1, s.feed_rate, hdr->peak);
}
// phase A of the callback for source j (AudioThread.cpp:113-228)
SDR_HD inline void mix_walk_thread(const MixArgs &a, int j) {
    if (j >= a.n_src) return;
    MixSource &s = a.src[j];
    const int slots = 2 * a.n_frames;
    int *map = a.map + (size_t)j * slots;
    for (int k = 0; k < slots; ++k) map[k] = -1;
    a.mix_peak[j] = -1.0;
    if (s.q_count == 0 || !s.active) return;                                       // :122
    if (!s.has_cur) { s.ptr = 0; mix_try_pop(s); return; }                         // :126-134
    if (s.cur.rate != a.rate) {                                                    // :136-155
        s.has_cur = 0;
        while (mix_try_pop(s)) { if (s.cur.rate == a.rate) break; s.has_cur = 0; }
        s.ptr = 0;
        if (!s.has_cur) return;
    }
    if (s.cur.channels == 0 || s.cur.n == 0) {                                     // :158-171
        if (s.q_count) { s.ptr = 0; s.has_cur = 0; mix_try_pop(s); }
        return;
    }
    double mixPeak = (double)(s.cur.peak * s.gain);                                // :173 (float product, widened)
    if (s.cur.channels == 1) {                                                     // :175-202
        for (int i = 0; i < a.n_frames; ++i) {
            if (s.ptr >= s.cur.n) {
                s.ptr = 0; s.has_cur = 0;
                if (!mix_try_pop(s)) break;
                const double srcPeak = (double)(s.cur.peak * s.gain);
                if (mixPeak < srcPeak) mixPeak = srcPeak;
            }
            if (s.has_cur && s.cur.n > 0) { map[2 * i] = s.cur.start + s.ptr; map[2 * i + 1] = s.cur.start + s.ptr; }
            s.ptr += 1;
        }
    } else {                                                                       // :203-226
        for (int i = 0, iMax = s.cur.channels * a.n_frames; i < iMax; ++i) {
            if (s.ptr >= s.cur.n) {
                s.ptr = 0; s.has_cur = 0;
                if (!mix_try_pop(s)) break;
                const double srcPeak = (double)(s.cur.peak * s.gain);
                if (mixPeak < srcPeak) mixPeak = srcPeak;
            }
            if (s.has_cur && s.cur.n > 0 && i < slots) map[i] = s.cur.start + s.ptr;
            s.ptr += 1;
        }
    }
    a.mix_peak[j] = mixPeak;                                                       // :228
}
// phase B: output float k, the sources in binding order
SDR_HD inline void mix_sum_thread(const MixArgs &a, int k) {
    const int slots = 2 * a.n_frames;
    if (k >= slots) return;
    float acc = 0.f;
    for (int o = 0; o < a.n_order; ++o) {
        const int j = a.order[o];
        const int idx = a.map[(size_t)j * slots + k];
        if (idx >= 0) {
            const MixSource &s = a.src[j];
#if defined(__CUDA_ARCH__) && !defined(SDR_EMU)
            acc = __fadd_rn(acc, __fmul_rn(a.rings[s.ring_off + idx], s.gain));
#else
            volatile float v = a.rings[s.ring_off + idx] * s.gain;
            acc = acc + v;
#endif
        }
    }
    a.out[k] = acc;
}
// phases C + D: peak over the sources that mixed (:228), normalisation (:231-238); returns the scale (1 = none)
SDR_HD inline float mix_scale(const MixArgs &a) {
    double peak = 0.0;
    for (int o = 0; o < a.n_order; ++o) { const int j = a.order[o]; if (a.mix_peak[j] >= 0.0) peak += a.mix_peak[j]; }
    return peak > 1.0 ? (float)(1.0 / peak) : 1.0f;
}

// ---- 16-bit PCM of a block's audio (AudioFileWAV.cpp:143-165): intScale = peak < 1 ? 32767 : 32767 / peak, int(x * intScale) ----
struct PcmArgs { const ChanStatic *cs; const unsigned char *results; int chan; int channels; short *out; int cap; int *n_out; };
SDR_HD inline void pcm16_thread(const PcmArgs &a, int i) {
    const ChanStatic &cs = a.cs[a.chan];
    const ChanResultHeader *hdr = reinterpret_cast<const ChanResultHeader *>(a.results + cs.res_off);
    const float *audio = reinterpret_cast<const float *>(a.results + cs.res_off + kResHdrBytes);
    int n = (int)hdr->n_audio;
    if (n > cs.res_audio_cap) n = cs.res_audio_cap;
    if (a.channels == 2) n &= ~1;                                                  // the reference writes whole {left, right} pairs
    if (n > a.cap) n = a.cap;
    if (i == 0) a.n_out[0] = hdr->skipped ? 0 : n;
    if (i >= n || hdr->skipped) return;
    const float peak = hdr->peak;
    const float intScale = (peak < 1.0f) ? 32767.0f : (32767.0f / peak);
#if defined(__CUDA_ARCH__) && !defined(SDR_EMU)
    const float v = __fmul_rn(audio[i], intScale);
#else
    volatile float v = audio[i] * intScale;
#endif
    a.out[i] = (short)(int)v;
}

// ---- scope (ScopeVisualProcessor.cpp:60-217) ---------------------------------------------------------------------------------
struct ScopeArgs {
    const ChanStatic *cs; const unsigned char *results; int chan; int channels;
    int sample_rate, input_rate;   // AudioThreadInput::sampleRate (audio) / inputRate (modem rate)
    int fft_size, max_samples;
    float2 *fft_in;                // [fft_size] the transform's input frame
    const float *mags;             // [fft_size] |X| fftshifted (the spectrum transform kernel's output layout)
    double2 *ema;                  // [fft_size / 2] {ma, maa}
    double *scalars;               // [4] ceil_ma, ceil_maa, floor_ma, floor_maa
    double rate;                   // fft_average_rate
    float *wave_xy, *spec_xy;      // [2 max_samples], [fft_size]
    int *counts;                   // [2] waveform points, spectrum points
    double *floor_ceil;            // [2]
    float *red;                    // scratch [2]: waveform peak
};
// waveform: peak over the first min(n, max_samples) samples (start 1), points {(i / iMax - 0.5) 2, data[i] / peak} (:73-122, type 0);
// and the transform's input frame: data (two channels: data[i] + data[iMax + i], :143-153), zero padded (:133-153)
SDR_HD inline void scope_prepare_thread(const ScopeArgs &a, int i, float wave_peak) {
    const ChanStatic &cs = a.cs[a.chan];
    const ChanResultHeader *hdr = reinterpret_cast<const ChanResultHeader *>(a.results + cs.res_off);
    const float *data = reinterpret_cast<const float *>(a.results + cs.res_off + kResHdrBytes);
    int n = hdr->skipped ? 0 : (int)hdr->n_audio;
    if (n > cs.res_audio_cap) n = cs.res_audio_cap;
    int iMax = n < a.max_samples ? n : a.max_samples;
    if (i == 0) a.counts[0] = iMax;
    if (i < iMax) {
        a.wave_xy[2 * i] = (float)((((double)i / (double)iMax) - 0.5) * 2.0);
        a.wave_xy[2 * i + 1] = data[i] / wave_peak;
    }
    if (i < a.fft_size) {
        float v = 0.f;
        if (a.channels == 1) { if (i < n) v = data[i]; }
        else { const int h = n / 2; if (i < h) v = data[i] + data[h + i]; }
        a.fft_in[i] = make_float2(v, 0.f);
    }
}
SDR_HD inline float scope_wave_peak_serial(const ScopeArgs &a) {
    const ChanStatic &cs = a.cs[a.chan];
    const ChanResultHeader *hdr = reinterpret_cast<const ChanResultHeader *>(a.results + cs.res_off);
    const float *data = reinterpret_cast<const float *>(a.results + cs.res_off + kResHdrBytes);
    int n = hdr->skipped ? 0 : (int)hdr->n_audio;
    if (n > cs.res_audio_cap) n = cs.res_audio_cap;
    const int iMax = n < a.max_samples ? n : a.max_samples;
    float peak = 1.0f;
    for (int i = 0; i < iMax; ++i) { const float p = fabsf(data[i]); if (p > peak) peak = p; }
    return peak;
}
// averages of bin i (:182-192); returns float-free doubles for the ceil / floor fold
SDR_HD SDR_FORCEINLINE double scope_ema_bin(const ScopeArgs &a, int i) {
    const int N = a.fft_size;
    const double x = (double)a.mags[(i + N / 2) & (N - 1)];
    double ma = a.ema[i].x, maa = a.ema[i].y;
    ma += (x - ma) * a.rate;                     // here ma moves FIRST and maa follows the NEW ma (:183-184; the wideband spectrum does the opposite)
    maa += (ma - maa) * a.rate;
    a.ema[i] = make_double2(ma, maa);
    return maa;
}
SDR_HD inline void scope_finish_serial(const ScopeArgs &a, double fft_ceil, double fft_floor) {      // :194-215, one thread
    double cma = a.scalars[0], cmaa = a.scalars[1], fma_ = a.scalars[2], fmaa = a.scalars[3];
    cma = cma + (fft_ceil - cma) * 0.05;
    cmaa = cmaa + (cma - cmaa) * 0.05;
    fma_ = fma_ + (fft_floor - fma_) * 0.05;
    fmaa = fmaa + (fma_ - fmaa) * 0.05;
    a.scalars[0] = cma; a.scalars[1] = cmaa; a.scalars[2] = fma_; a.scalars[3] = fmaa;
    unsigned outSize = (unsigned)a.fft_size / 2;
    if (a.sample_rate != a.input_rate) outSize = (unsigned)(int)floor((float)outSize * ((float)a.sample_rate / (float)a.input_rate));
    if (outSize > (unsigned)a.fft_size / 2) outSize = (unsigned)a.fft_size / 2;
    a.counts[1] = (int)outSize;
    a.floor_ceil[0] = fmaa; a.floor_ceil[1] = cmaa;
}
SDR_HD inline void scope_points_thread(const ScopeArgs &a, int i) {                                  // :207-211
    const int outSize = a.counts[1];
    if (i >= outSize) return;
    const double fmaa = a.floor_ceil[0], cmaa = a.floor_ceil[1];
    const float v = (float)(log10(a.ema[i].y + 0.25 - (fmaa - 0.75)) / log10((cmaa + 0.25) - (fmaa - 0.75)));
    a.spec_xy[2 * i] = (float)((double)i / (double)outSize);
    a.spec_xy[2 * i + 1] = v;
}

}  // namespace sdrb200
