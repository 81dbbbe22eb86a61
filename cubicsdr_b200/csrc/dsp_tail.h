// cubicsdr_b200/csrc/dsp_tail.h -- per-channel low-rate tail: remaining half-bands, arbitrary (polyphase) stage,
// demodulators, block auto-gain, audio resampler, metering.  One CTA per channel, phases separated by CTA barriers.
//
// Reference call sites (relative to /root/reference/src):
//   msresamp_crcf_execute          demod/DemodulatorPreThread.cpp:209   (stages not fused into the front kernel)
//   freqdem_demodulate_block       modules/modem/analog/ModemFM.cpp:36, ModemNBFM.cpp:36
//   envelope + firfilt DC blocker  modules/modem/analog/ModemAM.cpp:42-47
//   SSB chain                      modules/modem/analog/ModemUSB.cpp:54-61, ModemLSB.cpp:54-61
//   block auto-gain + msresamp_rrrf modules/modem/ModemAnalog.cpp:67-93
//   level / peak metering          demod/DemodulatorThread.cpp:142-163, 223-233
#pragma once
#include "dsp_core.h"

namespace sdrb200 {

constexpr int kTailStreams = kMaxTailHb + 1 + 3 + kMaxAudioHb + 1;   // complex levels, iq, w, d, audio levels

// singly rounded arithmetic for the few host-side expressions that are replayed on the device (the compiler may not fuse
// them: the host evaluates them with one rounding per operation)
#if defined(__CUDA_ARCH__) && !defined(SDR_EMU)
SDR_DEV SDR_FORCEINLINE double d_add(double a, double b) { return __dadd_rn(a, b); }
SDR_DEV SDR_FORCEINLINE double d_sub(double a, double b) { return __dsub_rn(a, b); }
SDR_DEV SDR_FORCEINLINE double d_mul(double a, double b) { return __dmul_rn(a, b); }
SDR_DEV SDR_FORCEINLINE double d_div(double a, double b) { return __ddiv_rn(a, b); }
SDR_DEV SDR_FORCEINLINE float f_add(float a, float b) { return __fadd_rn(a, b); }
#else
SDR_HD SDR_FORCEINLINE double d_add(double a, double b) { return a + b; }
SDR_HD SDR_FORCEINLINE double d_sub(double a, double b) { return a - b; }
SDR_HD SDR_FORCEINLINE double d_mul(double a, double b) { return a * b; }
SDR_HD SDR_FORCEINLINE double d_div(double a, double b) { return a / b; }
SDR_HD SDR_FORCEINLINE float f_add(float a, float b) { return a + b; }
#endif
#if defined(__CUDA_ARCH__) && !defined(SDR_EMU)
SDR_DEV SDR_FORCEINLINE float f_sub(float a, float b) { return __fsub_rn(a, b); }
SDR_DEV SDR_FORCEINLINE float f_mul(float a, float b) { return __fmul_rn(a, b); }
#else
SDR_HD SDR_FORCEINLINE float f_sub(float a, float b) { return a - b; }
SDR_HD SDR_FORCEINLINE float f_mul(float a, float b) { return a * b; }
#endif
// nco_crcf_pll_step's phase / frequency increment (liquid 1.5.0: the float argument goes through double for the division by
// 2 pi, back to float, loses its integer part towards zero, wraps into [0, 1) and is scaled to the 32-bit phase word)
SDR_HD SDR_FORCEINLINE unsigned pll_word(float z) {
    float x = (float)d_mul((double)z, 0.15915494309189535);
    x = f_sub(x, truncf(x));
    if (x < 0.0f) x = f_add(x, 1.0f);
    return (unsigned)(long long)f_mul(x, 4294967296.0f);
}

struct TailArgs {
    const ChanStatic *cs;
    const ChanDyn *cd;
    ChanState *st;
    const float *taps;
    const float2 *nco;
    unsigned char *results;
    int n;                        // wideband block length
    int first;                    // channel slot handled by CTA 0
    int view_pass;                // 0: the block's channels; 1: the launch of a zoomed view's hidden channel (ChanDyn::view)
    unsigned long long *dbg;      // diagnostics: per-phase end times of the first 64 channels (null = off)
};

struct TailShared {
    int lvl_n[kMaxTailHb + 1];            // new elements per complex level this block
    long long lvl_g0[kMaxTailHb + 1];     // global index of the first new element per level
    int n_iq;                             // decimated IQ samples this block
    unsigned long long arb_phase_after;
    int a_n[kMaxAudioHb + 1];             // new elements per audio level
    long long a_g0[kMaxAudioHb + 1];
    int n_audio;
    unsigned long long a_phase_after;
    float gain;
    float red_f[kTailThreads];
    double red_d[kTailThreads];
    float2 hist_tmp[kTailStreams][kHist];
};

// number of arbitrary-stage outputs for n_in inputs entering with `phase` (closed form of resamp_execute's loop)
SDR_HD inline int arb_count(unsigned step, unsigned long long phase, int n_in, unsigned long long *phase_after) {
    unsigned long long span = (unsigned long long)n_in << 24, cnt = 0;
    if (phase < span) cnt = (span - phase + step - 1) / step;
    *phase_after = phase + cnt * step - span;
    return (int)cnt;
}

enum TailPhase {
    TP_SETUP = 0,
    TP_HB0 = 1,                                   // kMaxTailHb half-band phases
    TP_ARB = TP_HB0 + kMaxTailHb,
    TP_DEMOD_A,                                   // FM / AM parallel; SSB serial IIR by thread 0
    TP_DEMOD_B,                                   // SSB Hilbert stage
    TP_AGC_A, TP_AGC_B, TP_AGC_C,
    TP_AUD0,                                      // audio: stage phases (kMaxAudioHb + 1)
    TP_CW_A = TP_AUD0 + kMaxAudioHb + 1,           // CW: beat mixer, Hilbert c2r, block gain (after the complex interpolator)
    TP_CW_B, TP_CW_C, TP_CW_D, TP_CW_E,
    TP_METER_A,
    TP_METER_B,
    TP_HIST_A, TP_HIST_B,
    TP_COUNT
};

SDR_HD inline float2 *chan_cbuf(const TailArgs &, const ChanStatic &cs, int off) {
    return reinterpret_cast<float2 *>(cs.arena + off);
}
SDR_HD inline float *chan_rbuf(const TailArgs &, const ChanStatic &cs, int off) { return cs.arena + off; }
// level j complex buffer of this block (level 0 is double-buffered against the next block's front kernel)
SDR_HD inline float2 *chan_lvl(const TailArgs &a, const ChanStatic &cs, const ChanDyn &cd, int j) {
    return chan_cbuf(a, cs, (j == 0 && cd.lvl0_sel) ? cs.lvl0_alt_off : cs.lvl_off[j]);
}

// Decimate-by-2 half-band over a [kHist | new] stream, parallel over outputs.  One output is a chain of 2m+1 dependent
// multiply-adds fed by 2m+1 loads; a thread working on one output at a time spends its life waiting for memory, so every
// thread carries kTailIlp outputs at once (independent accumulators, loads of all of them in flight) and the semi-length is
// a template parameter (liquid's msresamp2 only uses m = 3, 5, 10) so that the tap loop unrolls with the taps in registers.
// The order of the additions per output is the reference's (delay branch first, then taps 0..2m-1).  `in` was written by
// an earlier phase of the same launch, so it is read with ordinary (coherent) loads, never through the read-only path.
constexpr int kTailIlp = 4;
SDR_HD SDR_FORCEINLINE float tap_fma(float h, float x, float acc) { return fmaf(h, x, acc); }
SDR_HD SDR_FORCEINLINE float2 tap_fma(float h, float2 x, float2 acc) { return cf_fma(h, x, acc); }

template <int M, typename T>
SDR_HD inline void hb_decim_t(const T *in, long long in_g0, T *out, long long out_g0, int n_out,
                              const float *__restrict__ h1, int tid, int nt) {
    float h[2 * M];
#pragma unroll
    for (int t = 0; t < 2 * M; ++t) h[t] = sdr_ldg(h1 + t);
    const int base0 = (int)(2 * out_g0 - in_g0) + kHist;        // buffer position of in[2 * out_g0]
    int r = tid;
    for (; r + (kTailIlp - 1) * nt < n_out; r += kTailIlp * nt) {
        T acc[kTailIlp];
        const T *p[kTailIlp];
#pragma unroll
        for (int u = 0; u < kTailIlp; ++u) {
            const int base = base0 + 2 * (r + u * nt);
            acc[u] = in[base - 2 * M + 1];                       // delay branch: in[2k - 2m + 1]
            p[u] = in + base - 4 * M + 2;                        // filtered branch: in[2k - 4m + 2 + 2t]
        }
#pragma unroll
        for (int t = 0; t < 2 * M; ++t) {
#pragma unroll
            for (int u = 0; u < kTailIlp; ++u) acc[u] = tap_fma(h[t], p[u][2 * t], acc[u]);
        }
#pragma unroll
        for (int u = 0; u < kTailIlp; ++u) out[kHist + r + u * nt] = acc[u];
    }
    for (; r < n_out; r += nt) {
        const int base = base0 + 2 * r;
        T acc = in[base - 2 * M + 1];
        const T *q = in + base - 4 * M + 2;
#pragma unroll
        for (int t = 0; t < 2 * M; ++t) acc = tap_fma(h[t], q[2 * t], acc);
        out[kHist + r] = acc;
    }
}
template <typename T>
SDR_HD inline void hb_decim_any(const T *in, long long in_g0, T *out, long long out_g0, int n_out, int m, const float *h1, int tid, int nt) {
    if (m == 3) hb_decim_t<3, T>(in, in_g0, out, out_g0, n_out, h1, tid, nt);
    else if (m == 5) hb_decim_t<5, T>(in, in_g0, out, out_g0, n_out, h1, tid, nt);
    else if (m == 10) hb_decim_t<10, T>(in, in_g0, out, out_g0, n_out, h1, tid, nt);
    else {                                                       // not produced by the design code; kept for completeness
        for (int r = tid; r < n_out; r += nt) {
            const int base = (int)(2 * (out_g0 + r) - in_g0) + kHist;
            T acc = in[base - 2 * m + 1];
            const T *p = in + base - 4 * m + 2;
            for (int t = 0; t < 2 * m; ++t) acc = tap_fma(sdr_ldg(h1 + t), p[2 * t], acc);
            out[kHist + r] = acc;
        }
    }
}
SDR_HD inline void hb_decim_c(const float2 *in, long long in_g0, float2 *out, long long out_g0, int n_out, int m,
                              const float *h1, int tid, int nt) {
    hb_decim_any<float2>(in, in_g0, out, out_g0, n_out, m, h1, tid, nt);
}
SDR_HD inline void hb_decim_r(const float *in, long long in_g0, float *out, long long out_g0, int n_out, int m,
                              const float *h1, int tid, int nt) {
    hb_decim_any<float>(in, in_g0, out, out_g0, n_out, m, h1, tid, nt);
}
// interpolate-by-2 half-band (resamp2 interp_execute): Y[2k] = X[k-m], Y[2k+1] = sum h1[t] X[k-2m+1+t]
// (thread r handles input k = r: both outputs of the pair, kTailIlp pairs at a time)
SDR_HD SDR_FORCEINLINE void t_zero(float &v) { v = 0.f; }
SDR_HD SDR_FORCEINLINE void t_zero(float2 &v) { v = make_float2(0.f, 0.f); }
template <int M, typename T>
SDR_HD inline void hb_interp_t(const T *in, int n_in, T *out, const float *__restrict__ h1, int tid, int nt) {
    float h[2 * M];
#pragma unroll
    for (int t = 0; t < 2 * M; ++t) h[t] = sdr_ldg(h1 + t);
    int k = tid;
    for (; k + (kTailIlp - 1) * nt < n_in; k += kTailIlp * nt) {
        T v[kTailIlp], e[kTailIlp];
#pragma unroll
        for (int u = 0; u < kTailIlp; ++u) { t_zero(v[u]); e[u] = in[kHist + k + u * nt - M]; }
#pragma unroll
        for (int t = 0; t < 2 * M; ++t) {
#pragma unroll
            for (int u = 0; u < kTailIlp; ++u) v[u] = tap_fma(h[t], in[kHist + k + u * nt - 2 * M + 1 + t], v[u]);
        }
#pragma unroll
        for (int u = 0; u < kTailIlp; ++u) { out[kHist + 2 * (k + u * nt)] = e[u]; out[kHist + 2 * (k + u * nt) + 1] = v[u]; }
    }
    for (; k < n_in; k += nt) {
        T v;
        t_zero(v);
#pragma unroll
        for (int t = 0; t < 2 * M; ++t) v = tap_fma(h[t], in[kHist + k - 2 * M + 1 + t], v);
        out[kHist + 2 * k] = in[kHist + k - M];
        out[kHist + 2 * k + 1] = v;
    }
}
// complex samples (msresamp_cccf's half-band interpolators, ModemCW.cpp:158): the design only produces m = 3, 5, 10
SDR_HD inline void hb_interp_c(const float2 *in, int n_in, float2 *out, int m, const float *h1, int tid, int nt) {
    if (m == 3) hb_interp_t<3, float2>(in, n_in, out, h1, tid, nt);
    else if (m == 5) hb_interp_t<5, float2>(in, n_in, out, h1, tid, nt);
    else hb_interp_t<10, float2>(in, n_in, out, h1, tid, nt);
}
SDR_HD inline void hb_interp_r(const float *in, int n_in, float *out, int m, const float *h1, int tid, int nt) {
    if (m == 3) hb_interp_t<3, float>(in, n_in, out, h1, tid, nt);
    else if (m == 5) hb_interp_t<5, float>(in, n_in, out, h1, tid, nt);
    else if (m == 10) hb_interp_t<10, float>(in, n_in, out, h1, tid, nt);
    else {
        for (int r = tid; r < 2 * n_in; r += nt) {
            const int k = r >> 1;
            float v;
            if ((r & 1) == 0) v = in[kHist + k - m];
            else {
                const float *p = in + kHist + k - 2 * m + 1;
                v = 0.f;
                for (int t = 0; t < 2 * m; ++t) v = fmaf(sdr_ldg(h1 + t), p[t], v);
            }
            out[kHist + r] = v;
        }
    }
}

// kTailIlp outputs of the real 256-arm polyphase stage (audio resampler), loads of all of them in flight
SDR_HD SDR_FORCEINLINE void arb_r_group(const float *z, const float *pfb, unsigned long long ph0, unsigned step, int r0, int stride,
                                        float *acc) {
    const float *zq[kTailIlp];
    const float *h[kTailIlp];
#pragma unroll
    for (int u = 0; u < kTailIlp; ++u) {
        const unsigned long long psi = ph0 + (unsigned long long)(r0 + u * stride) * step;
        zq[u] = z + (int)(psi >> 24);
        h[u] = pfb + ((psi >> 16) & 255u) * kPfbTaps;
        acc[u] = 0.f;
    }
#pragma unroll
    for (int t = 0; t < kPfbTaps; ++t) {
#pragma unroll
        for (int u = 0; u < kTailIlp; ++u) acc[u] = fmaf(sdr_ldg(h[u] + t), zq[u][-t], acc[u]);
    }
}
SDR_HD SDR_FORCEINLINE float arb_r_one(const float *z, const float *pfb, unsigned long long ph0, unsigned step, int r) {
    const unsigned long long psi = ph0 + (unsigned long long)r * step;
    const float *zq = z + (int)(psi >> 24);
    const float *h = pfb + ((psi >> 16) & 255u) * kPfbTaps;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < kPfbTaps; ++t) acc = fmaf(sdr_ldg(h + t), zq[-t], acc);
    return acc;
}

SDR_HD SDR_FORCEINLINE float2 arb_c_one(const float2 *z, const float *pfb, unsigned long long ph0, unsigned step, int r) {
    const unsigned long long psi = ph0 + (unsigned long long)r * step;
    const float2 *zq = z + (int)(psi >> 24);
    const float *h = pfb + ((psi >> 16) & 255u) * kPfbTaps;
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < kPfbTaps; ++t) acc = cf_fma(sdr_ldg(h + t), zq[-t], acc);
    return acc;
}

SDR_HD inline void tail_phase(int ph, const TailArgs &a, int ch, int tid, int nt, TailShared *sm) {
    const ChanStatic &cs = a.cs[ch];
    const ChanDyn &cd = a.cd[ch];
    if (!cd.valid || cd.view != a.view_pass) return;
    ChanState &st = a.st[ch];
    ChanResultHeader *hdr = reinterpret_cast<ChanResultHeader *>(a.results + cs.res_off);
    float *res_audio = reinterpret_cast<float *>(a.results + cs.res_off + kResHdrBytes);
    float2 *res_iq = reinterpret_cast<float2 *>(res_audio + cs.res_audio_cap);
    if (!cd.active) {
        if (ph == TP_SETUP && tid == 0) {
            hdr->n_audio = 0; hdr->n_iq = 0; hdr->peak = 0.f; hdr->skipped = 1; hdr->level_db = 0.0;
            hdr->sig_level = st.sig_level; hdr->sig_floor = st.sig_floor; hdr->sig_ceil = st.sig_ceil;   // no block reached DemodulatorThread
            hdr->squelched = (cd.squelch_enabled && st.sig_level < cd.squelch_level) ? 1 : 0;
        }
        return;
    }
    // FM stereo (analog/ModemFMStereo.cpp:186-287) shares the discriminator and the metering of FM; its own stages reuse phases that
    // FM leaves empty: TP_DEMOD_B real -> complex, TP_AGC_A pilot band-pass (five sections as a systolic chain), TP_AGC_B phase-locked
    // loop + 38 kHz down-mix (serial), TP_AGC_C complex -> real, the audio phases run both resamplers, TP_CW_A de-emphasis, TP_CW_B low-pass pair, TP_CW_C/D histories
    const bool is_fms = cs.modem == MODEM_FMS;
    const bool is_fm = (cs.modem == MODEM_FM || cs.modem == MODEM_NBFM || is_fms);
    const bool is_ssb = (cs.modem == MODEM_USB || cs.modem == MODEM_LSB);
    // "I/Q" (analog/ModemIQ.cpp:39-56): no demodulator, no audio resampler -- the decimated IQ itself leaves as two-channel
    // audio {imag, real}; the level is metered on the IQ like FM's (useSignalOutput stays false)
    const bool is_iq = cs.modem == MODEM_IQ;
    const bool is_none = cs.modem == MODEM_NONE;             // resampler only (zoomed spectrum view): nothing after the IQ stage
    // CW (analog/ModemCW.cpp:155-209): the decimated IQ is interpolated to the audio rate as complex samples (msresamp_cccf),
    // mixed up by the beat oscillator, made real by firhilbf(5, 60) and scaled by its own block gain; no audio resampler after
    const bool is_cw = cs.modem == MODEM_CW;
    const bool has_agc = !is_fm && !is_iq && !is_none && !is_cw;   // ModemAM / USB / LSB / DSB pass autoGain = true to buildAudioOutput
    const bool lvl_from_iq = is_fm || is_iq || is_none;

    if (ph == TP_SETUP) {
        if (tid == 0) {
            const long long g0 = cd.g0, g1 = cd.g0 + a.n;
            for (int j = 0; j <= cs.n_tail; ++j) {
                const int sh = cs.K1 + j;
                sm->lvl_g0[j] = g0 >> sh;
                sm->lvl_n[j] = (int)((g1 >> sh) - (g0 >> sh));
            }
            sm->n_iq = arb_count(cs.arb_step, st.arb_phase, sm->lvl_n[cs.n_tail], &sm->arb_phase_after);
            if (cd.inject_n > 0) {              // Modem::demodulate seam: IQ already at the modem rate
                for (int j = 0; j <= cs.n_tail; ++j) sm->lvl_n[j] = 0;
                sm->n_iq = cd.inject_n;
                sm->arb_phase_after = st.arb_phase;
            }
            // audio resampler bookkeeping: its input is the demodulated stream, one sample per decimated IQ sample
            const long long d0 = (long long)st.iq_count, d1 = d0 + sm->n_iq;
            if (is_iq || is_none) {
                for (int j = 0; j <= cs.a_S; ++j) { sm->a_g0[j] = 0; sm->a_n[j] = 0; }
                sm->n_audio = is_iq ? 2 * sm->n_iq : 0;
                sm->a_phase_after = st.a_phase;
            } else if (!cs.a_interp) {
                for (int j = 0; j <= cs.a_S; ++j) { sm->a_g0[j] = d0 >> j; sm->a_n[j] = (int)((d1 >> j) - (d0 >> j)); }
                sm->n_audio = arb_count(cs.a_step, st.a_phase, sm->a_n[cs.a_S], &sm->a_phase_after);
            } else {
                int nb = arb_count(cs.a_step, st.a_phase, sm->n_iq, &sm->a_phase_after);
                for (int j = 0; j <= cs.a_S; ++j) { sm->a_g0[j] = 0; sm->a_n[j] = nb << j; }
                sm->n_audio = nb << cs.a_S;
            }
        }
        return;
    }

    // ---- remaining half-band decimators of the IQ resampler --------------------------------------------------------
    if (ph >= TP_HB0 && ph < TP_HB0 + kMaxTailHb) {
        const int j = ph - TP_HB0;
        if (j < cs.n_tail && cd.inject_n <= 0)
            hb_decim_c(chan_lvl(a, cs, cd, j), sm->lvl_g0[j], chan_cbuf(a, cs, cs.lvl_off[j + 1]), sm->lvl_g0[j + 1],
                       sm->lvl_n[j + 1], cs.tail_m[j], a.taps + cs.tail_h1[j], tid, nt);
        return;
    }

    if (ph == TP_ARB && cd.inject_n > 0) return;
    if (ph == TP_ARB) {      // 256-arm polyphase stage, nearest arm, 24-bit phase
        const float2 *z = chan_lvl(a, cs, cd, cs.n_tail) + kHist;
        float2 *iq = chan_cbuf(a, cs, cs.iq_off) + kHist;
        const float *pfb = a.taps + cs.pfb_off;
        const unsigned long long ph0 = st.arb_phase;
        const float zeta = cs.zeta;
        const int n_iq = sm->n_iq;
        const bool keep = cs.want_iq != 0;
        int r = tid;
        for (; r + (kTailIlp - 1) * nt < n_iq; r += kTailIlp * nt) {       // kTailIlp outputs in flight per thread
            float2 acc[kTailIlp];
            const float2 *zq[kTailIlp];
            const float *h[kTailIlp];
#pragma unroll
            for (int u = 0; u < kTailIlp; ++u) {
                const unsigned long long psi = ph0 + (unsigned long long)(r + u * nt) * cs.arb_step;
                zq[u] = z + (int)(psi >> 24);
                h[u] = pfb + ((psi >> 16) & 255u) * kPfbTaps;
                acc[u] = make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int t = 0; t < kPfbTaps; ++t) {
#pragma unroll
                for (int u = 0; u < kTailIlp; ++u) acc[u] = cf_fma(sdr_ldg(h[u] + t), zq[u][-t], acc[u]);
            }
#pragma unroll
            for (int u = 0; u < kTailIlp; ++u) {
                const float2 y = cf_scale(zeta, acc[u]);
                iq[r + u * nt] = y;
                if (keep && r + u * nt < cs.res_iq_cap) res_iq[r + u * nt] = y;
            }
        }
        for (; r < n_iq; r += nt) {
            const unsigned long long psi = ph0 + (unsigned long long)r * cs.arb_step;
            const int q = (int)(psi >> 24);
            const float *h = pfb + ((psi >> 16) & 255u) * kPfbTaps;
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 0; t < kPfbTaps; ++t) acc = cf_fma(sdr_ldg(h + t), z[q - t], acc);
            acc = cf_scale(zeta, acc);
            iq[r] = acc;
            if (keep && r < cs.res_iq_cap) res_iq[r] = acc;
        }
        return;
    }

    float *d = chan_rbuf(a, cs, cs.d_off) + kHist;
    const float2 *iq = chan_cbuf(a, cs, cs.iq_off) + kHist;

    if (ph == TP_DEMOD_A) {
        const int n = sm->n_iq;
        if (is_none || is_cw) {
        } else if (is_iq) {
            for (int i = tid; i < n; i += nt) {
                const float2 x = iq[i];
                if (2 * i + 1 < cs.res_audio_cap) { res_audio[2 * i] = x.y; res_audio[2 * i + 1] = x.x; }
            }
        } else if (is_fm) {
            const float ref = 0.31830987f;                       // (float)(1 / (2 pi kf)), kf = 0.5  (freqdem_create(0.5))
            for (int i = tid; i < n; i += nt) {
                float2 x = iq[i], p = iq[i - 1];
                float re = fmaf(p.x, x.x, p.y * x.y);            // conj(p) * x
                float im = fmaf(p.x, x.y, -(p.y * x.x));
                d[i] = atan2f(im, re) * ref;
            }
        } else if (cs.modem == MODEM_AM) {
            const float *h = a.taps + cs.am_taps_off;
            for (int i = tid; i < n; i += nt) {
                float acc = 0.f;
                for (int k = 0; k < 51; ++k) {
                    float2 x = iq[i - k];
                    acc = fmaf(sdr_ldg(h + k), sqrtf(x.x * x.x + x.y * x.y), acc);
                }
                d[i] = acc;
            }
        } else if (cs.modem == MODEM_DSB) {
            // ampmodem_demod_dsb_pll_costas (liquid 1.5.0 binary, read from its disassembly): per sample
            //   v = x * conj(osc)                       nco_crcf_mix_down, table oscillator, four products and two sums, no fusing
            //   e = Re v > 0 ? Im v : -Im v             hard-limited Costas error
            //   d_theta += w(alpha e), theta += w(beta e)   nco_crcf_pll_step, alpha = 0.001f, beta = sqrtf(alpha),
            //                                           w(z) = (uint32)(int64)(frac((float)((double)z / 2 pi)) * 2^32)
            //   theta += d_theta                        nco_crcf_step
            //   y = Re v / 0.5                          mod_index
            // A feedback loop through a quantised oscillator: strictly serial, one thread.
            if (tid == 0) {
                unsigned th = st.dsb_theta, dth = st.dsb_dtheta;
                const float alpha = 0.001f, beta = 0.0316227786f;    // sqrtf(0.001f) = 0x3d0186e3 (nco_crcf_pll_set_bandwidth)
                for (int i = 0; i < n; ++i) {
                    const float2 w = nco_lookup(a.nco, th);
                    const float2 x = iq[i];
                    const float vr = f_add(f_mul(x.x, w.x), f_mul(x.y, w.y));
                    const float vi = f_sub(f_mul(x.y, w.x), f_mul(x.x, w.y));
                    const float e = vr > 0.0f ? vi : -vi;
                    dth += pll_word(f_mul(alpha, e));
                    th += pll_word(f_mul(e, beta));
                    th += dth;
                    d[i] = vr / 0.5f;
                }
                st.dsb_theta = th; st.dsb_dtheta = dth;
            }
        } else if (tid == 0) {                                   // SSB: quarter-rate shift, Butterworth-6, shift back
            float2 *w = chan_cbuf(a, cs, cs.w_off) + kHist;
            const float *sos = a.taps + cs.sos_off;
            unsigned th = st.ssb_theta;
            float v[3][2][2];
            for (int s = 0; s < 3; ++s) for (int k = 0; k < 2; ++k) { v[s][k][0] = st.sos_v[s][k][0]; v[s][k][1] = st.sos_v[s][k][1]; }
            const float sgn = (cs.modem == MODEM_USB) ? -1.0f : 1.0f;   // USB mixes down first, LSB mixes up first
            for (int i = 0; i < n; ++i) {
                th += 0x40000000u;                                // nco_crcf_step before use (ModemUSB.cpp:55)
                float2 cs2 = nco_lookup(a.nco, th);
                float2 x = cf_rot(iq[i], cs2.x, sgn * cs2.y);
                for (int s = 0; s < 3; ++s) {                     // direct-form II second-order sections
                    const float *c = sos + 5 * s;                 // {b0, b1, b2, a1, a2}
                    float v0r = x.x - c[3] * v[s][0][0] - c[4] * v[s][1][0];
                    float v0i = x.y - c[3] * v[s][0][1] - c[4] * v[s][1][1];
                    float yr = c[0] * v0r + c[1] * v[s][0][0] + c[2] * v[s][1][0];
                    float yi = c[0] * v0i + c[1] * v[s][0][1] + c[2] * v[s][1][1];
                    v[s][1][0] = v[s][0][0]; v[s][1][1] = v[s][0][1];
                    v[s][0][0] = v0r; v[s][0][1] = v0i;
                    x = make_float2(yr, yi);
                }
                w[i] = cf_rot(x, cs2.x, -sgn * cs2.y);
            }
            st.ssb_theta = th;
            for (int s = 0; s < 3; ++s) for (int k = 0; k < 2; ++k) { st.sos_v[s][k][0] = v[s][k][0]; st.sos_v[s][k][1] = v[s][k][1]; }
        }
        return;
    }

    if (ph == TP_DEMOD_B) {
        if (is_fms) {                                            // firhilbf_r2c_execute (:232): {x[i - 10], sum hq x[i - 19 + 2 j]}
            float2 *w = chan_cbuf(a, cs, cs.w_off) + kHist;
            const float *hq = a.taps + cs.cw_hq_off;             // firhilbf_create(5, 60): the taps CW uses
            for (int i = tid; i < sm->n_iq; i += nt) {
                float yq = 0.f;
                for (int j = 0; j < 10; ++j) yq = fmaf(sdr_ldg(hq + j), d[i - 19 + 2 * j], yq);
                w[i] = make_float2(d[i - 10], yq);
            }
            return;
        }
        if (is_ssb && !is_none) {                                // firhilbf c2r, 21 taps
            const float2 *w = chan_cbuf(a, cs, cs.w_off) + kHist;
            const float *hq = a.taps + cs.hq_off;
            const float sgn = (cs.modem == MODEM_USB) ? -1.0f : 1.0f;   // usb = yi - yq, lsb = yi + yq
            for (int i = tid; i < sm->n_iq; i += nt) {
                float yq = 0.f;
                for (int j = 0; j < 10; ++j) yq = fmaf(sdr_ldg(hq + j), w[i - 19 + 2 * j].y, yq);
                d[i] = w[i - 10].x + sgn * yq;
            }
        }
        return;
    }

    // ---- block auto-gain (AM / SSB): ModemAnalog.cpp:70-86 ------------------------------------------------------------
    if (ph == TP_AGC_A && is_fms) {
        // :234, the 19 kHz pilot band-pass over the block: five second-order sections; liquid 1.5.0's iirfiltsos_execute_df2 as its
        // binary evaluates it -- v0 = x - (a1 v1 + a2 v2), y = (b1 v1 + b2 v2) + b0 v0, every operation singly rounded (the poles sit
        // 0.002 inside the unit circle).  The filter does not depend on the loop behind it, so it runs ahead of it as a systolic
        // chain: lane q of warp 0 is section q and works on sample t - q at step t (same arithmetic per section and sample as the
        // serial form, five sections in flight); the result goes through the x2 stream, which the loop then overwrites in place.
        const float2 *w = chan_cbuf(a, cs, cs.w_off) + kHist;
        float2 *x2 = chan_cbuf(a, cs, cs.fms_x2_off) + kHist;
        const float *sos = a.taps + cs.fms_sos_off;
        const int n = sm->n_iq;
#if defined(__CUDA_ARCH__) && !defined(SDR_EMU)
        if (tid < 32) {
            const int lane = tid, q = lane < 5 ? lane : 4;
            const float b0 = sos[5 * q], b1 = sos[5 * q + 1], b2 = sos[5 * q + 2], a1 = sos[5 * q + 3], a2 = sos[5 * q + 4];
            float v1r = st.fms_sos_v[q][0][0], v1i = st.fms_sos_v[q][0][1], v2r = st.fms_sos_v[q][1][0], v2i = st.fms_sos_v[q][1][1];
            float2 out = make_float2(0.f, 0.f);
            float2 chunk = lane < n ? w[lane] : make_float2(0.f, 0.f), next = chunk;
            for (int t = 0; t < n + 4; ++t) {
                if ((t & 31) == 0) {                             // 32 inputs at a time, the next 32 already on their way
                    chunk = next;
                    const int k = t + 32 + lane;
                    next = k < n ? w[k] : make_float2(0.f, 0.f);
                }
                const float xr = __shfl_sync(0xffffffffu, chunk.x, t & 31), xi = __shfl_sync(0xffffffffu, chunk.y, t & 31);
                const float ur = __shfl_up_sync(0xffffffffu, out.x, 1), ui = __shfl_up_sync(0xffffffffu, out.y, 1);
                const int i = t - lane;
                if (lane < 5 && i >= 0 && i < n) {
                    const float yr0 = lane == 0 ? xr : ur, yi0 = lane == 0 ? xi : ui;
                    const float v0r = f_sub(yr0, f_add(f_mul(a1, v1r), f_mul(v2r, a2)));
                    const float v0i = f_sub(yi0, f_add(f_mul(a1, v1i), f_mul(a2, v2i)));
                    out.x = f_add(f_add(f_mul(b1, v1r), f_mul(v2r, b2)), f_mul(v0r, b0));
                    out.y = f_add(f_add(f_mul(b1, v1i), f_mul(b2, v2i)), f_mul(b0, v0i));
                    v2r = v1r; v2i = v1i; v1r = v0r; v1i = v0i;
                    if (lane == 4) x2[i] = out;
                }
            }
            if (lane < 5) { st.fms_sos_v[q][0][0] = v1r; st.fms_sos_v[q][0][1] = v1i; st.fms_sos_v[q][1][0] = v2r; st.fms_sos_v[q][1][1] = v2i; }
        }
#else
        if (tid == 0) {
            float c[5][5], v[5][2][2];
            for (int q = 0; q < 5; ++q) {
                for (int k = 0; k < 5; ++k) c[q][k] = sos[5 * q + k];
                for (int k = 0; k < 2; ++k) { v[q][k][0] = st.fms_sos_v[q][k][0]; v[q][k][1] = st.fms_sos_v[q][k][1]; }
            }
            for (int i = 0; i < n; ++i) {
                float yr = w[i].x, yi = w[i].y;
                for (int q = 0; q < 5; ++q) {
                    const float v0r = f_sub(yr, f_add(f_mul(c[q][3], v[q][0][0]), f_mul(v[q][1][0], c[q][4])));
                    const float v0i = f_sub(yi, f_add(f_mul(c[q][3], v[q][0][1]), f_mul(c[q][4], v[q][1][1])));
                    yr = f_add(f_add(f_mul(c[q][1], v[q][0][0]), f_mul(v[q][1][0], c[q][2])), f_mul(v0r, c[q][0]));
                    yi = f_add(f_add(f_mul(c[q][1], v[q][0][1]), f_mul(c[q][2], v[q][1][1])), f_mul(c[q][0], v0i));
                    v[q][1][0] = v[q][0][0]; v[q][1][1] = v[q][0][1];
                    v[q][0][0] = v0r; v[q][0][1] = v0i;
                }
                x2[i] = make_float2(yr, yi);
            }
            for (int q = 0; q < 5; ++q) for (int k = 0; k < 2; ++k) { st.fms_sos_v[q][k][0] = v[q][k][0]; st.fms_sos_v[q][k][1] = v[q][k][1]; }
        }
#endif
        return;
    }
    if (ph == TP_AGC_B && is_fms) {
        // :235-255, one sample after the other: phase detector against the loop oscillator (table oscillator: nco_crcf ignores its
        // LIQUID_VCO type), nco_crcf_pll_step with bandwidth 0.25 (alpha = 0.25, beta = sqrtf(alpha) = 0.5), one oscillator step, two
        // mix-downs by the stepped phase.
        if (tid == 0) {
            const float2 *w = chan_cbuf(a, cs, cs.w_off) + kHist;
            float2 *x2 = chan_cbuf(a, cs, cs.fms_x2_off) + kHist;
            unsigned th = st.fms_theta, dth = st.fms_dtheta;
            const int n = sm->n_iq;
            for (int i = 0; i < n; ++i) {
                const float2 x = w[i];
                const float yr = x2[i].x, yi = x2[i].y;          // the pilot filter's output
                const float2 o = nco_lookup(a.nco, th);          // nco_crcf_cexpf: {cos, sin}
                const float wi = -o.y;                            // conjf(w)
                const float ur = f_sub(f_mul(yr, o.x), f_mul(yi, wi));
                const float ui = f_add(f_mul(yr, wi), f_mul(yi, o.x));
                const float e = atan2f(ui, ur);
                dth += pll_word(f_mul(e, 0.25f));
                th += pll_word(f_mul(e, 0.5f));
                th += dth;                                        // nco_crcf_step
                const float2 p = nco_lookup(a.nco, th);
                const float y1r = f_add(f_mul(x.x, p.x), f_mul(x.y, p.y)), y1i = f_sub(f_mul(x.y, p.x), f_mul(x.x, p.y));   // nco_crcf_mix_down
                x2[i] = make_float2(f_add(f_mul(y1r, p.x), f_mul(y1i, p.y)), f_sub(f_mul(y1i, p.x), f_mul(y1r, p.y)));
            }
            st.fms_theta = th; st.fms_dtheta = dth;
        }
        return;
    }
    if (ph == TP_AGC_C && is_fms) {                              // firhilbf_c2r_execute (:258): the lower-sideband output
        const float2 *x2 = chan_cbuf(a, cs, cs.fms_x2_off) + kHist;
        float *ds = chan_rbuf(a, cs, cs.fms_ds_off) + kHist;
        const float *hq = a.taps + cs.cw_hq_off;
        for (int i = tid; i < sm->n_iq; i += nt) {
            float yq = 0.f;
            for (int j = 0; j < 10; ++j) yq = fmaf(sdr_ldg(hq + j), x2[i - 19 + 2 * j].y, yq);
            ds[i] = x2[i - 10].x + yq;
        }
        return;
    }
    if (ph == TP_AGC_A) {
        if (has_agc) {
            float mx = 0.f;
            for (int i = tid; i < sm->n_iq; i += nt) mx = fmaxf(mx, d[i]);
            sm->red_f[tid] = mx;
        }
        return;
    }
    if (ph == TP_AGC_B) {
        // a block that yields no decimated samples never reaches buildAudioOutput (`if (!bufSize) return;`,
        // ModemAM.cpp:35, ModemUSB/LSB likewise): the ceiling averages must not advance
        if (has_agc && tid == 0 && sm->n_iq > 0) {
            float mx = 0.f;
            for (int t = 0; t < nt; ++t) mx = fmaxf(mx, sm->red_f[t]);
            float ma = st.agc_ma + (st.agc_ceil - st.agc_ma) * 0.025f;
            float maa = st.agc_maa + (ma - st.agc_maa) * 0.025f;
            st.agc_ma = ma; st.agc_maa = maa; st.agc_ceil = mx;
            sm->gain = 0.5f / maa;
        }
        return;
    }
    if (ph == TP_AGC_C) {
        if (has_agc) { const float g = sm->gain; for (int i = tid; i < sm->n_iq; i += nt) d[i] *= g; }
        return;
    }

    // ---- audio resampler msresamp_rrrf ---------------------------------------------------------------------------------
    if (ph >= TP_AUD0 && ph < TP_AUD0 + kMaxAudioHb + 1) {
        const int j = ph - TP_AUD0;
        if (is_iq || is_none) return;
        if (is_cw) {                                             // msresamp_cccf interpolator: arbitrary stage, then a_S half-bands (:158)
            if (j == 0) {
                const float2 *z = iq;
                float2 *out = chan_cbuf(a, cs, cs.al_off[0]) + kHist;
                const float *pfb = a.taps + cs.a_pfb_off;
                const unsigned long long ph0 = st.a_phase;
                for (int r = tid; r < sm->a_n[0]; r += nt) out[r] = arb_c_one(z, pfb, ph0, cs.a_step, r);
            } else if (j <= cs.a_S) {
                hb_interp_c(chan_cbuf(a, cs, cs.al_off[j - 1]), sm->a_n[j - 1], chan_cbuf(a, cs, cs.al_off[j]), cs.a_m[j - 1],
                            a.taps + cs.a_h1[j - 1], tid, nt);
            }
            return;
        }
        if (!cs.a_interp) {
            // decimator: a_S half-bands (cascade position g = a_S-1-j for phase j), then the arbitrary stage.  FM stereo runs
            // two of them in lock step (audioResampler on the sum signal, stereoResampler on the difference signal: same ratio,
            // same input counts, hence the same phases) and keeps both outputs for its back end.
            for (int strm = 0; strm < (is_fms ? 2 : 1); ++strm) {
                const int s_d = strm ? cs.fms_ds_off : cs.d_off;
                const int *s_al = strm ? cs.fms_al_off : cs.al_off;
                if (j < cs.a_S) {
                    const int g = cs.a_S - 1 - j;
                    const float *in = (j == 0) ? chan_rbuf(a, cs, s_d) : chan_rbuf(a, cs, s_al[j]);
                    hb_decim_r(in, sm->a_g0[j], chan_rbuf(a, cs, s_al[j + 1]), sm->a_g0[j + 1], sm->a_n[j + 1], cs.a_m[g],
                               a.taps + cs.a_h1[g], tid, nt);
                } else if (j == cs.a_S) {
                    const float *z = ((cs.a_S == 0) ? chan_rbuf(a, cs, s_d) : chan_rbuf(a, cs, s_al[cs.a_S])) + kHist;
                    const float *pfb = a.taps + cs.a_pfb_off;
                    const unsigned long long ph0 = st.a_phase;
                    const int na = sm->n_audio;
                    float *dst = is_fms ? chan_rbuf(a, cs, strm ? cs.fms_as_off : cs.fms_am_off) : res_audio;
                    const int cap = is_fms ? 0x7fffffff : cs.res_audio_cap;
                    int r = tid;
                    for (; r + (kTailIlp - 1) * nt < na; r += kTailIlp * nt) {
                        float acc[kTailIlp];
                        arb_r_group(z, pfb, ph0, cs.a_step, r, nt, acc);
#pragma unroll
                        for (int u = 0; u < kTailIlp; ++u) if (r + u * nt < cap) dst[r + u * nt] = acc[u] * cs.a_zeta;
                    }
                    for (; r < na; r += nt) {
                        const float acc = arb_r_one(z, pfb, ph0, cs.a_step, r);
                        if (r < cap) dst[r] = acc * cs.a_zeta;
                    }
                }
            }
        } else {
            // interpolator: arbitrary stage first (1..2 outputs per input), then a_S half-band interpolators (g = j-1)
            if (j == 0) {
                const float *z = chan_rbuf(a, cs, cs.d_off) + kHist;
                float *out = (cs.a_S == 0) ? nullptr : chan_rbuf(a, cs, cs.al_off[0]) + kHist;
                const float *pfb = a.taps + cs.a_pfb_off;
                const unsigned long long ph0 = st.a_phase;
                const int nb = sm->a_n[0];
                int r = tid;
                for (; r + (kTailIlp - 1) * nt < nb; r += kTailIlp * nt) {
                    float acc[kTailIlp];
                    arb_r_group(z, pfb, ph0, cs.a_step, r, nt, acc);
#pragma unroll
                    for (int u = 0; u < kTailIlp; ++u) {
                        if (out) out[r + u * nt] = acc[u]; else if (r + u * nt < cs.res_audio_cap) res_audio[r + u * nt] = acc[u];
                    }
                }
                for (; r < nb; r += nt) {
                    const float acc = arb_r_one(z, pfb, ph0, cs.a_step, r);
                    if (out) out[r] = acc; else if (r < cs.res_audio_cap) res_audio[r] = acc;
                }
            } else if (j <= cs.a_S) {
                const int g = j - 1;
                const float *in = chan_rbuf(a, cs, cs.al_off[j - 1]);
                float *out = chan_rbuf(a, cs, cs.al_off[j]);
                hb_interp_r(in, sm->a_n[j - 1], out, cs.a_m[g], a.taps + cs.a_h1[g], tid, nt);
                if (j == cs.a_S) {
                    // final level: also the audio result (each thread copies what it just wrote)
                    for (int k = tid; k < sm->a_n[j - 1]; k += nt) {      // thread k wrote outputs 2k and 2k + 1
                        for (int r = 2 * k; r < 2 * k + 2; ++r)
                            if (r < sm->n_audio && r < cs.res_audio_cap) res_audio[r] = out[kHist + r];
                    }
                }
            }
        }
        return;
    }

    // ---- CW back end (ModemCW.cpp:176-208) -----------------------------------------------------------------------------
    if (ph >= TP_CW_A && ph <= TP_CW_E && is_fms) {
        // ---- FM stereo back end (ModemFMStereo.cpp:262-286) at the audio rate ------------------------------------------------
        const int na = sm->n_audio;                              // samples per channel (doubled for the result header in TP_CW_E)
        float *ld = chan_rbuf(a, cs, cs.fms_ld_off) + kFmsFirHist, *rd = chan_rbuf(a, cs, cs.fms_rd_off) + kFmsFirHist;
        if (ph == TP_CW_A) {
            // 0.568 (sum -+ difference), then iirfilt_rrrf {b0, b1} / {1, a1}: v0 = x - a1 v1, y = b0 v0 + b1 v1 (two serial
            // chains: thread 0 left, thread 32 right)
            if (tid == 0 || tid == 32) {
                const bool right = tid == 32;
                const float *m = chan_rbuf(a, cs, cs.fms_am_off), *sd = chan_rbuf(a, cs, cs.fms_as_off);
                float *out = right ? rd : ld;
                float v1 = right ? st.fms_dr : st.fms_dl;
                for (int i = 0; i < na; ++i) {
                    const float x = f_mul(0.568f, right ? f_add(m[i], sd[i]) : f_sub(m[i], sd[i]));
                    if (cs.fms_demph) {
                        const float v0 = f_sub(x, f_mul(cs.fms_da1, v1));
                        out[i] = f_add(f_mul(cs.fms_db0, v0), f_mul(cs.fms_db1, v1));
                        v1 = v0;
                    } else out[i] = x;
                }
                if (right) st.fms_dr = v1; else st.fms_dl = v1;
            }
        } else if (ph == TP_CW_B) {                              // firfilt_rrrf_push + execute on both: y[t] = sum h[k] x[t - k]
            const float *h = a.taps + cs.fms_fir_off;
            const int nh = cs.fms_fir_n;
            for (int r = tid; r < 2 * na; r += nt) {
                const int i = r >> 1;
                const float *x = (r & 1) ? rd : ld;
                float acc = 0.f;
                for (int k = 0; k < nh; ++k) acc = fmaf(sdr_ldg(h + k), x[i - k], acc);
                if (r < cs.res_audio_cap) res_audio[r] = acc;   // interleaved {left, right} (:284-285)
            }
        } else if (ph == TP_CW_C || ph == TP_CW_D) {
            // histories of the streams only this modem has; staged through the shared scratch like the common ones
            float *scr = reinterpret_cast<float *>(&sm->hist_tmp[0][0]);       // kTailStreams * kHist float2 >= 2 * 256 + 128 + 64 * 10
            const bool save = ph == TP_CW_C;
            for (int t = tid; t < kFmsFirHist; t += nt) {
                if (save) { scr[t] = ld[na - kFmsFirHist + t]; scr[kFmsFirHist + t] = rd[na - kFmsFirHist + t]; }
                else { ld[t - kFmsFirHist] = scr[t]; rd[t - kFmsFirHist] = scr[kFmsFirHist + t]; }
            }
            float *q = scr + 2 * kFmsFirHist;
            if (tid < kHist) {
                float2 *x2b = chan_cbuf(a, cs, cs.fms_x2_off);
                float2 *q2 = reinterpret_cast<float2 *>(q);
                if (save) q2[tid] = x2b[sm->n_iq + tid]; else x2b[tid] = q2[tid];
                float *dsb = chan_rbuf(a, cs, cs.fms_ds_off);
                float *q3 = q + 2 * kHist;
                if (save) q3[tid] = dsb[sm->n_iq + tid]; else dsb[tid] = q3[tid];
                for (int j = 1; j <= cs.a_S; ++j) {                       // the difference signal's decimator levels
                    float *b = chan_rbuf(a, cs, cs.fms_al_off[j]);
                    float *qj = q + 3 * kHist + (j - 1) * kHist;
                    if (save) qj[tid] = b[sm->a_n[j] + tid]; else b[tid] = qj[tid];
                }
            }
        } else if (tid == 0) {
            sm->n_audio = 2 * na;                                 // audioOut->channels = 2 (:264): the header counts floats
        }
        return;
    }
    if (ph >= TP_CW_A && ph <= TP_CW_E) {
        if (!is_cw) return;
        const int na = sm->n_audio < cs.res_audio_cap ? sm->n_audio : cs.res_audio_cap;
        float2 *w = chan_cbuf(a, cs, cs.cw_w_off) + kHist;
        if (ph == TP_CW_A) {                                     // nco_crcf_mix_up with the current phase, then step (:181-182)
            const float2 *x = chan_cbuf(a, cs, cs.al_off[cs.a_S]) + kHist;
            const unsigned th0 = st.ssb_theta;
            for (int i = tid; i < na; i += nt) {
                const float2 o = nco_lookup(a.nco, th0 + (unsigned)i * cs.cw_dtheta);
                w[i] = cf_rot(x[i], o.x, o.y);
            }
        } else if (ph == TP_CW_B) {                              // firhilbf_c2r_execute(mToReal, sig, &lsb, &out): out = the usb branch (:183)
            const float *hq = a.taps + cs.cw_hq_off;
            for (int i = tid; i < na; i += nt) {
                float yq = 0.f;
                for (int j = 0; j < 10; ++j) yq = fmaf(sdr_ldg(hq + j), w[i - 19 + 2 * j].y, yq);
                res_audio[i] = w[i - 10].x - yq;
            }
        } else if (ph == TP_CW_C) {                              // :187-197 aOutputCeil = max over the block (positive peaks)
            float mx = 0.f;
            for (int i = tid; i < na; i += nt) mx = fmaxf(mx, res_audio[i]);
            sm->red_f[tid] = mx;
        } else if (ph == TP_CW_D) {
            if (tid == 0 && na > 0) {
                float mx = 0.f;
                for (int t = 0; t < nt; ++t) mx = fmaxf(mx, sm->red_f[t]);
                const float ma = st.agc_ma + (st.agc_ceil - st.agc_ma) * 0.025f;
                const float maa = st.agc_maa + (ma - st.agc_maa) * 0.025f;
                st.agc_ma = ma; st.agc_maa = maa; st.agc_ceil = mx;
                const float gain_db = 10.0f * log10f(0.5f / maa);            // mGain (:199)
                sm->gain = powf(10.0f, gain_db / 10.0f);                     // :204
                st.ssb_theta += (unsigned)sm->n_audio * cs.cw_dtheta;
            }
        } else {
            const float g = sm->gain;
            for (int i = tid; i < na; i += nt) res_audio[i] *= g;
        }
        return;
    }

    // ---- metering: DemodulatorThread.cpp:142-163 (level), 223-233 (peak) ---------------------------------------------
    if (ph == TP_METER_A) {
        if (is_none) { sm->red_f[tid] = 0.f; sm->red_d[tid] = 0.0; return; }
        const int na = sm->n_audio < cs.res_audio_cap ? sm->n_audio : cs.res_audio_cap;
        float pk = 0.f;
        double acc = 0.0;
        for (int i = tid; i < na; i += nt) { float v = fabsf(res_audio[i]); pk = fmaxf(pk, v); if (!lvl_from_iq) acc += (double)v; }
        if (lvl_from_iq) for (int i = tid; i < sm->n_iq; i += nt) { double re = iq[i].x, im = iq[i].y; acc += sqrt(re * re + im * im); }
        sm->red_f[tid] = pk;
        sm->red_d[tid] = acc;
        return;
    }
    if (ph == TP_METER_B) {
        if (tid == 0) {
            float pk = 0.f;
            double acc = 0.0;
            for (int t = 0; t < nt; ++t) { pk = fmaxf(pk, sm->red_f[t]); acc += sm->red_d[t]; }
            const int cnt = lvl_from_iq ? sm->n_iq : sm->n_audio;
            double lvl = 0.0;
            if (sm->n_audio > 0 && cnt > 0) {
                double lin = acc / (double)cnt;
                if (lin <= 1e-20) lin = 1e-20;
                lvl = 20.0 * log10(lin);
            }
            hdr->n_audio = (unsigned)sm->n_audio; hdr->n_iq = (unsigned)sm->n_iq; hdr->peak = pk; hdr->skipped = 0; hdr->level_db = lvl;
            // Floor / ceiling trackers, smoothed level and the squelch decision, DemodulatorThread.cpp:142-198.  The members
            // are std::atomic<float>, the expressions mix float and double exactly as written there; every operation is
            // issued singly rounded (no contraction into fused multiply-adds) so that the doubles match the host's.
            // A block without decimated samples never gets this far in the reference (`if (!bufSize) continue`, :96-99).
            if (sm->n_iq > 0) {
                const double csl = lvl;                                  // 0 when the block gave no audio (:142, :145)
                const double sample_time = d_div((double)sm->n_iq, (double)cs.bandwidth);        // :143
                if (sm->n_audio > 0) {
                    float sf = st.sig_floor, sc = st.sig_ceil;
                    const float sl = cd.squelch_level;
                    if (csl > (double)sc) sc = (float)csl;
                    if (csl < (double)sf) sf = (float)csl;
                    if (f_add(sl, 1.0f) > sc) sc = f_add(sl, 1.0f);
                    if (f_add(sf, 2.0f) > sc) sc = f_add(sf, 2.0f);
                    // sc -= (sc - (csl + 2.0f)) * sampleTime * 0.05f;   sf += ((csl - 5.0f) - sf) * sampleTime * 0.15f;
                    sc = (float)d_sub((double)sc, d_mul(d_mul(d_sub((double)sc, d_add(csl, 2.0)), sample_time), (double)0.05f));
                    sf = (float)d_add((double)sf, d_mul(d_mul(d_sub(d_sub(csl, 5.0), (double)sf), sample_time), (double)0.15f));
                    st.sig_floor = sf; st.sig_ceil = sc;
                }
                const double slv = (double)st.sig_level;
                if (csl > slv) st.sig_level = (float)d_add(slv, d_mul(d_sub(csl, slv), 0.5));                             // :192-193
                else st.sig_level = (float)d_add(slv, d_mul(d_mul(d_mul(d_sub(csl, slv), 0.05), sample_time), 30.0));      // :195
            }
            hdr->sig_level = st.sig_level; hdr->sig_floor = st.sig_floor; hdr->sig_ceil = st.sig_ceil;
            hdr->squelched = (cd.squelch_enabled && st.sig_level < cd.squelch_level) ? 1 : 0;                             // :198
        }
        return;
    }

    // ---- slide every stream's history: new_hist[t] = combined[n + t] -----------------------------------------------------
    if (ph == TP_HIST_A || ph == TP_HIST_B) {
        int s = 0;
        for (int j = 0; j <= cs.n_tail; ++j, ++s) {
            float2 *b = chan_lvl(a, cs, cd, j);
            // level 0: the slid history goes in front of the OTHER buffer, where the next block's tail will look for it
            float2 *dst = (j == 0) ? chan_cbuf(a, cs, cd.lvl0_sel ? cs.lvl_off[0] : cs.lvl0_alt_off) : b;
            if (tid < kHist) { if (ph == TP_HIST_A) sm->hist_tmp[s][tid] = b[sm->lvl_n[j] + tid]; else dst[tid] = sm->hist_tmp[s][tid]; }
        }
        {
            float2 *b = chan_cbuf(a, cs, cs.iq_off);
            if (tid < kHist) { if (ph == TP_HIST_A) sm->hist_tmp[s][tid] = b[sm->n_iq + tid]; else b[tid] = sm->hist_tmp[s][tid]; }
            ++s;
        }
        if (is_ssb) {
            float2 *b = chan_cbuf(a, cs, cs.w_off);
            if (tid < kHist) { if (ph == TP_HIST_A) sm->hist_tmp[s][tid] = b[sm->n_iq + tid]; else b[tid] = sm->hist_tmp[s][tid]; }
        }
        if (is_cw) {                                             // the beat mixer's stream feeds the 21-tap Hilbert filter at the audio rate
            float2 *b = chan_cbuf(a, cs, cs.cw_w_off);
            if (tid < kHist) { if (ph == TP_HIST_A) sm->hist_tmp[s][tid] = b[sm->n_audio + tid]; else b[tid] = sm->hist_tmp[s][tid]; }
        }
        ++s;
        {
            float *b = chan_rbuf(a, cs, cs.d_off);
            if (tid < kHist) { if (ph == TP_HIST_A) sm->hist_tmp[s][tid].x = b[sm->n_iq + tid]; else b[tid] = sm->hist_tmp[s][tid].x; }
            ++s;
        }
        const int n_al = cs.a_S;      // decim: levels 1..a_S ; interp: levels 0..a_S-1 are inputs of later stages
        for (int j = 0; j <= n_al; ++j, ++s) {
            if (!cs.a_interp && j == 0) continue;               // decimator level 0 is the d stream itself
            if (cs.a_interp && j == cs.a_S && cs.a_S > 0) continue;   // final interpolator level feeds nothing
            if (cs.a_interp && cs.a_S == 0) continue;
            if (is_cw) {                                          // complex levels
                float2 *bc = chan_cbuf(a, cs, cs.al_off[j]);
                if (tid < kHist) { if (ph == TP_HIST_A) sm->hist_tmp[s][tid] = bc[sm->a_n[j] + tid]; else bc[tid] = sm->hist_tmp[s][tid]; }
                continue;
            }
            float *b = chan_rbuf(a, cs, cs.al_off[j]);
            if (tid < kHist) { if (ph == TP_HIST_A) sm->hist_tmp[s][tid].x = b[sm->a_n[j] + tid]; else b[tid] = sm->hist_tmp[s][tid].x; }
        }
        if (ph == TP_HIST_B && tid == 0) {
            st.arb_phase = sm->arb_phase_after;
            st.iq_count += (unsigned long long)sm->n_iq;
            st.a_phase = sm->a_phase_after;
        }
        return;
    }
}

}  // namespace sdrb200
