// cubicsdr_b200/csrc/dsp_pfb.h -- polyphase filterbank channelizer (analysis), SURVEY.md section 8(f) row 1.
//
// Reference: SDRPostThread::runPFBCH (src/sdr/SDRPostThread.cpp:416-455) calls liquid-dsp's
// firpfbch_crcf_analyzer_execute once per M input samples; the object comes from
// firpfbch_crcf_create_kaiser(LIQUID_ANALYZER, M, m = 4, As = 60) (:406).  liquid 1.5.0 is present only as a binary; its
// published algorithm, pinned here against that binary (tests/test_pfb.py), is for frame f (input samples x[fM .. fM+M-1]):
//     X[c]  = sum_{q=0}^{2m-1} h[(M-1-c) + qM] * x[(f-q)M + c]            c = 0..M-1    (M branch filters, 2m taps each)
//     y[k]  = sum_{c=0}^{M-1} X[c] * exp(-j 2 pi c k / M)                   k = 0..M-1    (forward DFT)
// with h = Kaiser-windowed sinc of length 2Mm+1, cutoff 0.5/M (only the first 2Mm taps are used).  Output y is frame-major,
// M interleaved channel samples per frame, exactly the layout runDemodChannels de-interleaves (:366-382).
//
// GPU form: a CTA owns a tile of kPfbTileFrames frames.  Phase 0 stages the tile's input (plus 2m-1 frames of history)
// in shared memory with coalesced loads; phase 1 runs the branch filters (one thread per (frame, branch), 2m multiply-adds
// from shared memory); phase 2 evaluates the M-point DFT of every frame directly from a shared twiddle table (M is even
// but rarely a power of two -- 122 at 61.44 MSps -- and M^2 work per frame is small next to the per-channel demodulators).
// HBM traffic: 8 B in + 8 B out per sample.
#pragma once
#include "dsp_core.h"

namespace sdrb200 {

constexpr int kPfbTileFrames = 16;
constexpr int kPfbThreads = 256;
constexpr int kPfbMaxChannels = 256;

struct PfbArgs {
    const float2 *x;       // x[0] = first sample of this call; x[-(p-1)*M .. -1] = the previous (p-1) frames
    float2 *y;             // [F][M]
    const float *hq;       // [M][p]: hq[c*p + q] = h[(M-1-c) + q*M]
    const float2 *tw;      // [M]: exp(-j 2 pi n / M)
    int M, p, F;
};

// shared memory layout (float2): sx[(T + p - 1) * M] | su[T * M] | stw[M]
SDR_HD inline size_t pfb_smem_elems(int M, int p) { return (size_t)(kPfbTileFrames + p - 1) * M + (size_t)kPfbTileFrames * M + M; }

SDR_HD inline void pfb_phase(int ph, const PfbArgs &a, int tile, int tid, int nt, float2 *sm) {
    const int M = a.M, p = a.p;
    const int f0 = tile * kPfbTileFrames;
    int T = a.F - f0;
    if (T > kPfbTileFrames) T = kPfbTileFrames;
    if (T <= 0) return;
    float2 *sx = sm;
    float2 *su = sm + (size_t)(kPfbTileFrames + p - 1) * M;
    float2 *stw = su + (size_t)kPfbTileFrames * M;
    if (ph == 0) {
        const float2 *src = a.x + (long long)(f0 - (p - 1)) * M;           // history frames first
        const int n = (T + p - 1) * M;
        for (int i = tid; i < n; i += nt) sx[i] = src[i];
        for (int i = tid; i < M; i += nt) stw[i] = a.tw[i];
    } else if (ph == 1) {
        for (int o = tid; o < T * M; o += nt) {
            const int fl = o / M, c = o - fl * M;
            const float *h = a.hq + (size_t)c * p;
            const float2 *xc = sx + (size_t)(fl + p - 1) * M + c;           // x[(f - q) M + c] = xc[-q M]
            float2 acc = make_float2(0.f, 0.f);
            for (int q = 0; q < p; ++q) acc = cf_fma(sdr_ldg(h + q), xc[-(long long)q * M], acc);
            su[o] = acc;
        }
    } else {
        for (int o = tid; o < T * M; o += nt) {
            const int fl = o / M, k = o - fl * M;
            const float2 *u = su + (size_t)fl * M;
            float re = 0.f, im = 0.f;
            int idx = 0;                                                    // (c * k) mod M, kept incrementally
            for (int c = 0; c < M; ++c) {
                const float2 w = stw[idx], v = u[c];
                re = fmaf(v.x, w.x, re); re = fmaf(-v.y, w.y, re);
                im = fmaf(v.x, w.y, im); im = fmaf(v.y, w.x, im);
                idx += k; if (idx >= M) idx -= M;
            }
            a.y[(size_t)(f0 + fl) * M + k] = make_float2(re, im);
        }
    }
}

}  // namespace sdrb200
