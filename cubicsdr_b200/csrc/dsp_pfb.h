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
// from shared memory); phase 2 is one radix-2 decimation-in-frequency step (M is always even: getOptimalChannelCount),
//     A[c] = X[c] + X[c + M/2],   B[c] = (X[c] - X[c + M/2]) exp(-j 2 pi c / M),      y[2r] = DFT_{M/2}(A)[r],  y[2r+1] = DFT_{M/2}(B)[r]
// and phase 3 evaluates the two M/2-point DFTs directly from the shared twiddle table, four frames per thread so that every
// twiddle fetched is used four times (M/2 is 61 at 61.44 MSps: prime, so no further cheap split; M^2/2 multiply-adds per
// frame are small next to the per-channel demodulators).  HBM traffic: 8 B in + 8 B out per sample.
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
    int osf;               // 1: firpfbch (critically sampled, one frame per M inputs); 2: firpfbch2 (one frame per M/2 inputs)
    long long frame_base;  // osf 2: index of this call's first output frame since reset (its parity selects the branch phase)
    long long chan_stride; // 0: y is frame-major [F][M] (liquid's layout); else channel-major: y[k * chan_stride + f]
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
        // input tile: osf 1 -> frames of M samples, p-1 frames of history; osf 2 -> half-frames of M/2 samples, 2p-1 of history
        const int D = a.osf == 2 ? M / 2 : M, hist = a.osf == 2 ? 2 * p - 1 : p - 1;
        const float2 *src = a.x + (long long)(f0 - hist) * D;
        const int n = (T + hist) * D;
        for (int i = tid; i < n; i += nt) sx[i] = src[i];
        for (int i = tid; i < M; i += nt) stw[i] = a.tw[i];
    } else if (ph == 1 && a.osf == 2) {
        // firpfbch2 (liquid 1.5.0, pinned against the binary): with D = M/2, frame n, branch c (cm = c mod D, pc = c >= D):
        //   X[c] = sum_{q=0}^{2m-1} g[cm + D d] x[(n - d) D + (D-1-cm)],  d = ((n + pc) & 1) + 2q,   g = h / sum(h), fc = 1/M
        // (a sample pushed in half-frame e only ever feeds branches of parity e; the output transform is the inverse DFT / M)
        const int D = M / 2, hist = 2 * p - 1;
        for (int o = tid; o < T * M; o += nt) {
            const int fl = o / M, c = o - fl * M;
            const int cm = c >= D ? c - D : c, pc = c >= D ? 1 : 0;
            const int d0 = (int)((a.frame_base + f0 + fl + pc) & 1);
            const float *h = a.hq + (size_t)cm * (2 * p) + d0;             // hq[cm][d], d = d0 + 2q
            const float2 *xc = sx + (size_t)(fl + hist - d0) * D + (D - 1 - cm);   // x[(n - d0) D + ...]; next tap 2 D earlier
            float2 acc = make_float2(0.f, 0.f);
            for (int q = 0; q < p; ++q) acc = cf_fma(sdr_ldg(h + 2 * q), xc[-(long long)(2 * q) * D], acc);
            su[o] = acc;
        }
    } else if (ph == 1) {
        for (int o = tid; o < T * M; o += nt) {
            const int fl = o / M, c = o - fl * M;
            const float *h = a.hq + (size_t)c * p;
            const float2 *xc = sx + (size_t)(fl + p - 1) * M + c;           // x[(f - q) M + c] = xc[-q M]
            float2 acc = make_float2(0.f, 0.f);
            for (int q = 0; q < p; ++q) acc = cf_fma(sdr_ldg(h + q), xc[-(long long)q * M], acc);
            su[o] = acc;
        }
    } else if (ph == 2) {
        const int H = M >> 1;
        for (int o = tid; o < T * H; o += nt) {
            const int fl = o / H, c = o - fl * H;
            float2 *u = su + (size_t)fl * M;
            const float2 x0 = u[c], x1 = u[c + H], w = stw[c];
            const float dr = x0.x - x1.x, di = x0.y - x1.y;
            u[c] = make_float2(x0.x + x1.x, x0.y + x1.y);
            u[c + H] = make_float2(dr * w.x - di * w.y, dr * w.y + di * w.x);
        }
    } else {
        const int H = M >> 1;
        const int groups = (T + 3) >> 2;                                    // four frames per work item
        for (int o = tid; o < groups * M; o += nt) {
            const int g = o / M, k = o - g * M;
            const int r = k >> 1;
            const int fl0 = g * 4;
            const float2 *u0 = su + (size_t)fl0 * M + ((k & 1) ? H : 0);
            const bool v1 = fl0 + 1 < T, v2 = fl0 + 2 < T, v3 = fl0 + 3 < T;
            const float2 *u1 = v1 ? u0 + M : u0, *u2 = v2 ? u0 + 2 * M : u0, *u3 = v3 ? u0 + 3 * M : u0;
            float re0 = 0.f, im0 = 0.f, re1 = 0.f, im1 = 0.f, re2 = 0.f, im2 = 0.f, re3 = 0.f, im3 = 0.f;
            int idx = 0;                                                    // (2 c r) mod M, kept incrementally
            const int step = 2 * r;                                         // < M
            for (int c = 0; c < H; ++c) {
                const float2 w = stw[idx];
                const float2 a0 = u0[c], a1 = u1[c], a2 = u2[c], a3 = u3[c];
                re0 = fmaf(a0.x, w.x, re0); re0 = fmaf(-a0.y, w.y, re0); im0 = fmaf(a0.x, w.y, im0); im0 = fmaf(a0.y, w.x, im0);
                re1 = fmaf(a1.x, w.x, re1); re1 = fmaf(-a1.y, w.y, re1); im1 = fmaf(a1.x, w.y, im1); im1 = fmaf(a1.y, w.x, im1);
                re2 = fmaf(a2.x, w.x, re2); re2 = fmaf(-a2.y, w.y, re2); im2 = fmaf(a2.x, w.y, im2); im2 = fmaf(a2.y, w.x, im2);
                re3 = fmaf(a3.x, w.x, re3); re3 = fmaf(-a3.y, w.y, re3); im3 = fmaf(a3.x, w.y, im3); im3 = fmaf(a3.y, w.x, im3);
                idx += step; if (idx >= M) idx -= M;
            }
            // frame-major: M apart; channel-major (engine channelizer mode): the four frames of a channel are adjacent
            const long long fs = a.chan_stride ? 1 : M;
            float2 *y = a.chan_stride ? a.y + (long long)k * a.chan_stride + (f0 + fl0) : a.y + (size_t)(f0 + fl0) * M + k;
            y[0] = make_float2(re0, im0);
            if (v1) y[fs] = make_float2(re1, im1);
            if (v2) y[2 * fs] = make_float2(re2, im2);
            if (v3) y[3 * fs] = make_float2(re3, im3);
        }
    }
}

constexpr int kPfbPhases = 4;

// Channel 0 of the filterbank sits on the tuner's centre frequency and gets the DC blocker in channelizer mode
// (SDRPostThread::runDemodChannels, SDRPostThread.cpp:363-375): the same float recurrence as the wideband one, at the
// channel rate, so short enough to run serially (n = block / M samples).
SDR_HD inline void pfb_dc0_serial(const float2 *in, float2 *out, int n, float c, float2 *state) {
    float2 v = state[0], y;
    for (int i = 0; i < n; ++i) { dc_ref_step(c, in[i], v, y); out[i] = y; }
    state[0] = v;
}

}  // namespace sdrb200
