// cubicsdr_b200/csrc/dsp_spec.h -- spectrum / waterfall path.
//
// Reference: SpectrumVisualProcessor::process, non-view branch (src/process/SpectrumVisualProcessor.cpp:388-576) and
// WaterfallPanel::step (src/panel/WaterfallPanel.cpp:51-83).  Frames of N = 2 * fftSize samples (SPECTRUM_VZM = 2,
// SpectrumVisualProcessor.cpp:145), rectangular window, unnormalised forward FFT (fft_create_plan(.., LIQUID_FFT_FORWARD)
// :177), |X| in float, fftshift, per-bin double EMA pair, frame ceil/floor, scalar EMAs, log-normalised points, u8 line.
//
// Kernels:
//   spec_fft2    one CTA per (frame, sub-sequence): Stockham passes with the butterflies in registers (below); a frame longer than
//                kFft2MaxSub points is split P ways by a radix-P decimation-in-frequency stage applied while loading; |X| is written
//                fftshifted.
//   spec_ema2    one thread per bin, frames in order: maa += (ma - maa) r; ma += (x - ma) r (:494-497), fused with the frames'
//                ceil / floor of float(maa) (:499-504)
//   spec_points2 the four scalar averages up to the CTA's frame (:513-521), then one thread per display point: bin merge, log
//                normalisation, u8 (:532-576, WaterfallPanel.cpp:64-70)
#pragma once
#include "dsp_core.h"

namespace sdrb200 {

constexpr int kSpecThreads = 256;

struct SpecArgs {
    const float2 *x;          // wideband stream, x[0] = first sample of the current block (history at negative indices)
    const int *frame_start;   // [F] first sample of each frame relative to x
    int frame_stride;         // > 0: frame f starts at f * frame_stride instead (zoomed view: the hidden channel's decimated stream)
    int F;
    int N;                    // transform length (2 * fft_size)
    int fft_size;             // display points
    int P;                    // sub-transforms per frame (1, 2, 4)
    int nsub, log_nsub;       // N / P and its log2
    const float2 *tw;         // [N] exp(-j 2 pi k / N)
    float *mags;              // [F][N]  fftshifted |X|
    float *maa_f;             // [F][N]  float(maa) after each frame
    double2 *ema;             // [N] {ma, maa}
    double rate;              // (double)(float)fft_average_rate
    float2 *frame_cf;         // [F] {ceil, floor} of the frame (float, as the reference's locals)
    double *scalars;          // [4] ceil_ma, ceil_maa, floor_ma, floor_maa  (persistent)
    double *scalars_out;      // fused tail (k_spec_points2): the four averages after this block go here (the other half of a pair)
    double2 *frame_cf_maa;    // [F] {ceil_maa, floor_maa} after the frame's scalar update
    float *points;            // [F][fft_size]
    unsigned char *lines;     // [F][fft_size]
    float scale_factor;
    // bins merged into display point x: round(visual_start + i) for i = pt_i0[x] .. pt_i0[x] + pt_cnt[x] - 1
    // (the reference's visualAccum loop, :538-556, evaluated once on the host for the current view)
    const int *pt_i0, *pt_cnt;
    double visual_start;
    // zoomed view, applied to {ma, maa} before the first frame of the block: rebin < 0 / > 0 = zoomed in / out (:454-492),
    // shift = bins the averages move by after the view's centre moved (> 0: towards lower indices, :317-329)
    int rebin, shift;
    double2 *ema_tmp;         // [N] scratch for the two moves above
};

// ---- zoomed view (SpectrumVisualProcessor.cpp:283-386): the frames handed to the view are mixed down by the view's oscillator
//      (nco_crcf_mix_block_up / _down over each frame, the phase running on from frame to frame) straight into the level-0
//      stream of the view's hidden channel, whose resampler (msresamp_crcf(resampleBw / sampleRate)) is the channel tail's
struct ViewMixArgs {
    const float2 *x;          // block stream (history at negative indices)
    const int *frame_start;   // [F]
    int F, D;                 // frames this block, samples per frame
    const float2 *nco;        // [1024] {cos, sin}
    unsigned theta0, dtheta;  // oscillator phase at the first sample, phase step
    int dir;                  // +1 mix up, -1 mix down, 0 copy (:347)
    float2 *out;              // level-0 stream of the hidden channel: out[0 .. F D)
};
SDR_HD inline void view_mix_thread(const ViewMixArgs &a, long long g) {
    if (g >= (long long)a.F * a.D) return;
    const int f = (int)(g / a.D), i = (int)(g - (long long)f * a.D);
    const float2 v = a.x[a.frame_start[f] + i];
    if (!a.dir) { a.out[g] = v; return; }
    const float2 w = a.nco[(a.theta0 + (unsigned)g * a.dtheta + 0x200000u) >> 22];
    a.out[g] = cf_rot(v, w.x, a.dir < 0 ? -w.y : w.y);
}
// the two re-arrangements of the running averages; phase 0 copies into the scratch, phase 1 writes back (grid-wide barrier
// between them = two launches)
SDR_HD inline void view_ema_move_thread(const SpecArgs &a, int ph, int i) {
    if (i >= a.N) return;
    if (ph == 0) { a.ema_tmp[i] = a.ema[i]; return; }
    const int N = a.N;
    double2 v = a.ema_tmp[i];
    if (a.shift > 0) { if (i + a.shift < N) v = a.ema_tmp[i + a.shift]; }             // memmove(&ma[0], &ma[numShift], N - numShift)
    else if (a.shift < 0) { if (i + a.shift >= 0) v = a.ema_tmp[i + a.shift]; }       // memmove(&ma[numShift], &ma[0], N - numShift)
    if (a.rebin) {
        // the shift (in process(): before the transform) and the re-binning (after it) compose: rebin reads the shifted array
        auto shifted = [&](int k) -> double2 {
            double2 u = a.ema_tmp[k];
            if (a.shift > 0) { if (k + a.shift < N) u = a.ema_tmp[k + a.shift]; }
            else if (a.shift < 0) { if (k + a.shift >= 0) u = a.ema_tmp[k + a.shift]; }
            return u;
        };
        if (a.rebin < 0) v = shifted((N / 4) + (i / 2));
        else if (i < N / 4 || i >= N - N / 4) v = make_double2(0.0, 0.0);
        else v = shifted((i - N / 4) * 2);
    }
    a.ema[i] = v;
}

SDR_HD SDR_FORCEINLINE float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}

// =====================================================================================================================
// Spectrum transform: Stockham autosort passes with the butterflies in registers.
//   A frame of N points is split P ways by a radix-P decimation-in-frequency stage applied while the samples are read from
//   global memory (P = N / nsub, nsub <= 8192); one CTA transforms one sub-sequence of nsub = r0 r1 r2 points in up to three
//   passes of radix 16 / 16 / {2..32}.  In a pass thread j reads its r inputs x[j + t n / r] (consecutive threads, consecutive
//   addresses), multiplies by exp(-2 pi i t k / (s r)), k = j mod s, transforms them in registers and writes them back to
//   y[(j - k) r + k + t s]; every thread has read before anyone writes (one buffer, two barriers per pass).  The stride-r0 writes
//   of the first pass are skewed by one element per r0 so that the 32 lanes of a warp fall on distinct banks (a' = a + a / r0); the
//   later passes write runs of >= 16 consecutive elements.  The last pass writes |X| straight to global memory, fftshifted.
// =====================================================================================================================
constexpr int kFft2MaxSub = 8192;
constexpr int kFft2Threads = 512;
constexpr int kFft2SmemElems = kFft2MaxSub + kFft2MaxSub / 16;

SDR_HD SDR_FORCEINLINE float2 w32_const(int m) {          // exp(-2 pi i m / 32); m is a compile-time constant after unrolling
    switch (m & 31) {
        case 0: return make_float2(1.000000000e+00f, 0.0f);
        case 1: return make_float2(9.807852804e-01f, -1.950903220e-01f);
        case 2: return make_float2(9.238795325e-01f, -3.826834324e-01f);
        case 3: return make_float2(8.314696123e-01f, -5.555702330e-01f);
        case 4: return make_float2(7.071067812e-01f, -7.071067812e-01f);
        case 5: return make_float2(5.555702330e-01f, -8.314696123e-01f);
        case 6: return make_float2(3.826834324e-01f, -9.238795325e-01f);
        case 7: return make_float2(1.950903220e-01f, -9.807852804e-01f);
        case 8: return make_float2(0.0f, -1.000000000e+00f);
        case 9: return make_float2(-1.950903220e-01f, -9.807852804e-01f);
        case 10: return make_float2(-3.826834324e-01f, -9.238795325e-01f);
        case 11: return make_float2(-5.555702330e-01f, -8.314696123e-01f);
        case 12: return make_float2(-7.071067812e-01f, -7.071067812e-01f);
        case 13: return make_float2(-8.314696123e-01f, -5.555702330e-01f);
        case 14: return make_float2(-9.238795325e-01f, -3.826834324e-01f);
        case 15: return make_float2(-9.807852804e-01f, -1.950903220e-01f);
        case 16: return make_float2(-1.000000000e+00f, 0.0f);
        case 17: return make_float2(-9.807852804e-01f, 1.950903220e-01f);
        case 18: return make_float2(-9.238795325e-01f, 3.826834324e-01f);
        case 19: return make_float2(-8.314696123e-01f, 5.555702330e-01f);
        case 20: return make_float2(-7.071067812e-01f, 7.071067812e-01f);
        case 21: return make_float2(-5.555702330e-01f, 8.314696123e-01f);
        case 22: return make_float2(-3.826834324e-01f, 9.238795325e-01f);
        case 23: return make_float2(-1.950903220e-01f, 9.807852804e-01f);
        case 24: return make_float2(0.0f, 1.000000000e+00f);
        case 25: return make_float2(1.950903220e-01f, 9.807852804e-01f);
        case 26: return make_float2(3.826834324e-01f, 9.238795325e-01f);
        case 27: return make_float2(5.555702330e-01f, 8.314696123e-01f);
        case 28: return make_float2(7.071067812e-01f, 7.071067812e-01f);
        case 29: return make_float2(8.314696123e-01f, 5.555702330e-01f);
        case 30: return make_float2(9.238795325e-01f, 3.826834324e-01f);
        case 31: return make_float2(9.807852804e-01f, 1.950903220e-01f);
    }
    return make_float2(1.0f, 0.0f);
}
template <int R> SDR_HD SDR_FORCEINLINE int bitrev_r(int i) {
    int o = 0;
    for (int b = 1; b < R; b <<= 1) { o = (o << 1) | (i & 1); i >>= 1; }
    return o;
}
// in-register decimation-in-frequency transform of R points: X[bitrev_r<R>(i)] = u[i] afterwards
template <int R> SDR_HD SDR_FORCEINLINE void reg_fft(float2 (&u)[R]) {
#pragma unroll
    for (int len = R; len >= 2; len >>= 1) {
        const int half = len >> 1;
#pragma unroll
        for (int blk = 0; blk < R; blk += len) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const float2 p = u[blk + j], q = u[blk + j + half];
                u[blk + j] = make_float2(p.x + q.x, p.y + q.y);
                const float2 d = make_float2(p.x - q.x, p.y - q.y);
                const int m = j * (32 / len);                     // W_len^j = W_32^(j 32 / len)
                if (m == 0) u[blk + j + half] = d;
                else if (m == 8) u[blk + j + half] = make_float2(d.y, -d.x);          // times -j
                else u[blk + j + half] = cmul(d, w32_const(m));
            }
        }
    }
}
SDR_HD SDR_FORCEINLINE int fft2_pad(int a, int lpr) { return lpr ? a + (a >> lpr) : a; }        // lpr = log2 of the skew period, 0 = plain

// the radices of the passes for a sub-sequence of n points (n a power of two, 2 <= n <= kFft2MaxSub)
struct Fft2Plan { int passes; int r[3]; };
SDR_HD inline Fft2Plan fft2_plan(int n) {
    Fft2Plan p; p.passes = 0; p.r[0] = p.r[1] = p.r[2] = 1;
    int rem = n;
    while (rem > 32) { p.r[p.passes++] = 16; rem /= 16; }
    if (rem > 1 || p.passes == 0) p.r[p.passes++] = rem;
    return p;
}

// sample n of sub-sequence psub of frame x (radix-P DIF stage on load): (sum_r x[n + r nsub] W_P^(r psub)) W_N^(n psub)
SDR_HD SDR_FORCEINLINE float2 fft2_split_load(const SpecArgs &a, const float2 *x, int psub, int n) {
    if (a.P == 1) return x[n];
    const int nsub = a.nsub, N = a.N;
    float2 v;
    if (a.P == 2) {
        const float2 x0 = x[n], x1 = x[n + nsub];
        v = psub == 0 ? make_float2(x0.x + x1.x, x0.y + x1.y) : make_float2(x0.x - x1.x, x0.y - x1.y);
    } else if (a.P == 4) {
        const float2 x0 = x[n], x1 = x[n + nsub], x2 = x[n + 2 * nsub], x3 = x[n + 3 * nsub];
        const float2 s02 = make_float2(x0.x + x2.x, x0.y + x2.y), d02 = make_float2(x0.x - x2.x, x0.y - x2.y);
        const float2 s13 = make_float2(x1.x + x3.x, x1.y + x3.y), d13 = make_float2(x1.x - x3.x, x1.y - x3.y);
        if (psub == 0) v = make_float2(s02.x + s13.x, s02.y + s13.y);
        else if (psub == 1) v = make_float2(d02.x + d13.y, d02.y - d13.x);      // d02 - j d13
        else if (psub == 2) v = make_float2(s02.x - s13.x, s02.y - s13.y);
        else v = make_float2(d02.x - d13.y, d02.y + d13.x);                      // d02 + j d13
    } else {                                                                     // any power of two: direct sum
        v = x[n];
        for (int r = 1; r < a.P; ++r) {
            const float2 w = sdr_ldg(a.tw + ((r * psub) & (a.P - 1)) * nsub);       // W_P^(r psub) = W_N^(r psub N / P), N / P = nsub
            const float2 t = cmul(x[n + r * nsub], w);
            v.x += t.x; v.y += t.y;
        }
    }
    if (psub) v = cmul(v, sdr_ldg(a.tw + ((n * psub) & (N - 1))));
    return v;
}

// one pass for butterfly j: read + twiddle + transform into u
template <int R>
SDR_HD SDR_FORCEINLINE void fft2_pass_load(const SpecArgs &a, const float2 *buf, int pr_in, int s, int j, float2 (&u)[R]) {
    const int n = a.nsub, q = n / R;
#pragma unroll
    for (int t = 0; t < R; ++t) u[t] = buf[fft2_pad(j + t * q, pr_in)];
    if (s > 1) {
        // exp(-2 pi i t k / (s R)) = tw[t k N / (s R)], t = 1 .. R - 1.  The table is read only at t = 1, 2, 4, .. (the reads of a
        // warp are scattered: k differs per lane), the other powers are products of at most log2 R of those (<= 3e-7 off)
        const int k = j & (s - 1);
        const int tstep = k * (a.N / (s * R));
        float2 w[R];
#pragma unroll
        for (int b = 1; b < R; b <<= 1) w[b] = sdr_ldg(a.tw + b * tstep);
#pragma unroll
        for (int t = 3; t < R; ++t) {
            const int low = t & (-t);
            if (low != t) w[t] = cmul(w[t - low], w[low]);
        }
#pragma unroll
        for (int t = 1; t < R; ++t) u[t] = cmul(u[t], w[t]);
    }
    reg_fft<R>(u);
}
template <int R>
SDR_HD SDR_FORCEINLINE void fft2_pass_store(float2 *buf, int pr_out, int s, int j, const float2 (&u)[R]) {
    const int k = j & (s - 1), base = (j - k) * R + k;
#pragma unroll
    for (int i = 0; i < R; ++i) buf[fft2_pad(base + bitrev_r<R>(i) * s, pr_out)] = u[i];
}
// last pass: |X[k]| fftshifted straight to global memory: result[i] = |X[(i + N/2) mod N]|   (SpectrumVisualProcessor.cpp:441-452)
template <int R>
SDR_HD SDR_FORCEINLINE void fft2_pass_store_mags(const SpecArgs &a, int f, int psub, int s, int j, const float2 (&u)[R]) {
    const int k = j & (s - 1), base = (j - k) * R + k, N = a.N;
    float *out = a.mags + (size_t)f * N;
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int ks = base + bitrev_r<R>(i) * s;               // index inside the sub-sequence's transform
        const int kk = ks * a.P + psub;                          // bin of the frame
        out[(kk + N / 2) & (N - 1)] = sqrtf(u[i].x * u[i].x + u[i].y * u[i].y);
    }
}
// The whole transform of one (frame, sub-sequence) for thread tid, as a sequence of half-phases separated by barriers:
//   hp = 2 i     : pass i reads (+ transforms) into the thread's registers `u`
//   hp = 2 i + 1 : pass i writes
// (the GPU kernel keeps `u` in registers across the barrier; the emulator keeps one array per thread)
struct Fft2Regs { float2 u[32]; };
template <int R>
SDR_HD SDR_FORCEINLINE void fft2_half(const SpecArgs &a, int f, int psub, float2 *buf, int pass, int passes, int s, int pr_prev,
                                      int pr_this, bool write_half, int j, Fft2Regs &rg) {
    float2 (&u)[R] = reinterpret_cast<float2 (&)[R]>(rg.u);
    if (j >= a.nsub / R) return;
    if (!write_half) {
        fft2_pass_load<R>(a, buf, pr_prev, s, j, u);
    } else {
        if (pass == passes - 1) fft2_pass_store_mags<R>(a, f, psub, s, j, u);
        else fft2_pass_store<R>(buf, pr_this, s, j, u);
    }
}
// load phase: the sub-sequence (radix-P stage applied) into the work buffer, plain layout; nt threads, a few elements each with
// their loads independent of one another (the passes then never wait on global memory)
SDR_HD inline void spec_fft2_load(const SpecArgs &a, int bid, int tid, int nt, float2 *buf) {
    const int f = bid / a.P, psub = bid - f * a.P;
    const float2 *x = a.x + (a.frame_stride > 0 ? f * a.frame_stride : a.frame_start[f]);
#pragma unroll 4
    for (int n = tid; n < a.nsub; n += nt) buf[n] = fft2_split_load(a, x, psub, n);
}
SDR_HD inline void spec_fft2_half(int hp, const SpecArgs &a, int bid, int tid, float2 *buf, Fft2Regs &rg) {
    const int f = bid / a.P, psub = bid - f * a.P;
    const Fft2Plan pl = fft2_plan(a.nsub);
    const int pass = hp >> 1;
    if (pass >= pl.passes) return;
    int s = 1;
    for (int i = 0; i < pass; ++i) s *= pl.r[i];
    // layout of the buffer: skewed by r0 after pass 0 (stride-r0 writes), plain after pass 1
    int lr0 = 0;
    while ((1 << lr0) < pl.r[0]) ++lr0;
    const int pr_prev = (pass == 1) ? lr0 : 0, pr_this = (pass == 0) ? lr0 : 0;
    const bool wh = (hp & 1) != 0;
    switch (pl.r[pass]) {
        case 32: fft2_half<32>(a, f, psub, buf, pass, pl.passes, s, pr_prev, pr_this, wh, tid, rg); break;
        case 16: fft2_half<16>(a, f, psub, buf, pass, pl.passes, s, pr_prev, pr_this, wh, tid, rg); break;
        case 8: fft2_half<8>(a, f, psub, buf, pass, pl.passes, s, pr_prev, pr_this, wh, tid, rg); break;
        case 4: fft2_half<4>(a, f, psub, buf, pass, pl.passes, s, pr_prev, pr_this, wh, tid, rg); break;
        default: fft2_half<2>(a, f, psub, buf, pass, pl.passes, s, pr_prev, pr_this, wh, tid, rg); break;
    }
}

// ---- fused tail of the spectrum path -------------------------------------------------------------------------------------
// (1) per-bin averages over the block's frames and the frames' ceil / floor in one pass: every thread owns a bin and walks the
//     frames in order (:494-497); the ceil / floor of float(maa) (:499-504: ceil starts at 0, floor at 1) are folded per warp and
//     merged with one atomic max / min per warp and frame (non-negative floats order like their bit patterns).
//     frame_cf must hold {0, 1} per frame on entry (the transform kernel's CTA 0 sees to that).
SDR_HD SDR_FORCEINLINE void spec_ema2_step(const SpecArgs &a, int f, int i, double &ma, double &maa, float &out) {
    const double xv = (double)a.mags[(size_t)f * a.N + i];
    maa += (ma - maa) * a.rate;                 // uses the OLD ma (:496)
    ma += (xv - ma) * a.rate;
    out = (float)maa;
    a.maa_f[(size_t)f * a.N + i] = out;
}
SDR_HD inline void spec_ema2_thread_serial(const SpecArgs &a, int i) {          // emulator form
    if (i >= a.N) return;
    double ma = a.ema[i].x, maa = a.ema[i].y;
    for (int f = 0; f < a.F; ++f) {
        float v;
        spec_ema2_step(a, f, i, ma, maa, v);
        if (v > a.frame_cf[f].x) a.frame_cf[f].x = v;
        if (v < a.frame_cf[f].y) a.frame_cf[f].y = v;
    }
    a.ema[i] = make_double2(ma, maa);
}
// (2) scalar averages up to and including frame f (:513-521), from the block's entry values
SDR_HD inline double2 spec_scalar_upto(const SpecArgs &a, int f, double *after /* [4] or null */) {
    double cma = a.scalars[0], cmaa = a.scalars[1], fma_ = a.scalars[2], fmaa = a.scalars[3];
    for (int g = 0; g <= f; ++g) {
        const double c = (double)a.frame_cf[g].x, fl = (double)a.frame_cf[g].y;
        cma = cma + (c - cma) * 0.05;
        cmaa = cmaa + (cma - cmaa) * 0.05;
        fma_ = fma_ + (fl - fma_) * 0.05;
        fmaa = fmaa + (fma_ - fmaa) * 0.05;
    }
    if (after) { after[0] = cma; after[1] = cmaa; after[2] = fma_; after[3] = fmaa; }
    return make_double2(cmaa, fmaa);
}
//     display point xp of frame f (:532-576, WaterfallPanel.cpp:64-70)
SDR_HD inline void spec_points2_thread(const SpecArgs &a, int f, int xp, double point_ceil, double point_floor) {
    if (xp >= a.fft_size) return;
    const float *m = a.maa_f + (size_t)f * a.N;
    double acc = 0.0;
    const int i0 = a.pt_i0[xp], cnt = a.pt_cnt[xp];
    for (int j = 0; j < cnt; ++j) {
        const unsigned idx = (unsigned)round(a.visual_start + (double)(i0 + j));
        acc += (idx > 0 && idx < (unsigned)a.N) ? (double)m[idx] : point_floor;
    }
    const double num = log10((acc / (double)cnt) + 0.25 - (point_floor - 0.75));
    const double den = log10((point_ceil + 0.25) - (point_floor - 0.75));
    const float y = (float)((num / den) * (double)a.scale_factor);
    const size_t gid = (size_t)f * a.fft_size + xp;
    a.points[gid] = y;
    const float wv = y < 0 ? 0.f : (y > 0.99 ? (float)0.99 : y);
    a.lines[gid] = (unsigned char)floor((double)wv * 255.0);
}

}  // namespace sdrb200
