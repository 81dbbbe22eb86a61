// cubicsdr_b200/csrc/dsp_spec.h -- spectrum / waterfall path.
//
// Reference: SpectrumVisualProcessor::process, non-view branch (src/process/SpectrumVisualProcessor.cpp:388-576) and
// WaterfallPanel::step (src/panel/WaterfallPanel.cpp:51-83).  Frames of N = 2 * fftSize samples (SPECTRUM_VZM = 2,
// SpectrumVisualProcessor.cpp:145), rectangular window, unnormalised forward FFT (fft_create_plan(.., LIQUID_FFT_FORWARD)
// :177), |X| in float, fftshift, per-bin double EMA pair, frame ceil/floor, scalar EMAs, log-normalised points, u8 line.
//
// Kernels (phases as in dsp_core.h):
//   spec_fft   one CTA per (frame, sub-transform): decimation-in-frequency FFT in shared memory.  A frame longer than
//              kFftMaxSub points is split P ways by a radix-P first stage applied while loading, so each CTA still
//              transforms <= 16384 points in place; magnitudes are written fftshifted.
//   spec_ema   one thread per bin, frames in order: maa += (ma - maa) r; ma += (x - ma) r   (:494-497)
//   spec_minmax one CTA per frame: ceil/floor of float(maa)                               (:499-504)
//   spec_scalar one thread: the four scalar EMAs per frame                                (:513-521)
//   spec_points one thread per display point: bin merge, log normalisation, u8             (:532-576, WaterfallPanel.cpp:64-70)
#pragma once
#include "dsp_core.h"

namespace sdrb200 {

constexpr int kFftMaxSub = 16384;     // points one CTA transforms in shared memory (128 KB of float2)
constexpr int kFftThreads = 512;
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

// position p in the in-place DIF result holds X[k]; stages are: one radix-2 (if log_nsub odd) then radix-4's
SDR_HD inline int dif_output_index(int p, int log_n) {
    int k = 0, w = 1, rem_bits = log_n;
    if (log_n & 1) { k = (p >> (log_n - 1)) & 1; w = 2; rem_bits = log_n - 1; }
    while (rem_bits > 0) {
        int d = (p >> (rem_bits - 2)) & 3;
        k += d * w;
        w <<= 2;
        rem_bits -= 2;
    }
    return k;
}
// inverse map: which position holds output k
SDR_HD inline int dif_position_of(int k, int log_n) {
    int p = 0, rem_bits = log_n;
    if (log_n & 1) { p |= (k & 1) << (log_n - 1); k >>= 1; rem_bits = log_n - 1; }
    while (rem_bits > 0) {
        p |= (k & 3) << (rem_bits - 2);
        k >>= 2;
        rem_bits -= 2;
    }
    return p;
}

SDR_HD SDR_FORCEINLINE float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -(a.y * b.y)), fmaf(a.x, b.y, a.y * b.x));
}

// number of barrier-separated phases of the FFT kernel for a given log_nsub: load, stages, store
SDR_HD inline int fft_phase_count(int log_nsub) { return 1 + (log_nsub & 1) + (log_nsub >> 1) + 1; }

SDR_HD inline void spec_fft_phase(int ph, const SpecArgs &a, int bid, int tid, int nt, float2 *sm) {
    const int f = bid / a.P, psub = bid - f * a.P;
    const int nsub = a.nsub, N = a.N;
    const float2 *x = a.x + (a.frame_stride > 0 ? f * a.frame_stride : a.frame_start[f]);
    const int n_r2 = a.log_nsub & 1, n_r4 = a.log_nsub >> 1;
    if (ph == 0) {
        // load with the radix-P decimation-in-frequency first stage: y_p[n] = (sum_r x[n + r N/P] W_P^{r p}) W_N^{n p}
        for (int n = tid; n < nsub; n += nt) {
            float2 v;
            if (a.P == 1) v = x[n];
            else if (a.P == 2) {
                float2 x0 = x[n], x1 = x[n + nsub];
                v = psub == 0 ? make_float2(x0.x + x1.x, x0.y + x1.y) : make_float2(x0.x - x1.x, x0.y - x1.y);
                if (psub) v = cmul(v, sdr_ldg(a.tw + n));
            } else {
                float2 x0 = x[n], x1 = x[n + nsub], x2 = x[n + 2 * nsub], x3 = x[n + 3 * nsub];
                float2 s02 = make_float2(x0.x + x2.x, x0.y + x2.y), d02 = make_float2(x0.x - x2.x, x0.y - x2.y);
                float2 s13 = make_float2(x1.x + x3.x, x1.y + x3.y), d13 = make_float2(x1.x - x3.x, x1.y - x3.y);
                if (psub == 0) v = make_float2(s02.x + s13.x, s02.y + s13.y);
                else if (psub == 1) v = make_float2(d02.x + d13.y, d02.y - d13.x);      // d02 - j d13
                else if (psub == 2) v = make_float2(s02.x - s13.x, s02.y - s13.y);
                else v = make_float2(d02.x - d13.y, d02.y + d13.x);                      // d02 + j d13
                if (psub) v = cmul(v, sdr_ldg(a.tw + ((long long)n * psub) % N));
            }
            sm[n] = v;
        }
        return;
    }
    const int last = 1 + n_r2 + n_r4;
    if (ph < last) {
        int stage = ph - 1;
        if (n_r2 && stage == 0) {
            const int half = nsub >> 1, tstride = N / nsub;       // W_nsub^j = W_N^{j * N / nsub}
            for (int j = tid; j < half; j += nt) {
                float2 x0 = sm[j], x1 = sm[j + half];
                sm[j] = make_float2(x0.x + x1.x, x0.y + x1.y);
                sm[j + half] = cmul(make_float2(x0.x - x1.x, x0.y - x1.y), sdr_ldg(a.tw + j * tstride));
            }
            return;
        }
        if (n_r2) stage -= 1;
        // radix-4 stage `stage` over blocks of length L
        const int L = (nsub >> n_r2) >> (2 * stage), Q = L >> 2, tstride = N / L;
        const int n_bfly = nsub >> 2;
        for (int id = tid; id < n_bfly; id += nt) {
            const int blk = id / Q, j = id - blk * Q, i0 = blk * L + j;
            float2 x0 = sm[i0], x1 = sm[i0 + Q], x2 = sm[i0 + 2 * Q], x3 = sm[i0 + 3 * Q];
            float2 s02 = make_float2(x0.x + x2.x, x0.y + x2.y), d02 = make_float2(x0.x - x2.x, x0.y - x2.y);
            float2 s13 = make_float2(x1.x + x3.x, x1.y + x3.y), d13 = make_float2(x1.x - x3.x, x1.y - x3.y);
            float2 y0 = make_float2(s02.x + s13.x, s02.y + s13.y);
            float2 y1 = make_float2(d02.x + d13.y, d02.y - d13.x);
            float2 y2 = make_float2(s02.x - s13.x, s02.y - s13.y);
            float2 y3 = make_float2(d02.x - d13.y, d02.y + d13.x);
            if (Q > 1) {
                y1 = cmul(y1, sdr_ldg(a.tw + j * tstride));
                y2 = cmul(y2, sdr_ldg(a.tw + 2 * j * tstride));
                y3 = cmul(y3, sdr_ldg(a.tw + 3 * j * tstride));
            }
            sm[i0] = y0; sm[i0 + Q] = y1; sm[i0 + 2 * Q] = y2; sm[i0 + 3 * Q] = y3;
        }
        return;
    }
    // store: |X[k]| fftshifted:  result[i] = |X[(i + N/2) mod N]|           (SpectrumVisualProcessor.cpp:441-452)
    float *out = a.mags + (size_t)f * N;
    for (int ks = tid; ks < nsub; ks += nt) {
        float2 v = sm[dif_position_of(ks, a.log_nsub)];
        const int k = ks * a.P + psub;
        out[(k + N / 2) & (N - 1)] = sqrtf(v.x * v.x + v.y * v.y);
    }
}

// one thread per bin
SDR_HD inline void spec_ema_thread(const SpecArgs &a, int i) {
    if (i >= a.N) return;
    double ma = a.ema[i].x, maa = a.ema[i].y;
    const double r = a.rate;
    for (int f = 0; f < a.F; ++f) {
        const double xv = (double)a.mags[(size_t)f * a.N + i];
        maa += (ma - maa) * r;                 // uses the OLD ma (:496)
        ma += (xv - ma) * r;
        a.maa_f[(size_t)f * a.N + i] = (float)maa;
    }
    a.ema[i] = make_double2(ma, maa);
}

struct SpecRedShared { float mx[kSpecThreads]; float mn[kSpecThreads]; };
// one CTA per frame: ceil = max float(maa) (start 0), floor = min float(maa) (start 1)       (:437, 499-504)
SDR_HD inline void spec_minmax_phase(int ph, const SpecArgs &a, int f, int tid, int nt, SpecRedShared *sm) {
    if (ph == 0) {
        float mx = 0.f, mn = 1.f;
        const float *m = a.maa_f + (size_t)f * a.N;
        for (int i = tid; i < a.N; i += nt) { float v = m[i]; mx = fmaxf(mx, v); mn = fminf(mn, v); }
        sm->mx[tid] = mx; sm->mn[tid] = mn;
    } else if (tid == 0) {
        float mx = 0.f, mn = 1.f;
        for (int t = 0; t < nt; ++t) { mx = fmaxf(mx, sm->mx[t]); mn = fminf(mn, sm->mn[t]); }
        a.frame_cf[f] = make_float2(mx, mn);
    }
}

// single thread: scalar EMAs in frame order (:513-521)
SDR_HD inline void spec_scalar_thread(const SpecArgs &a) {
    double cma = a.scalars[0], cmaa = a.scalars[1], fma_ = a.scalars[2], fmaa = a.scalars[3];
    for (int f = 0; f < a.F; ++f) {
        const double c = (double)a.frame_cf[f].x, fl = (double)a.frame_cf[f].y;
        cma = cma + (c - cma) * 0.05;
        cmaa = cmaa + (cma - cmaa) * 0.05;
        fma_ = fma_ + (fl - fma_) * 0.05;
        fmaa = fmaa + (fma_ - fmaa) * 0.05;
        a.frame_cf_maa[f] = make_double2(cmaa, fmaa);
    }
    a.scalars[0] = cma; a.scalars[1] = cmaa; a.scalars[2] = fma_; a.scalars[3] = fmaa;
}

// one thread per (frame, point)
SDR_HD inline void spec_points_thread(const SpecArgs &a, long long gid) {
    const long long total = (long long)a.F * a.fft_size;
    if (gid >= total) return;
    const int f = (int)(gid / a.fft_size), xp = (int)(gid - (long long)f * a.fft_size);
    const float *m = a.maa_f + (size_t)f * a.N;
    const double point_ceil = a.frame_cf_maa[f].x, point_floor = a.frame_cf_maa[f].y;
    // bins of this point (:538-556): out-of-range bins -- and bin 0 -- count as the floor (`idx > 0 && idx < fftSizeInternal`)
    double acc = 0.0;
    const int i0 = a.pt_i0[xp], cnt = a.pt_cnt[xp];
    for (int j = 0; j < cnt; ++j) {
        const unsigned idx = (unsigned)round(a.visual_start + (double)(i0 + j));
        acc += (idx > 0 && idx < (unsigned)a.N) ? (double)m[idx] : point_floor;
    }
    const double num = log10((acc / (double)cnt) + 0.25 - (point_floor - 0.75));
    const double den = log10((point_ceil + 0.25) - (point_floor - 0.75));
    const float y = (float)((num / den) * (double)a.scale_factor);
    a.points[gid] = y;
    const float wv = y < 0 ? 0.f : (y > 0.99 ? (float)0.99 : y);
    a.lines[gid] = (unsigned char)floor((double)wv * 255.0);
}

}  // namespace sdrb200
