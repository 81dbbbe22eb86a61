// cubicsdr_b200/csrc/dsp_core.h -- every kernel body of the engine, written as per-thread "phases".
//
// A kernel is a list of phases; on the GPU all threads of a CTA run phase p and then __syncthreads(); the test-only
// emulator (SDR_EMU, see backend.h) runs the same phases serially.  The maths follows what the reference obtains
// from liquid-dsp 1.5.0 on CubicSDR's hot path; citations name the reference call sites.
//
// HBM layout (all little-endian, cf32 = interleaved {re, im} float pairs as liquid_float_complex):
//   wideband stream   float2 blk[kBlkHist + Nmax]      DC-blocked samples, history of the previous block in front
//   per channel       a float arena holding one [kHist | new] stream buffer per resampler level (ChanLayout)
//   per channel       a result slot {ChanResultHeader, audio[cap], iq[cap]} that is copied back to pinned host memory
#pragma once
#include <math.h>

#include "backend.h"

namespace sdrb200 {

constexpr int kHist = 64;         // history elements in front of every low-rate stream buffer (>= 51-tap AM FIR, 4m = 40)
constexpr int kBlkHist = 2048;    // wideband history kept in front of the current block (>= 12 * 2^kFrontMaxK + 2^kFrontMaxK)
constexpr int kFrontMaxK = 6;     // most half-band stages the front kernel fuses with the mixer
constexpr int kFrontWarm = 12;    // warm-up super-steps that flush a 13-tap (m = 3) half-band cascade
constexpr int kMaxTailHb = 16;    // msresamp2 never has more than 16 stages
constexpr int kMaxAudioHb = 8;
constexpr int kPfbTaps = 14;      // 2 * m of the arbitrary resampler (m = 7)
constexpr int kTailThreads = 256;

enum ModemKind { MODEM_FM = 0, MODEM_NBFM = 1, MODEM_AM = 2, MODEM_USB = 3, MODEM_LSB = 4, MODEM_IQ = 5,
                 MODEM_NONE = 6,    // no modem: the decimated IQ is the product (the zoomed spectrum view's hidden channel)
                 MODEM_DSB = 7, MODEM_CW = 8,
                 MODEM_FMS = 9 };   // FM stereo (analog/ModemFMStereo.cpp): two-channel audio {left, right}
constexpr int kFmsFirMax = 256;   // most taps of the stereo low-pass pair (firfilt_rrrf, estimate_req_filter_len(1000 / audio rate, 60))
constexpr int kFmsFirHist = 256;  // history kept in front of its two input streams

// ---- small complex helpers ------------------------------------------------------------------------------------------
SDR_HD SDR_FORCEINLINE float2 cf_make(float a, float b) { return make_float2(a, b); }
SDR_HD SDR_FORCEINLINE float2 cf_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SDR_HD SDR_FORCEINLINE float2 cf_fma(float h, float2 x, float2 acc) { return make_float2(fmaf(h, x.x, acc.x), fmaf(h, x.y, acc.y)); }
SDR_HD SDR_FORCEINLINE float2 cf_scale(float h, float2 x) { return make_float2(h * x.x, h * x.y); }
SDR_HD SDR_FORCEINLINE float flip_sign(float v, unsigned mask) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(__float_as_uint(v) ^ mask);
#else
    uint32_t u; memcpy(&u, &v, 4); u ^= mask; memcpy(&v, &u, 4); return v;
#endif
}
// x * (c + j s)
SDR_HD SDR_FORCEINLINE float2 cf_rot(float2 x, float c, float s) {
    return make_float2(fmaf(x.x, c, -(x.y * s)), fmaf(x.x, s, x.y * c));
}

// ---- packed complex arithmetic: one Blackwell FFMA2 / FMUL2 / FADD2 (PTX fma/mul/add .f32x2, sm_100+) handles the real and
//      imaginary parts together, halving the issue slots of the half-band cascade.  Bit-identical to two scalar ops.
#if defined(__CUDA_ARCH__) && !defined(SDR_EMU)
SDR_DEV SDR_FORCEINLINE float2 pk_fma(float2 h2, float2 x, float2 acc) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(h2.x), "f"(h2.y), "f"(x.x), "f"(x.y), "f"(acc.x), "f"(acc.y));
    return d;
}
SDR_DEV SDR_FORCEINLINE float2 pk_mul(float2 h2, float2 x) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(h2.x), "f"(h2.y), "f"(x.x), "f"(x.y));
    return d;
}
SDR_DEV SDR_FORCEINLINE float2 pk_add(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
#else
SDR_HD SDR_FORCEINLINE float2 pk_fma(float2 h2, float2 x, float2 acc) { return make_float2(fmaf(h2.x, x.x, acc.x), fmaf(h2.y, x.y, acc.y)); }
SDR_HD SDR_FORCEINLINE float2 pk_mul(float2 h2, float2 x) { return make_float2(h2.x * x.x, h2.y * x.y); }
SDR_HD SDR_FORCEINLINE float2 pk_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
#endif

// ---- table oscillator (nco_crcf: SURVEY.md A.1) -----------------------------------------------------------------------
// nco[idx] = {cos, sin} of table entry idx; idx = (theta + 2^21) >> 22
SDR_HD SDR_FORCEINLINE float2 nco_lookup(const float2 *nco, uint32_t theta) {
    return nco[(theta + 0x200000u) >> 22];
}

// =====================================================================================================================
// 1. DC blocker: iirfilt_crcf_create_dc_blocker(0.0005f) + execute_block            src/sdr/SDRPostThread.cpp:29,284
//    v[n] = x[n] + a v[n-1],  y[n] = v[n] - v[n-1],  a = 1 - alpha (as float).  Linear recurrence -> tiled scan in
//    double (the reference runs it in float direct-form II; its own rounding noise, ~4.6e-6 relative, is the parity
//    floor for this stage -- see DESIGN.md).
// =====================================================================================================================
constexpr int kDcTile = 1024;      // samples per scan tile (one CTA of 64 threads, 16 samples each)
constexpr int kDcThreads = 64;     // small CTAs (64 x 32 registers): they fit on an SM next to a resident channel-front CTA,
                                   // so the DC blocker of the next block really runs underneath the front kernel
constexpr int kDcPer = kDcTile / kDcThreads;

struct DcArgs {
    const float2 *x;      // raw block [n]
    float2 *y;            // DC-blocked block [n]
    int n;
    double a;             // feedback coefficient as double((float)(1 - alpha))
    double2 *tile_sum;    // [tiles]  zero-state end value of each tile
    double2 *tile_carry;  // [tiles]  state entering each tile
    double2 *state;       // [1]      v[-1] carried across blocks
    const double *apow;   // [kDcTile + 1] a^k
    float2 *copy_to;      // non-null: the tile-sum pass also copies x there (the block's way into the raw stream buffer)
};

// phase 0 of kernel A: each thread folds 4 consecutive samples, writes partial to shared; phase 1: thread 0 combines.
struct DcShared { double2 part[kDcThreads]; };

// a thread's kDcPer consecutive samples (128 contiguous bytes): 16-byte loads when the block is 16-byte aligned
SDR_HD SDR_FORCEINLINE void dc_load(const float2 *x, int base, int n, float2 *s) {
    if (base + kDcPer <= n && ((reinterpret_cast<size_t>(x + base) & 15) == 0)) {
        const float4 *p = reinterpret_cast<const float4 *>(x + base);
#pragma unroll
        for (int k = 0; k < kDcPer / 2; ++k) {
            const float4 u = sdr_ldg(p + k);
            s[2 * k] = make_float2(u.x, u.y); s[2 * k + 1] = make_float2(u.z, u.w);
        }
    } else {
        for (int k = 0; k < kDcPer; ++k) s[k] = (base + k < n) ? x[base + k] : make_float2(0.f, 0.f);
    }
}

SDR_HD inline void dc_tile_sum_phase(int ph, const DcArgs &a, int bid, int tid, DcShared *sm) {
    const int per = kDcTile / kDcThreads;
    const int base = bid * kDcTile + tid * per;
    if (ph == 0) {
        double vr = 0.0, vi = 0.0;
        float2 s[kDcPer];
        dc_load(a.x, base, a.n, s);
        if (a.copy_to) {
            if (base + kDcPer <= a.n && ((reinterpret_cast<size_t>(a.copy_to + base) & 15) == 0)) {
                float4 *p = reinterpret_cast<float4 *>(a.copy_to + base);
#pragma unroll
                for (int k = 0; k < kDcPer / 2; ++k) p[k] = make_float4(s[2 * k].x, s[2 * k].y, s[2 * k + 1].x, s[2 * k + 1].y);
            } else {
                for (int k = 0; k < per; ++k) if (base + k < a.n) a.copy_to[base + k] = s[k];
            }
        }
        for (int k = 0; k < per; ++k)      // beyond the block end the recurrence is simply not advanced
            if (base + k < a.n) { vr = fma(a.a, vr, (double)s[k].x); vi = fma(a.a, vi, (double)s[k].y); }
        sm->part[tid] = make_double2(vr, vi);
    } else if (tid == 0) {
        // combine the thread partials in order (Horner): V <- V * a^(samples in chunk t) + part[t]
        int tile_len = a.n - bid * kDcTile;
        if (tile_len > kDcTile) tile_len = kDcTile;
        double pw[kDcPer + 1];
        pw[0] = 1.0;
        for (int k = 1; k <= per; ++k) pw[k] = pw[k - 1] * a.a;
        double vr = 0.0, vi = 0.0;
        for (int t = 0; t < kDcThreads; ++t) {
            const int start = t * per;
            if (start >= tile_len) break;
            const int len = (tile_len - start < per) ? tile_len - start : per;
            const double w = pw[len];
            vr = fma(vr, w, sm->part[t].x);
            vi = fma(vi, w, sm->part[t].y);
        }
        a.tile_sum[bid] = make_double2(vr, vi);
    }
}

// kernel B (one CTA): carry entering every tile.  Each thread composes the affine maps of its run of tiles, thread 0
// chains the 256 thread-level maps, then every thread replays its run from its exact entering state.
struct DcCarryShared { double ar[kDcThreads]; double2 b[kDcThreads]; double2 enter[kDcThreads]; };
SDR_HD inline void dc_carry_phase(int ph, const DcArgs &a, int tiles, int tid, DcCarryShared *sm) {
    const int per = (tiles + kDcThreads - 1) / kDcThreads;
    const int t0 = tid * per, t1 = (t0 + per < tiles) ? t0 + per : tiles;
    if (ph == 0) {
        double A = 1.0, br = 0.0, bi = 0.0;
        for (int t = t0; t < t1; ++t) {
            int tile_len = a.n - t * kDcTile;
            if (tile_len > kDcTile) tile_len = kDcTile;
            const double w = a.apow[tile_len];
            A *= w;
            br = fma(br, w, a.tile_sum[t].x);
            bi = fma(bi, w, a.tile_sum[t].y);
        }
        sm->ar[tid] = A; sm->b[tid] = make_double2(br, bi);
    } else if (ph == 1) {
        if (tid == 0) {
            double cr = a.state[0].x, ci = a.state[0].y;
            for (int t = 0; t < kDcThreads; ++t) {
                sm->enter[t] = make_double2(cr, ci);
                cr = fma(cr, sm->ar[t], sm->b[t].x);
                ci = fma(ci, sm->ar[t], sm->b[t].y);
            }
            a.state[0] = make_double2(cr, ci);
        }
    } else {
        double cr = sm->enter[tid].x, ci = sm->enter[tid].y;
        for (int t = t0; t < t1; ++t) {
            a.tile_carry[t] = make_double2(cr, ci);
            int tile_len = a.n - t * kDcTile;
            if (tile_len > kDcTile) tile_len = kDcTile;
            const double w = a.apow[tile_len];
            cr = fma(cr, w, a.tile_sum[t].x);
            ci = fma(ci, w, a.tile_sum[t].y);
        }
    }
}

// kernel C: phase 0 thread partial prefix (same as A), phase 1 serial prefix over thread partials by thread 0,
// phase 2 every thread replays its 4 samples from its exact entering state and writes y.
struct DcShared2 { double2 part[kDcThreads]; double2 enter[kDcThreads]; };
SDR_HD inline void dc_apply_phase(int ph, const DcArgs &a, int bid, int tid, DcShared2 *sm) {
    const int per = kDcTile / kDcThreads;
    const int base = bid * kDcTile + tid * per;
    if (ph == 0) {
        double vr = 0.0, vi = 0.0;
        float2 s[kDcPer];
        dc_load(a.x, base, a.n, s);
        for (int k = 0; k < per; ++k)
            if (base + k < a.n) { vr = fma(a.a, vr, (double)s[k].x); vi = fma(a.a, vi, (double)s[k].y); }
        sm->part[tid] = make_double2(vr, vi);
    } else if (ph == 1) {
        if (tid == 0) {
            double cr = a.tile_carry[bid].x, ci = a.tile_carry[bid].y;
            const double w = a.apow[per];
            for (int t = 0; t < kDcThreads; ++t) {
                sm->enter[t] = make_double2(cr, ci);
                cr = fma(cr, w, sm->part[t].x);
                ci = fma(ci, w, sm->part[t].y);
            }
        }
    } else {
        double vr = sm->enter[tid].x, vi = sm->enter[tid].y;
        float2 s[kDcPer], o[kDcPer];
        dc_load(a.x, base, a.n, s);
        for (int k = 0; k < per; ++k) {
            const double nr = fma(a.a, vr, (double)s[k].x), ni = fma(a.a, vi, (double)s[k].y);
            o[k] = make_float2((float)(nr - vr), (float)(ni - vi));
            vr = nr; vi = ni;
        }
        if (base + kDcPer <= a.n && ((reinterpret_cast<size_t>(a.y + base) & 15) == 0)) {
            float4 *p = reinterpret_cast<float4 *>(a.y + base);
#pragma unroll
            for (int k = 0; k < kDcPer / 2; ++k) p[k] = make_float4(o[2 * k].x, o[2 * k].y, o[2 * k + 1].x, o[2 * k + 1].y);
        } else {
            for (int k = 0; k < per; ++k) if (base + k < a.n) a.y[base + k] = o[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// 1b. Bit-exact replay of the reference's float32 recurrence.
//     The reference evaluates, per component, t = fl(c v), v' = fl(x + t), y = fl(v' - v) with c = 0.9995f (iirfilt
//     direct-form II, norm path; checked bit-for-bit against the reference binary, tests/test_oracle.py).  Its rounding
//     noise (~5e-6 .. 2e-5 of the block RMS) would otherwise be the parity floor of everything downstream.
//     The recurrence is strictly sequential, but it is a contraction on the float lattice: two trajectories that start a
//     few ulps apart differ by an integer number of ulps that shrinks by one with probability ~|d|/2000 per step and
//     never grows, so they coalesce (median ~4k steps, exponential tail with a ~2k-step time constant) and are identical from then on.
//     Hence: cut the block into segments of L samples, give each a warm-up of W samples started from the EXACT state
//     (the double-precision scan above) rounded to float, run all segments in parallel, then check every junction
//     (end state of segment s-1 == warmed-up state of segment s, bitwise).  Matching junctions prove, by induction
//     from the true carried state of segment 0, that every output is the reference's; a failing junction (rare) is
//     replayed serially by dc_exact_verify.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kDcWarm = 24576;      // warm-up samples (a multiple of kDcTile); P(candidate not merged with the true trajectory) ~ 1.5e-4.
                                    // (12288 was measured: 2.8 % wrong candidates, and they come in runs -- chains that pass a region
                                    // where merging is slow are all late -- so a parallel fix-up needs one round per run member.)
constexpr int kDcMinSpec = 32768;   // blocks shorter than this are replayed serially (the warm-up alone is 24576 steps)

struct DcExactArgs {
    const float2 *xraw;              // raw stream: xraw[0] = first sample of the block, history at negative indices
    float2 *y;                       // DC-blocked block
    int n, n_seg;
    int seg_len, warm;
    float c;                         // feedback coefficient as the reference holds it (float)
    float one;                       // 1.0f, opaque to the compiler (k_dc_lanes_tma's FMA form of x + t)
    const double2 *tile_carry;       // exact state entering tile t of the current block (kDcTile samples per tile)
    const double2 *prev_tile_carry;  // same for the previous block
    int prev_n;                      // length of the previous block, 0 if there was none
    int raw_hist;                    // samples of raw history readable in front of xraw[0] (GPU staging bound)
    long long raw_end;               // samples readable from xraw[0] on (block + padding)
    float2 *seg_in, *seg_out;        // [n_seg] warmed-up state entering / true-candidate state leaving each segment
    float2 *fstate;                  // [1] the reference's float state carried across blocks
    unsigned *fixups;                // [1] junctions that had to be replayed (diagnostic)
    const float2 *start;             // non-null: the true state entering every segment is known (dsp_dcwin.h), no warm-up
    unsigned *junction_bad;          // with `start`: set by a segment whose end state is not the next segment's start (k_dc_verify then runs)
    int states_only;                 // warm-up kernels: leave the warmed-up state in seg_in[s] and stop (the tiles are replayed by k_dc_exact)
    int no_publish;                  // with `start`: candidates, not proven states -- record seg_in / seg_out only
    const int *fix_mask;             // non-null: re-run only segments s with fix_mask[s] != 0, each from seg_out[s - 1]
};

// one step of the reference recurrence for both components; no fused multiply-add allowed
SDR_HD SDR_FORCEINLINE void dc_ref_step(float c, float2 x, float2 &v, float2 &y) {
#if defined(__CUDA_ARCH__) && !defined(SDR_EMU)
    const float tr = __fmul_rn(c, v.x), ti = __fmul_rn(c, v.y);
    const float nr = __fadd_rn(x.x, tr), ni = __fadd_rn(x.y, ti);
    y = make_float2(__fsub_rn(nr, v.x), __fsub_rn(ni, v.y));
#else
    volatile float tr = c * v.x, ti = c * v.y;            // volatile: keep the two roundings apart on any host compiler
    volatile float nr = x.x + tr, ni = x.y + ti;
    y = make_float2(nr - v.x, ni - v.y);
#endif
    v = make_float2(nr, ni);
}

// Where segment s starts its warm-up (index relative to the block start, possibly negative) and from which state.
// seg_len and warm are multiples of kDcTile, so d = s * seg_len - warm is a tile boundary of this block when it is >= 0;
// when it reaches back into the previous block, the warm-up starts at that block's first tile boundary at or after d
// (its tile grid is anchored at ITS start), or at this block's start if there is none.  Every lane of the GPU kernel
// therefore starts at an index in [d, d + kDcTile).
SDR_HD inline void dc_exact_plan(const DcExactArgs &a, int s, int &i_start, float2 &v0) {
    if (a.fix_mask) { i_start = s * a.seg_len; v0 = a.seg_out[s > 0 ? s - 1 : 0]; return; }
    if (a.start) { i_start = s * a.seg_len; v0 = a.start[s]; return; }       // (segment 0: the walker put the carried state there)
    if (s == 0) { i_start = 0; v0 = a.fstate[0]; return; }
    const long long d = (long long)s * a.seg_len - a.warm;
    if (d >= 0) {
        const int t = (int)(d / kDcTile);
        i_start = t * kDcTile;
        v0 = make_float2((float)a.tile_carry[t].x, (float)a.tile_carry[t].y);
    } else if (a.prev_n > 0) {
        long long pos = (long long)a.prev_n + d;             // position inside the previous block
        if (pos < 0) pos = 0;
        const int t = (int)((pos + kDcTile - 1) / kDcTile);
        if ((long long)t * kDcTile >= a.prev_n) {            // no boundary left in the previous block: start with this one
            i_start = 0;
            v0 = make_float2((float)a.tile_carry[0].x, (float)a.tile_carry[0].y);
        } else {
            i_start = t * kDcTile - a.prev_n;
            v0 = make_float2((float)a.prev_tile_carry[t].x, (float)a.prev_tile_carry[t].y);
        }
    } else {
        i_start = 0;                                         // first block of the stream: everything before it is zero
        v0 = make_float2(0.f, 0.f);
    }
}

// end of a segment's replay: record the states; with known start states also prove the junction on the spot and, for the
// last segment, publish the carried state (k_dc_verify then has nothing to do)
SDR_HD SDR_FORCEINLINE void dc_segment_done(const DcExactArgs &a, int s, float2 v_in, float2 v_out) {
    a.seg_in[s] = v_in; a.seg_out[s] = v_out;
    if (a.start && !a.no_publish) {
        if (s + 1 < a.n_seg) {
            const float2 nx = a.start[s + 1];
            if (nx.x != v_out.x || nx.y != v_out.y) a.junction_bad[0] = 1u;
        } else a.fstate[0] = v_out;
    }
}
// plain serial form of one segment (emulator; also the replay used by dc_exact_verify)
SDR_HD inline void dc_exact_segment_serial(const DcExactArgs &a, int s, bool use_v, float2 v_in) {
    int i; float2 v;
    dc_exact_plan(a, s, i, v);
    const int p0 = s * a.seg_len;
    int p1 = p0 + a.seg_len; if (p1 > a.n) p1 = a.n;
    float2 y;
    if (use_v) { i = p0; v = v_in; }
    for (; i < p0; ++i) dc_ref_step(a.c, a.xraw[i], v, y);
    a.seg_in[s] = v;
    for (; i < p1; ++i) { dc_ref_step(a.c, a.xraw[i], v, y); a.y[i] = y; }
    a.seg_out[s] = v;
}

// single thread: prove the junctions, replay what does not match, publish the carried state
SDR_HD inline void dc_exact_verify(const DcExactArgs &a) {
    float2 truth = a.start ? a.start[0] : a.fstate[0];      // (with known starts the last segment has already overwritten fstate)
    for (int s = 0; s < a.n_seg; ++s) {
        if (s > 0) {
            const float2 w = a.seg_in[s];
            if (w.x != truth.x || w.y != truth.y) {
                dc_exact_segment_serial(a, s, true, truth);
                a.fixups[0] += 1u;
            }
        }
        truth = a.seg_out[s];
    }
    a.fstate[0] = truth;
}

// =====================================================================================================================
// 2. Channel descriptors
// =====================================================================================================================
struct ChanStatic {
    int modem;
    int S;                        // half-band stages of the IQ resampler (msresamp_crcf)
    int K1;                       // how many of them (all m = 3) the front kernel fuses with the mixer
    int n_tail;                   // S - K1 stages left for the tail, executed in order j = 0..n_tail-1
    int tail_m[kMaxTailHb];       // semi-length of tail stage j
    int tail_h1[kMaxTailHb];      // offset (floats) of its 2m branch taps in the taps pool
    float zeta;                   // 1 / 2^S
    unsigned arb_step;            // fixed-point stride of the arbitrary stage
    int pfb_off;                  // offset of its [256][14] taps in the taps pool
    // audio resampler (msresamp_rrrf, ModemAnalog.cpp:30)
    int a_interp, a_S;
    int a_m[kMaxAudioHb];         // semi-lengths by cascade position g (g = 0 is the m = 10 stage)
    int a_h1[kMaxAudioHb];
    float a_zeta;
    unsigned a_step;
    int a_pfb_off;
    // modem filters (offsets into the taps pool)
    int am_taps_off;              // 51 taps
    int hq_off;                   // 10 Hilbert quadrature taps
    int sos_off;                  // 3 x {b0,b1,b2,a1,a2}
    int cw_hq_off;                // CW: quadrature taps of firhilbf_create(5, 60) (ModemCW.cpp:23)
    unsigned cw_dtheta;           // CW: phase step of the beat oscillator, 2 pi 650 / audio rate (ModemCW.cpp:176)
    int cw_w_off;                 // CW: complex stream after the beat mixer, at the audio rate ([kHist | new])
    // FM stereo (ModemFMStereo.cpp:89-167): pilot band-pass sections 5 x {b0,b1,b2,a1,a2}, the low-pass pair's taps, de-emphasis
    int fms_sos_off, fms_fir_off, fms_fir_n, fms_demph;
    float fms_db0, fms_db1, fms_da1;
    int fms_x2_off;               // complex [kHist | new]: the pilot-squared down-mix feeding firhilbf c2r
    int fms_ds_off;               // real [kHist | new]: the difference signal at the modem rate (demodStereoData)
    int fms_al_off[kMaxAudioHb + 1];   // its audio-resampler levels (stereoResampler)
    int fms_am_off, fms_as_off;   // real: sum / difference signal at the audio rate (resampledOutputData / resampledStereoData)
    int fms_ld_off, fms_rd_off;   // real [kFmsFirHist | new]: de-emphasised left / right feeding the low-pass pair
    // stream buffers inside the channel arena (offsets in floats); each is [kHist | data]
    float *arena;                 // this channel's arena in HBM
    int lvl_off[kMaxTailHb + 1];  // complex level buffers: level 0 = front output, level j = output of tail stage j-1
    int lvl0_alt_off;             // second level-0 buffer: the front kernel of block b+1 writes one while the tail of block b
                                  // still reads the other (ChanDyn::lvl0_sel says which one a block uses)
    int iq_off;                   // complex: arbitrary-stage output (decimated IQ handed to the modem)
    int w_off;                    // complex: SSB post-filter stream feeding the Hilbert stage
    int d_off;                    // real: demodulated samples (after auto-gain) feeding the audio resampler
    int al_off[kMaxAudioHb + 1];  // real: audio resampler level buffers (decim: after each half-band; interp: arb out, then each half-band)
    int want_iq;                  // copy decimated IQ into the result slot
    int bandwidth;                // DemodulatorThreadPostIQData::sampleRate (the bandwidth asked for): sampleTime of a block = n_iq / bandwidth
    int res_audio_cap, res_iq_cap;
    long long res_off;            // byte offset of the result slot in the results buffer
};

struct ChanDyn {                  // written by the host for every block
    long long g0;                 // wideband samples this channel consumed before this block
    unsigned theta0;              // oscillator phase at the first sample of the block
    unsigned dtheta, dtheta_prev; // phase step for this block / for the history in front of it
    int dir, dir_prev;            // +1 mix up, -1 mix down, 0 no mixing   (DemodulatorPreThread.cpp:186-195)
    int active;                   // 0: channel skips this block
    int valid;                    // 0: free slot, nothing may be touched
    int inject_n;                 // > 0: the decimated IQ of this call was uploaded by the host (Modem::demodulate seam):
                                  //      skip the IQ resampler and demodulate inject_n samples
    int lvl0_sel;                 // which level-0 buffer this block uses (0: lvl_off[0], 1: lvl0_alt_off)
    int squelch_enabled;          // DemodulatorThread::squelchEnabled for this block
    float squelch_level;          // DemodulatorThread::squelchLevel (dB)
    int view;                     // 1: hidden channel of a zoomed spectrum view -- processed by that spectrum's own tail launch
                                  //    (its level-0 stream is written by the view mixer, its block length is the frames' total)
    const float2 *hist;           // non-null: history for samples i < 0 is hist[i] (the channel skipped the previous block(s),
                                  // so its delay lines must see the tail of the last block it DID consume)
    const float2 *xin;            // non-null: this channel's own input stream (channelizer mode: the samples of its filterbank
                                  // channel, xin[0] = first sample of the block, history at negative indices) instead of FrontArgs::x
};

struct ChanState {                // device-resident, advanced by the tail kernel
    unsigned long long arb_phase;     // `phase` of the IQ arbitrary resampler between blocks
    unsigned long long iq_count;      // decimated IQ samples produced so far (= demod samples)
    unsigned long long a_phase;       // audio arbitrary resampler phase
    unsigned long long a_lvl_count;   // real samples that have entered the audio arbitrary stage (decim) -- informational
    unsigned ssb_theta;               // SSB quarter-rate oscillator phase
    float sos_v[3][2][2];             // [section][v1,v2][re,im]
    float agc_ceil, agc_ma, agc_maa;  // ModemAnalog.cpp:6,70-72
    float sig_level, sig_floor, sig_ceil;   // DemodulatorThread::signalLevel / signalFloor / signalCeil (std::atomic<float>)
    unsigned dsb_theta, dsb_dtheta;         // ampmodem's Costas-loop oscillator (nco_crcf theta / d_theta), ModemDSB
    unsigned fms_theta, fms_dtheta;         // FM stereo: the pilot's phase-locked oscillator (stereoPilot)
    float fms_sos_v[5][2][2];               // pilot band-pass [section][v1,v2][re,im]
    float fms_dl, fms_dr;                   // de-emphasis delay elements (iirDemphL / iirDemphR)
};

struct ChanResultHeader {
    unsigned n_audio, n_iq;
    float peak;
    int skipped;
    double level_db;
    float sig_level, sig_floor, sig_ceil;   // after this block (DemodulatorThread.cpp:165-196)
    int squelched;                          // AudioThreadInput::is_squelch_active (:198, :235-238)
};
constexpr int kResHdrBytes = 64;            // the header's room in front of a result slot's audio samples

// =====================================================================================================================
// 3. Front kernel: table-oscillator mix + K half-band (m = 3) decimators, one thread per (channel, time span)
//    nco_crcf_mix_block_up/down  DemodulatorPreThread.cpp:186-195 ; first K stages of msresamp_crcf_execute :209
//    Each thread is a streaming decimator whose delay lines live in registers (transposed-form FIR: no shifts).
// =====================================================================================================================
struct FrontArgs {
    const float2 *x;              // x[0] = first sample of the block, x[-kBlkHist..-1] = history
    int n;
    const float2 *nco;            // [1024] {cos, sin}
    const ChanStatic *cs;
    const ChanDyn *cd;
    const int *group;             // channel ids handled by this launch (all with the same K1)
    int cg;                       // number of channels in the group
    int span;                     // super-steps per time span
    int n_span;                   // spans per block
    float c0, c1, c2;             // m = 3 branch taps (symmetric: c[d] = c[5-d])
    unsigned long long *dbg;      // optional [4 * CTAs]: {smid, t_start, t_after_table, t_end} in ns (globaltimer), else null
};

struct Hb3State { float2 s0, s1, s2, s3, s4, s5; };

SDR_HD SDR_FORCEINLINE void hb3_reset(Hb3State &h) {
    h.s0 = h.s1 = h.s2 = h.s3 = h.s4 = h.s5 = make_float2(0.f, 0.f);
}
// first sample of a pair (filtered branch): returns the decimated output that becomes available with it
// (c0, c1, c2 are the taps duplicated into both halves of a packed pair)
SDR_HD SDR_FORCEINLINE float2 hb3_push_even(Hb3State &h, float2 x, float2 c0, float2 c1, float2 c2) {
    h.s0 = pk_fma(c0, x, h.s1);
    h.s1 = pk_fma(c1, x, h.s2);
    h.s2 = pk_fma(c2, x, h.s3);
    h.s3 = pk_fma(c2, x, h.s4);
    h.s4 = pk_fma(c1, x, h.s5);
    h.s5 = pk_mul(c0, x);
    return h.s0;
}
// second sample of a pair (pure delay branch, emerges m = 3 outputs later)
SDR_HD SDR_FORCEINLINE void hb3_push_odd(Hb3State &h, float2 x) { h.s3 = pk_add(h.s3, x); }

// Feed one mixed sample into the cascade.  The sample with in-super-step index t enters stage l only if bits 0..l-1 of
// t are 0 (it was the first sample of a pair at every earlier stage); bit l then says which branch of stage l it lands
// on.  Returns the stage-(K-1) output, which exists only for t == 0.  With t a compile-time constant (unrolled caller)
// everything folds to straight-line code.
template <int K>
SDR_HD SDR_FORCEINLINE float2 front_feed(Hb3State *hb, int t, float2 u, float2 c0, float2 c1, float2 c2) {
    float2 y = u;
    bool alive = true;
#pragma unroll
    for (int l = 0; l < K; ++l) {
        if (alive) {
            if ((t >> l) & 1) { hb3_push_odd(hb[l], y); alive = false; }
            else y = hb3_push_even(hb[l], y, c0, c1, c2);
        }
    }
    return y;
}
// Same with a run-time t (slow path: few super-steps per span, keeps the code small).  hb[] is indexed with compile-time
// constants only, so it stays in registers.
template <int K>
SDR_HD inline float2 front_feed_dyn(Hb3State *hb, int t, float2 u, float2 c0, float2 c1, float2 c2) {
    float2 y = u;
    bool alive = true;
#pragma unroll
    for (int l = 0; l < K; ++l) {
        if (alive) {
            if ((t >> l) & 1) { hb3_push_odd(hb[l], y); alive = false; }
            else y = hb3_push_even(hb[l], y, c0, c1, c2);
        }
    }
    return y;
}

// Where the samples of a super-step that lies entirely inside the current block come from.
// XDirect: straight from global memory (one address per warp -> a broadcast load); used by the emulator, by warps whose
// channels are not in lock step, and for K < 4.
struct XDirect {
    const float2 *p;
    SDR_HD SDR_FORCEINLINE void prefetch(const float2 *) {}
    SDR_HD SDR_FORCEINLINE void commit(const float2 *src) { p = src; }
    SDR_HD SDR_FORCEINLINE float2 get(int t) const { return sdr_ldg(p + t); }
};
#ifndef SDR_EMU
// XWarpStage: the warp fetches the next super-step (M samples, M/2 lanes x one 16-byte pair each) while it computes the
// current one, parks it in its private double-buffered shared-memory slot, and every lane then reads the samples back
// as broadcast 16-byte loads.  No per-sample global load, no CTA barrier.
template <int M>
struct XWarpStage {
    float4 *buf;          // [2][32] float4, private to this warp
    int lane, cur;
    float2 pa, pb;
    __device__ __forceinline__ void prefetch(const float2 *src) {
        if (lane < M / 2) { pa = __ldg(src + 2 * lane); pb = __ldg(src + 2 * lane + 1); }
    }
    __device__ __forceinline__ void commit(const float2 *) {
        cur ^= 1;
        if (lane < M / 2) buf[cur * 32 + lane] = make_float4(pa.x, pa.y, pb.x, pb.y);
        __syncwarp();
    }
    __device__ __forceinline__ float2 get(int t) const {
        const float4 v = buf[cur * 32 + (t >> 1)];
        return (t & 1) ? make_float2(v.z, v.w) : make_float2(v.x, v.y);
    }
};
#endif

// Oscillator table access.  TabPlain: the 8 KB {cos, sin} table as is.  TabRep16 (GPU): the table replicated 16 times
// in shared memory, entry i copy r at ((i * 16 + r) * 8) bytes with r = lane & 15, so the 16 lanes of a half-warp always
// hit 16 different bank pairs whatever their phases: 2 wavefronts per lookup instead of ~5.6 with random conflicts.
struct TabPlain {
    const float2 *t;
    SDR_HD SDR_FORCEINLINE float2 get(unsigned th_biased) const { return t[th_biased >> 22]; }
};
#ifndef SDR_EMU
struct TabRep16 {
    const char *base;        // shared memory
    unsigned lane_off;       // (lane & 15) * 8
    __device__ __forceinline__ float2 get(unsigned th_biased) const {
        return *reinterpret_cast<const float2 *>(base + (((th_biased >> 15) & 0x1FF80u) | lane_off));
    }
};
#endif

// One thread = one channel over one time span.  `store` is false for padding lanes that only keep a warp converged.
template <int K, class XS, class TAB>
SDR_HD inline void front_span(const FrontArgs &a, int ch, int span_idx, bool store, const TAB &tabf, XS &xs) {
    const ChanStatic &cs = a.cs[ch];
    const ChanDyn cd = a.cd[ch];                           // by value: keep the per-block parameters in registers
    constexpr int M = 1 << K;
    const long long g0 = cd.g0, g1 = cd.g0 + a.n;
    const long long qa = g0 >> K, qb = g1 >> K;            // outputs [qa, qb) complete inside this block
    long long q_lo = qa + (long long)span_idx * a.span;
    long long q_hi = q_lo + a.span;
    if (span_idx == a.n_span - 1 || q_hi > qb) q_hi = qb;
    if (q_lo >= q_hi) return;
    float2 *out = reinterpret_cast<float2 *>(cs.arena + (cd.lvl0_sel ? cs.lvl0_alt_off : cs.lvl_off[0])) + kHist;
    const float2 *xb = cd.xin ? cd.xin : a.x;             // the block this channel reads: the shared wideband block, or its own stream
    const float2 *hx = cd.hist ? cd.hist : xb;
    // table oscillator: entry = (theta + 2^21) >> 22; mixing down conjugates (sign on the sine); no mixing = constant
    // entry 0 = (1, 0).  The *_prev values describe the block whose tail forms the history in front of this one.
    // the conjugation is a sign flip of the sine: done as an integer XOR (bit-identical to multiplying by -1.0f, and it
    // runs on the integer pipe instead of the FP32 pipe the kernel is bound by)
    const unsigned sg = cd.dir < 0 ? 0x80000000u : 0u, sg_prev = cd.dir_prev < 0 ? 0x80000000u : 0u;
    const unsigned dth = cd.dir ? cd.dtheta : 0u, dth_prev = cd.dir_prev ? cd.dtheta_prev : 0u;
    const unsigned thb = (cd.dir ? cd.theta0 : 0u) + 0x200000u;          // biased phase at sample 0 of the block
    const unsigned thb_prev = (cd.dir_prev ? cd.theta0 : 0u) + 0x200000u;

    if (K == 0) {
        for (long long q = q_lo; q < q_hi; ++q) {
            const int i = (int)(q - g0);
            const float2 w = tabf.get(thb + (unsigned)i * dth);
            if (store) out[q - qa] = cf_rot(sdr_ldg(xb + i), w.x, flip_sign(w.y, sg));
        }
        return;
    }

    Hb3State hb[K > 0 ? K : 1];
#pragma unroll
    for (int l = 0; l < K; ++l) hb3_reset(hb[l]);
    const float2 c0 = make_float2(a.c0, a.c0), c1 = make_float2(a.c1, a.c1), c2 = make_float2(a.c2, a.c2);
    bool prefetched = false;

    for (long long q = q_lo - kFrontWarm; q < q_hi; ++q) {
        const long long gq = q * M;                         // channel sample index of the super-step's first input
        const int i0 = (int)(gq - g0);                      // same, relative to the block start
        // a super-step is "whole" when all its samples come from one place with one oscillator setting: the current
        // block (i0 >= 0) or the history in front of it (i0 + M <= 0), and the channel already existed (gq >= 0)
        const bool in_block = i0 >= 0, in_hist = (i0 + M <= 0) && (i0 >= -kBlkHist);
        if ((in_block || in_hist) && gq >= 0) {
            const float2 *src = (in_block ? xb : hx) + i0;
            if (!prefetched) xs.prefetch(src);
            xs.commit(src);
            // the next super-step can be prefetched only if it is whole and comes from the same buffer
            prefetched = (q + 1 < q_hi) && (in_block || (i0 + 2 * M <= 0));
            if (prefetched) xs.prefetch(src + M);
            const unsigned step = in_block ? dth : dth_prev;
            const unsigned s_sign = in_block ? sg : sg_prev;
            unsigned th = in_block ? thb + (unsigned)i0 * dth : thb_prev - (unsigned)(-i0) * dth_prev;
#pragma unroll
            for (int t = 0; t < M; ++t) {
                const float2 w = tabf.get(th);
                th += step;
                const float2 y = front_feed<K>(hb, t, cf_rot(xs.get(t), w.x, flip_sign(w.y, s_sign)), c0, c1, c2);
                if (t == 0 && store && q >= q_lo) out[q - qa] = y;
            }
        } else {
            // ---- rare: a super-step straddling the block start, or samples from before the channel existed (its delay
            // lines then held zeros)
            prefetched = false;
#pragma unroll 2
            for (int t = 0; t < M; ++t) {
                const int i = i0 + t;
                float2 u = make_float2(0.f, 0.f);
                if (gq + t >= 0 && i >= -kBlkHist) {
                    if (i >= 0) {
                        const float2 w = tabf.get(thb + (unsigned)i * dth);
                        u = cf_rot(sdr_ldg(xb + i), w.x, flip_sign(w.y, sg));
                    } else {
                        const float2 w = tabf.get(thb_prev - (unsigned)(-i) * dth_prev);
                        u = cf_rot(sdr_ldg(hx + i), w.x, flip_sign(w.y, sg_prev));
                    }
                }
                const float2 y = front_feed_dyn<K>(hb, t, u, c0, c1, c2);
                if (t == 0 && store && q >= q_lo) out[q - qa] = y;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Rare path: history that spans more than one oscillator setting.  The front kernel rebuilds its delay lines at a block
// start from up to kFrontWarm * 2^K history samples, re-mixed with ONE setting (that of the previous block).  When the blocks
// are shorter than that window and the channel was retuned within it, the host instead has the history mixed sample by
// sample with the settings that were actually in effect (this function, one thread per sample) into the channel's private
// history, and tells the front kernel that the history needs no mixing (dir_prev = 0: table entry 0 = (1, 0) is exact).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kPremixSegs = 8;
struct PremixSeg { int n; unsigned dtheta; int dir; };          // newest block first: length, phase step, +1 up / -1 down / 0 none
struct PremixArgs {
    const float2 *x;              // raw stream of the channel: x[0] = first sample of the new block, history at negative indices
    float2 *out;                  // out[-count .. -1] receive the mixed history
    const float2 *nco;            // [1024] {cos, sin}
    unsigned theta0;              // oscillator phase at x[0]
    int count, nseg;
    PremixSeg seg[kPremixSegs];
};
SDR_HD inline void premix_thread(const PremixArgs &a, int i) {      // i = 1 .. count: sample x[-i]
    if (i < 1 || i > a.count) return;
    unsigned theta = a.theta0;
    int rem = i, dir = 0;
    bool found = false;
    for (int k = 0; k < a.nseg && !found; ++k) {
        const PremixSeg sg = a.seg[k];
        if (rem <= sg.n) {                                           // x[-i] is the rem-th sample from the end of block k
            if (sg.dir) theta -= (unsigned)rem * sg.dtheta;          // mix_block steps once per sample; an unmixed block does not step
            dir = sg.dir; found = true;
        } else {
            if (sg.dir) theta -= (unsigned)sg.n * sg.dtheta;
            rem -= sg.n;
        }
    }
    const float2 x = a.x[-i];
    if (!found || dir == 0) { a.out[-i] = x; return; }              // older than every tracked block, or never mixed
    const float2 w = a.nco[(theta + 0x200000u) >> 22];
    a.out[-i] = cf_rot(x, w.x, flip_sign(w.y, dir < 0 ? 0x80000000u : 0u));
}

// job -> (span, channel chunk): jobs are dealt round-robin over the CTAs so that a small job count still lands on
// every SM (job = cta + n_cta * warp_in_cta)
SDR_HD SDR_FORCEINLINE void front_job(const FrontArgs &a, int job, int &span_idx, int &chunk) {
    const int n_chunks = (a.cg + 31) / 32;
    span_idx = job / n_chunks;
    chunk = job - span_idx * n_chunks;
}

// generic entry (emulator): every lane on its own, samples from global memory
template <int K>
SDR_HD inline void front_thread(const FrontArgs &a, int job, int lane) {
    int span_idx, chunk;
    front_job(a, job, span_idx, chunk);
    const int ci = chunk * 32 + lane;
    if (span_idx >= a.n_span || ci >= a.cg) return;
    XDirect xs;
    xs.p = a.x;
    TabPlain tp;
    tp.t = a.nco;
    front_span<K, XDirect, TabPlain>(a, a.group[ci], span_idx, true, tp, xs);
}

}  // namespace sdrb200
