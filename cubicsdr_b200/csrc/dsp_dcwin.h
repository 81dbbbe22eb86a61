// cubicsdr_b200/csrc/dsp_dcwin.h -- bit-exact DC blocker without a warm-up: candidate windows + table composition.
//
// Reference: iirfilt_crcf_create_dc_blocker(0.0005f) + iirfilt_crcf_execute_block, src/sdr/SDRPostThread.cpp:29,284.
// The float recurrence  t = fl(c v), v' = fl(x + t), y = fl(v' - v)  (dsp_core.h 1b) is serial, but
//   (1) the two components are independent chains;
//   (2) the exact state e_b at every tile boundary b (1024 samples) is known from the double-precision scan, and the true
//       float state stays within a few dozen float steps of it (its own rounding noise: sigma ~ 13 ulps of the running
//       magnitude, measured; profiles/micro/dc_window_sim.c);
//   (3) a tile is a deterministic map  start state -> end state.
// So instead of warming every segment up for 24576 steps until it has forgotten its start (k_dc_lanes*), the map of every tile is
// TABULATED for the kDwCand consecutive floats around e_b (pass 1: kDwCand independent chains per tile and component, all
// tiles in parallel -- throughput work, no serial latency beyond one tile), the tables are composed chunk by chunk (pass 2),
// one thread per component walks the chunk maps from the true carried state (pass 3), and every tile's true start state
// falls out by a gather.  The replay kernel then runs every tile from its true state (1024 steps) and stores y.
// Serial depth: 1024 steps (tables) + ~2 x 32 table look-ups + 1024 steps (replay) instead of 24576 + segment.
//
// A boundary where |e_b| sits more than one binade under the recent magnitude of the state is not enumerable (the floats
// near zero are much denser than the noise of the trajectory, which is set by the ulp at the LARGE values): such a
// boundary is "bad", no window is opened there and the candidate chains of the last good boundary simply run on through
// it, leaving their states at the bad boundary in the table as well.  If the true state ever leaves a window (never seen
// in 40 M-sample simulations with a 256-wide window; possible in principle) the walker falls back to running that stretch
// of the chain serially -- the result is bit-exact in every case, and k_dc_verify still proves every junction.
#pragma once
#include "dsp_core.h"

namespace sdrb200 {

constexpr int kDwCand = 256;              // candidates per good boundary and component (consecutive floats)
constexpr int kDwHalf = kDwCand / 2;
constexpr int kDwChunk = 32;              // boundaries per composition chunk
constexpr int kDwMinTiles = 3;            // shorter blocks take the serial replay

struct DcWinArgs {
    const float2 *xraw;            // raw block, xraw[0] = first sample
    int n, tiles;                  // tiles = ceil(n / kDcTile); boundary b = start of tile b, b = 0 .. tiles - 1
    float c;                       // feedback coefficient (float)
    const double2 *tile_carry;     // exact state entering tile b
    const float2 *fstate;          // the reference's float state entering the block
    float *st;                     // [tiles][2][kDwCand] state at boundary p of candidate i of the run that covers tile p - 1
    float *pst;                    // [tiles][2][kDwCand] the same, re-indexed by the candidate index at the chunk's entry
    short *gmap;                   // [chunks][2][kDwCand] candidate index after the chunk, -1 = left a window on the way
    int *entry;                    // [chunks][2] true candidate index entering each chunk (-2: chunk written by the slow walker)
    float2 *seg_in;                // out [tiles]: true state entering tile b
    unsigned *stats;               // [0] stretches replayed serially by the walker (diagnostic)
};

// ---- float <-> position on the ordered float line (-0 and +0 share position 0) ------------------------------------------
SDR_HD SDR_FORCEINLINE int dw_ord(float f) {
    uint32_t u;
#ifdef __CUDA_ARCH__
    u = __float_as_uint(f);
#else
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? -(int)(u & 0x7fffffffu) : (int)u;
}
SDR_HD SDR_FORCEINLINE float dw_unord(int o) {
    const uint32_t u = o >= 0 ? (uint32_t)o : (0x80000000u | (uint32_t)(-o));
    float f;
#ifdef __CUDA_ARCH__
    f = __uint_as_float(u);
#else
    memcpy(&f, &u, 4);
#endif
    return f;
}
SDR_HD SDR_FORCEINLINE int dw_expo(float f) { return (int)(((uint32_t)dw_ord(fabsf(f)) >> 23) & 0xffu); }

SDR_HD SDR_FORCEINLINE float dw_comp(const float2 &v, int k) { return k ? v.y : v.x; }
SDR_HD SDR_FORCEINLINE float dw_compd(const double2 &v, int k) { return (float)(k ? v.y : v.x); }

// centre of the window at boundary b: the exact state rounded to float (the block's first boundary: the true state itself)
SDR_HD SDR_FORCEINLINE float dw_center(const DcWinArgs &a, int b, int k) {
    return b == 0 ? dw_comp(a.fstate[0], k) : dw_compd(a.tile_carry[b], k);
}
// is boundary b enumerable for component k?  Its exact state must lie within one binade of the largest exact state at the
// boundaries around it (four back, one ahead): then kDwHalf float steps at b are >= 64 ulps of that magnitude ~ 5 sigma.
SDR_HD inline bool dw_good(const DcWinArgs &a, int b, int k) {
    if (b == 0) return true;
    int lo = b - 4, hi = b + 1;
    if (lo < 0) lo = 0;
    if (hi > a.tiles - 1) hi = a.tiles - 1;
    int em = 0;
    for (int q = lo; q <= hi; ++q) { const int e = dw_expo(dw_compd(a.tile_carry[q], k)); if (e > em) em = e; }
    return dw_expo(dw_compd(a.tile_carry[b], k)) >= em - 1;
}
SDR_HD SDR_FORCEINLINE float dw_step(float c, float x, float v) {
#if defined(__CUDA_ARCH__) && !defined(SDR_EMU)
    return __fadd_rn(x, __fmul_rn(c, v));
#else
    volatile float t = c * v;
    volatile float nv = x + t;
    return nv;
#endif
}
// the chain over tile p from state v (serial form: the walker's fallback and the emulator)
SDR_HD inline float dw_run_tile(const DcWinArgs &a, int p, int k, float v) {
    const int i0 = p * kDcTile;
    int i1 = i0 + kDcTile; if (i1 > a.n) i1 = a.n;
    for (int i = i0; i < i1; ++i) v = dw_step(a.c, dw_comp(a.xraw[i], k), v);
    return v;
}
SDR_HD SDR_FORCEINLINE float *dw_row(float *t, int p, int k) { return t + ((size_t)p * 2 + k) * kDwCand; }

// ---- pass 1, serial form of one candidate (emulator; the GPU kernel stages the tile in shared memory) --------------------
SDR_HD inline void dw_window_thread(const DcWinArgs &a, int b, int k, int i) {
    if (!dw_good(a, b, k)) return;
    float v = dw_unord(dw_ord(dw_center(a, b, k)) + i - kDwHalf);
    for (int p = b; p < a.tiles;) {
        v = dw_run_tile(a, p, k, v);
        ++p;
        if (p >= a.tiles) break;
        dw_row(a.st, p, k)[i] = v;
        if (dw_good(a, p, k)) break;
    }
}

// ---- pass 2: chunk c = boundaries 32 c + 1 .. 32 c + 32 (those below `tiles`); thread i = candidate i of the run current at
//      boundary 32 c.  rows = the chunk's st rows (shared memory on the GPU), good / centre per boundary precomputed ------------
struct DwChunkInfo { int n; unsigned char good[kDwChunk]; int ord_center[kDwChunk]; };
SDR_HD inline void dw_chunk_info(const DcWinArgs &a, int c, int k, int j, DwChunkInfo *ci) {     // thread j < kDwChunk
    const int p = c * kDwChunk + 1 + j;
    int n = a.tiles - 1 - c * kDwChunk; if (n > kDwChunk) n = kDwChunk; if (n < 0) n = 0;
    if (j == 0) ci->n = n;
    if (j < n) { ci->good[j] = dw_good(a, p, k) ? 1 : 0; ci->ord_center[j] = dw_ord(dw_center(a, p, k)); }
}
SDR_HD inline void dw_compose_thread(const DcWinArgs &a, int c, int k, int i, const float *rows, const DwChunkInfo *ci) {
    int idx = i;
    for (int j = 0; j < ci->n; ++j) {
        const int p = c * kDwChunk + 1 + j;
        float s = 0.f;
        if (idx >= 0) {
            s = rows[j * kDwCand + idx];
            if (ci->good[j]) {
                const int q = dw_ord(s) - ci->ord_center[j] + kDwHalf;
                idx = (q >= 0 && q < kDwCand) ? q : -1;
            }
        }
        dw_row(a.pst, p, k)[i] = s;           // valid up to and including the boundary where the path left a window
    }
    a.gmap[((size_t)c * 2 + k) * kDwCand + i] = (short)idx;
}

// ---- pass 3a: one thread per component walks the chunks from the true state; stretches the tables cannot answer are run
//      serially (dw_run_tile).  Chunks it had to open are written here and flagged -2 for the gather. -------------------------
SDR_HD inline void dw_walk(const DcWinArgs &a, int k, const short *gmap /* may be a shared-memory copy */) {
    float *segk = reinterpret_cast<float *>(a.seg_in) + k;           // component k of seg_in[p] at segk[2 p]
    const int chunks = (a.tiles - 1 + kDwChunk - 1) / kDwChunk;
    int idx = kDwHalf;                                               // boundary 0: the window is centred on the true state
    bool indexed = true;
    float v = 0.f;
    segk[0] = dw_comp(a.fstate[0], k);
    for (int c = 0; c < chunks; ++c) {
        if (indexed) {
            const int g = gmap[((size_t)c * 2 + k) * kDwCand + idx];
            if (g >= 0) { a.entry[c * 2 + k] = idx; idx = g; continue; }
        }
        a.entry[c * 2 + k] = -2;
        const int p0 = c * kDwChunk;
        int pend = p0 + kDwChunk; if (pend > a.tiles - 1) pend = a.tiles - 1;
        for (int p = p0 + 1; p <= pend; ++p) {
            const float s = indexed ? dw_row(a.st, p, k)[idx] : dw_run_tile(a, p - 1, k, v);
            segk[2 * p] = s;
            v = s;
            if (dw_good(a, p, k)) {
                const int q = dw_ord(s) - dw_ord(dw_center(a, p, k)) + kDwHalf;
                if (q >= 0 && q < kDwCand) { idx = q; indexed = true; }
                else {
                    if (indexed) {
#if defined(__CUDA_ARCH__) && !defined(SDR_EMU)
                        atomicAdd(a.stats, 1u);                      // (the two components walk concurrently)
#else
                        a.stats[0] += 1u;
#endif
                    }
                    indexed = false;
                }
            }
        }
    }
}
// ---- pass 3b: gather.  thread t handles boundary p = t + 1 ----------------------------------------------------------------
SDR_HD inline void dw_gather_thread(const DcWinArgs &a, int p) {
    if (p < 1 || p >= a.tiles) return;
    const int c = (p - 1) / kDwChunk;
    float *seg = reinterpret_cast<float *>(a.seg_in);
    for (int k = 0; k < 2; ++k) {
        const int e = a.entry[c * 2 + k];
        if (e >= 0) seg[2 * p + k] = dw_row(a.pst, p, k)[e];
    }
}

}  // namespace sdrb200
