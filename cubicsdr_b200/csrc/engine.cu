// cubicsdr_b200/csrc/engine.cu -- engine object, kernel launches and the C-ABI of include/sdr_b200.h.
//
// One engine per GPU.  A submit enqueues, without any host synchronisation, on five streams tied together by events:
//   dc    (highest priority)  history slides -> block copy-in -> DC blocker (double-precision scan + bit-exact float replay)
//                             or, in channelizer mode, the analysis filterbank -> upload of the block's parameter lists;
//                             runs one block ahead of the channel kernels (double-buffered block)
//   main                      channel front kernel(s): oscillator mix + first half-band stages
//   side  (high priority)     spectrum: FFT / averaging / waterfall line, forked after the front launch, joined before the next
//   tail                      per-channel low-rate chain (remaining half-bands, polyphase arm, demodulator, audio, metering)
//   res                       every device -> host result copy (ring of three pinned result buffers: sdr_engine_wait_block)
// Host-side logic that the reference keeps in its threads is restated here next to the launch it controls:
//   DemodulatorPreThread shift / skip rule      src/demod/DemodulatorPreThread.cpp:154-163
//   FFTDataDistributor framing / line limiter   src/process/FFTDataDistributor.cpp:28-144
//   SDRPostThread channelizer dispatch, channel centres, demodulator routing   src/sdr/SDRPostThread.cpp:115-139,178-187,303-455
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#ifndef SDR_EMU
#include <cuda.h>          // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, nothing links libcuda)
#endif
#include "../../include/sdr_b200.h"
#include "design.h"
#include "dsp_core.h"
#include "dsp_dcwin.h"
#include "dsp_spec.h"
#include "dsp_pfb.h"
#include "dsp_tail.h"
#include "dsp_sink.h"

using namespace sdrb200;

// ---------------------------------------------------------------------------------------------------------------------
// error text
// ---------------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }

#ifndef SDR_EMU
#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess) return fail(SDR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#endif

// ---------------------------------------------------------------------------------------------------------------------
// kernels (GPU) / serial replays (test-only emulator)
// ---------------------------------------------------------------------------------------------------------------------
// SMs the channel front kernel leaves free for the DC blocker of the next block (the front CTAs take a whole SM's registers, so
// nothing else runs next to them): 8.  The warm-up replay puts 1024 lanes there (1024-sample segments at 61.44 MSps), the scan's
// small CTAs and the candidate-window kernels start there and spread out when the front kernel ends.  Measured at 512 channels:
// 4 SMs give the front kernel 3 % more room but stretch the DC stage to 0.55 ms (its scan CTAs queue on four SMs), which then
// is the block period and costs the host entry 12 %; 8 SMs: DC 0.33 ms, hidden.  SDR_B200_DC_SMS overrides (experiments).
// With fewer channels more SMs are left free (below).
static inline int dc_reserved_sms(size_t live_channels) {
    static int forced = -1;
    if (forced < 0) { const char *v = getenv("SDR_B200_DC_SMS"); forced = v ? atoi(v) : 0; }
    if (forced > 0) return forced > 100 ? 100 : forced;
    // fewer channels: the front kernel has less work than SMs, and everything else of the block (DC tables and replay, tail,
    // spectrum) is queued behind its SM-exclusive CTAs -- give those more room (measured on 61.44 MSps blocks, profiles/gpu_exp5.sh:
    // 64 channels 3973 -> 4265 MS/s with 74 SMs, 128 channels 3106 -> 3416 with 32, 256 channels 2578 -> 2624 with 16)
    if (live_channels >= 384) return 8;
    if (live_channels >= 192) return 16;
    if (live_channels >= 96) return 32;
    return 74;
}
constexpr int kResBuf = 3;           // host result buffers (ring): blocks a pipelined consumer may keep in flight + 1
constexpr unsigned kDlWarpsHost = 4;   // warps per CTA of k_dc_lanes (one per scheduler)
constexpr int kFrontWarps = 16;      // warps per CTA of the channel front kernel (one CTA per SM): 16 x 32 x 128 registers = the whole
                                     // register file; the bit-exact DC replay has its own SMs (dc_split_for)
#ifndef SDR_EMU
__global__ void __launch_bounds__(kDcThreads) k_dc_tile_sum(DcArgs a) {
    __shared__ DcShared sm;
    dc_tile_sum_phase(0, a, blockIdx.x, threadIdx.x, &sm);
    __syncthreads();
    dc_tile_sum_phase(1, a, blockIdx.x, threadIdx.x, &sm);
}
__global__ void __launch_bounds__(kDcThreads) k_dc_carry(DcArgs a, int tiles) {
    __shared__ DcCarryShared sm;
    for (int ph = 0; ph < 3; ++ph) {
        dc_carry_phase(ph, a, tiles, threadIdx.x, &sm);
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kDcThreads) k_dc_apply(DcArgs a) {
    __shared__ DcShared2 sm;
    for (int ph = 0; ph < 3; ++ph) {
        dc_apply_phase(ph, a, blockIdx.x, threadIdx.x, &sm);
        __syncthreads();
    }
}
// Channel front: 16 warps per CTA (one CTA per SM), one warp = 32 channels over one time span.  Dynamic shared memory:
// the oscillator table replicated 16 times (128 KB, bank-conflict-free lookups for arbitrary phases) followed by each
// warp's private double-buffered slot for the wideband samples of the super-step it is working on (XWarpStage).
constexpr int kFrontSmemTab = 1024 * 16 * (int)sizeof(float2);                    // 131072
constexpr int kFrontSmemBytes = kFrontSmemTab + kFrontWarps * 2 * 32 * (int)sizeof(float4);   // + 16384
// Bit-exact DC blocker, serial form (short blocks, or no usable history for a warm-up): one warp, every lane runs the
// (strictly serial) chain redundantly; the warp prefetches the raw samples three batches ahead into shared memory and
// stores the outputs coalesced.
constexpr int kDcBatch = 64;     // samples per cp.async batch (two 8-byte copies per lane)
constexpr int kDcRing = 4;       // batches in flight: ~1700 cycles of lead over the serial chain
__global__ void __launch_bounds__(32) k_dc_exact(DcExactArgs a) {
    __shared__ __align__(16) float2 sx[kDcRing][kDcBatch];
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x;
    if (s >= a.n_seg) return;
    if (a.fix_mask && !a.fix_mask[s]) return;                        // fix-up round: only the tiles whose start state was wrong
    int i0;
    float2 v;
    dc_exact_plan(a, s, i0, v);
    const int p0 = s * a.seg_len;
    int p1 = p0 + a.seg_len;
    if (p1 > a.n) p1 = a.n;
    const int nb = (p1 - i0 + kDcBatch - 1) / kDcBatch;
    const float c = a.c;
    float2 *const y_out = a.y;
    // cp.async ring: the global -> shared copy of batch k + kDcRing - 1 is issued before batch k is consumed, so the
    // serial chain never waits on memory.  `src` / `fidx` walk ahead of the chain by kDcRing - 1 batches.
    const float2 *src = a.xraw + i0 + lane;
    int fidx = i0 + lane;
    const unsigned sx_base = (unsigned)__cvta_generic_to_shared(&sx[0][0]) + (unsigned)lane * 8u;
    int fk = 0;
    auto fetch = [&]() {
        const unsigned dst = sx_base + (unsigned)(fk & (kDcRing - 1)) * (kDcBatch * 8u);
        if (fidx < p1) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src));
        if (fidx + 32 < p1) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 256u), "l"(src + 32));
        asm volatile("cp.async.commit_group;");
        src += kDcBatch; fidx += kDcBatch; ++fk;
    };
#pragma unroll
    for (int k = 0; k < kDcRing - 1; ++k) fetch();
    float2 v_in = v, y;
    int base = i0;
    for (int k = 0; k < nb; ++k, base += kDcBatch) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kDcRing - 2) : "memory");   // batch k has landed (for this lane)
        __syncwarp();                                          // ... and for every lane; batch k-1 is fully consumed
        fetch();
        const float2 *b = sx[k & (kDcRing - 1)];
        if (base + kDcBatch <= p0) {
            // pure warm-up batch: nothing but the chain
#pragma unroll
            for (int j = 0; j < kDcBatch; ++j) dc_ref_step(c, b[j], v, y);
            if (base + kDcBatch == p0) v_in = v;
        } else if (base >= p0 && base + kDcBatch <= p1) {
            // pure output batch
            if (base == p0) v_in = v;
            // every lane runs the chain; lane l keeps outputs l and l + 32 in registers (selects, off the chain), so
            // the batch is stored coalesced without a shared-memory round trip between the steps
            float2 k0 = make_float2(0.f, 0.f), k1 = k0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                dc_ref_step(c, b[j], v, y);
                if (lane == j) k0 = y;
            }
#pragma unroll
            for (int j = 32; j < kDcBatch; ++j) {
                dc_ref_step(c, b[j], v, y);
                if (lane == j - 32) k1 = y;
            }
            y_out[base + lane] = k0;
            y_out[base + 32 + lane] = k1;
        } else {
            // boundary batch (segment start not aligned to the batch grid, or the end of the block)
            for (int j = 0; j < kDcBatch; ++j) {
                const int idx = base + j;
                if (idx < p1) {
                    if (idx == p0) v_in = v;
                    dc_ref_step(c, b[j], v, y);
                    if (idx >= p0 && lane == 0) y_out[idx] = y;
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (lane == 0) dc_segment_done(a, s, v_in, v);
}

// Bit-exact DC blocker, speculative form: ONE LANE PER SEGMENT.  The chain of a segment is strictly serial (FMUL -> FADD,
// ~9 cycles a step) but segments are independent once each has been warmed up from the exact state, so a warp advances 32
// segments in lock step -- lane l of warp w owns segment 32 w + l and walks the W + L samples [s L - W, (s + 1) L) of its
// own stream -- instead of 32 lanes repeating one chain.  A 61.44 MSps block is ~1000 segments of 1024 samples = 32 warps
// on 8 SMs (one warp per scheduler: nothing else competes for its issue slots), and the latency of the whole replay is
// W + L = 25600 steps instead of W + block / 64.
// Staging: the 32 streams of a warp are 8 L bytes apart, so every batch of 32 steps is moved through shared memory
// transposed: half a warp copies 32 consecutive samples of one stream with 16-byte cp.async (coalesced 256 B), five
// batches in flight; rows are 34 float2 apart, which keeps rows 16-byte aligned and makes the per-lane 128-bit reads down
// a column conflict-free (quarter-warp word offsets 68 l mod 32 = 4 l).  Outputs leave the same way in reverse.
constexpr int kDlBatch = 32;
constexpr int kDlRing = 5;
constexpr int kDlRow = 34;
constexpr int kDlWarps = 4;
constexpr int kDlSmemWarp = (kDlRing + 1) * 32 * kDlRow * (int)sizeof(float2);     // input ring + one output tile: 52224 B
constexpr int kDlSmem = kDlWarps * kDlSmemWarp;                                    // 208896 B: the CTA owns its SM
__global__ void __launch_bounds__(32 * kDlWarps, 1) k_dc_lanes(DcExactArgs a) {
    extern __shared__ __align__(16) unsigned char dl_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s0 = (blockIdx.x * kDlWarps + warp) * 32;              // first segment of this warp
    if (s0 >= a.n_seg) return;
    float2 *ring = reinterpret_cast<float2 *>(dl_sm + (size_t)warp * kDlSmemWarp);   // [kDlRing][32][kDlRow]
    float2 *so = ring + kDlRing * 32 * kDlRow;                                        // [32][kDlRow]
    const int L = a.seg_len, W = a.warm, n = a.n;
    const int s = s0 + lane;
    int i_begin = 0x7fffffff;                                        // lanes past the last segment never become active
    float2 v = make_float2(0.f, 0.f);
    if (s < a.n_seg) dc_exact_plan(a, s, i_begin, v);
    const int p0 = s * L;                                            // first output sample of this lane
    const int p1 = (s < a.n_seg) ? ((p0 + L < n) ? p0 + L : n) : p0;
    const int d = p0 - W;                                            // this lane's sample at step 0
    const int d0 = s0 * L - W;                                       // same for lane 0 (streams are L apart)
    const int nb = a.states_only ? W / kDlBatch : (W + L) / kDlBatch;
    const long long lo = -(long long)a.raw_hist, hi = a.raw_end;     // readable extent of the raw stream: [lo, hi)
    const float c = a.c;
    // transposed copies: lanes 0-15 move stream 2 it, lanes 16-31 stream 2 it + 1, two samples (16 bytes) each
    const int fj = lane >> 4, ft = (lane & 15) * 2;
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    int fk = 0;
    auto fetch = [&]() {
        if (fk < nb) {
            const unsigned dst0 = ring_s + (unsigned)(((fk % kDlRing) * 32 + fj) * kDlRow + ft) * 8u;
            const int i00 = d0 + fk * kDlBatch;                      // lane 0's first sample of this batch
            const float2 *src0 = a.xraw + i00 + fj * L + ft;
            // the engine pads the raw buffer so that the streams of lanes past the end of the block stay readable (their
            // samples are never used): normally every tile takes the unguarded path
            if (i00 >= lo && (long long)i00 + 31ll * L + kDlBatch <= hi) {
#pragma unroll
                for (int it = 0; it < 16; ++it)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (unsigned)(it * 2 * kDlRow * 8)),
                                 "l"(src0 + (size_t)it * 2 * L));
            } else {
                for (int it = 0; it < 16; ++it) {
                    const long long idx = (long long)i00 + (long long)(2 * it + fj) * L + ft;
                    if (idx >= lo && idx + 2 <= hi)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (unsigned)(it * 2 * kDlRow * 8)),
                                     "l"(src0 + (size_t)it * 2 * L));
                }
            }
        }
        asm volatile("cp.async.commit_group;");
        ++fk;
    };
#pragma unroll
    for (int k = 0; k < kDlRing - 1; ++k) fetch();
    float2 v_in = v, y;
    const int kw = W / kDlBatch;                                     // batches [0, kw) warm up, [kw, nb) produce output
    for (int k = 0; k < nb; ++k) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kDlRing - 2) : "memory");   // batch k has landed (this lane's copies)
        __syncwarp();                                                // ... and every lane's; batch k - 1 and its outputs are consumed
        fetch();
        const float4 *row = reinterpret_cast<const float4 *>(ring + ((k % kDlRing) * 32 + lane) * kDlRow);
        const int base = d + k * kDlBatch;
        // A lane is `on` (all 32 steps count), `off` (none does: it has not started, has finished, or owns no segment) or
        // partial.  Without partial lanes the warp runs the unrolled chain on a copy of the state, which off lanes discard.
        if (k < kw) {
            const bool on = base >= i_begin, off = (long long)base + kDlBatch <= (long long)i_begin;
            if (__all_sync(0xffffffffu, on || off)) {
                float2 vv = v;
#pragma unroll
                for (int t = 0; t < kDlBatch / 2; ++t) {
                    const float4 q = row[t];
                    dc_ref_step(c, make_float2(q.x, q.y), vv, y);
                    dc_ref_step(c, make_float2(q.z, q.w), vv, y);
                }
                if (on) v = vv;
            } else {                                                 // a lane starts inside this batch (previous block's tile grid)
                for (int t = 0; t < kDlBatch / 2; ++t) {
                    const float4 q = row[t];
                    if (base + 2 * t >= i_begin) dc_ref_step(c, make_float2(q.x, q.y), v, y);
                    if (base + 2 * t + 1 >= i_begin) dc_ref_step(c, make_float2(q.z, q.w), v, y);
                }
            }
            if (k == kw - 1) v_in = v;
        } else {
            float4 *orow = reinterpret_cast<float4 *>(so + lane * kDlRow);
            const bool on = base + kDlBatch <= p1, off = base >= p1;
            if (__all_sync(0xffffffffu, on || off)) {
                float2 vv = v;
#pragma unroll
                for (int t = 0; t < kDlBatch / 2; ++t) {
                    const float4 q = row[t];
                    float2 y0, y1;
                    dc_ref_step(c, make_float2(q.x, q.y), vv, y0);
                    dc_ref_step(c, make_float2(q.z, q.w), vv, y1);
                    orow[t] = make_float4(y0.x, y0.y, y1.x, y1.y);
                }
                if (on) v = vv;
            } else {                                                 // the block ends inside this batch
                for (int t = 0; t < kDlBatch / 2; ++t) {
                    const float4 q = row[t];
                    float2 y0 = make_float2(0.f, 0.f), y1 = y0;
                    if (base + 2 * t < p1) dc_ref_step(c, make_float2(q.x, q.y), v, y0);
                    if (base + 2 * t + 1 < p1) dc_ref_step(c, make_float2(q.z, q.w), v, y1);
                    orow[t] = make_float4(y0.x, y0.y, y1.x, y1.y);
                }
            }
            __syncwarp();
            // transposed store: half a warp writes 32 consecutive outputs of one stream (nothing at or past the block end)
            const int ob = (k - kw) * kDlBatch + ft;                  // offset of this lane's pair inside its stream's segment
#pragma unroll 4
            for (int it = 0; it < 16; ++it) {
                const int j = 2 * it + fj;
                const long long idx = (long long)(s0 + j) * L + ob;
                const float4 q = *reinterpret_cast<const float4 *>(so + j * kDlRow + ft);
                if (idx + 2 <= n) *reinterpret_cast<float4 *>(a.y + idx) = q;
                else if (idx < n) a.y[idx] = make_float2(q.x, q.y);
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (a.states_only) { if (s < a.n_seg) a.seg_in[s] = v_in; return; }
    if (s < a.n_seg) dc_segment_done(a, s, v_in, v);
}
// ---- TMA + mbarrier helpers (cp.async.bulk.tensor: SASS UTMALDG / UTMASTG; mbarrier: SYNCS) -----------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap *map, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ unsigned mbar_test(unsigned bar, unsigned parity) {       // non-blocking probe
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}

// The same replay fed by TMA.  The raw stream (from sample -W on) is described to the copy engine as a 2-D tensor of rows of
// L samples: the lanes of a warp walk 32 consecutive rows in lock step, so ONE cp.async.bulk.tensor instruction per 16 steps
// brings the next 16 samples of all 32 segments into a shared-memory box of 32 rows x 128 B; completion is counted on an
// mbarrier per ring slot (a slot = 64 steps = four boxes; three slots in flight).  The 128-byte swizzle of the tensor map
// (16-byte chunk index XOR row mod 8) is what makes the per-lane row reads conflict-free: lane l reads chunk c of its row at
// l * 128 + ((c ^ (l & 7)) << 4).  Results leave through a small padded tile as full 128-byte lines.  The serial chain is
// the only thing the warp really does:
// per 64 steps it spends a few dozen instructions on the ring (no divisions, no per-sample copies), the barrier of the next
// slot is probed while the current one is being consumed, and the first samples of the next slot are loaded before the
// current chain ends.
constexpr int kDtBox = 16;                                    // samples per box row (128 B)
constexpr int kDtBoxBytes = 32 * kDtBox * (int)sizeof(float2);   // 4096
constexpr int kDtBoxes = 4;                                   // boxes per ring slot
constexpr int kDtBatch = kDtBoxes * kDtBox;                   // 64 steps per slot
constexpr int kDtSlotBytes = kDtBoxes * kDtBoxBytes;          // 16384
constexpr int kDtRing = 3;
constexpr int kDtSmemWarp = kDtRing * kDtSlotBytes + 2 * kDtBoxBytes;   // input ring + an output tile of 32 steps: 57344 B
constexpr int kDtSmem = kDlWarps * kDtSmemWarp + 1024 + 256;  // + alignment slack + mbarriers = 230656 B
__device__ __forceinline__ bool elect_one() {                  // one lane of the (converged) warp; the compiler treats what it guards as uniform
    unsigned p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
    return p != 0;
}
// One step of the recurrence for the replay kernels.  FORM 0: FMUL, FADD.  FORM 1: FMUL, then x * one + t as a fused multiply-add
// with one = 1.0f read from the argument block (an exact identity: the same single rounding of x + t; it keeps both operations
// of the chain on the FMA pipe, and `one` must be opaque to the compiler or it folds the operation back into an FADD).
template <int FORM>
__device__ __forceinline__ void dt_step(float c, float one, float2 x, float2 &v, float2 &y) {
    const float tr = __fmul_rn(c, v.x), ti = __fmul_rn(c, v.y);
    float nr, ni;
    if (FORM == 1) {
        asm("fma.rn.f32 %0, %1, %2, %3;" : "=f"(nr) : "f"(x.x), "f"(one), "f"(tr));
        asm("fma.rn.f32 %0, %1, %2, %3;" : "=f"(ni) : "f"(x.y), "f"(one), "f"(ti));
    } else {
        nr = __fadd_rn(x.x, tr); ni = __fadd_rn(x.y, ti);
    }
    y = make_float2(__fsub_rn(nr, v.x), __fsub_rn(ni, v.y));
    v = make_float2(nr, ni);
}
// 64 steps (a whole ring slot) without output
template <int FORM>
__device__ __forceinline__ void dt_slot(const unsigned char *row, unsigned swz, float c, float one, float2 &v, float4 q0, float4 q1) {
#pragma unroll
    for (int t = 0; t < 32; ++t) {
        const unsigned o = (unsigned)(t >> 3) * kDtBoxBytes + (((unsigned)(t & 7) << 4) ^ swz);
        const float4 q = t == 0 ? q0 : (t == 1 ? q1 : *reinterpret_cast<const float4 *>(row + o));
        float2 y;
        dt_step<FORM>(c, one, make_float2(q.x, q.y), v, y);
        dt_step<FORM>(c, one, make_float2(q.z, q.w), v, y);
    }
}
template <int FORM>
__global__ void __launch_bounds__(32 * kDlWarps, 1)
k_dc_lanes_tma(DcExactArgs a, const __grid_constant__ CUtensorMap in_map) {
    extern __shared__ unsigned char dt_sm_raw[];
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform also in the compiler's eyes
    const int lane = threadIdx.x & 31;
    const int s0 = (blockIdx.x * kDlWarps + warp) * 32;              // first segment (tensor row) of this warp
    if (s0 >= a.n_seg) return;
    // swizzled boxes must sit on 1024-byte boundaries of the shared window
    const unsigned raw_s = (unsigned)__cvta_generic_to_shared(dt_sm_raw);
    const unsigned pad = (1024u - (raw_s & 1023u)) & 1023u;
    unsigned char *wsm = dt_sm_raw + pad + (size_t)warp * kDtSmemWarp;
    const unsigned ring_s = raw_s + pad + (unsigned)warp * kDtSmemWarp;                // [kDtRing][4 boxes][4096 B]
    const unsigned bar_s = raw_s + pad + kDlWarps * kDtSmemWarp + (unsigned)warp * 64u;   // [kDtRing] mbarriers
    const unsigned swz = (unsigned)(lane & 7) << 4;
    const unsigned char *my_in = wsm + lane * 128;
    unsigned char *out_p = wsm + kDtRing * kDtSlotBytes;              // output tile [32 rows][144 B]
    unsigned char *my_out_p = out_p + lane * 144;
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < kDtRing; ++r) mbar_init(bar_s + 8u * r, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const bool leader = elect_one();                                 // the one lane that issues every TMA operation and owns their groups
    const int L = a.seg_len, W = a.warm, n = a.n;
    const int s = s0 + lane;
    const bool live = s < a.n_seg;
    int i_begin = 0x7fffffff;                                        // lanes past the last segment never become active
    float2 v = make_float2(0.f, 0.f);
    if (live) dc_exact_plan(a, s, i_begin, v);
    const int p0 = s * L;
    const int p1 = live ? ((p0 + L < n) ? p0 + L : n) : p0;
    const int d = p0 - W;                                            // this lane's sample at step 0
    const int kw = W / kDtBatch;                                     // batches [0, kw) warm up, [kw, nb) produce output
    const int nb = a.states_only ? kw : (W + L) / kDtBatch;
    const float c = a.c, one = a.one;
    // A lane starts at i_begin.  Lanes whose warm-up begins in the previous block start at one of ITS tile boundaries, less than a
    // tile after step 0 and anywhere inside a batch: the first k1 batches test every step.  Every other start (tile boundaries
    // of this block, the block start) falls on a batch boundary: from batch k1 on a lane is simply on from batch k_on.
    const int k1 = kw < kDcTile / kDtBatch ? kw : kDcTile / kDtBatch;
    const int k_on = live ? (int)(((long long)i_begin - d + kDtBatch - 1) / kDtBatch) : 0x7fffffff;
    // only the block's last segment can end inside a box
    const bool warp_partial = __any_sync(0xffffffffu, live && ((p1 - p0) % kDtBox) != 0);
    // ---- producer side (one elected lane issues; the running coordinates are warp-uniform): step j of segment r is sample
    //      r L - W + j = element (row r + j / L, column j mod L) of the stream viewed as rows of L samples starting at sample -W;
    //      a box never straddles rows (L is a multiple of 1024)
    int fk = 0, f_slot = 0, f_col = 0, f_row = s0;
    auto fetch = [&]() {
        if (fk < nb && leader) {
            const unsigned bar = bar_s + 8u * (unsigned)f_slot, dst = ring_s + (unsigned)(f_slot * kDtSlotBytes);
            mbar_expect_tx(bar, (unsigned)kDtSlotBytes);
#pragma unroll
            for (int b = 0; b < kDtBoxes; ++b) tma_load_2d(dst + b * kDtBoxBytes, &in_map, 2 * (f_col + b * kDtBox), f_row, bar);   // columns are floats
        }
        ++fk;
        f_col += kDtBatch;
        if (f_col >= L) { f_col = 0; ++f_row; }
        f_slot = (f_slot + 1 == kDtRing) ? 0 : f_slot + 1;
    };
#pragma unroll 1
    for (int k = 0; k < kDtRing - 1; ++k) fetch();
    // ---- consumer side: per batch -- make sure it has landed, let the producer refill the slot consumed last, probe the next
    //      slot's barrier (the answer is needed only after this batch's chain), run the chain, load the next slot's first samples
    int c_slot = 0, n_slot = 0;
    unsigned c_par = 0, n_par = 0, ready = 0, ready_next = 0;
    float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
    bool have_q = false;
    const unsigned char *row = my_in;
    auto begin_batch = [&](int k) {
        if (!ready) mbar_wait(bar_s + 8u * (unsigned)c_slot, c_par);
        __syncwarp();                                                // every lane is done with batch k - 1: its slot may be refilled
        fetch();
        n_slot = (c_slot + 1 == kDtRing) ? 0 : c_slot + 1;
        n_par = n_slot == 0 ? (c_par ^ 1u) : c_par;
        ready_next = (k + 1 < nb) ? mbar_test(bar_s + 8u * (unsigned)n_slot, n_par) : 0u;
        row = my_in + c_slot * kDtSlotBytes;
        if (!have_q) {
            q0 = *reinterpret_cast<const float4 *>(row + swz);
            q1 = *reinterpret_cast<const float4 *>(row + (16u ^ swz));
        }
    };
    auto end_batch = [&]() {
        have_q = false;
        if (ready_next) {                                            // known to have landed: its first samples, ahead of the chain
            const unsigned char *nrow = my_in + n_slot * kDtSlotBytes;
            q0 = *reinterpret_cast<const float4 *>(nrow + swz);
            q1 = *reinterpret_cast<const float4 *>(nrow + (16u ^ swz));
            have_q = true;
        }
        ready = ready_next;
        c_slot = n_slot; c_par = n_par;
    };
    int k = 0;
    // ---- warm-up, first tile: starts may fall anywhere
#pragma unroll 1
    for (; k < k1; ++k) {
        begin_batch(k);
        const int base = d + k * kDtBatch;
        const bool on = base >= i_begin, off = (long long)base + kDtBatch <= (long long)i_begin;
        if (__all_sync(0xffffffffu, on || off)) {
            float2 vv = v;
            dt_slot<FORM>(row, swz, c, one, vv, q0, q1);
            if (on) v = vv;
        } else {
            float2 y;
            for (int t = 0; t < kDtBatch / 2; ++t) {
                const float4 q = *reinterpret_cast<const float4 *>(row + (t >> 3) * kDtBoxBytes + (((unsigned)(t & 7) << 4) ^ swz));
                if (base + 2 * t >= i_begin) dt_step<FORM>(c, one, make_float2(q.x, q.y), v, y);
                if (base + 2 * t + 1 >= i_begin) dt_step<FORM>(c, one, make_float2(q.z, q.w), v, y);
            }
        }
        end_batch();
    }
    // ---- warm-up, steady state: nothing but the ring and the chain; a lane that has not started discards its batch
#pragma unroll 1
    for (; k < kw; ++k) {
        begin_batch(k);
        float2 vv = v;
        dt_slot<FORM>(row, swz, c, one, vv, q0, q1);
        if (k >= k_on) v = vv;
        end_batch();
    }
    const float2 v_in = v;
    if (a.states_only) { if (live) a.seg_in[s] = v_in; return; }    // the tiles themselves are replayed one warp each (k_dc_exact)
    // ---- output: 16 steps at a time through a padded tile (rows 144 B apart: the per-lane 16-byte writes of a quarter warp and
    //      the transposed reads -- eight lanes per row -- are both conflict-free); every row leaves as one full 128-byte line.
    //      (A TMA store per box was measured slower here: the proxy fence and the read-wait cost ~800 cycles per round.)
    constexpr int kORow = 144;
    const int orow = lane >> 3, ochunk = lane & 7;
    float2 *const y_lane = a.y + (long long)(s0 + orow) * L;          // segment start of this lane's first row of a tile
#pragma unroll 1
    for (; k < nb; ++k) {
        begin_batch(k);
        const int base = d + k * kDtBatch;
#pragma unroll 1
        for (int h = 0; h < kDtBoxes; ++h) {
            const int hb = base + h * kDtBox;
            const unsigned char *rowh = row + h * kDtBoxBytes;
            const bool on = hb + kDtBox <= p1;
            __syncwarp();                                            // the previous round's rows have been read
            if (!warp_partial || __all_sync(0xffffffffu, on || hb >= p1)) {
                float2 vv = v;
                float4 q = h == 0 ? q0 : *reinterpret_cast<const float4 *>(rowh + swz);
                float4 qn = h == 0 ? q1 : *reinterpret_cast<const float4 *>(rowh + (16u ^ swz));
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    float2 y0, y1;
                    dt_step<FORM>(c, one, make_float2(q.x, q.y), vv, y0);
                    dt_step<FORM>(c, one, make_float2(q.z, q.w), vv, y1);
                    *reinterpret_cast<float4 *>(my_out_p + t * 16) = make_float4(y0.x, y0.y, y1.x, y1.y);
                    q = qn;
                    if (t + 2 < 8) qn = *reinterpret_cast<const float4 *>(rowh + (((unsigned)(t + 2) << 4) ^ swz));
                }
                if (on) v = vv;
            } else {                                                 // the block ends inside these 16 steps
                for (int t = 0; t < kDtBox / 2; ++t) {
                    const float4 q = *reinterpret_cast<const float4 *>(rowh + (((unsigned)t << 4) ^ swz));
                    float2 y0 = make_float2(0.f, 0.f), y1 = y0;
                    if (hb + 2 * t < p1) dt_step<FORM>(c, one, make_float2(q.x, q.y), v, y0);
                    if (hb + 2 * t + 1 < p1) dt_step<FORM>(c, one, make_float2(q.z, q.w), v, y1);
                    *reinterpret_cast<float4 *>(my_out_p + t * 16) = make_float4(y0.x, y0.y, y1.x, y1.y);
                }
            }
            __syncwarp();
            const int ob = (k - kw) * kDtBatch + h * kDtBox + 2 * ochunk;     // this lane's pair inside the segment of its row
            const float4 *tp = reinterpret_cast<const float4 *>(out_p + orow * kORow + ochunk * 16);
            if ((long long)(s0 + 31) * L + (k - kw) * kDtBatch + (h + 1) * kDtBox <= (long long)n) {   // the whole tile lies inside the block
                float4 *gp = reinterpret_cast<float4 *>(y_lane + ob);
#pragma unroll
                for (int it = 0; it < 8; ++it) { *gp = tp[it * (4 * kORow / 16)]; gp += 2 * (size_t)L; }   // four rows down: 4 L samples = 2 L float4
            } else {
#pragma unroll 1
                for (int it = 0; it < 8; ++it) {
                    const long long idx = (long long)(s0 + 4 * it + orow) * L + ob;
                    const float4 q = tp[it * (4 * kORow / 16)];
                    if (idx + 2 <= n) *reinterpret_cast<float4 *>(a.y + idx) = q;
                    else if (idx < n) a.y[idx] = make_float2(q.x, q.y);
                }
            }
        }
        end_batch();
    }
    if (live) dc_segment_done(a, s, v_in, v);
}
// which tiles started from a state that is not their neighbour's end state (stats[2] counts them: parallel re-runs)
__global__ void __launch_bounds__(256) k_dc_mark(DcExactArgs a, int *mask, unsigned *stats) {
    const int s = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (s >= a.n_seg) return;
    int bad = 0;
    if (s > 0) { const float2 w = a.seg_in[s], t = a.seg_out[s - 1]; bad = (w.x != t.x) | (w.y != t.y); }
    mask[s] = bad;
    if (bad) atomicAdd(stats + 2, 1u);
}
// All junctions are compared in parallel; only when one of them fails (rare) does thread 0 walk them serially and replay.
__global__ void __launch_bounds__(256) k_dc_verify(DcExactArgs a) {
    if (a.start) {                                                   // every segment has checked its own junction (dc_segment_done)
        if (a.junction_bad[0] == 0u) return;
        __syncthreads();
        if (threadIdx.x == 0) a.junction_bad[0] = 0u;
    }
    int bad = 0;
    for (int s = 1 + (int)threadIdx.x; s < a.n_seg; s += (int)blockDim.x) {
        const float2 w = a.seg_in[s], t = a.seg_out[s - 1];
        bad |= (w.x != t.x) | (w.y != t.y);
    }
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0) {
        if (bad) dc_exact_verify(a);
        else a.fstate[0] = a.seg_out[a.n_seg - 1];
    }
}

// ---- bit-exact DC blocker without a warm-up (dsp_dcwin.h): candidate windows, chunk composition, walk + gather --------------
// pass 1: CTA = (boundary b, component k), thread = candidate.  The tile's samples of that component sit in shared memory (one
// LDS.128 per four steps, broadcast); the chain is two dependent FP32 operations a step and eight warps a CTA / eight CTAs an SM
// keep the issue slots busy: this is throughput work spread over the whole chip, not a latency chain on a few SMs.
__global__ void __launch_bounds__(kDwCand) k_dc_window(DcWinArgs a) {
    __shared__ __align__(16) float sx[kDcTile];
    const int b = (int)(blockIdx.x >> 1), k = (int)(blockIdx.x & 1), i = (int)threadIdx.x;
    if (!dw_good(a, b, k)) return;                                   // uniform: the run of an earlier boundary covers this tile
    float v = dw_unord(dw_ord(dw_center(a, b, k)) + i - kDwHalf);
    const float c = a.c;
    const float *xk = reinterpret_cast<const float *>(a.xraw) + k;
    for (int p = b; p < a.tiles;) {
        const int i0 = p * kDcTile;
        int len = a.n - i0; if (len > kDcTile) len = kDcTile;
        __syncthreads();                                             // the previous tile has been consumed
        for (int j = i; j < len; j += kDwCand) sx[j] = __ldg(xk + 2 * (size_t)(i0 + j));
        __syncthreads();
        if (len == kDcTile) {
            const float4 *s4 = reinterpret_cast<const float4 *>(sx);
#pragma unroll 8
            for (int j = 0; j < kDcTile / 4; ++j) {
                const float4 q = s4[j];
                v = __fadd_rn(q.x, __fmul_rn(c, v));
                v = __fadd_rn(q.y, __fmul_rn(c, v));
                v = __fadd_rn(q.z, __fmul_rn(c, v));
                v = __fadd_rn(q.w, __fmul_rn(c, v));
            }
        } else {
            for (int j = 0; j < len; ++j) v = __fadd_rn(sx[j], __fmul_rn(c, v));
        }
        ++p;
        if (p >= a.tiles) break;
        dw_row(a.st, p, k)[i] = v;
        if (dw_good(a, p, k)) break;                                 // uniform
    }
}
// pass 2: CTA = (chunk, component): the chunk's table rows in shared memory, thread = candidate at the chunk's entry
__global__ void __launch_bounds__(kDwCand) k_dc_compose(DcWinArgs a) {
    __shared__ float rows[kDwChunk * kDwCand];
    __shared__ DwChunkInfo ci;
    const int c = (int)(blockIdx.x >> 1), k = (int)(blockIdx.x & 1), i = (int)threadIdx.x;
    if (i < kDwChunk) dw_chunk_info(a, c, k, i, &ci);
    __syncthreads();
    for (int j = 0; j < ci.n; ++j) rows[j * kDwCand + i] = dw_row(a.st, c * kDwChunk + 1 + j, k)[i];
    __syncthreads();
    dw_compose_thread(a, c, k, i, rows, &ci);
}
// pass 3: the chunk maps in shared memory, one thread per component walks them (two warps: no divergence between the
// components), then everybody gathers the true state of every boundary
__global__ void __launch_bounds__(1024) k_dc_resolve(DcWinArgs a, int gmap_in_smem) {
    extern __shared__ __align__(16) unsigned char rs_sm[];
    short *sg = reinterpret_cast<short *>(rs_sm);
    const int chunks = (a.tiles - 1 + kDwChunk - 1) / kDwChunk;
    const short *gm = a.gmap;
    if (gmap_in_smem) {
        const int words = chunks * kDwCand;                          // 2 components x kDwCand shorts = kDwCand 32-bit words a chunk
        for (int w = (int)threadIdx.x; w < words; w += (int)blockDim.x)
            reinterpret_cast<unsigned *>(sg)[w] = reinterpret_cast<const unsigned *>(a.gmap)[w];
        gm = sg;
    }
    __syncthreads();
    if (threadIdx.x == 0) dw_walk(a, 0, gm);
    else if (threadIdx.x == 32) dw_walk(a, 1, gm);
    __syncthreads();
    for (int p = 1 + (int)threadIdx.x; p < a.tiles; p += (int)blockDim.x) dw_gather_thread(a, p);
}

// ---- audio-side consumers (dsp_sink.h) ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mix_feed(MixArgs a) { mix_feed_thread(a, (int)(blockIdx.x * blockDim.x + threadIdx.x)); }
__global__ void __launch_bounds__(128) k_mix_walk(MixArgs a) { mix_walk_thread(a, (int)(blockIdx.x * blockDim.x + threadIdx.x)); }
__global__ void __launch_bounds__(256) k_mix_sum(MixArgs a) { mix_sum_thread(a, (int)(blockIdx.x * blockDim.x + threadIdx.x)); }
__global__ void __launch_bounds__(256) k_mix_scale(MixArgs a) {
    __shared__ float sc;
    if (threadIdx.x == 0) sc = mix_scale(a);
    __syncthreads();
    const int k = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (k < 2 * a.n_frames && sc != 1.0f) a.out[k] = __fmul_rn(a.out[k], sc);
}
// a host-fed AudioThreadInput (tests, or audio that does not come from this engine): samples staged in `data`
__global__ void k_mix_push(MixArgs a, int j, const float *data, int n, int channels, float peak, int rate) {
    __shared__ int w_s;
    MixSource &s = a.src[j];
    if (threadIdx.x == 0) w_s = mix_ring_take(s, n);
    __syncthreads();
    const int w = w_s;
    if (w < 0) return;
    float *ring = a.rings + s.ring_off;
    for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) ring[w + i] = data[i];
    __syncthreads();
    if (threadIdx.x == 0) mix_queue_push(s, w, n, channels, rate, peak);
}
__global__ void __launch_bounds__(256) k_pcm16(PcmArgs a) { pcm16_thread(a, (int)(blockIdx.x * blockDim.x + threadIdx.x)); }
__global__ void __launch_bounds__(256) k_scope_prepare(ScopeArgs a) {
    __shared__ float pk;
    if (threadIdx.x == 0) pk = scope_wave_peak_serial(a);          // <= max_samples comparisons; every CTA repeats them
    __syncthreads();
    scope_prepare_thread(a, (int)(blockIdx.x * blockDim.x + threadIdx.x), pk);
}
// averages, ceil / floor, scalar averages and points of the audio spectrum: one CTA (fft_size / 2 <= 4096 bins)
__global__ void __launch_bounds__(1024) k_scope_finish(ScopeArgs a) {
    __shared__ double smx[32], smn[32];
    double mx = 0.0, mn = 1.0;
    for (int i = (int)threadIdx.x; i < a.fft_size / 2; i += (int)blockDim.x) {
        const double maa = scope_ema_bin(a, i);
        if (maa > mx) mx = maa;
        if (maa < mn) mn = maa;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ox = __shfl_xor_sync(0xffffffffu, mx, o), on = __shfl_xor_sync(0xffffffffu, mn, o);
        if (ox > mx) mx = ox;
        if (on < mn) mn = on;
    }
    if ((threadIdx.x & 31) == 0) { smx[threadIdx.x >> 5] = mx; smn[threadIdx.x >> 5] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { if (smx[w] > mx) mx = smx[w]; if (smn[w] < mn) mn = smn[w]; }
        scope_finish_serial(a, mx, mn);
    }
    __syncthreads();
    for (int i = (int)threadIdx.x; i < a.fft_size / 2; i += (int)blockDim.x) scope_points_thread(a, i);
}

template <int K>
__global__ void __launch_bounds__(kFrontWarps * 32, 1) k_chan_front(FrontArgs a, int jobs) {
    extern __shared__ __align__(16) unsigned char front_sm[];
    float2 *rep = reinterpret_cast<float2 *>(front_sm);
    float4 *s_x = reinterpret_cast<float4 *>(front_sm + kFrontSmemTab);
    unsigned long long t_start = 0;
    if (a.dbg && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    for (int i = threadIdx.x; i < 1024 * 16; i += blockDim.x) rep[i] = __ldg(a.nco + (i >> 4));
    __syncthreads();
    if (a.dbg && threadIdx.x == 0) {
        unsigned smid; unsigned long long t1;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        a.dbg[4 * blockIdx.x] = smid; a.dbg[4 * blockIdx.x + 1] = t_start; a.dbg[4 * blockIdx.x + 2] = t1;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int job = blockIdx.x + gridDim.x * warp;
    if (job < jobs) {
        int span_idx, chunk;
        front_job(a, job, span_idx, chunk);
        int ci = chunk * 32 + lane;
        const bool real = ci < a.cg;
        if (!real) ci = a.cg - 1;                              // padding lanes shadow the last channel, without stores
        const int ch = a.group[ci];
        const long long g0 = a.cd[ch].g0;
        // warp-staged samples need every lane to read the same samples: same position in the same stream
        // (a lane with a private history -- a channel that skipped blocks, or whose history was pre-mixed -- reads other
        // samples than its neighbours for i < 0, so such a warp takes the per-lane path as well)
        const unsigned long long xin = (unsigned long long)a.cd[ch].xin, hst = (unsigned long long)a.cd[ch].hist;
        const bool lock_step = __all_sync(0xffffffffu, g0 == __shfl_sync(0xffffffffu, g0, 0)) &&
                               __all_sync(0xffffffffu, xin == 0ull) && __all_sync(0xffffffffu, hst == 0ull);
        TabRep16 tab;
        tab.base = reinterpret_cast<const char *>(front_sm);
        tab.lane_off = (unsigned)(lane & 15) * 8u;
        if (K >= 4 && lock_step) {
            XWarpStage<(1 << K)> xs;
            xs.buf = s_x + warp * 64; xs.lane = lane; xs.cur = 0;
            front_span<K, XWarpStage<(1 << K)>, TabRep16>(a, ch, span_idx, real, tab, xs);
        } else if (real) {
            XDirect xs;
            xs.p = a.x;
            front_span<K, XDirect, TabRep16>(a, ch, span_idx, true, tab, xs);
        }
    }
    if (a.dbg) {
        unsigned long long t2;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t2));
        atomicMax(a.dbg + 4 * blockIdx.x + 3, t2);
    }
}
__global__ void k_premix_hist(PremixArgs a) { premix_thread(a, 1 + (int)(blockIdx.x * blockDim.x + threadIdx.x)); }
__global__ void __launch_bounds__(kTailThreads, 4) k_chan_tail(TailArgs a) {
    __shared__ TailShared sm;
    const bool dbg = a.dbg && threadIdx.x == 0 && blockIdx.x < 64;
    if (dbg) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.dbg[blockIdx.x * 64 + 63] = t; }
    for (int ph = 0; ph < TP_COUNT; ++ph) {
        tail_phase(ph, a, a.first + blockIdx.x, threadIdx.x, blockDim.x, &sm);
        __syncthreads();
        if (dbg) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.dbg[blockIdx.x * 64 + ph] = t; }
    }
}
// Spectrum transform, Stockham passes with register butterflies (dsp_spec.h, second form).  CTA = (frame, sub-sequence).
__global__ void __launch_bounds__(kFft2Threads) k_spec_fft2(SpecArgs a) {
    extern __shared__ __align__(16) unsigned char fft2_sm[];
    float2 *buf = reinterpret_cast<float2 *>(fft2_sm);
    if (blockIdx.x == 0) for (int f = (int)threadIdx.x; f < a.F; f += (int)blockDim.x) a.frame_cf[f] = make_float2(0.f, 1.f);
    const Fft2Plan pl = fft2_plan(a.nsub);
    Fft2Regs rg;
    spec_fft2_load(a, (int)blockIdx.x, (int)threadIdx.x, (int)blockDim.x, buf);
    __syncthreads();
    for (int pass = 0; pass < pl.passes; ++pass) {
        spec_fft2_half(2 * pass, a, (int)blockIdx.x, (int)threadIdx.x, buf, rg);
        __syncthreads();                                            // everybody has read
        spec_fft2_half(2 * pass + 1, a, (int)blockIdx.x, (int)threadIdx.x, buf, rg);
        __syncthreads();
    }
}
// averages + frame ceil / floor (fused): one thread per bin, frames in order; the frame's ceil / floor are folded per warp, then per
// CTA through shared memory (double-buffered by frame parity: one barrier per frame), and merged with one atomic pair per CTA and frame
__global__ void __launch_bounds__(kSpecThreads) k_spec_ema2(SpecArgs a) {
    __shared__ float s_mx[2][kSpecThreads / 32], s_mn[2][kSpecThreads / 32];
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    const bool live = i < a.N;
    double ma = 0.0, maa = 0.0;
    if (live) { ma = a.ema[i].x; maa = a.ema[i].y; }
    unsigned *cf = reinterpret_cast<unsigned *>(a.frame_cf);
    const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
    for (int f = 0; f < a.F; ++f) {
        float v = 0.f;
        if (live) spec_ema2_step(a, f, i, ma, maa, v);
        float mx = live ? v : 0.f, mn = live ? v : 1.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        }
        if (lane == 0) { s_mx[f & 1][warp] = mx; s_mn[f & 1][warp] = mn; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float cx = 0.f, cn = 1.f;
#pragma unroll
            for (int w = 0; w < kSpecThreads / 32; ++w) { cx = fmaxf(cx, s_mx[f & 1][w]); cn = fminf(cn, s_mn[f & 1][w]); }
            atomicMax(cf + 2 * f, __float_as_uint(cx));
            atomicMin(cf + 2 * f + 1, __float_as_uint(cn));
        }
    }
    if (live) a.ema[i] = make_double2(ma, maa);
}
// scalar averages + display points + waterfall line (fused): CTA = (256 points, frame); thread 0 replays the scalar averages up to
// its frame (a few dozen dependent operations), everybody then maps its point
__global__ void __launch_bounds__(kSpecThreads) k_spec_points2(SpecArgs a) {
    __shared__ double2 cfm;
    const int f = (int)blockIdx.y, xp = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (threadIdx.x == 0) {
        const bool last = f == a.F - 1 && blockIdx.x == 0;
        double after[4];
        cfm = spec_scalar_upto(a, f, after);
        if (blockIdx.x == 0) a.frame_cf_maa[f] = cfm;
        if (last) { a.scalars_out[0] = after[0]; a.scalars_out[1] = after[1]; a.scalars_out[2] = after[2]; a.scalars_out[3] = after[3]; }
    }
    __syncthreads();
    spec_points2_thread(a, f, xp, cfm.x, cfm.y);
}
__global__ void __launch_bounds__(256) k_view_mix(ViewMixArgs a) {
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < (long long)a.F * a.D; g += (long long)gridDim.x * blockDim.x)
        view_mix_thread(a, g);
}
__global__ void __launch_bounds__(kSpecThreads) k_view_ema_move(SpecArgs a, int ph) {
    view_ema_move_thread(a, ph, blockIdx.x * blockDim.x + threadIdx.x);
}
// serial float recurrence at the channel rate.  The warp stages 256 samples at a time in shared memory; every lane runs the
// chain (fully unrolled 32 steps, inputs fetched ahead of it) and lane l keeps output l, so a batch is stored coalesced and
// the chain runs at its latency (~10 cycles a step) instead of waiting for shared memory every step.
__global__ void __launch_bounds__(kPfbThreads) k_pfb_analyze(PfbArgs a) {
    extern __shared__ float2 pfb_sm[];
    for (int ph = 0; ph < kPfbPhases; ++ph) {
        pfb_phase(ph, a, blockIdx.x, threadIdx.x, blockDim.x, pfb_sm);
        __syncthreads();
    }
}
#endif

// Per-block parameter blocks (a few KB: channel descriptors, group lists, frame offsets) are pulled from their pinned,
// device-visible host slots by a tiny kernel instead of cudaMemcpyAsync: a memcpy would queue on the host-to-device copy
// engine behind the caller's multi-megabyte input transfers and stall the compute stream that waits for it.
#ifndef SDR_EMU
__global__ void k_upload_words(const unsigned *__restrict__ src_host, unsigned *__restrict__ dst, unsigned n_words) {
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += gridDim.x * blockDim.x) dst[i] = src_host[i];
}
#endif
static void upload_params(void *dst, const void *src_pinned, size_t bytes, sdr_stream_t s, unsigned long long &launches) {
    if (!bytes) return;
#ifndef SDR_EMU
    const unsigned words = (unsigned)((bytes + 3) / 4);          // all parameter blocks are arrays of 4-byte-aligned structs
    unsigned blocks = (words + 255u) / 256u;
    if (blocks > 64u) blocks = 64u;
    k_upload_words<<<blocks, 256, 0, s>>>(reinterpret_cast<const unsigned *>(src_pinned), reinterpret_cast<unsigned *>(dst), words);
    launches += 1;
#else
    (void)s; (void)launches;
    memcpy(dst, src_pinned, bytes);
#endif
}

#ifndef SDR_EMU
// cuTensorMapEncodeTiled is the one driver-API call the library needs: fetched through the runtime, nothing links libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tma_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else cudaGetLastError();
    }
    return fn;
}
// rows of `row_samples` complex samples (as pairs of floats) starting at `base`; boxes of 16 samples x 32 rows, 128-byte swizzle
static bool tma_rows_map(CUtensorMap *m, float2 *base, unsigned row_samples, unsigned long long rows) {
    EncodeTiledFn enc = tma_encoder();
    if (!enc || rows == 0) return false;
    const cuuint64_t gdim[2] = {2ull * row_samples, rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)row_samples * sizeof(float2)};
    const cuuint32_t box[2] = {2u * kDtBox, 32u};
    const cuuint32_t estr[2] = {1u, 1u};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
#endif

#ifndef SDR_EMU
// cudaFuncSetAttribute is per device: a process that drives several GPUs (sdr_group_create) must raise the shared-memory limit of a
// kernel on every one of them.  `flags` is the caller's static table; returns true the first time on the current device.
static inline bool first_on_device(bool (&flags)[64]) {
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (flags[d]) return false;
    flags[d] = true;
    return true;
}
#endif
// candidate windows -> chunk maps -> walk + gather: the true state entering every tile lands in a.seg_in (dsp_dcwin.h)
struct sdr_engine;
static void mixers_feed(sdr_engine *e, sdr_stream_t ts);
static int launch_dc_window(const DcWinArgs &a, sdr_stream_t s, unsigned long long &launches) {
    const int chunks = (a.tiles - 1 + kDwChunk - 1) / kDwChunk;
#ifndef SDR_EMU
    k_dc_window<<<2 * a.tiles, kDwCand, 0, s>>>(a);
    k_dc_compose<<<2 * chunks, kDwCand, 0, s>>>(a);
    const size_t gbytes = (size_t)chunks * 2 * kDwCand * sizeof(short);
    const bool in_smem = gbytes <= 200 * 1024;
    static bool attr_set[64];
    if (first_on_device(attr_set)) CUDA_TRY(cudaFuncSetAttribute(k_dc_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_dc_resolve<<<1, 1024, in_smem ? gbytes : 0, s>>>(a, in_smem ? 1 : 0);
#else
    (void)s;
    for (int b = 0; b < a.tiles; ++b) for (int k = 0; k < 2; ++k) for (int i = 0; i < kDwCand; ++i) dw_window_thread(a, b, k, i);
    for (int c = 0; c < chunks; ++c) for (int k = 0; k < 2; ++k) {
        DwChunkInfo ci;
        std::vector<float> rows((size_t)kDwChunk * kDwCand);
        for (int j = 0; j < kDwChunk; ++j) dw_chunk_info(a, c, k, j, &ci);
        for (int j = 0; j < ci.n; ++j) for (int i = 0; i < kDwCand; ++i) rows[(size_t)j * kDwCand + i] = dw_row(a.st, c * kDwChunk + 1 + j, k)[i];
        for (int i = 0; i < kDwCand; ++i) dw_compose_thread(a, c, k, i, rows.data(), &ci);
    }
    dw_walk(a, 0, a.gmap); dw_walk(a, 1, a.gmap);
    for (int p = 1; p < a.tiles; ++p) dw_gather_thread(a, p);
#endif
    launches += 3;
    return SDR_OK;
}

// tma_rows: samples readable from xraw[-warm] on (rows of the input tensor = that / seg_len); 0 = use the cp.async kernel
static int launch_dc_exact(const DcExactArgs &a, unsigned long long tma_rows, sdr_stream_t s, unsigned long long &launches, int *fix_mask, unsigned *stats) {
#ifndef SDR_EMU
    if (a.start) {
        // the true state entering every tile is known (candidate windows): one warp per tile, every lane runs the 1024-step chain
        // and keeps two outputs per batch -- a few hundred light CTAs anywhere on the chip, no warm-up, no staging kernel
        k_dc_exact<<<a.n_seg, 32, 0, s>>>(a);
    } else if (a.n_seg > 1) {
        // Warm-up form: one lane per tile (SM-exclusive CTAs of four warps on the SMs the front kernel leaves free) runs the
        // chain over the kDcWarm samples in front of its tile from the exact state and leaves a CANDIDATE start state; the tiles
        // are then replayed one warp each; tiles whose candidate was not their neighbour's end state (1.5e-4 of them) are re-run
        // from it in one parallel round; k_dc_verify proves every junction and replays serially whatever is left (a run of
        // consecutive wrong candidates).
        DcExactArgs w = a;
        w.states_only = 1;
        const int warps = (a.n_seg + 31) / 32;
        static int use_tma = -1;
        if (use_tma < 0) { const char *v = getenv("SDR_B200_DC_TMA"); use_tma = (v && v[0] == '0') ? 0 : 1; }
        CUtensorMap in_map;
        if (use_tma && tma_rows) {
            if (!tma_rows_map(&in_map, const_cast<float2 *>(a.xraw) - a.warm, (unsigned)a.seg_len, tma_rows / (unsigned)a.seg_len))
                return fail(SDR_ERR_CUDA, "cuTensorMapEncodeTiled failed for the DC replay's stream tensors");
            static bool attr_set[64];
            static int form = -1;
            if (first_on_device(attr_set)) {
                CUDA_TRY(cudaFuncSetAttribute(k_dc_lanes_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDtSmem));
                CUDA_TRY(cudaFuncSetAttribute(k_dc_lanes_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDtSmem));
            }
            if (form < 0) { const char *f = getenv("SDR_B200_DC_FORM"); form = (f && f[0] == '0') ? 0 : 1; }
            if (form) k_dc_lanes_tma<1><<<(warps + kDlWarps - 1) / kDlWarps, 32 * kDlWarps, kDtSmem, s>>>(w, in_map);
            else k_dc_lanes_tma<0><<<(warps + kDlWarps - 1) / kDlWarps, 32 * kDlWarps, kDtSmem, s>>>(w, in_map);
        } else {
            static bool attr_set[64];
            if (first_on_device(attr_set)) CUDA_TRY(cudaFuncSetAttribute(k_dc_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, kDlSmem));
            k_dc_lanes<<<(warps + kDlWarps - 1) / kDlWarps, 32 * kDlWarps, kDlSmem, s>>>(w);
        }
        DcExactArgs r = a;                                           // every tile from its candidate
        r.start = a.seg_in; r.no_publish = 1;
        k_dc_exact<<<a.n_seg, 32, 0, s>>>(r);
        DcExactArgs f = a;                                           // wrong candidates: from the neighbour's end state
        f.fix_mask = fix_mask;
        k_dc_mark<<<(a.n_seg + 255) / 256, 256, 0, s>>>(a, fix_mask, stats);
        k_dc_exact<<<a.n_seg, 32, 0, s>>>(f);
        launches += 4;
    } else {
        k_dc_exact<<<1, 32, 0, s>>>(a);
    }
    k_dc_verify<<<1, 256, 0, s>>>(a);
#else
    (void)s; (void)tma_rows; (void)fix_mask; (void)stats;
    for (int g = 0; g < a.n_seg; ++g) dc_exact_segment_serial(a, g, false, make_float2(0.f, 0.f));
    dc_exact_verify(a);
#endif
    launches += 2;
    return SDR_OK;
}

// double-precision scan of the recurrence: tile sums + carries always (they seed the bit-exact replay), and the direct
// evaluation of y from them when `apply` is set (SDR_B200_DC=scan: the exact-arithmetic variant, not bit-exact)
static void launch_dc(const DcArgs &a, bool apply, sdr_stream_t s, unsigned long long &launches) {
    const int tiles = (a.n + kDcTile - 1) / kDcTile;
    if (tiles == 0) return;
#ifndef SDR_EMU
    k_dc_tile_sum<<<tiles, kDcThreads, 0, s>>>(a);
    k_dc_carry<<<1, kDcThreads, 0, s>>>(a, tiles);
    if (apply) k_dc_apply<<<tiles, kDcThreads, 0, s>>>(a);
#else
    (void)s;
    for (int b = 0; b < tiles; ++b) {
        DcShared sm;
        for (int ph = 0; ph < 2; ++ph) for (int t = 0; t < kDcThreads; ++t) dc_tile_sum_phase(ph, a, b, t, &sm);
    }
    {
        DcCarryShared cs;
        for (int ph = 0; ph < 3; ++ph) for (int t = 0; t < kDcThreads; ++t) dc_carry_phase(ph, a, tiles, t, &cs);
    }
    if (apply) for (int b = 0; b < tiles; ++b) {
        DcShared2 sm;
        for (int ph = 0; ph < 3; ++ph) for (int t = 0; t < kDcThreads; ++t) dc_apply_phase(ph, a, b, t, &sm);
    }
#endif
    launches += apply ? 3 : 2;
}

template <int K>
static int launch_front_k(const FrontArgs &a, int jobs, int reserved_sms, sdr_stream_t s) {
#ifndef SDR_EMU
    static bool attr_set[64];
    if (first_on_device(attr_set)) CUDA_TRY(cudaFuncSetAttribute(k_chan_front<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFrontSmemBytes));
    // one CTA per SM while the jobs fit; reserved_sms SMs stay free for the DC blocker's serial chains
    const int sms = 148 - reserved_sms;
    int n_cta = (jobs + kFrontWarps - 1) / kFrontWarps;
    if (jobs <= sms * kFrontWarps) n_cta = jobs < sms ? jobs : sms;
    k_chan_front<K><<<n_cta, kFrontWarps * 32, kFrontSmemBytes, s>>>(a, jobs);
#else
    (void)s; (void)reserved_sms;
    for (int j = 0; j < jobs; ++j) for (int l = 0; l < 32; ++l) front_thread<K>(a, j, l);
#endif
    return 0;
}
static int launch_front(int K, const FrontArgs &a, int jobs, int reserved_sms, sdr_stream_t s, unsigned long long &launches) {
    launches += 1;
    switch (K) {
    case 0: return launch_front_k<0>(a, jobs, reserved_sms, s);
    case 1: return launch_front_k<1>(a, jobs, reserved_sms, s);
    case 2: return launch_front_k<2>(a, jobs, reserved_sms, s);
    case 3: return launch_front_k<3>(a, jobs, reserved_sms, s);
    case 4: return launch_front_k<4>(a, jobs, reserved_sms, s);
    case 5: return launch_front_k<5>(a, jobs, reserved_sms, s);
    default: return launch_front_k<6>(a, jobs, reserved_sms, s);
    }
}

static void launch_tail(const TailArgs &a, int n_chan, sdr_stream_t s, unsigned long long &launches) {
    if (n_chan == 0) return;
#ifndef SDR_EMU
    k_chan_tail<<<n_chan, kTailThreads, 0, s>>>(a);
#else
    (void)s;
    TailShared *sm = new TailShared;
    for (int c = 0; c < n_chan; ++c)
        for (int ph = 0; ph < TP_COUNT; ++ph) for (int t = 0; t < kTailThreads; ++t) tail_phase(ph, a, a.first + c, t, kTailThreads, sm);
    delete sm;
#endif
    launches += 1;
}

static int launch_spectrum(const SpecArgs &a, sdr_stream_t s, unsigned long long &launches) {
    if (a.F == 0) return 0;
#ifndef SDR_EMU
    const Fft2Plan pl = fft2_plan(a.nsub);
    const size_t smem = (size_t)(a.nsub + a.nsub / pl.r[0]) * sizeof(float2);
    static bool attr_set[64];
    if (first_on_device(attr_set)) CUDA_TRY(cudaFuncSetAttribute(k_spec_fft2, cudaFuncAttributeMaxDynamicSharedMemorySize, kFft2SmemElems * (int)sizeof(float2)));
    k_spec_fft2<<<a.F * a.P, kFft2Threads, smem, s>>>(a);
    if (a.rebin || a.shift) {
        k_view_ema_move<<<(a.N + kSpecThreads - 1) / kSpecThreads, kSpecThreads, 0, s>>>(a, 0);
        k_view_ema_move<<<(a.N + kSpecThreads - 1) / kSpecThreads, kSpecThreads, 0, s>>>(a, 1);
        launches += 2;
    }
    k_spec_ema2<<<(a.N + kSpecThreads - 1) / kSpecThreads, kSpecThreads, 0, s>>>(a);
    k_spec_points2<<<dim3((unsigned)((a.fft_size + kSpecThreads - 1) / kSpecThreads), (unsigned)a.F), kSpecThreads, 0, s>>>(a);
#else
    (void)s;
    {
        const Fft2Plan pl = fft2_plan(a.nsub);
        std::vector<float2> buf((size_t)a.nsub + a.nsub / 2 + 16);
        std::vector<Fft2Regs> regs(kFft2Threads);
        for (int f = 0; f < a.F; ++f) a.frame_cf[f] = make_float2(0.f, 1.f);
        for (int b = 0; b < a.F * a.P; ++b) {
            for (int t = 0; t < kFft2Threads; ++t) spec_fft2_load(a, b, t, kFft2Threads, buf.data());
            for (int hp = 0; hp < 2 * pl.passes; ++hp) for (int t = 0; t < kFft2Threads; ++t) spec_fft2_half(hp, a, b, t, buf.data(), regs[t]);
        }
    }
    if (a.rebin || a.shift) for (int ph = 0; ph < 2; ++ph) for (int i = 0; i < a.N; ++i) view_ema_move_thread(a, ph, i);
    for (int i = 0; i < a.N; ++i) spec_ema2_thread_serial(a, i);
    for (int f = 0; f < a.F; ++f) {
        double after[4];
        const double2 cfm = spec_scalar_upto(a, f, after);
        a.frame_cf_maa[f] = cfm;
        for (int xp = 0; xp < a.fft_size; ++xp) spec_points2_thread(a, f, xp, cfm.x, cfm.y);
        if (f == a.F - 1) for (int k = 0; k < 4; ++k) a.scalars_out[k] = after[k];
    }
#endif
    launches += 3;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// host objects
// ---------------------------------------------------------------------------------------------------------------------
static int modem_from_name(const char *n) {
    if (!n) return -1;
    std::string s(n);
    if (s == "FM") return MODEM_FM;
    if (s == "NBFM") return MODEM_NBFM;
    if (s == "AM") return MODEM_AM;
    if (s == "USB") return MODEM_USB;
    if (s == "LSB") return MODEM_LSB;
    if (s == "I/Q") return MODEM_IQ;
    if (s == "DSB") return MODEM_DSB;
    if (s == "CW") return MODEM_CW;
    if (s == "FMS") return MODEM_FMS;
    return -1;
}

// Modem::checkSampleRate: ModemAnalog.cpp:14-19 (MIN_BANDWIDTH 500), ModemUSB.cpp:29-37 / ModemLSB.cpp:29-37
static int check_sample_rate(int modem, long long bw) {
    if (modem == MODEM_FMS) return bw < 100000 ? 100000 : (int)bw;      // ModemFMStereo.cpp:27-35
    if (bw < 500) return 500;
    if ((modem == MODEM_USB || modem == MODEM_LSB) && (bw % 2)) return (int)(bw + 1);
    return (int)bw;
}

struct sdr_chan {
    sdr_engine *eng = nullptr;
    int index = -1;                 // slot in the engine's arrays
    int modem = 0;
    long long sample_rate = 0;
    int bandwidth = 0, audio_rate = 0;
    int asked_bandwidth = 0;        // what the caller asked for: DemodulatorThreadPostIQData::sampleRate (DemodulatorPreThread.cpp:222)
    ResamplerPlan iq_plan, audio_plan;
    ChanStatic cs;
    float *d_arena = nullptr;
    size_t arena_floats = 0;
    int lvl0_sel = 0;               // level-0 buffer the channel's next processed block will use
    bool standalone = false;        // sdr_modem: driven by sdr_modem_demodulate, ignored by sdr_engine_submit
    unsigned inject_cap = 0;
    ChanDyn *h_dyn = nullptr;       // pinned, standalone only
    unsigned char *h_slot = nullptr; // pinned result slot, standalone only
    float2 *d_priv_hist = nullptr;  // tail of the last consumed block, kept while the channel skips blocks
    bool priv_valid = false;
    int roll_pending = 0;           // > 0: the last consumed block was this short and read its history from d_priv_hist: the
                                    // private tail must be extended by that block before it is used again
    int pfb_slot_prev = -1;         // channelizer mode: stream slot this channel read in the last block it consumed
    // DemodulatorPreThread state
    long long frequency = 0;
    long long shift_frequency = 0;  // last applied shift (starts at 0 like the reference member)
    unsigned theta = 0, dtheta = 0, dtheta_prev = 0;
    int dir_prev = 0;
    long long consumed = 0;         // wideband samples consumed
    std::vector<PremixSeg> recent;  // oscillator settings of the last consumed blocks, newest first (rare-path history mixing)
    size_t slot_bytes = 0;
    // lifecycle (see "channels" below): device memory comes from the stream-ordered allocator on the engine's admin stream
    bool squelch_enabled = false;   // DemodulatorThread::squelchEnabled / squelchLevel (DemodulatorThread.cpp:25-26, 384-397)
    float squelch_level = -100.f;
    bool fresh = true;              // the first submit that sees the channel makes its main stream wait for ev_ready
    bool cs_dirty = false;          // descriptor changed after creation (sdr_chan_want_iq): re-uploaded with the next block
    bool async_mem = false;         // arena / history came from be_malloc_async
#ifndef SDR_EMU
    cudaEvent_t ev_ready = nullptr; // arena cleared, descriptor / state / filter taps on the device
    cudaEvent_t ev_dead = nullptr;  // destroyed: every block that could still touch the channel's memory has drained
#endif
};

struct sdr_spectrum {
    sdr_engine *eng = nullptr;
    unsigned fft_size = 0, N = 0, lps = 0;
    int P = 1, nsub = 0, log_nsub = 0;
    int scalar_sel = 0;                          // which half of d_scalars holds the current scalar averages
    float average_rate = 0.65f, scale_factor = 1.0f;
    // framing state (FFTDataDistributor)
    long long buffered = 0, last_rate = 0, last_center = 0;
    double line_accum = 0.0;
    // zoomed view (SpectrumVisualProcessor::setView, .cpp:57-75, process() :283-386): members as in the reference
    bool is_view = false, last_view = false;
    long long view_center = 0; long view_bw = 0;
    long shift_frequency = 0, last_bandwidth = 0, last_input_bw = 0;
    unsigned nco_theta = 0, nco_dtheta = 0;         // freqShifter
    sdr_chan *vchan = nullptr;                      // the view's msresamp_crcf (a hidden, modem-less channel)
    double visual_ratio = 1.0;                      // bandwidth / resampleBw of the point table on the device
    int *d_pt = nullptr, *h_pt = nullptr;           // [2][fft_size]: first bin offset and bin count of every display point
    double2 *d_ema_tmp = nullptr;
    // device
    float2 *d_tw = nullptr; float *d_mags = nullptr, *d_maa = nullptr; double2 *d_ema = nullptr; double *d_scalars = nullptr;
    float2 *d_frame_cf = nullptr; double2 *d_frame_cf_maa = nullptr; float *d_points = nullptr; unsigned char *d_lines = nullptr;
    int *d_frame_start = nullptr;
    // pinned host
    float *h_points = nullptr; unsigned char *h_lines = nullptr; double2 *h_cf = nullptr; int *h_frame_start = nullptr;
    float *h_dbg = nullptr;
    unsigned f_max = 0, frames_last = 0;
    unsigned frames_buf[kResBuf] = {0, 0, 0};   // frames of the blocks whose results sit in each host buffer
    float *hp(int q) const { return h_points + (size_t)q * f_max * fft_size; }
    unsigned char *hl(int q) const { return h_lines + (size_t)q * f_max * fft_size; }
    double2 *hc(int q) const { return h_cf + (size_t)q * f_max; }
};

struct sdr_engine {
    int device = 0;
    sdr_stream_t stream = nullptr;
    bool own_stream = false;
    unsigned max_block = 0;
    unsigned chan_block_max = 0;                // most samples one demodulator receives per block (= max_block unless the channelizer divides it)
    std::mutex mu;
    unsigned long long launches = 0;
    // wideband buffers
    size_t hist_cap = 0, raw_pad = 0;
    float2 *d_raw = nullptr, *d_scratch = nullptr;
    float2 *d_blkbuf[2] = {nullptr, nullptr};   // DC-blocked stream, double-buffered: [hist_cap | max_block] each
    float2 *d_rawbuf = nullptr;                 // raw stream [hist_cap | max_block] (the bit-exact DC replay warms up on history)
    // channelizer mode (SDRPostThread::runPFBCH): filterbank tables for the current channel count, the channel streams
    // [2 parities][M + 1 slots][kBlkHist history | frames] (slot M = DC-corrected channel 0) and the channel-0 DC state
    int chan_mode = 0;                          // 0: runSingleCH for every rate; 1: runPFBCH when the rate gives more than one channel
    int pfb_M = 0, pfb_osf = 0; long long pfb_frame_base = 0; float *d_pfb_hq = nullptr; float2 *d_pfb_tw = nullptr;
    float2 *d_pfb_streams[2] = {nullptr, nullptr}; size_t pfb_stride = 0; size_t pfb_slots = 0;
    float2 *d_pfb_dc_state = nullptr;
    unsigned pfb_last_frames = 0; bool pfb_have_block = false;
    int parity = 0;                             // buffer holding the block submitted last
    bool exact_dc = true;                       // bit-exact replay of the reference's float DC blocker (SDR_B200_DC=scan disables)
    double2 *d_tile_carry2[2] = {nullptr, nullptr};
    float2 *d_seg_in = nullptr, *d_seg_out = nullptr, *d_fstate = nullptr;
    unsigned *d_fixups = nullptr;
    // candidate-window DC replay (dsp_dcwin.h): tables and walker outputs; dc_window: -1 auto (by channel load), 0 off, 1 on
    float *d_dw_st = nullptr, *d_dw_pst = nullptr;
    short *d_dw_gmap = nullptr;
    int *d_dw_entry = nullptr;
    unsigned *d_dw_stats = nullptr;
    int *d_dc_fix_mask = nullptr;            // warm-up replay: tiles to re-run in a fix-up round
    int dc_window = -1;
    std::vector<struct sdr_mixer *> mixers;   // device-side audio mixers fed after every block (dsp_sink.h)
    std::mutex mu_pcm;                        // sdr_chan_fetch_pcm16 calls, one at a time
    short *d_pcm = nullptr, *h_pcm = nullptr; // its staging: [count | words]
    size_t pcm_cap = 0;
#ifndef SDR_EMU
    cudaEvent_t ev_pcm = nullptr;
#endif
    unsigned long long dw_blocks = 0;
    // host blocks: ring of pinned staging buffers -> DMA on stream_h2d -> ring of device buffers -> the DC stream
    static const int kInRing = 3;
    sdr_cf32 *h_in[kInRing] = {nullptr, nullptr, nullptr};
    float2 *d_in[kInRing] = {nullptr, nullptr, nullptr};
    bool in_rec[kInRing] = {false, false, false};
    int in_sel = 0;
    bool poisoned = false;              // a CUDA error after host-side state had advanced: the engine refuses further blocks
    unsigned last_n = 0;
    bool have_block = false;
    // DC blocker
    double2 *d_tile_sum = nullptr, *d_dc_state = nullptr; double *d_apow = nullptr;
    double dc_a = 0.0;
    float2 *d_nco = nullptr;
    unsigned long long *d_front_dbg = nullptr;   // diagnostics (sdr_engine_debug_front_timeline)
    // channels
    static const int kMaxChan = 2048;
    std::vector<sdr_chan *> chans;     // by slot; nullptr = free
    ChanStatic *d_cs = nullptr; ChanState *d_st = nullptr;
    ChanDyn *d_cd2[2] = {nullptr, nullptr};   // per-block parameters, double-buffered: the tail of block b reads its copy
                                              // while the next submit already uploads block b+1's
    // small per-block parameter blocks live in a ring of pinned buffers so that submits can be enqueued back to back
    static const int kRing = 8;
    ChanDyn *h_cd = nullptr;            // pinned [kRing][kMaxChan]
    int *d_group = nullptr; int *h_group = nullptr;   // h_group pinned [kRing][kMaxChan]
    unsigned long long ring_seq = 0;
#ifndef SDR_EMU
    cudaEvent_t ring_ev[kRing];
#endif
    std::vector<float> taps_host; float *d_taps = nullptr; size_t taps_cap = 0, taps_uploaded = 0;
    std::map<std::string, int> taps_cache;
    unsigned char *d_results = nullptr, *h_results = nullptr; size_t results_bytes = 0;
    // Channel lifecycle: sdr_chan_create / sdr_chan_destroy run on other threads than the submitter (DemodulatorWorkerThread,
    // DemodulatorMgr) and must never make it wait (SDRPostThread only try_push()es, SDRPostThread.cpp:238-298).  `mu` (the
    // submit lock) is therefore only ever taken for a few host instructions by them; allocations, uploads and frees go
    // through the stream-ordered allocator on `stream_admin` under `mu_cfg`, and a destroyed channel's memory, slot and
    // result region are recycled only after an event proves that the blocks in flight at destroy time have drained.
    std::mutex mu_cfg;                  // taps pool, slot / result-region allocators, graveyard; never held across a device sync
    std::vector<char> slot_used;        // by slot: owned by a live or not-yet-reaped channel
    std::vector<std::pair<size_t, size_t>> res_free;   // free extents (offset, bytes) of the result buffer below res_hwm
    size_t res_hwm = 0;                 // layout high-water mark; the buffers grow (the one remaining sync) if it passes results_bytes
    std::vector<sdr_chan *> graveyard;
    ChanStatic *h_cs_stage = nullptr;   // pinned [kMaxChan]: descriptors re-uploaded by the submitter (cs_dirty)
#ifndef SDR_EMU
    cudaStream_t stream_admin = nullptr;
    cudaStream_t stream_h2d = nullptr;
    cudaEvent_t ev_h2d[kInRing] = {nullptr, nullptr, nullptr}, ev_in_free[kInRing] = {nullptr, nullptr, nullptr};
#endif
    bool results_on_host = false;
    // host results live in a ring of kResBuf buffers so that a caller may fetch block b while up to kResBuf - 1 newer
    // blocks are in flight (sdr_engine_wait_block)
    int res_sel = 0, fetch_sel = 0;
    int last_res_q = -1;                 // host buffer of the previous to_host block (its D2H reads d_points / d_results)
    bool res_on_host[kResBuf] = {false, false, false};
    unsigned char *hres(int q) const { return h_results + (size_t)q * results_bytes; }
#ifndef SDR_EMU
    cudaEvent_t ev_res_tail[kResBuf] = {nullptr, nullptr, nullptr}, ev_res_spec[kResBuf] = {nullptr, nullptr, nullptr};
    bool res_tail_rec[kResBuf] = {false, false, false}, res_spec_rec[kResBuf] = {false, false, false};
#endif
    std::vector<sdr_spectrum *> spectra;
    // optional per-stage device timing (CUDA events on the engine stream): 0 dc, 1 spectrum, 2 front, 3 tail
    bool timing = false;
    static const int kTimeSets = 64;
    int time_pending = 0;
    double stage_ms[4] = {0, 0, 0, 0};
    float timeline[16][8];
    int timeline_n = 0;
    unsigned long long stage_count = 0;
#ifndef SDR_EMU
    cudaEvent_t tev[kTimeSets][5];
    cudaEvent_t tev_spec[kTimeSets][2];
    bool tev_made = false;
    // the spectrum kernels run on a side stream, concurrently with the channel kernels (both only read the block)
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // the DC blocker of block b+1 runs on its own stream underneath the channel kernels of block b
    // the channel tail of block b runs on its own stream underneath the front kernel of block b+1
    cudaStream_t stream_tail = nullptr;
    cudaStream_t stream_res = nullptr;   // device -> host copies of results, off the compute streams
    cudaEvent_t ev_spec_k = nullptr, ev_tail_k = nullptr;   // spectrum / tail kernels of the current block are done
    cudaEvent_t ev_front_done = nullptr, ev_tail_done[2] = {nullptr, nullptr};
    cudaEvent_t tev_tail[kTimeSets][2];
    cudaStream_t stream_dc = nullptr;
    cudaEvent_t ev_input = nullptr, ev_dc_done[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    cudaEvent_t ev_params[2] = {nullptr, nullptr};   // per-block parameter blocks of this parity are on the device
    cudaEvent_t tev_dc[kTimeSets][2];
#endif
};

// several engines of one process may sit on different GPUs (sdr_group): every entry point that issues CUDA work first makes
// its engine's device current
static inline void use_device(const sdr_engine *e) {
#ifndef SDR_EMU
    cudaSetDevice(e->device);
#else
    (void)e;
#endif
}

static void time_mark(sdr_engine *e, int which) {
#ifndef SDR_EMU
    if (e->timing && e->time_pending < sdr_engine::kTimeSets) cudaEventRecord(e->tev[e->time_pending][which], e->stream);
#else
    (void)e; (void)which;
#endif
}
static void time_collect(sdr_engine *e) {
#ifndef SDR_EMU
    // diagnostics: absolute timeline (ms since the first block's DC start) of the first 16 blocks of this batch
    e->timeline_n = 0;
    for (int i = 0; i < e->time_pending && i < 16; ++i) {
        float t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        cudaEventElapsedTime(&t[0], e->tev_dc[0][0], e->tev_dc[i][0]);
        cudaEventElapsedTime(&t[1], e->tev_dc[0][0], e->tev_dc[i][1]);
        cudaEventElapsedTime(&t[2], e->tev_dc[0][0], e->tev[i][2]);
        cudaEventElapsedTime(&t[3], e->tev_dc[0][0], e->tev[i][3]);
        cudaEventElapsedTime(&t[4], e->tev_dc[0][0], e->tev_tail[i][1]);
        cudaEventElapsedTime(&t[5], e->tev_dc[0][0], e->tev_spec[i][0]);     // may fail when no spectrum is attached: stays 0
        cudaEventElapsedTime(&t[6], e->tev_dc[0][0], e->tev_spec[i][1]);
        cudaEventElapsedTime(&t[7], e->tev_dc[0][0], e->tev_tail[i][0]);
        cudaGetLastError();
        for (int k = 0; k < 8; ++k) e->timeline[i][k] = t[k];
        e->timeline_n = i + 1;
    }
    for (int i = 0; i < e->time_pending; ++i) {
        for (int k = 0; k < 4; ++k) {
            float ms = 0.f;
            if (k == 0) { if (cudaEventElapsedTime(&ms, e->tev_dc[i][0], e->tev_dc[i][1]) == cudaSuccess) e->stage_ms[k] += ms; continue; }
            if (k == 1) { if (cudaEventElapsedTime(&ms, e->tev_spec[i][0], e->tev_spec[i][1]) == cudaSuccess) e->stage_ms[k] += ms; continue; }
            if (k == 3) { if (cudaEventElapsedTime(&ms, e->tev_tail[i][0], e->tev_tail[i][1]) == cudaSuccess) e->stage_ms[k] += ms; continue; }
            if (cudaEventElapsedTime(&ms, e->tev[i][k], e->tev[i][k + 1]) == cudaSuccess) e->stage_ms[k] += ms;
        }
        e->stage_count++;
    }
#endif
    e->time_pending = 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// taps pool
// ---------------------------------------------------------------------------------------------------------------------
static int pool_add(sdr_engine *e, const std::string &key, const float *v, size_t n) {
    auto it = e->taps_cache.find(key);
    if (it != e->taps_cache.end()) return it->second;
    int off = (int)e->taps_host.size();
    e->taps_host.insert(e->taps_host.end(), v, v + n);
    while (e->taps_host.size() % 4) e->taps_host.push_back(0.f);
    e->taps_cache[key] = off;
    return off;
}
static std::string fkey(const char *tag, float v, int extra = 0) {
    char b[64];
    uint32_t u; memcpy(&u, &v, 4);
    snprintf(b, sizeof b, "%s:%08x:%d", tag, u, extra);
    return b;
}

// ---------------------------------------------------------------------------------------------------------------------
// C-ABI: misc
// ---------------------------------------------------------------------------------------------------------------------
extern "C" const char *sdr_last_error(void) { return g_err.c_str(); }
extern "C" const char *sdr_version(void) { return "cubicsdr_b200 0.1 (sm_100a)"; }
extern "C" int sdr_is_device_build(void) {
#ifdef SDR_EMU
    return 0;
#else
    return 1;
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int sdr_engine_create(int device, unsigned max_block, void *stream, sdr_engine **out) {
    if (!out || max_block == 0) return fail(SDR_ERR_ARG, "sdr_engine_create: bad arguments");
#ifndef SDR_EMU
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(SDR_ERR_CUDA, "sdr_engine_create: no CUDA device (this library has no CPU path)");
    CUDA_TRY(cudaSetDevice(device));
#endif
    sdr_engine *e = new sdr_engine();
    e->device = device;
    e->max_block = max_block;
    e->chan_block_max = max_block;
#ifndef SDR_EMU
    if (stream) { e->stream = (cudaStream_t)stream; e->own_stream = false; }
    else { CUDA_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)); e->own_stream = true; }
#else
    (void)stream;
#endif
    // history kept in front of the block: the DC replay's warm-up, and a zoomed spectrum view's frame (up to one block long)
    // history kept in front of a block: the spectrum frames that straddle blocks; a zoomed view's frames are the transform length
    // times its decimation (262144 samples for a 16384-point transform zoomed 16x), however short the blocks are
    e->hist_cap = (std::max<size_t>((size_t)1 << 19, max_block) + kBlkHist + 1) & ~(size_t)1;
    int bad = 0;
    bad |= be_malloc((void **)&e->d_raw, (size_t)max_block * sizeof(float2));
    e->raw_pad = (size_t)max_block / 8 + 2 * 65536;
    // (padded like the raw buffer: the TMA store of the DC replay writes the last, partial segment as a whole row)
    for (int b = 0; b < 2; ++b) bad |= be_malloc((void **)&e->d_blkbuf[b], (e->hist_cap + max_block + e->raw_pad) * sizeof(float2));
    // the DC replay's lanes past the end of a block still stage (unused) samples of their streams: pad instead of guarding
    e->raw_pad = (size_t)max_block / 8 + 2 * 65536;
    bad |= be_malloc((void **)&e->d_rawbuf, (e->hist_cap + max_block + e->raw_pad) * sizeof(float2));
    bad |= be_malloc((void **)&e->d_scratch, e->hist_cap * sizeof(float2));
    for (int k = 0; k < sdr_engine::kInRing; ++k) {
        bad |= be_host_alloc((void **)&e->h_in[k], (size_t)max_block * sizeof(sdr_cf32));
#ifndef SDR_EMU
        bad |= be_malloc((void **)&e->d_in[k], (size_t)max_block * sizeof(float2));
#endif
    }
    const int tiles = (max_block + kDcTile - 1) / kDcTile;
    bad |= be_malloc((void **)&e->d_tile_sum, tiles * sizeof(double2));
    for (int b = 0; b < 2; ++b) bad |= be_malloc((void **)&e->d_tile_carry2[b], tiles * sizeof(double2));
    const int max_seg = (int)(max_block / kDcTile) + 2;
    bad |= be_malloc((void **)&e->d_seg_in, max_seg * sizeof(float2));
    bad |= be_malloc((void **)&e->d_seg_out, max_seg * sizeof(float2));
    bad |= be_malloc((void **)&e->d_fstate, sizeof(float2));
    bad |= be_malloc((void **)&e->d_fixups, sizeof(unsigned));
    {
        const size_t rows = (size_t)max_seg * 2 * kDwCand;
        const size_t chunks = (size_t)max_seg / kDwChunk + 2;
        bad |= be_malloc((void **)&e->d_dw_st, rows * sizeof(float));
        bad |= be_malloc((void **)&e->d_dw_pst, rows * sizeof(float));
        bad |= be_malloc((void **)&e->d_dw_gmap, chunks * 2 * kDwCand * sizeof(short));
        bad |= be_malloc((void **)&e->d_dw_entry, chunks * 2 * sizeof(int));
        bad |= be_malloc((void **)&e->d_dw_stats, 4 * sizeof(unsigned));
        bad |= be_malloc((void **)&e->d_dc_fix_mask, (size_t)max_seg * sizeof(int));
        const char *w = getenv("SDR_B200_DC_WINDOW");
        if (w) e->dc_window = atoi(w) ? 1 : 0;
    }
    { const char *m = getenv("SDR_B200_DC"); if (m && std::string(m) == "scan") e->exact_dc = false; }
    bad |= be_malloc((void **)&e->d_dc_state, sizeof(double2));
    bad |= be_malloc((void **)&e->d_apow, (kDcTile + 1) * sizeof(double));
    bad |= be_malloc((void **)&e->d_nco, 1024 * sizeof(float2));

    bad |= be_malloc((void **)&e->d_cs, sdr_engine::kMaxChan * sizeof(ChanStatic));
    for (int b = 0; b < 2; ++b) bad |= be_malloc((void **)&e->d_cd2[b], sdr_engine::kMaxChan * sizeof(ChanDyn));
    bad |= be_malloc((void **)&e->d_st, sdr_engine::kMaxChan * sizeof(ChanState));
    bad |= be_host_alloc((void **)&e->h_cd, (size_t)sdr_engine::kRing * sdr_engine::kMaxChan * sizeof(ChanDyn));
    bad |= be_malloc((void **)&e->d_group, 2 * sdr_engine::kMaxChan * sizeof(int));   // double-buffered by block parity
    bad |= be_host_alloc((void **)&e->h_group, (size_t)sdr_engine::kRing * sdr_engine::kMaxChan * sizeof(int));
#ifndef SDR_EMU
    for (int i = 0; i < sdr_engine::kRing; ++i) CUDA_TRY(cudaEventCreateWithFlags(&e->ring_ev[i], cudaEventDisableTiming));
    {   // highest priority: the DC replay's CTAs should slot in the moment an SM is free.  Next: the spectrum, which the
        // next block's front kernel waits for (join): its FFT CTAs must get their SMs before the 512 tail CTAs that become
        // runnable at the same instant fill the GPU (nothing waits for the tail until a block later)
        int lo = 0, hi = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&e->stream_dc, cudaStreamNonBlocking, hi));
        CUDA_TRY(cudaStreamCreateWithPriority(&e->stream2, cudaStreamNonBlocking, hi < lo ? hi + 1 : hi));
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&e->stream_tail, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&e->stream_res, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&e->stream_admin, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&e->stream_h2d, cudaStreamNonBlocking));
    for (int k = 0; k < sdr_engine::kInRing; ++k) {
        CUDA_TRY(cudaEventCreateWithFlags(&e->ev_h2d[k], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&e->ev_in_free[k], cudaEventDisableTiming));
    }
    {   // channel memory comes from the stream-ordered pool: keep what is freed cached instead of returning it to the driver
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_spec_k, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_tail_k, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_front_done, cudaEventDisableTiming));
    for (int b = 0; b < 2; ++b) CUDA_TRY(cudaEventCreateWithFlags(&e->ev_tail_done[b], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_input, cudaEventDisableTiming));
    for (int b = 0; b < 2; ++b) {
        CUDA_TRY(cudaEventCreateWithFlags(&e->ev_dc_done[b], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&e->ev_params[b], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&e->ev_consumed[b], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    for (int b = 0; b < kResBuf; ++b) {
        CUDA_TRY(cudaEventCreateWithFlags(&e->ev_res_tail[b], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&e->ev_res_spec[b], cudaEventDisableTiming));
    }
#endif
    e->taps_cap = 4u << 20;   // floats
    bad |= be_malloc((void **)&e->d_taps, e->taps_cap * sizeof(float));
    e->slot_used.assign(sdr_engine::kMaxChan, 0);
    bad |= be_host_alloc((void **)&e->h_cs_stage, (size_t)sdr_engine::kMaxChan * sizeof(ChanStatic));
    e->results_bytes = (size_t)16 << 20;                        // grows (ensure_results) if a layout ever needs more
    bad |= be_malloc((void **)&e->d_results, e->results_bytes);
    bad |= be_host_alloc((void **)&e->h_results, kResBuf * e->results_bytes);
    if (bad) { delete e; return fail(SDR_ERR_NOMEM, "sdr_engine_create: allocation failed"); }
    for (int b = 0; b < 2; ++b) be_memset(e->d_blkbuf[b], 0, (e->hist_cap + max_block + e->raw_pad) * sizeof(float2), e->stream);
    be_memset(e->d_rawbuf, 0, (e->hist_cap + max_block + e->raw_pad) * sizeof(float2), e->stream);
    be_memset(e->d_fstate, 0, sizeof(float2), e->stream);
    be_memset(e->d_fixups, 0, sizeof(unsigned), e->stream);
    if (!bad) be_memset(e->d_dw_stats, 0, 4 * sizeof(unsigned), e->stream);
    be_memset(e->d_dc_state, 0, sizeof(double2), e->stream);
    // DC blocker coefficient: a = -(-1 + alpha) with alpha = 0.0005f, evaluated in float as the reference does
    const float a_f = 1.0f - 0.0005f;
    e->dc_a = (double)a_f;
    std::vector<double> apow(kDcTile + 1);
    apow[0] = 1.0;
    for (int k = 1; k <= kDcTile; ++k) apow[k] = apow[k - 1] * e->dc_a;
    be_h2d(e->d_apow, apow.data(), apow.size() * sizeof(double), e->stream);
    float tab[1024];
    nco_sine_table(tab);
    std::vector<float2> nco(1024);
    for (int i = 0; i < 1024; ++i) nco[i] = make_float2(tab[(i + 256) & 1023], tab[i]);
    be_h2d(e->d_nco, nco.data(), 1024 * sizeof(float2), e->stream);
    // entry 0 must be exactly (1, 0): "no mixing" is expressed as a constant phase of 0 (front_span)
    if (nco[0].x != 1.0f || nco[0].y != 0.0f) { delete e; return fail(SDR_ERR_STATE, "oscillator table entry 0 is not (1, 0)"); }

    be_sync(e->stream);
    *out = e;
    return SDR_OK;
}

static void spectrum_free(sdr_spectrum *sp);
static void sync_all(sdr_engine *e);
static void reap_graveyard(sdr_engine *e, bool wait);
static void chan_release_memory(sdr_engine *e, sdr_chan *c, sdr_stream_t s);
static sdr_stream_t admin_stream(sdr_engine *e);

extern "C" int sdr_engine_destroy(sdr_engine *e) {
    if (!e) return SDR_OK;
#ifndef SDR_EMU
    cudaSetDevice(e->device);
#endif
    sync_all(e);
#ifndef SDR_EMU
    if (e->stream_admin) cudaStreamSynchronize(e->stream_admin);
#endif
    reap_graveyard(e, true);
    for (auto *c : e->chans) if (c) {
        chan_release_memory(e, c, admin_stream(e));
        be_host_free(c->h_dyn); be_host_free(c->h_slot);
#ifndef SDR_EMU
        if (c->ev_ready) cudaEventDestroy(c->ev_ready);
#endif
        delete c;
    }
    be_host_free(e->h_cs_stage);
    for (auto *s : e->spectra) if (s) { spectrum_free(s); delete s; }
    be_free(e->d_pfb_hq); be_free(e->d_pfb_tw); be_free(e->d_pfb_streams[0]); be_free(e->d_pfb_streams[1]); be_free(e->d_pfb_dc_state);
    be_free(e->d_raw); be_free(e->d_blkbuf[0]); be_free(e->d_blkbuf[1]); be_free(e->d_rawbuf); be_free(e->d_scratch);
    for (int k = 0; k < sdr_engine::kInRing; ++k) { be_host_free(e->h_in[k]); be_free(e->d_in[k]); }
    be_free(e->d_tile_sum); be_free(e->d_tile_carry2[0]); be_free(e->d_tile_carry2[1]); be_free(e->d_dc_state);
    be_free(e->d_dc_fix_mask);
    be_free(e->d_pcm); be_host_free(e->h_pcm);
#ifndef SDR_EMU
    if (e->ev_pcm) cudaEventDestroy(e->ev_pcm);
#endif
    be_free(e->d_dw_st); be_free(e->d_dw_pst); be_free(e->d_dw_gmap); be_free(e->d_dw_entry); be_free(e->d_dw_stats);
    be_free(e->d_seg_in); be_free(e->d_seg_out); be_free(e->d_fstate); be_free(e->d_fixups); be_free(e->d_apow); be_free(e->d_nco);
    be_free(e->d_cs); be_free(e->d_cd2[0]); be_free(e->d_cd2[1]); be_free(e->d_st); be_host_free(e->h_cd); be_free(e->d_group); be_host_free(e->h_group);
    be_free(e->d_taps); be_free(e->d_results); be_host_free(e->h_results);
#ifndef SDR_EMU
    for (int i = 0; i < sdr_engine::kRing; ++i) cudaEventDestroy(e->ring_ev[i]);
    if (e->stream2) { cudaStreamSynchronize(e->stream2); cudaStreamDestroy(e->stream2); }
    if (e->stream_dc) { cudaStreamSynchronize(e->stream_dc); cudaStreamDestroy(e->stream_dc); }
    if (e->stream_tail) { cudaStreamSynchronize(e->stream_tail); cudaStreamDestroy(e->stream_tail); }
    if (e->stream_res) { cudaStreamSynchronize(e->stream_res); cudaStreamDestroy(e->stream_res); }
    if (e->stream_admin) { cudaStreamSynchronize(e->stream_admin); cudaStreamDestroy(e->stream_admin); }
    if (e->stream_h2d) { cudaStreamSynchronize(e->stream_h2d); cudaStreamDestroy(e->stream_h2d); }
    for (int k = 0; k < sdr_engine::kInRing; ++k) { if (e->ev_h2d[k]) cudaEventDestroy(e->ev_h2d[k]); if (e->ev_in_free[k]) cudaEventDestroy(e->ev_in_free[k]); }
    if (e->ev_spec_k) cudaEventDestroy(e->ev_spec_k);
    if (e->ev_tail_k) cudaEventDestroy(e->ev_tail_k);
    if (e->ev_front_done) cudaEventDestroy(e->ev_front_done);
    for (int b = 0; b < 2; ++b) if (e->ev_tail_done[b]) cudaEventDestroy(e->ev_tail_done[b]);
    if (e->ev_input) cudaEventDestroy(e->ev_input);
    for (int b = 0; b < 2; ++b) { if (e->ev_dc_done[b]) cudaEventDestroy(e->ev_dc_done[b]); if (e->ev_consumed[b]) cudaEventDestroy(e->ev_consumed[b]); if (e->ev_params[b]) cudaEventDestroy(e->ev_params[b]); }
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    for (int b = 0; b < kResBuf; ++b) { if (e->ev_res_tail[b]) cudaEventDestroy(e->ev_res_tail[b]); if (e->ev_res_spec[b]) cudaEventDestroy(e->ev_res_spec[b]); }
    if (e->own_stream) cudaStreamDestroy(e->stream);
#endif
    delete e;
    return SDR_OK;
}

extern "C" unsigned long long sdr_engine_launch_count(sdr_engine *e) { return e ? e->launches : 0; }

// diagnostics: {dc_start, dc_end, front_start, front_end, tail_end} in ms for up to 16 blocks of the last timed batch
extern "C" int sdr_engine_debug_timeline(sdr_engine *e, float *out128, int *n_blocks) {   // [16][8], see time_collect
    float *out80 = out128;
    if (!e || !out80) return SDR_ERR_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    memcpy(out80, e->timeline, sizeof e->timeline);
    if (n_blocks) *n_blocks = e->timeline_n;
    return SDR_OK;
}

// diagnostics: number of segment junctions the bit-exact DC replay had to recompute serially so far
extern "C" int sdr_engine_debug_dc_fixups(sdr_engine *e, unsigned *fixups, int *exact_mode) {
    if (!e) return SDR_ERR_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    be_sync(e->stream);
    unsigned v = 0;
    be_d2h(&v, e->d_fixups, sizeof v, e->stream);
    be_sync(e->stream);
    if (fixups) *fixups = v;
    if (exact_mode) *exact_mode = e->exact_dc ? 1 : 0;
    return SDR_OK;
}
// diagnostics: candidate-window replay -- stretches its walker had to run serially so far; blocks that took the window path
// diagnostics: tiles the warm-up replay re-ran in its parallel fix-up rounds so far
extern "C" int sdr_engine_debug_dc_reruns(sdr_engine *e, unsigned *reruns) {
    if (!e || !reruns) return SDR_ERR_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    be_sync(e->stream);
    unsigned v[4] = {0, 0, 0, 0};
    be_d2h(v, e->d_dw_stats, sizeof v, e->stream);
    be_sync(e->stream);
    *reruns = v[2];
    return SDR_OK;
}
extern "C" int sdr_engine_debug_dc_window(sdr_engine *e, unsigned *serial_stretches, unsigned long long *window_blocks) {
    if (!e) return SDR_ERR_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    be_sync(e->stream);
    unsigned v = 0;
    be_d2h(&v, e->d_dw_stats, sizeof v, e->stream);
    be_sync(e->stream);
    if (serial_stretches) *serial_stretches = v;
    if (window_blocks) *window_blocks = e->dw_blocks;
    return SDR_OK;
}
// select the DC replay: -1 automatic (windows unless the channel load hides a warm-up replay), 0 warm-up replay, 1 windows
extern "C" int sdr_engine_set_dc_window(sdr_engine *e, int mode) {
    if (!e) return SDR_ERR_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    e->dc_window = mode < 0 ? -1 : (mode ? 1 : 0);
    return SDR_OK;
}

// diagnostics: per-CTA {smid, t_start, t_after_table, t_end} of the LAST front-kernel launch (not part of the public header)
extern "C" int sdr_engine_debug_front_timeline(sdr_engine *e, int enable, unsigned long long *out, unsigned max_ctas) {
    if (!e) return SDR_ERR_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    be_sync(e->stream);
    if (enable && !e->d_front_dbg) {
        if (be_malloc((void **)&e->d_front_dbg, (4 * 1024 + 64 * 64) * sizeof(unsigned long long))) return SDR_ERR_NOMEM;
        be_memset(e->d_front_dbg, 0, (4 * 1024 + 64 * 64) * sizeof(unsigned long long), e->stream);
    }
    if (out && e->d_front_dbg) {
        be_d2h(out, e->d_front_dbg, (size_t)4 * std::min(max_ctas, 1024u) * sizeof(unsigned long long), e->stream);
        be_sync(e->stream);
        be_memset(e->d_front_dbg, 0, (4 * 1024 + 64 * 64) * sizeof(unsigned long long), e->stream);
    }
    if (!enable && e->d_front_dbg) { be_free(e->d_front_dbg); e->d_front_dbg = nullptr; }
    return SDR_OK;
}

static void sync_all(sdr_engine *e);
// diagnostics: globaltimer at the end of every tail-kernel phase for the first 64 channels ([64][64], slot 63 = start);
// enabled together with sdr_engine_debug_front_timeline(e, 1, ...)
extern "C" int sdr_engine_debug_tail_phases(sdr_engine *e, unsigned long long *out, int *n_phases) {
    if (!e || !out || !e->d_front_dbg) return SDR_ERR_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    sync_all(e);
    be_d2h(out, e->d_front_dbg + 4 * 1024, (size_t)64 * 64 * sizeof(unsigned long long), e->stream);
    be_sync(e->stream);
    if (n_phases) *n_phases = TP_COUNT;
    return SDR_OK;
}

extern "C" int sdr_engine_set_timing(sdr_engine *e, int enable) {
    if (!e) return fail(SDR_ERR_ARG, "null engine");
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
#ifndef SDR_EMU
    if (enable && !e->tev_made) {
        for (int i = 0; i < sdr_engine::kTimeSets; ++i) for (int k = 0; k < 5; ++k) CUDA_TRY(cudaEventCreate(&e->tev[i][k]));
        for (int i = 0; i < sdr_engine::kTimeSets; ++i) for (int k = 0; k < 2; ++k) CUDA_TRY(cudaEventCreate(&e->tev_spec[i][k]));
        for (int i = 0; i < sdr_engine::kTimeSets; ++i) for (int k = 0; k < 2; ++k) CUDA_TRY(cudaEventCreate(&e->tev_dc[i][k]));
        for (int i = 0; i < sdr_engine::kTimeSets; ++i) for (int k = 0; k < 2; ++k) CUDA_TRY(cudaEventCreate(&e->tev_tail[i][k]));
        e->tev_made = true;
    }
#endif
    e->timing = enable != 0;
    e->time_pending = 0;
    for (int k = 0; k < 4; ++k) e->stage_ms[k] = 0.0;
    e->stage_count = 0;
    return SDR_OK;
}

extern "C" int sdr_engine_get_timing(sdr_engine *e, double *stage_ms4, unsigned long long *blocks) {
    if (!e) return fail(SDR_ERR_ARG, "null engine");
    std::lock_guard<std::mutex> lk(e->mu);
    if (stage_ms4) for (int k = 0; k < 4; ++k) stage_ms4[k] = e->stage_ms[k];
    if (blocks) *blocks = e->stage_count;
    return SDR_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// channels
// ---------------------------------------------------------------------------------------------------------------------
static int add_halfband_taps(sdr_engine *e, int m) {
    float h1[2 * 10];
    if (!halfband_branch_taps(m, 65.0f, h1)) return -1;
    return pool_add(e, fkey("hb", (float)m), h1, 2 * m);
}

static int ensure_results(sdr_engine *e);
static void sync_all(sdr_engine *e) {
    be_sync(e->stream);
#ifndef SDR_EMU
    cudaStreamSynchronize(e->stream_tail);
    cudaStreamSynchronize(e->stream_res);
    cudaStreamSynchronize(e->stream_dc);
    cudaStreamSynchronize(e->stream2);
    if (e->stream_h2d) cudaStreamSynchronize(e->stream_h2d);
#endif
}

// ---- allocators of the configuration side (all under e->mu_cfg) --------------------------------------------------------------
static sdr_stream_t admin_stream(sdr_engine *e) {
#ifndef SDR_EMU
    return e->stream_admin;
#else
    return e->stream;
#endif
}
static void chan_release_memory(sdr_engine *e, sdr_chan *c, sdr_stream_t s) {
    if (c->async_mem) { be_free_async(c->d_arena, s); be_free_async(c->d_priv_hist, s); }
    else { be_free(c->d_arena); be_free(c->d_priv_hist); }
    c->d_arena = nullptr; c->d_priv_hist = nullptr;
    (void)e;
}
// result-buffer regions: first fit over the free extents, else extend the high-water mark
static size_t res_region_take(sdr_engine *e, size_t bytes) {
    for (size_t k = 0; k < e->res_free.size(); ++k) {
        if (e->res_free[k].second >= bytes) {
            const size_t off = e->res_free[k].first;
            e->res_free[k].first += bytes; e->res_free[k].second -= bytes;
            if (!e->res_free[k].second) e->res_free.erase(e->res_free.begin() + (long)k);
            return off;
        }
    }
    const size_t off = e->res_hwm;
    e->res_hwm += bytes;
    return off;
}
static void res_region_give(sdr_engine *e, size_t off, size_t bytes) {
    auto it = e->res_free.begin();
    while (it != e->res_free.end() && it->first < off) ++it;
    it = e->res_free.insert(it, std::make_pair(off, bytes));
    if (it + 1 != e->res_free.end() && it->first + it->second == (it + 1)->first) { it->second += (it + 1)->second; e->res_free.erase(it + 1); }
    if (it != e->res_free.begin() && (it - 1)->first + (it - 1)->second == it->first) { (it - 1)->second += it->second; it = e->res_free.erase(it) - 1; }
    if (it->first + it->second == e->res_hwm) { e->res_hwm = it->first; e->res_free.erase(it); }
}
// recycle what destroyed channels held once their event has fired (never waits)
static void reap_graveyard(sdr_engine *e, bool wait) {
    for (size_t k = 0; k < e->graveyard.size();) {
        sdr_chan *g = e->graveyard[k];
        bool done = true;
#ifndef SDR_EMU
        if (g->ev_dead) {
            if (wait) cudaEventSynchronize(g->ev_dead);
            done = cudaEventQuery(g->ev_dead) == cudaSuccess;
        }
#endif
        (void)wait;
        if (!done) { ++k; continue; }
        if (g->index >= 0 && (size_t)g->index < e->slot_used.size()) e->slot_used[(size_t)g->index] = 0;
        if (g->slot_bytes) res_region_give(e, (size_t)g->cs.res_off, g->slot_bytes);
#ifndef SDR_EMU
        if (g->ev_dead) cudaEventDestroy(g->ev_dead);
        if (g->ev_ready) cudaEventDestroy(g->ev_ready);
#endif
        delete g;
        e->graveyard.erase(e->graveyard.begin() + (long)k);
    }
}

// DemodulatorWorkerThread builds the resampler for a new demodulator off the hot threads (DemodulatorWorkerThread.cpp:63-108);
// so does this: design and layout without any lock, allocation / clearing / uploads on the admin stream under mu_cfg, and the
// submit lock only to publish the finished channel.
// view_cap != 0: the hidden channel of a zoomed spectrum view (SpectrumVisualProcessor's own msresamp_crcf): no modem, every
// half-band stage in the tail kernel (its level-0 stream is written by the view mixer, not by the front kernel), buffers for
// view_cap samples per block; called from a submit, i.e. with e->mu already held.
static int chan_create_impl(sdr_engine *e, const char *modem, long long sample_rate, long long bandwidth, int audio_rate,
                            unsigned inject_cap, sdr_chan **out, unsigned view_cap = 0) {
    if (!e || !out || sample_rate <= 0 || audio_rate <= 0) return fail(SDR_ERR_ARG, "sdr_chan_create: bad arguments");
    const int mk = view_cap ? (int)MODEM_NONE : modem_from_name(modem);
    if (mk < 0) return fail(SDR_ERR_UNSUPPORTED, std::string("sdr_chan_create: unknown modem '") + (modem ? modem : "") + "'");
    sdr_chan *c = new sdr_chan();
    c->eng = e; c->index = -1; c->modem = mk; c->sample_rate = sample_rate;
    c->bandwidth = check_sample_rate(mk, bandwidth);
    c->asked_bandwidth = (int)bandwidth;
    if (view_cap) c->bandwidth = (int)bandwidth;                        // a power-of-two fraction of the rate, as it comes
    if (mk == MODEM_IQ) c->bandwidth = audio_rate;                      // ModemIQ::checkSampleRate returns audioSampleRate (ModemIQ.cpp:31-33)
    c->audio_rate = audio_rate;
    const float As = 60.0f;                                             // DemodulatorWorkerThread.cpp:95, ModemAnalog.cpp:25
    const double iq_ratio = (double)c->bandwidth / (double)sample_rate; // :97-100
    if (!(iq_ratio > 0.004 / 65536.0) || iq_ratio > 1.0) { delete c; return fail(SDR_ERR_ARG, "sdr_chan_create: bandwidth/sample_rate out of range"); }
    c->iq_plan = make_resampler_plan((float)iq_ratio, As);
    const double a_ratio = double(audio_rate) / double(c->bandwidth);   // ModemAnalog.cpp:29-30
    c->audio_plan = make_resampler_plan((float)a_ratio, As);
    if (c->iq_plan.interp) { delete c; return fail(SDR_ERR_UNSUPPORTED, "sdr_chan_create: bandwidth above the sample rate"); }
    if (mk == MODEM_CW && !c->audio_plan.interp) { delete c; return fail(SDR_ERR_UNSUPPORTED, "sdr_chan_create: CW bandwidth above the audio rate"); }
    if (mk == MODEM_FMS && c->audio_plan.interp) { delete c; return fail(SDR_ERR_UNSUPPORTED, "sdr_chan_create: FM stereo with an audio rate above the modem rate"); }
    if (c->iq_plan.num_halfband > kMaxTailHb + kFrontMaxK || c->audio_plan.num_halfband > kMaxAudioHb) {
        delete c; return fail(SDR_ERR_UNSUPPORTED, "sdr_chan_create: too many half-band stages");
    }
    ChanStatic &cs = c->cs;
    memset(&cs, 0, sizeof cs);
    cs.modem = mk;
    cs.S = c->iq_plan.num_halfband;
    cs.K1 = std::max(0, std::min(cs.S - 2, kFrontMaxK));               // only the m = 3 stages can be fused
    if (view_cap) cs.K1 = 0;
    cs.n_tail = cs.S - cs.K1;
    if (cs.n_tail > kMaxTailHb) { delete c; return fail(SDR_ERR_UNSUPPORTED, "sdr_chan_create: too many tail stages"); }
    for (int g = cs.S - cs.K1; g < cs.S; ++g)
        if (c->iq_plan.hb.m[g] != 3) { delete c; return fail(SDR_ERR_STATE, "front stage is not m = 3"); }
    cs.zeta = c->iq_plan.zeta;
    cs.arb_step = c->iq_plan.arb_step;
    cs.a_interp = c->audio_plan.interp ? 1 : 0;
    cs.a_S = c->audio_plan.num_halfband;
    cs.a_zeta = c->audio_plan.interp ? 1.0f : c->audio_plan.zeta;
    cs.a_step = c->audio_plan.arb_step;
    float am[51]; am_dc_block_taps(25, 30.0f, am);                      // ModemAM.cpp:9
    float hq[10]; hilbert_quadrature_taps(5, 90.0f, hq);                // ModemUSB.cpp:11
    float hq_cw[10]; hilbert_quadrature_taps(5, 60.0f, hq_cw);          // ModemCW.cpp:23
    std::vector<Sos> sos; butterworth_lowpass_sos(6, 0.25f, sos);       // ModemUSB.cpp:8
    float flat[15];
    for (int k = 0; k < 3; ++k) { flat[5 * k] = sos[k].b[0]; flat[5 * k + 1] = sos[k].b[1]; flat[5 * k + 2] = sos[k].b[2]; flat[5 * k + 3] = sos[k].a[1]; flat[5 * k + 4] = sos[k].a[2]; }
    // ---- arena layout (floats)
    size_t off = 0;
    auto take = [&](size_t floats) { size_t o = off; off += (floats + 3) & ~size_t(3); return (int)o; };
    const size_t nb = view_cap ? view_cap : (inject_cap ? inject_cap : e->chan_block_max);
    c->standalone = inject_cap != 0 || view_cap != 0;                   // neither is part of the block's channel set
    c->inject_cap = inject_cap;
    for (int j = 0; j <= cs.n_tail; ++j) cs.lvl_off[j] = take(2 * (kHist + (nb >> (cs.K1 + j)) + 4));
    cs.lvl0_alt_off = take(2 * (kHist + (nb >> cs.K1) + 4));
    const size_t cap_z = (nb >> cs.S) + 4;
    const size_t cap_iq = (size_t)std::ceil((double)cap_z * (double)c->iq_plan.rate_arb) + 8;
    cs.iq_off = take(2 * (kHist + cap_iq));
    cs.w_off = take(2 * (kHist + cap_iq));
    cs.d_off = take(kHist + cap_iq);
    size_t cap_audio;
    if (!cs.a_interp) {
        for (int j = 1; j <= cs.a_S; ++j) cs.al_off[j] = take(kHist + (cap_iq >> j) + 4);
        cap_audio = (size_t)std::ceil((double)((cap_iq >> cs.a_S) + 4) * (double)c->audio_plan.rate_arb) + 8;
    } else {
        const size_t cap_b = (size_t)std::ceil((double)cap_iq * (double)c->audio_plan.rate_arb) + 8;
        const size_t width = mk == MODEM_CW ? 2 : 1;                    // CW interpolates complex samples (msresamp_cccf)
        for (int j = 0; j <= cs.a_S; ++j) cs.al_off[j] = take(width * (kHist + (cap_b << j) + 4));
        cap_audio = cap_b << cs.a_S;
    }
    if (mk == MODEM_CW) {
        cs.cw_w_off = take(2 * (kHist + cap_audio + 4));
        // nco_crcf_set_frequency(mLO, 2.0f * (float)M_PI * mBeepFrequency / kit->audioSampleRate), mBeepFrequency = 650 (ModemCW.cpp:17,176)
        cs.cw_dtheta = nco_phase_step(2.0f * (float)3.14159265358979323846 * 650.0f / (float)audio_rate);
    }
    // FM stereo (ModemFMStereo.cpp:89-167): the second resampler's levels, the streams between its stages, the filters
    std::vector<float> fms_fir;
    float fms_sos_flat[25];
    if (mk == MODEM_FMS) {
        cs.fms_x2_off = take(2 * (kHist + cap_iq));
        cs.fms_ds_off = take(kHist + cap_iq);
        for (int j = 1; j <= cs.a_S; ++j) cs.fms_al_off[j] = take(kHist + (cap_iq >> j) + 4);
        cs.fms_am_off = take(cap_audio + 4);
        cs.fms_as_off = take(cap_audio + 4);
        cs.fms_ld_off = take(kFmsFirHist + cap_audio + 4);
        cs.fms_rd_off = take(kFmsFirHist + cap_audio + 4);
        float cutoff = 16000.0f / float(audio_rate);                   // :103-115
        const float ft = 1000.0f / float(audio_rate);
        if (cutoff < 0) cutoff = 0;
        if (cutoff > 0.5f) cutoff = 0.5f;
        const unsigned h_len = filter_len_estimate(ft, As);            // estimate_req_filter_len (:117)
        if (h_len < 1 || h_len > (unsigned)kFmsFirMax) { delete c; return fail(SDR_ERR_UNSUPPORTED, "sdr_chan_create: FM stereo low-pass too long for this audio rate"); }
        fms_fir.resize(h_len);
        kaiser_lowpass(h_len, cutoff, As, 0.0f, fms_fir.data());       // liquid_firdes_kaiser (:119)
        cs.fms_fir_n = (int)h_len;
        float bwf = float(c->bandwidth);                                // :126-136
        if (bwf < 100000.0) bwf = 100000.0;
        std::vector<Sos> ps;
        cheby2_bandpass5_sos((float)19500 / bwf, (float)19000 / bwf, ps);
        for (auto &sec : ps) sos_section_object(sec);                   // iirfilt_crcf_create_sos -> iirfiltsos_crcf_create
        for (int k = 0; k < 5; ++k) { fms_sos_flat[5 * k] = ps[k].b[0]; fms_sos_flat[5 * k + 1] = ps[k].b[1]; fms_sos_flat[5 * k + 2] = ps[k].b[2];
                                      fms_sos_flat[5 * k + 3] = ps[k].a[1]; fms_sos_flat[5 * k + 4] = ps[k].a[2]; }
        cs.fms_demph = 75;                                              // _demph = 75 (:8); other settings are a writeSetting away (:78-83)
        float b2[2], a1;
        fm_deemphasis(cs.fms_demph, audio_rate, b2, &a1);               // :149-159
        cs.fms_db0 = b2[0]; cs.fms_db1 = b2[1]; cs.fms_da1 = a1;
        cap_audio *= 2;                                                 // two-channel output {left, right}
    }
    if (mk == MODEM_IQ) cap_audio = 2 * cap_iq;                         // two-channel output, one {imag, real} pair per decimated sample
    cs.res_audio_cap = (int)((cap_audio + 3) & ~size_t(3));
    cs.res_iq_cap = (int)cap_iq;
    cs.want_iq = 0;
    // DemodulatorPreThread hands the demodulator thread `currentBandwidth` -- the bandwidth that was asked for, not what
    // Modem::checkSampleRate made of it (:222) -- and DemodulatorThread's sampleTime divides by it (:143)
    cs.bandwidth = c->asked_bandwidth > 0 ? c->asked_bandwidth : c->bandwidth;
    c->arena_floats = off;
    c->slot_bytes = kResHdrBytes + (size_t)cs.res_audio_cap * 4 + (size_t)cs.res_iq_cap * 8;
    c->slot_bytes = (c->slot_bytes + 31) & ~size_t(31);
    ChanState st;
    memset(&st, 0, sizeof st);
    st.agc_ceil = st.agc_ma = st.agc_maa = 1.0f;                        // ModemAnalog.cpp:6
    st.sig_level = -100.f; st.sig_floor = -30.f; st.sig_ceil = 30.f;    // DemodulatorThread.cpp:25-26
    if (inject_cap) {
        if (be_host_alloc((void **)&c->h_dyn, sizeof(ChanDyn)) || be_host_alloc((void **)&c->h_slot, c->slot_bytes)) {
            be_host_free(c->h_dyn); be_host_free(c->h_slot); delete c;
            return fail(SDR_ERR_NOMEM, "sdr_modem_create: host buffers");
        }
    }
    // ---- configuration lock: filter taps, slot, result region, device memory and uploads (admin stream; no device-wide sync)
    {
        std::lock_guard<std::mutex> lk(e->mu_cfg);
#ifndef SDR_EMU
        cudaSetDevice(e->device);                                       // DemodulatorWorkerThread is not the thread that made the engine
#endif
        reap_graveyard(e, false);
        sdr_stream_t as = admin_stream(e);
        auto bail = [&](int code, const char *msg) {
            chan_release_memory(e, c, as);
            if (c->index >= 0) e->slot_used[(size_t)c->index] = 0;
            be_host_free(c->h_dyn); be_host_free(c->h_slot);
#ifndef SDR_EMU
            if (c->ev_ready) cudaEventDestroy(c->ev_ready);
#endif
            delete c;
            return fail(code, msg);
        };
        for (int j = 0; j < cs.n_tail; ++j) {
            const int g = cs.S - cs.K1 - 1 - j;                         // cascade position executed at tail step j
            cs.tail_m[j] = c->iq_plan.hb.m[g];
            cs.tail_h1[j] = add_halfband_taps(e, cs.tail_m[j]);
            if (cs.tail_h1[j] < 0) return bail(SDR_ERR_UNSUPPORTED, "half-band prototype not available");
        }
        cs.pfb_off = pool_add(e, fkey("pfb", c->iq_plan.rate_arb), c->iq_plan.pfb.data(), c->iq_plan.pfb.size());
        for (int g = 0; g < cs.a_S; ++g) {
            cs.a_m[g] = c->audio_plan.hb.m[g];
            cs.a_h1[g] = add_halfband_taps(e, cs.a_m[g]);
            if (cs.a_h1[g] < 0) return bail(SDR_ERR_UNSUPPORTED, "half-band prototype not available");
        }
        cs.a_pfb_off = pool_add(e, fkey("pfb", c->audio_plan.rate_arb), c->audio_plan.pfb.data(), c->audio_plan.pfb.size());
        // firfilt runs y[t] = sum h[k] x[t-k]; the AM notch is symmetric so no reversal is needed
        cs.am_taps_off = pool_add(e, "am51", am, 51);
        cs.hq_off = pool_add(e, "hq10", hq, 10);
        cs.sos_off = pool_add(e, "sos6", flat, 15);
        cs.cw_hq_off = pool_add(e, "hq10_60", hq_cw, 10);
        if (mk == MODEM_FMS) {
            cs.fms_sos_off = pool_add(e, fkey("fms_sos", (float)c->bandwidth), fms_sos_flat, 25);
            cs.fms_fir_off = pool_add(e, fkey("fms_fir", (float)audio_rate), fms_fir.data(), fms_fir.size());
        }
        if (e->taps_host.size() > e->taps_cap) return bail(SDR_ERR_NOMEM, "taps pool exhausted");
        int slot = -1;
        for (size_t i = 0; i < e->slot_used.size(); ++i) if (!e->slot_used[i]) { slot = (int)i; break; }
        if (slot < 0) return bail(SDR_ERR_STATE, "sdr_chan_create: too many channels");
        e->slot_used[(size_t)slot] = 1;
        c->index = slot;
        c->async_mem = true;
        if (be_malloc_async((void **)&c->d_arena, off * sizeof(float), as)) return bail(SDR_ERR_NOMEM, "sdr_chan_create: arena");
        // [raw tail of the consumed stream | the same, mixed (rare path) | scratch for rolling the raw tail]
        if (be_malloc_async((void **)&c->d_priv_hist, 3 * (size_t)kBlkHist * sizeof(float2), as)) return bail(SDR_ERR_NOMEM, "sdr_chan_create: history");
        be_memset(c->d_arena, 0, off * sizeof(float), as);
        cs.arena = c->d_arena;
        cs.res_off = (long long)res_region_take(e, c->slot_bytes);
        // pageable sources: the runtime stages them before returning, so the locals may go out of scope
        be_h2d(e->d_st + slot, &st, sizeof st, as);
        be_h2d(e->d_cs + slot, &cs, sizeof cs, as);
        if (e->taps_host.size() > e->taps_uploaded) {
            be_h2d(e->d_taps + e->taps_uploaded, e->taps_host.data() + e->taps_uploaded,
                   (e->taps_host.size() - e->taps_uploaded) * sizeof(float), as);
            e->taps_uploaded = e->taps_host.size();
        }
#ifndef SDR_EMU
        if (cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming) != cudaSuccess) return bail(SDR_ERR_CUDA, "sdr_chan_create: event");
        cudaEventRecord(c->ev_ready, as);
        if (inject_cap) cudaEventSynchronize(c->ev_ready);              // sdr_modem_*: the caller's own thread, nobody else waits
#endif
    }
    if (view_cap) {                                                     // e->mu is held by the submit that makes the view
        if ((size_t)c->index >= e->chans.size()) e->chans.resize((size_t)c->index + 1, nullptr);
        e->chans[(size_t)c->index] = c;
    } else {   // publish: the submitter sees the channel from its next block on
        std::lock_guard<std::mutex> lk(e->mu);
        if ((size_t)c->index >= e->chans.size()) e->chans.resize((size_t)c->index + 1, nullptr);
        e->chans[(size_t)c->index] = c;
    }
    *out = c;
    return SDR_OK;
}

extern "C" int sdr_chan_create(sdr_engine *e, const char *modem, long long sample_rate, long long bandwidth, int audio_rate,
                               sdr_chan **out) {
    return chan_create_impl(e, modem, sample_rate, bandwidth, audio_rate, 0, out);
}

// ---- Modem plugin seam (S1): demodulate already-decimated IQ, one call = Modem::demodulate(kit, input, audioOut) ---------
extern "C" int sdr_modem_check_sample_rate(const char *modem, long long sample_rate) {
    const int mk = modem_from_name(modem);
    if (mk < 0) return fail(SDR_ERR_UNSUPPORTED, "unknown modem");
    return check_sample_rate(mk, sample_rate);
}
// the two-argument form of Modem::checkSampleRate(sampleRate, audioSampleRate): "I/Q" answers with the audio rate
extern "C" int sdr_modem_check_sample_rate2(const char *modem, long long sample_rate, int audio_rate) {
    const int mk = modem_from_name(modem);
    if (mk < 0) return fail(SDR_ERR_UNSUPPORTED, "unknown modem");
    return mk == MODEM_IQ ? audio_rate : check_sample_rate(mk, sample_rate);
}

extern "C" int sdr_modem_create(sdr_engine *e, const char *modem, long long sample_rate, int audio_rate, unsigned max_samples,
                                sdr_chan **out) {
    if (max_samples == 0) return fail(SDR_ERR_ARG, "sdr_modem_create: max_samples must be > 0");
    // the kit's input rate IS the modem rate: no IQ resampler in front (bandwidth == sample_rate after checkSampleRate)
    const int mk = modem_from_name(modem);
    if (mk < 0) return fail(SDR_ERR_UNSUPPORTED, "sdr_modem_create: unknown modem");
    const long long rate = check_sample_rate(mk, sample_rate);
    return chan_create_impl(e, modem, rate, rate, audio_rate, max_samples, out);
}

extern "C" int sdr_modem_demodulate(sdr_chan *c, const sdr_cf32 *iq, unsigned n, float *audio, unsigned cap, unsigned *n_audio,
                                    double *level_db, float *peak) {
    if (!c || !c->standalone) return fail(SDR_ERR_ARG, "sdr_modem_demodulate: not a modem handle");
    if (n_audio) *n_audio = 0;
    if (n == 0) return SDR_OK;                                          // ModemAnalog::initOutputBuffers: `if (!bufSize) return`
    if (!iq || n > c->inject_cap) return fail(SDR_ERR_ARG, "sdr_modem_demodulate: bad buffer / too many samples");
    sdr_engine *e = c->eng;
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    { int r = ensure_results(e); if (r) return r; }
    sync_all(e);                                                        // h_dyn / h_slot are reused call to call
    float *d_iq = c->cs.arena + c->cs.iq_off + 2 * kHist;
    be_h2d(d_iq, iq, (size_t)n * sizeof(sdr_cf32), e->stream);
    memset(c->h_dyn, 0, sizeof(ChanDyn));
    c->h_dyn->valid = 1; c->h_dyn->active = 1; c->h_dyn->inject_n = (int)n;
    be_h2d(e->d_cd2[0] + c->index, c->h_dyn, sizeof(ChanDyn), e->stream);
    TailArgs ta;
    ta.cs = e->d_cs; ta.cd = e->d_cd2[0]; ta.st = e->d_st; ta.taps = e->d_taps; ta.nco = e->d_nco;
    ta.results = e->d_results; ta.n = 0; ta.first = c->index; ta.view_pass = 0; ta.dbg = nullptr;
    launch_tail(ta, 1, e->stream, e->launches);
    be_d2h(c->h_slot, e->d_results + c->cs.res_off, c->slot_bytes, e->stream);
    if (be_sync(e->stream)) return fail(SDR_ERR_CUDA, std::string("sdr_modem_demodulate: ") + be_last_error());
    const ChanResultHeader *h = reinterpret_cast<const ChanResultHeader *>(c->h_slot);
    unsigned m = std::min(cap, std::min(h->n_audio, (unsigned)c->cs.res_audio_cap));
    if (audio) memcpy(audio, c->h_slot + kResHdrBytes, (size_t)m * sizeof(float));
    if (n_audio) *n_audio = h->n_audio;
    if (level_db) *level_db = h->level_db;
    if (peak) *peak = h->peak;
    return SDR_OK;
}

// DemodulatorMgr::deleteThread happens while blocks are in flight.  The channel disappears from the submitter's list at once
// (a few host instructions under the submit lock); its memory is handed back on the admin stream once events recorded on
// every engine stream at this moment have fired, and its slot / result region are recycled when that is known to have happened.
static void mixers_forget_chan(sdr_engine *e, int index);
static void chan_destroy_core(sdr_chan *c, bool have_mu) {
    sdr_engine *e = c->eng;
#ifndef SDR_EMU
    cudaSetDevice(e->device);
    cudaEvent_t fence[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    for (auto &f : fence) cudaEventCreateWithFlags(&f, cudaEventDisableTiming);
#endif
    {
        std::unique_lock<std::mutex> lk(e->mu, std::defer_lock);
        if (!have_mu) lk.lock();
        if (c->index >= 0 && (size_t)c->index < e->chans.size() && e->chans[(size_t)c->index] == c) e->chans[(size_t)c->index] = nullptr;
        if (c->index >= 0) mixers_forget_chan(e, c->index);      // a mixer source bound to it keeps what it has queued and is fed no more
#ifndef SDR_EMU
        // everything already enqueued may still touch the channel; nothing enqueued later will
        cudaStream_t ss[5] = {e->stream, e->stream_tail, e->stream_res, e->stream_dc, e->stream2};
        for (int k = 0; k < 5; ++k) if (fence[k]) cudaEventRecord(fence[k], ss[k]);
#endif
    }
    {
        std::lock_guard<std::mutex> lk(e->mu_cfg);
        sdr_stream_t as = admin_stream(e);
#ifndef SDR_EMU
        for (int k = 0; k < 5; ++k) if (fence[k]) cudaStreamWaitEvent(as, fence[k], 0);
        if (c->ev_ready) cudaStreamWaitEvent(as, c->ev_ready, 0);
#endif
        chan_release_memory(e, c, as);
#ifndef SDR_EMU
        if (cudaEventCreateWithFlags(&c->ev_dead, cudaEventDisableTiming) == cudaSuccess) cudaEventRecord(c->ev_dead, as);
        if (c->h_dyn && c->ev_dead) cudaEventSynchronize(c->ev_dead);   // sdr_modem_*: pinned host buffers go now, the caller's own thread waits
#endif
        be_host_free(c->h_dyn); be_host_free(c->h_slot);
        c->h_dyn = nullptr; c->h_slot = nullptr;
        e->graveyard.push_back(c);
        reap_graveyard(e, false);
    }
#ifndef SDR_EMU
    for (auto &f : fence) if (f) cudaEventDestroy(f);                   // a recorded event may be destroyed: the wait stays valid
#endif
}
static void chan_destroy_locked(sdr_chan *c) { chan_destroy_core(c, true); }

// DemodulatorMgr::deleteThread happens while blocks are in flight.  The channel disappears from the submitter's list at once
// (a few host instructions under the submit lock); its memory is handed back on the admin stream once events recorded on
// every engine stream at this moment have fired, and its slot / result region are recycled when that is known to have happened.
extern "C" int sdr_chan_destroy(sdr_chan *c) {
    if (!c) return SDR_OK;
    chan_destroy_core(c, false);
    return SDR_OK;
}

extern "C" int sdr_chan_set_frequency(sdr_chan *c, long long f) {
    if (!c) return fail(SDR_ERR_ARG, "null channel");
    std::lock_guard<std::mutex> lk(c->eng->mu);
    c->frequency = f;
    return SDR_OK;
}
extern "C" int sdr_chan_get_bandwidth(sdr_chan *c, int *bw) {
    if (!c || !bw) return fail(SDR_ERR_ARG, "null argument");
    *bw = c->bandwidth;
    return SDR_OK;
}
// DemodulatorThread::setSquelchEnabled / setSquelchLevel (src/demod/DemodulatorThread.cpp:384-397): takes effect with the next block
extern "C" int sdr_chan_set_squelch(sdr_chan *c, int enabled, float level_db) {
    if (!c) return fail(SDR_ERR_ARG, "null channel");
    std::lock_guard<std::mutex> lk(c->eng->mu);
    c->squelch_enabled = enabled != 0;
    c->squelch_level = level_db;
    return SDR_OK;
}
// AudioThreadInput::channels of this channel's output: 2 for "I/Q" (interleaved {imag, real}, ModemIQ.cpp:46-55), else 1
extern "C" int sdr_chan_audio_channels(sdr_chan *c, int *channels) {
    if (!c || !channels) return fail(SDR_ERR_ARG, "null argument");
    *channels = (c->modem == MODEM_IQ || c->modem == MODEM_FMS) ? 2 : 1;
    return SDR_OK;
}
extern "C" int sdr_chan_result_caps(sdr_chan *c, unsigned *audio_cap, unsigned *iq_cap) {
    if (!c) return fail(SDR_ERR_ARG, "null channel");
    if (audio_cap) *audio_cap = (unsigned)c->cs.res_audio_cap;
    if (iq_cap) *iq_cap = (unsigned)c->cs.res_iq_cap;
    return SDR_OK;
}
extern "C" int sdr_chan_want_iq(sdr_chan *c, int enable) {
    if (!c) return fail(SDR_ERR_ARG, "null channel");
    std::lock_guard<std::mutex> lk(c->eng->mu);
    c->cs.want_iq = enable ? 1 : 0;
    c->cs_dirty = true;                                                 // re-uploaded with the next block's parameters
    return SDR_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// spectrum
// ---------------------------------------------------------------------------------------------------------------------
// The reference's point loop (SpectrumVisualProcessor.cpp:532-556) decides with a running double accumulator how many bins merge
// into each display point; the same loop, run once on the host for a view ratio, gives every point its first bin offset and
// bin count (the device then does the sums).
static void spectrum_point_table(sdr_spectrum *s, double visualRatio) {
    double visualAccum = 0, i = 0;
    for (unsigned x = 0; x < s->fft_size; ++x) {
        visualAccum += visualRatio * double(2);                               // SPECTRUM_VZM
        int cnt = 0;
        const int i0 = (int)i;
        while (visualAccum >= 1.0) { cnt++; visualAccum -= 1.0; i++; }
        s->h_pt[x] = i0; s->h_pt[s->fft_size + x] = cnt < 1 ? 1 : cnt;
    }
    s->visual_ratio = visualRatio;
}
static void chan_destroy_locked(sdr_chan *c);
static void spectrum_free(sdr_spectrum *s) {
    be_free(s->d_pt); be_host_free(s->h_pt); be_free(s->d_ema_tmp);
    be_free(s->d_tw); be_free(s->d_mags); be_free(s->d_maa); be_free(s->d_ema); be_free(s->d_scalars); be_free(s->d_frame_cf);
    be_free(s->d_frame_cf_maa); be_free(s->d_points); be_free(s->d_lines); be_free(s->d_frame_start);
    be_host_free(s->h_points); be_host_free(s->h_lines); be_host_free(s->h_cf); be_host_free(s->h_frame_start); be_host_free(s->h_dbg);
}

extern "C" int sdr_spectrum_create(sdr_engine *e, unsigned fft_size, unsigned lps, sdr_spectrum **out) {
    if (!e || !out) return fail(SDR_ERR_ARG, "sdr_spectrum_create: null argument");
    if (fft_size < 16 || (fft_size & (fft_size - 1)) || fft_size > 32768)
        return fail(SDR_ERR_UNSUPPORTED, "sdr_spectrum_create: fft_size must be a power of two in [16, 32768]");
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    sdr_spectrum *s = new sdr_spectrum();
    s->eng = e; s->fft_size = fft_size; s->N = 2 * fft_size; s->lps = lps;     // SPECTRUM_VZM = 2 (SpectrumVisualProcessor.cpp:145)
    // sub-sequences per frame: at most kFft2MaxSub points per CTA (a CTA's transform takes a few microseconds: frames are not
    // split further just to occupy more SMs)
    s->P = 1;
    while ((int)(s->N / s->P) > kFft2MaxSub) s->P *= 2;
    s->nsub = s->N / s->P;
    s->log_nsub = 0;
    while ((1 << s->log_nsub) < s->nsub) ++s->log_nsub;
    // frames per block: the block's own, plus those that fit into what the distributor may still hold from earlier blocks -- less
    // than one frame of the settings of THEN, and a zoomed view's frames are up to hist_cap samples long (leaving a view, or
    // zooming out, turns that remainder into many short frames at once)
    s->f_max = (unsigned)((e->max_block + e->hist_cap) / s->N + 2);
    const size_t N = s->N, F = s->f_max;
    int bad = 0;
    bad |= be_malloc((void **)&s->d_tw, N * sizeof(float2));
    bad |= be_malloc((void **)&s->d_mags, F * N * sizeof(float));
    bad |= be_malloc((void **)&s->d_maa, F * N * sizeof(float));
    bad |= be_malloc((void **)&s->d_ema, N * sizeof(double2));
    bad |= be_malloc((void **)&s->d_scalars, 8 * sizeof(double));          // two sets: a block reads one and writes the other
    bad |= be_malloc((void **)&s->d_frame_cf, F * sizeof(float2));
    bad |= be_malloc((void **)&s->d_frame_cf_maa, F * sizeof(double2));
    bad |= be_malloc((void **)&s->d_points, F * fft_size * sizeof(float));
    bad |= be_malloc((void **)&s->d_lines, F * fft_size);
    bad |= be_malloc((void **)&s->d_frame_start, 2 * F * sizeof(int));   // double-buffered by block parity
    bad |= be_host_alloc((void **)&s->h_points, kResBuf * F * fft_size * sizeof(float));
    bad |= be_host_alloc((void **)&s->h_lines, kResBuf * F * fft_size);
    bad |= be_host_alloc((void **)&s->h_cf, kResBuf * F * sizeof(double2));
    bad |= be_host_alloc((void **)&s->h_frame_start, (size_t)sdr_engine::kRing * F * sizeof(int));
    bad |= be_host_alloc((void **)&s->h_dbg, 2 * N * sizeof(float));
    bad |= be_malloc((void **)&s->d_pt, 2 * (size_t)fft_size * sizeof(int));
    bad |= be_host_alloc((void **)&s->h_pt, (size_t)(1 + sdr_engine::kRing) * 2 * (size_t)fft_size * sizeof(int));   // [table | one upload slot per parameter-ring entry]
    bad |= be_malloc((void **)&s->d_ema_tmp, N * sizeof(double2));
    if (bad) { spectrum_free(s); delete s; return fail(SDR_ERR_NOMEM, "sdr_spectrum_create: allocation failed"); }
    std::vector<float2> tw(N);
    for (size_t k = 0; k < N; ++k) {
        double ang = -2.0 * 3.14159265358979323846 * (double)k / (double)N;
        tw[k] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    be_h2d(s->d_tw, tw.data(), N * sizeof(float2), e->stream);
    be_memset(s->d_ema, 0, N * sizeof(double2), e->stream);                      // vectors resized to zeros (:252-264)
    const double sc[4] = {100.0, 100.0, 0.0, 0.0};                                 // fft_ceil_ma/maa, fft_floor_ma/maa (:32-33)
    be_h2d(s->d_scalars, sc, sizeof sc, e->stream);
    spectrum_point_table(s, 1.0);
    be_h2d(s->d_pt, s->h_pt, 2 * (size_t)fft_size * sizeof(int), e->stream);
    be_sync(e->stream);
    e->spectra.push_back(s);
    *out = s;
    return SDR_OK;
}

extern "C" int sdr_spectrum_destroy(sdr_spectrum *s) {
    if (!s) return SDR_OK;
    sdr_engine *e = s->eng;
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    sync_all(e);                                                        // the side / result streams may still use its buffers
    if (s->vchan) { chan_destroy_locked(s->vchan); s->vchan = nullptr; }
    e->spectra.erase(std::remove(e->spectra.begin(), e->spectra.end(), s), e->spectra.end());
    spectrum_free(s);
    delete s;
    return SDR_OK;
}

// SpectrumVisualProcessor::setView(bView, centerFreq, bandwidth) (:64-71); setCenterFrequency / setBandwidth are the same call
// with one argument changed.  Effective from the next block.
extern "C" int sdr_spectrum_set_view(sdr_spectrum *s, int is_view, long long center_freq, long bandwidth) {
    if (!s) return fail(SDR_ERR_ARG, "null spectrum");
    if (is_view && bandwidth <= 0) return fail(SDR_ERR_ARG, "sdr_spectrum_set_view: bandwidth must be positive");
    std::lock_guard<std::mutex> lk(s->eng->mu);
    s->is_view = is_view != 0; s->view_center = center_freq; s->view_bw = bandwidth;
    return SDR_OK;
}
// getDesiredInputSize (:134-138) for a sample rate: the frame length the view asks of the distributor
extern "C" int sdr_spectrum_desired_input(sdr_spectrum *s, long long sample_rate, unsigned *n) {
    if (!s || !n || sample_rate <= 0) return fail(SDR_ERR_ARG, "sdr_spectrum_desired_input: bad arguments");
    std::lock_guard<std::mutex> lk(s->eng->mu);
    if (!s->is_view) { *n = s->N; return SDR_OK; }
    long rb = (long)sample_rate;
    while (rb / 2 >= s->view_bw) rb /= 2;
    *n = (unsigned)(size_t)((double)s->N / ((double)rb / (double)sample_rate));
    return SDR_OK;
}
extern "C" int sdr_spectrum_set_average_rate(sdr_spectrum *s, float r) {
    if (!s) return fail(SDR_ERR_ARG, "null spectrum");
    std::lock_guard<std::mutex> lk(s->eng->mu);
    s->average_rate = r;
    return SDR_OK;
}

// FFTDataDistributor::process restated for frames that are transformed immediately (src/process/FFTDataDistributor.cpp:
// 41-53 reset on rate change, :89-97 line-rate step, :100-142 emit loop).  Returns frame starts relative to the block.
static void schedule_frames(sdr_spectrum *s, long long frame_len, unsigned n, long long sample_rate, long long center, std::vector<int> &starts) {
    starts.clear();
    // "Settings have changed ... dump all previous samples" (:41-53): a new rate OR a new tuner frequency empties the buffer;
    // lineRateAccum is left alone, as there
    if (s->last_rate != sample_rate || s->last_center != center) { s->buffered = 0; s->last_rate = sample_rate; s->last_center = center; }
    const long long before = s->buffered, N = frame_len;
    s->buffered += n;
    if (s->buffered < N) return;
    if (s->lps == 0) {                                  // full coverage: every non-overlapping frame
        long long i = 0;
        for (; i + N <= s->buffered; i += N) starts.push_back((int)(i - before));
        s->buffered -= i;
        return;
    }
    const double inputTime = (double)s->buffered / (double)sample_rate;
    const double inputLines = (double)s->buffered / (double)N;
    const double step = ((double)s->lps * inputTime) / inputLines;
    long long processed = 0;
    if (s->line_accum + (step * ((double)s->buffered / (double)N)) < 1.0) {
        s->line_accum += (step * ((double)s->buffered / (double)N));
        processed = s->buffered;
    } else {
        for (long long i = 0; i < s->buffered; i += N) {
            if (i + N > s->buffered) break;
            s->line_accum += step;
            if (s->line_accum >= 1.0) {
                starts.push_back((int)(i - before));
                while (s->line_accum >= 1.0) s->line_accum -= 1.0;
            }
            processed += N;
        }
    }
    s->buffered -= processed;
    if (s->buffered < 0) s->buffered = 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// submit
// ---------------------------------------------------------------------------------------------------------------------
// The result buffers (device + a ring of pinned host copies) are sized generously at engine creation; only when the layout
// outgrows them does the engine stop, once, to enlarge them.
static int ensure_results(sdr_engine *e) {
    size_t need;
    { std::lock_guard<std::mutex> lk(e->mu_cfg); need = e->res_hwm; }
    if (need <= e->results_bytes) return SDR_OK;
    sync_all(e);
    size_t total = std::max<size_t>(need * 2, (size_t)16 << 20);
    unsigned char *d_new = nullptr, *h_new = nullptr;
    if (be_malloc((void **)&d_new, total) || be_host_alloc((void **)&h_new, kResBuf * total)) {
        be_free(d_new); be_host_free(h_new);
        return fail(SDR_ERR_NOMEM, "results buffer");
    }
    be_free(e->d_results); be_host_free(e->h_results);
    e->d_results = d_new; e->h_results = h_new;
    e->results_bytes = total;
    for (int b = 0; b < kResBuf; ++b) e->res_on_host[b] = false;
    return SDR_OK;
}

// ready: 0 = the block becomes ready when the engine's main stream reaches this point (conservative),
//        1 = the block is ready now (resident data, or a host block that submit copies on the DC stream itself),
//        2 = the block is ready when `ready_event` fires
#ifndef SDR_EMU
// the filterbank kernel's dynamic shared memory grows with the channel count: raise the opt-in limit monotonically
static int pfb_allow_smem(size_t bytes) {
    static size_t allowed[64];                                     // per device (zero-initialised: anything above 48 KB asks once)
    int d = 0;
    cudaGetDevice(&d);
    d &= 63;
    if (bytes > 48 * 1024 && bytes > allowed[d]) {
        CUDA_TRY(cudaFuncSetAttribute(k_pfb_analyze, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        allowed[d] = bytes;
    }
    return SDR_OK;
}
#endif
// channelizer mode: (re)build the filterbank tables and the channel-stream buffers when the channel count changes
// (SDRPostThread::initPFBCH, SDRPostThread.cpp:401-414: a new count means a new channelizer, i.e. empty windows)
static int pfb_prepare(sdr_engine *e, int M, int osf) {
    if (e->pfb_M == M && e->pfb_osf == osf && e->d_pfb_streams[0]) return SDR_OK;
    sync_all(e);
    be_free(e->d_pfb_hq); be_free(e->d_pfb_tw); be_free(e->d_pfb_streams[0]); be_free(e->d_pfb_streams[1]); be_free(e->d_pfb_dc_state);
    e->d_pfb_hq = nullptr; e->d_pfb_tw = nullptr; e->d_pfb_streams[0] = e->d_pfb_streams[1] = nullptr; e->d_pfb_dc_state = nullptr;
    std::vector<float> hq, tw;
    if (osf == 2) pfb2_analysis_tables((unsigned)M, 4, 60.0f, hq, tw);       // firpfbch2_crcf_create_kaiser(LIQUID_ANALYZER, M, 4, 60)
    else pfb_analysis_tables((unsigned)M, 4, 60.0f, hq, tw);                  // firpfbch_crcf_create_kaiser(LIQUID_ANALYZER, M, 4, 60)
    e->pfb_slots = (size_t)M + 1;                                            // M channels + the DC-corrected copy of channel 0
    e->pfb_stride = ((size_t)kBlkHist + (size_t)osf * (e->max_block / (size_t)M) + 8 + 1) & ~(size_t)1;
    const size_t bytes = e->pfb_slots * e->pfb_stride * sizeof(float2);
    int bad = 0;
    bad |= be_malloc((void **)&e->d_pfb_hq, hq.size() * sizeof(float));
    bad |= be_malloc((void **)&e->d_pfb_tw, tw.size() * sizeof(float));
    for (int b = 0; b < 2; ++b) bad |= be_malloc((void **)&e->d_pfb_streams[b], bytes);
    bad |= be_malloc((void **)&e->d_pfb_dc_state, sizeof(float2));
    if (bad) return fail(SDR_ERR_NOMEM, "channelizer buffers");
    be_h2d(e->d_pfb_hq, hq.data(), hq.size() * sizeof(float), e->stream);
    be_h2d(e->d_pfb_tw, tw.data(), tw.size() * sizeof(float), e->stream);
    for (int b = 0; b < 2; ++b) be_memset(e->d_pfb_streams[b], 0, bytes, e->stream);
    be_memset(e->d_pfb_dc_state, 0, sizeof(float2), e->stream);
    be_sync(e->stream);
    e->pfb_M = M; e->pfb_osf = osf; e->pfb_frame_base = 0; e->pfb_have_block = false; e->pfb_last_frames = 0;
    for (auto *c : e->chans) if (c) c->pfb_slot_prev = -1;
    return SDR_OK;
}

static int submit_common(sdr_engine *e, const float2 *d_src, const sdr_cf32 *h_src, unsigned n, long long center, long long rate,
                         bool to_host, int ready, void *ready_event, void *consumed_event = nullptr) {
    if (e->poisoned) return fail(SDR_ERR_STATE, "engine is in a failed state after a CUDA error: destroy it");
#ifndef SDR_EMU
    cudaSetDevice(e->device);
#else
    (void)consumed_event;
#endif
    { int r = ensure_results(e); if (r) return r; }
    const int p = e->parity ^ 1;                               // this block's buffer; the other one holds the previous block
    float2 *blk = e->d_blkbuf[p] + e->hist_cap;
    float2 *raw = e->d_rawbuf + e->hist_cap;

    // SDRPostThread::run dispatch (SDRPostThread.cpp:178-187): blocks that come with more than one channel go through the
    // channelizer; the count is SDRThread::getOptimalChannelCount of the rate (SoapySDRThread.cpp:513,676-693)
    const int pfb_M = (e->chan_mode != 0) ? optimal_channel_count(rate) : 1;
    const int pfb_osf = e->chan_mode == 2 ? 2 : 1;                // SDRPostPFBCH2: firpfbch2, two output frames per numChannels inputs
    const bool pfb = pfb_M > 1;
    if (pfb) {
        if (pfb_M > kPfbMaxChannels) return fail(SDR_ERR_UNSUPPORTED, "channelizer mode: sample rate needs more channels than supported");
        if (n % (unsigned)pfb_M) return fail(SDR_ERR_ARG, "channelizer mode: block length must be a multiple of the channel count (SoapySDRThread.cpp:668-674)");
        if ((size_t)(2 * 8 - 1) * pfb_M > (size_t)kBlkHist) return fail(SDR_ERR_UNSUPPORTED, "channelizer history");
        int r = pfb_prepare(e, pfb_M, pfb_osf);
        if (r) return r;
    }
    // samples every demodulator sees this block, and at which rate: chanBw = sampleRate / numChannels (SDRPostThread.cpp:408),
    // twice that for the oversampled channelizer (runDemodChannels(chanBw * 2), :510)
    const unsigned n_ch = pfb ? (unsigned)pfb_osf * (n / (unsigned)pfb_M) : n;
    const long long rate_ch = pfb ? (long long)pfb_osf * (rate / pfb_M) : rate;
    if (n_ch > e->chan_block_max) return fail(SDR_ERR_ARG, "block exceeds sdr_engine_set_channel_block_max");
    // ---- the arguments are fine: from here on host-side state advances, so any failure below (a CUDA error) leaves the
    //      engine unusable (the C-ABI wrappers drain the streams and refuse further blocks)
    const int slot = (int)(e->ring_seq++ % sdr_engine::kRing);
#ifndef SDR_EMU
    CUDA_TRY(cudaEventSynchronize(e->ring_ev[slot]));      // the copies that last used this parameter slot are done
#endif
    ChanDyn *h_cd = e->h_cd + (size_t)slot * sdr_engine::kMaxChan;
    int *h_group = e->h_group + (size_t)slot * sdr_engine::kMaxChan;
    // channelizer mode: channel centre frequencies (SDRPostThread::updateChannels, SDRPostThread.cpp:115-124) and the channel
    // every attached demodulator lands on (getChannelAt, :127-139: nearest centre, the first one wins a tie).  Channel 0's DC
    // blocker only runs on blocks in which some demodulator sits on channel 0 (runDemodChannels skips idle channels,
    // :333-339), so its state must not advance on the others.
    std::vector<long long> chan_centers;
    std::vector<int> pfb_pick;
    bool pfb_ch0_active = false;
    if (pfb) {
        chan_centers.assign((size_t)pfb_M + 1, 0);
        const long long chan_bw = rate / pfb_M;
        for (int i = 0; i < pfb_M / 2; ++i) {
            const int ofs = (int)(chan_bw * i);
            chan_centers[i] = center + ofs;
            chan_centers[i + pfb_M / 2] = center - (rate / 2) + ofs;
        }
        chan_centers[pfb_M] = center + (rate / 2);
        pfb_pick.assign(e->chans.size(), -1);
        for (size_t i = 0; i < e->chans.size(); ++i) {
            const sdr_chan *c = e->chans[i];
            if (!c || c->standalone) continue;
            long long min_delta = rate;
            for (int k = 0; k <= pfb_M; ++k) {
                const long long fd = std::llabs(c->frequency - chan_centers[k]);
                if (fd < min_delta) { min_delta = fd; pfb_pick[i] = k; }
            }
            if (pfb_pick[i] == 0) pfb_ch0_active = true;
        }
    }

    // 1+2. DC blocker (SDRPostThread.cpp:284) on its own stream, so that it runs underneath the channel kernels of the
    //      previous block: raw history slide -> copy the block in -> double-precision scan -> bit-exact float replay
    //      into blk -> history of blk from the other buffer.
    sdr_stream_t dcs = e->stream;
#ifndef SDR_EMU
    dcs = e->stream_dc;
    if (ready == 0) {
        CUDA_TRY(cudaEventRecord(e->ev_input, e->stream));     // the input block is ready when the caller's stream gets here
        CUDA_TRY(cudaStreamWaitEvent(dcs, e->ev_input, 0));
    } else if (ready == 2 && ready_event) {
        CUDA_TRY(cudaStreamWaitEvent(dcs, (cudaEvent_t)ready_event, 0));
    }
    CUDA_TRY(cudaStreamWaitEvent(dcs, e->ev_consumed[p], 0));  // the kernels of two blocks ago are done with this buffer
    if (e->timing && e->time_pending < sdr_engine::kTimeSets) cudaEventRecord(e->tev_dc[e->time_pending][0], dcs);
#endif
    if (pfb) {
        // channelizer mode has no wideband DC blocker (runPFBCH pushes the block as it came to the visual queues and to the
        // filterbank, SDRPostThread.cpp:432-451): the block goes straight into the stream buffer
        if (h_src) be_h2d(blk, h_src, (size_t)n * sizeof(float2), dcs);
        else be_d2d(blk, d_src, (size_t)n * sizeof(float2), dcs);
#ifndef SDR_EMU
        if (consumed_event) CUDA_TRY(cudaEventRecord((cudaEvent_t)consumed_event, dcs));
#endif
    } else {
        const size_t need_raw = std::min(e->hist_cap, (size_t)kDcWarm + 2 * kDcTile);
        // which replay: candidate windows (no warm-up: a few dozen microseconds of chip-wide throughput work) unless enough
        // channels are attached for a warm-up replay on a few reserved SMs to hide under the channel front kernel
        size_t live = 0;
        for (auto *c : e->chans) live += c != nullptr;
        const unsigned tiles_n = (n + kDcTile - 1) / kDcTile;
        // automatic choice (measured, profiles/README.md): the windows' table pass is chip-wide work that competes with a long
        // front kernel -- at 256 channels and more the warm-up replay on the reserved SMs is the faster block period
        const bool window = e->exact_dc && tiles_n >= (unsigned)kDwMinTiles && (e->dc_window < 0 ? live < 256 : e->dc_window == 1);
        if (e->have_block && e->last_n > 0 && !window) {
            be_d2d(e->d_scratch, raw + (long long)e->last_n - (long long)need_raw, need_raw * sizeof(float2), dcs);
            be_d2d(raw - need_raw, e->d_scratch, need_raw * sizeof(float2), dcs);
        }
        DcArgs dc;
        dc.x = raw; dc.y = blk; dc.n = (int)n; dc.a = e->dc_a; dc.tile_sum = e->d_tile_sum; dc.tile_carry = e->d_tile_carry2[p];
        dc.state = e->d_dc_state; dc.apow = e->d_apow; dc.copy_to = nullptr;
        if (h_src) be_h2d(raw, h_src, (size_t)n * sizeof(float2), dcs);   // host block (emulator): straight into the raw stream buffer
        else if (e->exact_dc) { dc.x = d_src; dc.copy_to = raw; }         // the tile-sum pass of the scan brings the block in as it reads it
        else be_d2d(raw, d_src, (size_t)n * sizeof(float2), dcs);
        launch_dc(dc, !e->exact_dc, dcs, e->launches);
        dc.x = raw;
#ifndef SDR_EMU
        if (consumed_event) CUDA_TRY(cudaEventRecord((cudaEvent_t)consumed_event, dcs));   // the caller's buffer may be refilled
#endif
        if (e->exact_dc) {
            DcExactArgs xa;
            // one lane per segment on the SMs the channel front kernel leaves free: as many segments (multiples of the scan
            // tile) as those SMs hold lanes, each warmed up over kDcWarm samples
            const unsigned seg = (unsigned)kDcTile;                 // both replays work tile by tile
            xa.xraw = raw; xa.y = blk; xa.n = (int)n; xa.seg_len = (int)seg; xa.warm = window ? 0 : kDcWarm;
            xa.start = nullptr; xa.junction_bad = e->d_dw_stats + 1; xa.states_only = 0; xa.no_publish = 0; xa.fix_mask = nullptr;
            xa.raw_hist = (int)need_raw; xa.raw_end = (long long)e->max_block + (long long)e->raw_pad;
            xa.c = (float)e->dc_a; xa.one = 1.0f;
            xa.tile_carry = e->d_tile_carry2[p]; xa.prev_tile_carry = e->d_tile_carry2[p ^ 1];
            xa.prev_n = e->have_block ? (int)e->last_n : 0;
            xa.seg_in = e->d_seg_in; xa.seg_out = e->d_seg_out; xa.fstate = e->d_fstate; xa.fixups = e->d_fixups;
            // speculate only when there is parallelism to gain and the warm-up region lies inside the data we hold
            const bool spec = window || (n >= (unsigned)kDcMinSpec && n > seg && (!e->have_block || e->last_n >= (unsigned)(kDcWarm + kDcTile)));
            xa.n_seg = spec ? (int)((n + seg - 1) / seg) : 1;
            if (!spec) xa.seg_len = (int)n;
            if (window) {
                DcWinArgs wa;
                wa.xraw = raw; wa.n = (int)n; wa.tiles = (int)tiles_n; wa.c = xa.c; wa.tile_carry = e->d_tile_carry2[p]; wa.fstate = e->d_fstate;
                wa.st = e->d_dw_st; wa.pst = e->d_dw_pst; wa.gmap = e->d_dw_gmap; wa.entry = e->d_dw_entry; wa.seg_in = e->d_seg_in;
                wa.stats = e->d_dw_stats;
                { int r = launch_dc_window(wa, dcs, e->launches); if (r) return r; }
                xa.start = e->d_seg_in;
                e->dw_blocks += 1;
            }
            { int r = launch_dc_exact(xa, (unsigned long long)xa.warm + e->max_block + e->raw_pad, dcs, e->launches, e->d_dc_fix_mask, e->d_dw_stats); if (r) return r; }
        }
    }
    // history in front of the block: the tail of the previous block, which lives in the other buffer
    size_t need = kBlkHist;
    for (auto *s : e->spectra) {
        size_t fl = s->N;
        if (s->is_view && s->view_bw > 0) { long rb = (long)rate; while (rb / 2 >= s->view_bw) rb /= 2; fl = (size_t)((double)s->N / ((double)rb / (double)rate)); }
        need = std::max(need, fl + 16);
        // what the distributor still holds from earlier blocks was buffered under the frame length of THEN: after a zoom-out (or
        // on leaving a view) the frames get shorter and several whole frames start inside that remainder, further back than one
        // new frame
        if (s->buffered > 0) need = std::max(need, (size_t)s->buffered + 16);
    }
    if (need > e->hist_cap) need = e->hist_cap;
    if (e->have_block && e->last_n > 0) {
        const float2 *prev = e->d_blkbuf[p ^ 1] + e->hist_cap;
        be_d2d(blk - need, prev + (long long)e->last_n - (long long)need, need * sizeof(float2), dcs);
    }
    float2 *pfb_streams = nullptr;
    if (pfb) {
        // filterbank -> channel-major streams [slot][kBlkHist history | n_ch new]; history = tail of the previous block's
        // streams (other parity); slot M = channel 0 after the DC blocker (runDemodChannels, SDRPostThread.cpp:363-375)
        pfb_streams = e->d_pfb_streams[p];
        const size_t stride = e->pfb_stride;
        if (e->pfb_have_block)
            be_d2d_2d(pfb_streams, stride * sizeof(float2), e->d_pfb_streams[p ^ 1] + e->pfb_last_frames, stride * sizeof(float2),
                      (size_t)kBlkHist * sizeof(float2), e->pfb_slots, dcs);
        PfbArgs pa;
        pa.x = blk; pa.y = pfb_streams + kBlkHist; pa.hq = e->d_pfb_hq; pa.tw = e->d_pfb_tw; pa.M = pfb_M; pa.p = 8; pa.F = (int)n_ch;
        pa.chan_stride = (long long)stride; pa.osf = pfb_osf; pa.frame_base = e->pfb_frame_base;
        e->pfb_frame_base += n_ch;
        const int tiles = ((int)n_ch + kPfbTileFrames - 1) / kPfbTileFrames;
        float2 *ch0 = pfb_streams + kBlkHist, *ch0_dc = pfb_streams + (size_t)pfb_M * stride + kBlkHist;
#ifndef SDR_EMU
        const size_t smem = pfb_smem_elems(pa.M, pa.p) * sizeof(float2);
        { int r = pfb_allow_smem(smem); if (r) return r; }
        k_pfb_analyze<<<tiles, kPfbThreads, smem, dcs>>>(pa);
        if (pfb_ch0_active) {
            // channel 0's DC blocker (runDemodChannels, SDRPostThread.cpp:363-375): the same float recurrence at the channel rate,
            // one warp with the samples prefetched three batches ahead (k_dc_exact), the carried state published by k_dc_verify
            DcExactArgs xa;
            memset(&xa, 0, sizeof xa);
            xa.xraw = ch0; xa.y = ch0_dc; xa.n = (int)n_ch; xa.n_seg = 1; xa.seg_len = (int)n_ch; xa.warm = 0;
            xa.c = (float)e->dc_a; xa.one = 1.0f; xa.seg_in = e->d_seg_in; xa.seg_out = e->d_seg_out; xa.fstate = e->d_pfb_dc_state;
            xa.fixups = e->d_fixups; xa.start = nullptr; xa.junction_bad = e->d_dw_stats + 1;
            k_dc_exact<<<1, 32, 0, dcs>>>(xa);
            k_dc_verify<<<1, 256, 0, dcs>>>(xa);
        }
#else
        {
            std::vector<float2> sm(pfb_smem_elems(pa.M, pa.p));
            for (int t = 0; t < tiles; ++t)
                for (int ph = 0; ph < kPfbPhases; ++ph) for (int th = 0; th < kPfbThreads; ++th) pfb_phase(ph, pa, t, th, kPfbThreads, sm.data());
            if (pfb_ch0_active) pfb_dc0_serial(ch0, ch0_dc, (int)n_ch, (float)e->dc_a, e->d_pfb_dc_state);
        }
#endif
        e->launches += pfb_ch0_active ? 3 : 1;
        e->pfb_last_frames = n_ch; e->pfb_have_block = true;
    }
#ifndef SDR_EMU
    if (e->timing && e->time_pending < sdr_engine::kTimeSets) cudaEventRecord(e->tev_dc[e->time_pending][1], dcs);
    CUDA_TRY(cudaEventRecord(e->ev_dc_done[p], dcs));
    CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_dc_done[p], 0));
#endif
    time_mark(e, 0);
    time_mark(e, 1);
    // Per-block parameter blocks are uploaded on the DC stream, behind the replay: that stream runs a block ahead on its own
    // SMs, so the upload kernels never wait for a free SM slot right in front of the channel front kernel (the main stream
    // just waits for the event).  Their device buffers are double-buffered by block parity.
    sdr_stream_t par_stream = dcs;
    auto params_ready = [&]() -> int {
#ifndef SDR_EMU
        CUDA_TRY(cudaEventRecord(e->ev_params[p], par_stream));
        CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_params[p], 0));
#endif
        return SDR_OK;
    };
    e->parity = p;
    e->last_n = n; e->have_block = true;
    e->results_on_host = to_host;
    const int prev_q = e->last_res_q;
    const int q = e->res_sel = (e->res_sel + 1) % kResBuf;     // host result buffer of this block
    e->last_res_q = to_host ? q : -1;
    (void)prev_q;
    e->res_on_host[q] = to_host;
    e->fetch_sel = q;
#ifndef SDR_EMU
    e->res_tail_rec[q] = e->res_spec_rec[q] = false;
#endif

    // 3. spectra.  Frame scheduling and the (tiny) frame-start upload happen here; the kernels are enqueued on the side
    //    stream AFTER the channel front kernel has been launched (run_spectra below), so that the front kernel's single
    //    wave of 148 CTAs gets the SMs first and the spectrum fills in underneath the channel tail.
    std::vector<int> starts;
    struct SpecJob { sdr_spectrum *s; SpecArgs a; bool view; ViewMixArgs mix; sdr_chan *vc; int n_view; bool wait_ready; };
    std::vector<SpecJob> spec_jobs;
    std::vector<std::pair<int, ChanDyn>> view_dyn;             // per-block parameters of the views' hidden channels (patched in below)
    for (auto *s : e->spectra) {
        // frame length: the transform's, or what a zoomed view needs to end up with it after its resampler (:283-297)
        long resample_bw = (long)rate;
        long long D = s->N;
        if (s->is_view) {
            while (resample_bw / 2 >= s->view_bw) resample_bw /= 2;          // SPECTRUM_VZM = 2
            D = (long long)(size_t)((double)s->N / ((double)resample_bw / (double)rate));
            if ((size_t)D + 16 > e->hist_cap) return fail(SDR_ERR_UNSUPPORTED, "zoomed view: a frame of the view is longer than the engine's stream history (max(524288, max_block) samples)");
        }
        schedule_frames(s, D, n, rate, center, starts);
        s->frames_last = (unsigned)std::min<size_t>(starts.size(), s->f_max);
        s->frames_buf[q] = s->frames_last;
        if (!s->frames_last) continue;
        int *h_fs = s->h_frame_start + (size_t)slot * s->f_max;
        memcpy(h_fs, starts.data(), s->frames_last * sizeof(int));
        int *d_fs = s->d_frame_start + (size_t)p * s->f_max;
        upload_params(d_fs, h_fs, s->frames_last * sizeof(int), par_stream, e->launches);
        SpecJob job;
        job.s = s; job.view = false; job.vc = nullptr; job.n_view = 0; job.wait_ready = false;
        SpecArgs &a = job.a;
        a.x = blk; a.frame_start = d_fs; a.frame_stride = 0; a.F = (int)s->frames_last; a.N = (int)s->N; a.fft_size = (int)s->fft_size;
        a.P = s->P; a.nsub = s->nsub; a.log_nsub = s->log_nsub; a.tw = s->d_tw; a.mags = s->d_mags; a.maa_f = s->d_maa; a.ema = s->d_ema;
        a.rate = (double)s->average_rate; a.frame_cf = s->d_frame_cf; a.scalars = s->d_scalars + 4 * s->scalar_sel; a.scalars_out = s->d_scalars + 4 * (s->scalar_sel ^ 1); s->scalar_sel ^= 1; a.frame_cf_maa = s->d_frame_cf_maa;
        a.points = s->d_points; a.lines = s->d_lines; a.scale_factor = s->scale_factor;
        a.rebin = 0; a.shift = 0; a.ema_tmp = s->d_ema_tmp;
        double visual_ratio = 1.0;
        if (s->is_view) {
            // ---- SpectrumVisualProcessor::process, view branch, once for the block's frames (only its first frame can change anything)
            const long long N = s->N;
            bool mixing = false;
            if (s->view_center != center) {                                             // :304-347
                if ((s->view_center - center) != (long long)s->shift_frequency || s->last_input_bw != (long)rate) {
                    if (std::llabs(center - s->view_center) < (rate / 2)) {
                        const long last_shift = s->shift_frequency;
                        s->shift_frequency = (long)(s->view_center - center);
                        s->nco_dtheta = nco_phase_step(nco_shift_word_arg(s->shift_frequency, rate));
                        const long freq_diff = s->shift_frequency - last_shift;
                        if (s->last_bandwidth != 0) {                                   // the averages follow the view
                            const double bin_per_hz = double(s->last_bandwidth) / double(N);
                            const unsigned num_shift = (unsigned)std::floor(double(std::labs(freq_diff)) / bin_per_hz);
                            if (num_shift < (unsigned)(N / 2) && num_shift) a.shift = freq_diff > 0 ? (int)num_shift : -(int)num_shift;
                        }
                    }
                }
                mixing = true;
            }
            long bw_diff = 0;
            bool new_resampler = false;
            if (!s->vchan || resample_bw != s->last_bandwidth || s->last_input_bw != (long)rate) {   // :354-368
                if (s->vchan) { chan_destroy_locked(s->vchan); s->vchan = nullptr; }
                // one block's frames of a view: the block plus the distributor's remainder (less than one frame of the previous settings)
                const unsigned cap = (unsigned)std::min<size_t>(std::max<size_t>(2 * (size_t)e->max_block, (size_t)e->max_block + e->hist_cap) + 64, 0x7fffff00u);
                int r = chan_create_impl(e, nullptr, rate, resample_bw, 48000, 0, &s->vchan, cap);
                if (r) return r;
                bw_diff = resample_bw - s->last_bandwidth;
                s->last_bandwidth = resample_bw; s->last_input_bw = (long)rate;
                new_resampler = true;
            }
            if (new_resampler && s->last_view) a.rebin = bw_diff < 0 ? -1 : 1;           // :454
            visual_ratio = double(s->view_bw) / double(resample_bw);                    // :532
            sdr_chan *vc = s->vchan;
            const int n_view = (int)(s->frames_last * D);
            job.view = true; job.vc = vc; job.n_view = n_view; job.wait_ready = vc->fresh; vc->fresh = false;
            job.mix.x = blk; job.mix.frame_start = d_fs; job.mix.F = (int)s->frames_last; job.mix.D = (int)D; job.mix.nco = e->d_nco;
            job.mix.theta0 = s->nco_theta; job.mix.dtheta = s->nco_dtheta;
            job.mix.dir = mixing ? (s->shift_frequency < 0 ? +1 : -1) : 0;                // mix_block_up if shiftFrequency < 0 (:340-344)
            job.mix.out = reinterpret_cast<float2 *>(vc->cs.arena + (vc->lvl0_sel ? vc->cs.lvl0_alt_off : vc->cs.lvl_off[0])) + kHist;
            if (mixing) s->nco_theta += (unsigned)n_view * s->nco_dtheta;
            ChanDyn vd;
            memset(&vd, 0, sizeof vd);
            vd.valid = 1; vd.active = 1; vd.view = 1; vd.g0 = vc->consumed; vd.lvl0_sel = vc->lvl0_sel;
            vc->lvl0_sel ^= 1; vc->consumed += n_view;
            view_dyn.emplace_back(vc->index, vd);
            // every frame yields exactly N resampled samples (power-of-two ratio, frame = N / ratio): the transforms read the
            // hidden channel's decimated stream, frame f at f N
            a.x = reinterpret_cast<const float2 *>(vc->cs.arena + vc->cs.iq_off) + kHist;
            a.frame_stride = (int)N;
        }
        if (visual_ratio != s->visual_ratio) {                   // bins per display point for this view (the host-side loop, uploaded)
            spectrum_point_table(s, visual_ratio);
            int *h_tab = s->h_pt + (size_t)(1 + slot) * 2 * s->fft_size;
            memcpy(h_tab, s->h_pt, 2 * (size_t)s->fft_size * sizeof(int));
            upload_params(s->d_pt, h_tab, 2 * (size_t)s->fft_size * sizeof(int), par_stream, e->launches);
        }
        a.pt_i0 = s->d_pt; a.pt_cnt = s->d_pt + s->fft_size;
        a.visual_start = (double(s->N) / 2.0) - (double(s->N) * (visual_ratio / 2.0));
        s->last_view = s->is_view;                               // :636
        spec_jobs.push_back(job);
    }
    bool forked = false;
    auto run_spectra = [&]() -> int {
        sdr_stream_t sps = e->stream;
#ifndef SDR_EMU
        if (!spec_jobs.empty()) {
            CUDA_TRY(cudaEventRecord(e->ev_fork, e->stream));
            CUDA_TRY(cudaStreamWaitEvent(e->stream2, e->ev_fork, 0));
            sps = e->stream2;
            forked = true;
        }
        if (e->timing && e->time_pending < sdr_engine::kTimeSets) cudaEventRecord(e->tev_spec[e->time_pending][0], sps);
#endif
#ifndef SDR_EMU
        // the previous block's D2H (result stream) still reads the device outputs these kernels overwrite
        if (!spec_jobs.empty() && prev_q >= 0 && e->res_spec_rec[prev_q]) CUDA_TRY(cudaStreamWaitEvent(sps, e->ev_res_spec[prev_q], 0));
#endif
        for (auto &job : spec_jobs) {
            if (job.view) {
                // zoomed view: its frames -> oscillator mix into the hidden channel's level-0 stream -> that channel's half-bands
                // and polyphase arm (the tail kernel, on its own slot) -> the transforms read its decimated stream
#ifndef SDR_EMU
                if (job.wait_ready && job.vc->ev_ready) CUDA_TRY(cudaStreamWaitEvent(sps, job.vc->ev_ready, 0));
                const long long tot = (long long)job.mix.F * job.mix.D;
                k_view_mix<<<(unsigned)std::min<long long>((tot + 255) / 256, 2048), 256, 0, sps>>>(job.mix);
#else
                for (long long g = 0; g < (long long)job.mix.F * job.mix.D; ++g) view_mix_thread(job.mix, g);
#endif
                e->launches += 1;
                TailArgs tv;
                tv.cs = e->d_cs; tv.cd = e->d_cd2[p]; tv.st = e->d_st; tv.taps = e->d_taps; tv.nco = e->d_nco;
                tv.results = e->d_results; tv.n = job.n_view; tv.first = job.vc->index; tv.view_pass = 1; tv.dbg = nullptr;
                launch_tail(tv, 1, sps, e->launches);
            }
            int r = launch_spectrum(job.a, sps, e->launches);
            if (r) return r;
        }
#ifndef SDR_EMU
        if (e->timing && e->time_pending < sdr_engine::kTimeSets) cudaEventRecord(e->tev_spec[e->time_pending][1], sps);
        if (forked) CUDA_TRY(cudaEventRecord(e->ev_join, e->stream2));
#endif
        if (to_host && !spec_jobs.empty()) {
            sdr_stream_t rs = sps;
#ifndef SDR_EMU
            // results leave on their own stream: the compute streams (and the next block's front kernel behind the
            // join) do not wait for PCIe
            CUDA_TRY(cudaEventRecord(e->ev_spec_k, sps));
            rs = e->stream_res;
            CUDA_TRY(cudaStreamWaitEvent(rs, e->ev_spec_k, 0));
#endif
            for (auto &job : spec_jobs) {
                sdr_spectrum *s = job.s;
                be_d2h(s->hp(q), s->d_points, (size_t)s->frames_last * s->fft_size * sizeof(float), rs);
                be_d2h(s->hl(q), s->d_lines, (size_t)s->frames_last * s->fft_size, rs);
                be_d2h(s->hc(q), s->d_frame_cf_maa, (size_t)s->frames_last * sizeof(double2), rs);
            }
#ifndef SDR_EMU
            CUDA_TRY(cudaEventRecord(e->ev_res_spec[q], rs)); e->res_spec_rec[q] = true;
#endif
        }
        return SDR_OK;
    };
    // everything enqueued after this submit on the main stream (next block's history slide, DC blocker) must wait for the
    // side stream: join_side() is called on every successful return path below
    auto join_side = [&]() -> int {
#ifndef SDR_EMU
        if (forked) CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_join, 0));
        CUDA_TRY(cudaEventRecord(e->ev_consumed[p], e->stream));   // this block's readers are done once the stream gets here
#endif
        return SDR_OK;
    };

    time_mark(e, 2);
    // 4. channels: per-block oscillator bookkeeping (DemodulatorPreThread.cpp:154-163, 186-195)
    int n_slots = (int)e->chans.size();
    if (n_slots == 0) {
        { int r = params_ready(); if (r) return r; }
        { int r = run_spectra(); if (r) return r; }
        time_mark(e, 3); time_mark(e, 4);
#ifndef SDR_EMU
        if (e->timing && e->time_pending < sdr_engine::kTimeSets) {
            cudaEventRecord(e->tev_tail[e->time_pending][0], e->stream);
            cudaEventRecord(e->tev_tail[e->time_pending][1], e->stream);
        }
#endif
        if (e->timing && e->time_pending < sdr_engine::kTimeSets) e->time_pending++;
#ifndef SDR_EMU
        CUDA_TRY(cudaEventRecord(e->ring_ev[slot], e->stream));
#endif
        return join_side();
    }
    std::vector<int> by_k[kFrontMaxK + 1];
    size_t live_now = 0;
    for (auto *c : e->chans) live_now += c != nullptr;
    const int front_reserved = (e->exact_dc && !pfb) ? dc_reserved_sms(live_now) : 0;

    for (int i = 0; i < n_slots; ++i) {
        sdr_chan *c = e->chans[i];
        ChanDyn &d = h_cd[i];
        memset(&d, 0, sizeof d);
        if (!c || c->standalone) continue;
        d.valid = 1;
        if (c->fresh) {                                          // created since the last block: its memory is being prepared on
            c->fresh = false;                                    // the admin stream -- a device-side wait, the host never blocks
#ifndef SDR_EMU
            if (c->ev_ready) CUDA_TRY(cudaStreamWaitEvent(e->stream, c->ev_ready, 0));
#endif
        }
        if (c->cs_dirty) {                                       // sdr_chan_want_iq on a live channel: only that flag differs
            c->cs_dirty = false;
            e->h_cs_stage[i] = c->cs;
            upload_params(e->d_cs + i, e->h_cs_stage + i, sizeof(ChanStatic), par_stream, e->launches);
        }
        // Consumed history that is not contiguous in the stream (a skipped block inside the front kernel's history window,
        // only possible with blocks shorter than that window): keep extending the private raw tail block by block until a
        // consumed block is long enough to hold the whole window again.
        if (c && !c->standalone && c->roll_pending > 0) {
            const int n2 = c->roll_pending;
            const float2 *prev_tail = nullptr;                    // the n2 samples of the previous block = end of the stream history
            if (!pfb) prev_tail = blk - n2;
            else if (c->pfb_slot_prev >= 0) prev_tail = pfb_streams + (size_t)c->pfb_slot_prev * e->pfb_stride + kBlkHist - n2;
            if (prev_tail) {
                float2 *scr = c->d_priv_hist + 2 * (size_t)kBlkHist;
                be_d2d(scr, c->d_priv_hist + n2, (size_t)(kBlkHist - n2) * sizeof(float2), e->stream);
                be_d2d(c->d_priv_hist, scr, (size_t)(kBlkHist - n2) * sizeof(float2), e->stream);
                be_d2d(c->d_priv_hist + (kBlkHist - n2), prev_tail, (size_t)n2 * sizeof(float2), e->stream);
                c->priv_valid = true;
            }
            c->roll_pending = 0;
        }
        // a channel that sits this block out keeps the tail of the last block it consumed (its filters do not advance)
        // (in channelizer mode the tail lives in the stream the channel last read: the history region of that slot)
        auto keep_tail = [&]() {
            if (!c->priv_valid && c->consumed > 0) {
                const float2 *src = blk - kBlkHist;
                if (pfb) src = c->pfb_slot_prev >= 0 ? pfb_streams + (size_t)c->pfb_slot_prev * e->pfb_stride : nullptr;
                if (src) { be_d2d(c->d_priv_hist, src, kBlkHist * sizeof(float2), e->stream); c->priv_valid = true; }
            }
        };
        if (c->sample_rate != rate_ch) { keep_tail(); continue; }                  // `inp->sampleRate == currentSampleRate` (:167)
        long long in_freq = center;                                                 // inp->frequency
        int stream_slot = -1;
        if (pfb) {
            const int ci = pfb_pick[i];                                             // SDRPostThread::getChannelAt, computed above
            if (ci < 0) { keep_tail(); continue; }                                  // no channel: the demodulator gets no data this block
            in_freq = chan_centers[ci];
            // channel 0 is read after its DC blocker; the extra channel wraps onto numChannels/2 (:357-361)
            stream_slot = ci == 0 ? pfb_M : (ci == pfb_M ? pfb_M / 2 : ci);
            if (c->pfb_slot_prev >= 0 && c->pfb_slot_prev != stream_slot) keep_tail();   // it hops to another stream: remember the old one's tail
        }
        const long long shift = c->frequency - in_freq;
        const long long lim = (long long)(int)((double)(rate_ch / 2) * 1.5);
        if (shift != c->shift_frequency) {
            c->shift_frequency = shift;
            if (std::llabs(shift) <= lim) c->dtheta = nco_phase_step(nco_shift_word_arg(shift, rate_ch));
        }
        if (std::llabs(c->shift_frequency) > lim) { keep_tail(); continue; }   // block skipped for this channel (:161-164)
        const int dir = c->shift_frequency < 0 ? +1 : (c->shift_frequency > 0 ? -1 : 0);
        d.squelch_enabled = c->squelch_enabled ? 1 : 0; d.squelch_level = c->squelch_level;
        d.active = 1; d.g0 = c->consumed; d.theta0 = c->theta; d.dtheta = c->dtheta; d.dtheta_prev = c->dtheta_prev;
        d.dir = dir; d.dir_prev = c->dir_prev;
        d.hist = c->priv_valid ? c->d_priv_hist + kBlkHist : nullptr;
        c->priv_valid = false;
        const bool hopped = pfb && c->pfb_slot_prev >= 0 && c->pfb_slot_prev != stream_slot;
        if (pfb) { d.xin = pfb_streams + (size_t)stream_slot * e->pfb_stride + kBlkHist; c->pfb_slot_prev = stream_slot; }
        // Rare path (dsp_core.h, premix_thread): the history window of the front kernel reaches back over more than one
        // oscillator setting -- blocks shorter than kFrontWarm * 2^K samples around a retune.  Mix the history with the
        // settings that were in effect and hand it over as "already mixed".
        const bool hist_private_raw = d.hist != nullptr;
        if (!hopped && c->cs.K1 > 0 && !c->recent.empty()) {
            const int window = (kFrontWarm + 1) << c->cs.K1;
            int covered = 0;
            bool mixed_settings = false;
            for (const PremixSeg &sg : c->recent) {
                if (covered >= window) break;
                if (sg.dir != c->recent[0].dir || (sg.dir && sg.dtheta != c->recent[0].dtheta)) mixed_settings = true;
                covered += sg.n;
            }
            if (mixed_settings) {
                PremixArgs pm;
                pm.x = d.hist ? d.hist : (pfb ? d.xin : blk);       // private raw tail if the stream history is not the consumed one
                pm.out = c->d_priv_hist + 2 * (size_t)kBlkHist; pm.nco = e->d_nco; pm.theta0 = c->theta;
                pm.count = std::min<int>(kBlkHist, std::min<long long>(c->consumed, (long long)window + 64));
                pm.nseg = (int)std::min<size_t>(c->recent.size(), (size_t)kPremixSegs);
                for (int k = 0; k < pm.nseg; ++k) pm.seg[k] = c->recent[k];
#ifndef SDR_EMU
                k_premix_hist<<<(pm.count + 127) / 128, 128, 0, e->stream>>>(pm);
#else
                for (int t = 1; t <= pm.count; ++t) premix_thread(pm, t);
#endif
                e->launches += 1;
                d.hist = c->d_priv_hist + 2 * (size_t)kBlkHist;
                d.dir_prev = 0; d.dtheta_prev = 0;                       // the history is mixed already: read it through table entry 0
            }
        }
        d.lvl0_sel = c->lvl0_sel;                                       // consecutive processed blocks alternate level-0 buffers
        c->lvl0_sel ^= 1;
        if (dir) c->theta += n_ch * c->dtheta;                          // mix_block advances the phase by one step per input sample
        c->dtheta_prev = c->dtheta; c->dir_prev = dir;
        // the next block's window may reach back before this block: if this block's own history was private, it stays private
        if (hist_private_raw && (int)n_ch < ((kFrontWarm + 2) << kFrontMaxK)) c->roll_pending = (int)n_ch;
        c->consumed += n_ch;
        c->recent.insert(c->recent.begin(), PremixSeg{(int)n_ch, dir ? c->dtheta : 0u, dir});
        if (c->recent.size() > (size_t)kPremixSegs) c->recent.pop_back();
        by_k[c->cs.K1].push_back(i);
    }
    for (auto &vd : view_dyn) h_cd[vd.first] = vd.second;       // the zoomed views' hidden channels (their own tail launch)
    ChanDyn *d_cd = e->d_cd2[p];
#ifndef SDR_EMU
    // d_cd2[p] and the level-0 buffers this block's front kernel writes were last read by the tail of two blocks ago
    CUDA_TRY(cudaStreamWaitEvent(e->stream, e->ev_tail_done[p], 0));
#endif
#ifndef SDR_EMU
    CUDA_TRY(cudaStreamWaitEvent(par_stream, e->ev_tail_done[p], 0));   // d_cd2[p] was last read by the tail of two blocks ago
#endif
    upload_params(d_cd, h_cd, (size_t)n_slots * sizeof(ChanDyn), par_stream, e->launches);
    int goff = 0;
    float h3[6];
    halfband_branch_taps(3, 65.0f, h3);
    struct FrontGroup { int K, cg; int *d_grp; };
    std::vector<FrontGroup> groups;
    for (int K = 0; K <= kFrontMaxK; ++K) {
        const int cg = (int)by_k[K].size();
        if (!cg) continue;
        memcpy(h_group + goff, by_k[K].data(), cg * sizeof(int));
        int *d_grp = e->d_group + (size_t)p * sdr_engine::kMaxChan + goff;
        upload_params(d_grp, h_group + goff, cg * sizeof(int), par_stream, e->launches);
        groups.push_back(FrontGroup{K, cg, d_grp});
        goff += cg;
    }
    { int r = params_ready(); if (r) return r; }
    for (const FrontGroup &g : groups) {
        const int K = g.K, cg = g.cg;
        FrontArgs fa;
        fa.x = blk; fa.n = (int)n_ch; fa.nco = e->d_nco; fa.cs = e->d_cs; fa.cd = d_cd; fa.group = g.d_grp; fa.cg = cg;
        const int n_chunks = (cg + 31) / 32;
        const long long nq = ((long long)n_ch >> K) + 1;
        int n_span = std::max(1, ((148 - front_reserved) * kFrontWarps) / n_chunks);   // one wave of CTAs on the SMs the front uses
        // every span pays kFrontWarm warm-up super-steps; with few channels, parallelism is worth more than the overhead
        const long long min_span = (K == 0) ? 256 : kFrontWarm;
        if ((long long)n_span * min_span > nq) n_span = (int)std::max<long long>(1, nq / min_span);
        fa.n_span = n_span;
        fa.span = (int)((nq + n_span - 1) / n_span);
        fa.c0 = h3[0]; fa.c1 = h3[1]; fa.c2 = h3[2];
        fa.dbg = e->d_front_dbg;
        int fr = launch_front(K, fa, n_chunks * n_span, front_reserved, e->stream, e->launches);
        if (fr) return fr;
    }
    { int r = run_spectra(); if (r) return r; }                     // side stream, after the front kernel got the SMs
    TailArgs ta;
    ta.cs = e->d_cs; ta.cd = d_cd; ta.st = e->d_st; ta.taps = e->d_taps; ta.nco = e->d_nco;
    ta.results = e->d_results; ta.n = (int)n_ch; ta.first = 0; ta.view_pass = 0; ta.dbg = e->d_front_dbg ? e->d_front_dbg + 4 * 1024 : nullptr;
    // the tail indexes channels by slot; free slots carry valid = 0 and are not touched
    time_mark(e, 3);
    // the channel tail (and the results' D2H) go to their own stream: they overlap the next block's front kernel
    sdr_stream_t ts = e->stream;
#ifndef SDR_EMU
    CUDA_TRY(cudaEventRecord(e->ring_ev[slot], e->stream));
    CUDA_TRY(cudaEventRecord(e->ev_front_done, e->stream));
    ts = e->stream_tail;
    CUDA_TRY(cudaStreamWaitEvent(ts, e->ev_front_done, 0));
    if (e->timing && e->time_pending < sdr_engine::kTimeSets) cudaEventRecord(e->tev_tail[e->time_pending][0], ts);
#endif
#ifndef SDR_EMU
    if (prev_q >= 0 && e->res_tail_rec[prev_q]) CUDA_TRY(cudaStreamWaitEvent(ts, e->ev_res_tail[prev_q], 0));   // its D2H reads d_results
#endif
    launch_tail(ta, n_slots, ts, e->launches);
    mixers_feed(e, ts);                                            // bound channels' audio joins their mixer queues (device side)
    if (to_host && e->results_bytes) {
        size_t total = 0;
        for (auto *c : e->chans) if (c) total = std::max(total, (size_t)c->cs.res_off + c->slot_bytes);
        sdr_stream_t rs = ts;
#ifndef SDR_EMU
        CUDA_TRY(cudaEventRecord(e->ev_tail_k, ts));
        rs = e->stream_res;
        CUDA_TRY(cudaStreamWaitEvent(rs, e->ev_tail_k, 0));
#endif
        be_d2h(e->hres(q), e->d_results, total, rs);
#ifndef SDR_EMU
        CUDA_TRY(cudaEventRecord(e->ev_res_tail[q], rs)); e->res_tail_rec[q] = true;
#endif
    }
#ifndef SDR_EMU
    if (e->timing && e->time_pending < sdr_engine::kTimeSets) cudaEventRecord(e->tev_tail[e->time_pending][1], ts);
    CUDA_TRY(cudaEventRecord(e->ev_tail_done[p], ts));
#endif
    time_mark(e, 4);
    if (e->timing && e->time_pending < sdr_engine::kTimeSets) e->time_pending++;
#ifndef SDR_EMU
    {   // a kernel that could not be launched (shared-memory limit, bad configuration) must not pass silently
        const cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) return fail(SDR_ERR_CUDA, std::string("kernel launch failed in submit: ") + cudaGetErrorString(le));
    }
#endif
    return join_side();
}

// a failed submit after the argument checks means a CUDA error with host-side state already advanced: drain and refuse
static int submit_guard(sdr_engine *e, int r) {
    if (r == SDR_ERR_CUDA) {
        const std::string keep = g_err;
        sync_all(e);
#ifndef SDR_EMU
        cudaGetLastError();
#endif
        e->poisoned = true;
        g_err = keep;
    }
    return r;
}

// Pinned memory for the host side of the boundary (input blocks the producer fills, result arrays): a block submitted from
// such memory is copied by DMA straight from where it lies.
extern "C" int sdr_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(SDR_ERR_ARG, "sdr_host_alloc: null argument");
    *out = nullptr;
    if (be_host_alloc(out, bytes)) return fail(SDR_ERR_NOMEM, "sdr_host_alloc: pinned allocation failed");
    return SDR_OK;
}
extern "C" int sdr_host_free(void *p) { be_host_free(p); return SDR_OK; }

// The engine's own pinned input ring: the buffer the NEXT sdr_engine_submit will use.  A producer that fills it in place
// (SDRThread::readStream writes its samples there instead of into a std::vector) saves the staging copy; the buffer is
// free to write as soon as this call returns.
extern "C" int sdr_engine_host_buffer(sdr_engine *e, sdr_cf32 **buf, unsigned *cap) {
    if (!e || !buf) return fail(SDR_ERR_ARG, "sdr_engine_host_buffer: null argument");
    std::lock_guard<std::mutex> lk(e->mu);
    const int k = e->in_sel % sdr_engine::kInRing;
#ifndef SDR_EMU
    if (e->in_rec[k] && cudaEventSynchronize(e->ev_h2d[k]) != cudaSuccess) return fail(SDR_ERR_CUDA, "sdr_engine_host_buffer: wait failed");
#endif
    *buf = e->h_in[k];
    if (cap) *cap = e->max_block;
    return SDR_OK;
}

// SDRPostThread::run hands one block to the path (SDRPostThread.cpp:178-187).  Never waits for the GPU in steady state: the
// block is staged through a ring of pinned buffers (skipped when `iq` already is pinned memory: the engine's own ring
// buffer, sdr_host_alloc memory or anything cudaHostRegister'ed), copied by DMA on a dedicated stream into a ring of device
// buffers, and the DC stream picks it up from there.  A pinned source must stay untouched until two further blocks have been
// submitted or sdr_engine_wait / sdr_engine_wait_block(0) has returned for this one.
extern "C" int sdr_engine_submit(sdr_engine *e, const sdr_cf32 *iq, unsigned n, long long center, long long rate) {
    if (!e || !iq) return fail(SDR_ERR_ARG, "sdr_engine_submit: null argument");
    if (n == 0 || n > e->max_block) return fail(SDR_ERR_ARG, "sdr_engine_submit: block size out of range");
    std::lock_guard<std::mutex> lk(e->mu);
#ifndef SDR_EMU
    if (e->poisoned) return fail(SDR_ERR_STATE, "engine is in a failed state after a CUDA error: destroy it");
    cudaSetDevice(e->device);
    const int k = e->in_sel++ % sdr_engine::kInRing;
    // this slot's previous H2D (three submits ago) must have left the pinned buffer; in steady state it finished long ago
    if (e->in_rec[k]) CUDA_TRY(cudaEventSynchronize(e->ev_h2d[k]));
    const sdr_cf32 *src = iq;
    bool pinned = iq == e->h_in[k];
    if (!pinned) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, iq) == cudaSuccess) pinned = pa.type == cudaMemoryTypeHost;
        else cudaGetLastError();
    }
    if (!pinned) { memcpy(e->h_in[k], iq, (size_t)n * sizeof(sdr_cf32)); src = e->h_in[k]; }
    // the device buffer of this slot was read by the DC stream three blocks ago
    if (e->in_rec[k]) CUDA_TRY(cudaStreamWaitEvent(e->stream_h2d, e->ev_in_free[k], 0));
    CUDA_TRY(cudaMemcpyAsync(e->d_in[k], src, (size_t)n * sizeof(sdr_cf32), cudaMemcpyHostToDevice, e->stream_h2d));
    CUDA_TRY(cudaEventRecord(e->ev_h2d[k], e->stream_h2d));
    e->in_rec[k] = true;
    return submit_guard(e, submit_common(e, e->d_in[k], nullptr, n, center, rate, true, 2, e->ev_h2d[k], e->ev_in_free[k]));
#else
    return submit_common(e, nullptr, iq, n, center, rate, true, 1, nullptr);
#endif
}

extern "C" int sdr_engine_submit_device(sdr_engine *e, const void *d_iq, unsigned n, long long center, long long rate, int to_host) {
    if (!e || !d_iq) return fail(SDR_ERR_ARG, "sdr_engine_submit_device: null argument");
    if (n == 0 || n > e->max_block) return fail(SDR_ERR_ARG, "sdr_engine_submit_device: block size out of range");
    std::lock_guard<std::mutex> lk(e->mu);
    return submit_guard(e, submit_common(e, (const float2 *)d_iq, nullptr, n, center, rate, to_host != 0, 0, nullptr));
}

extern "C" int sdr_engine_submit_device_async(sdr_engine *e, const void *d_iq, unsigned n, long long center, long long rate,
                                              int to_host, void *ready_event) {
    if (!e || !d_iq) return fail(SDR_ERR_ARG, "sdr_engine_submit_device_async: null argument");
    if (n == 0 || n > e->max_block) return fail(SDR_ERR_ARG, "sdr_engine_submit_device_async: block size out of range");
    std::lock_guard<std::mutex> lk(e->mu);
    return submit_guard(e, submit_common(e, (const float2 *)d_iq, nullptr, n, center, rate, to_host != 0, ready_event ? 2 : 1, ready_event));
}

// SDRPostThread::setChannelizerType (SDRPostThread.cpp:142-144), restricted to what is built: 0 = every block through
// runSingleCH (default), 1 = SDRPostPFBCH: blocks whose rate gives more than one channel go through the analysis filterbank
// and every demodulator reads its filterbank channel at sampleRate / numChannels.  Set before the first block.
extern "C" int sdr_engine_set_channelizer(sdr_engine *e, int mode) {
    if (!e) return fail(SDR_ERR_ARG, "null engine");
    if (mode < 0 || mode > 2) return fail(SDR_ERR_UNSUPPORTED, "sdr_engine_set_channelizer: mode must be 0 (single channel), 1 (PFBCH) or 2 (PFBCH2)");
    std::lock_guard<std::mutex> lk(e->mu);
    if (e->have_block && mode != e->chan_mode) return fail(SDR_ERR_STATE, "sdr_engine_set_channelizer: set the mode before the first block");
    e->chan_mode = mode;
    return SDR_OK;
}

// Channelizer mode hands every demodulator block / numChannels samples (twice that for PFBCH2), not the whole block: tell the
// engine before creating channels so that their buffers and result slots are sized for that (default: max_block).
extern "C" int sdr_engine_set_channel_block_max(sdr_engine *e, unsigned n) {
    if (!e || n == 0) return fail(SDR_ERR_ARG, "sdr_engine_set_channel_block_max: bad argument");
    std::lock_guard<std::mutex> lk(e->mu);
    for (auto *c : e->chans) if (c) return fail(SDR_ERR_STATE, "sdr_engine_set_channel_block_max: call before creating channels");
    e->chan_block_max = n < e->max_block ? n : e->max_block;
    return SDR_OK;
}

// what the demodulators of a channelizer-mode engine must be created with: numChannels and chanBw = rate / numChannels
extern "C" int sdr_engine_channel_plan(long long sample_rate, int *num_channels, long long *channel_rate) {
    const int m = optimal_channel_count(sample_rate);
    if (num_channels) *num_channels = m;
    if (channel_rate) *channel_rate = sample_rate / m;
    return SDR_OK;
}

extern "C" int sdr_engine_wait(sdr_engine *e) {
    if (!e) return fail(SDR_ERR_ARG, "null engine");
    use_device(e);
    if (be_sync(e->stream)) return fail(SDR_ERR_CUDA, std::string("stream sync failed: ") + be_last_error());
#ifndef SDR_EMU
    if (cudaStreamSynchronize(e->stream_tail) != cudaSuccess) return fail(SDR_ERR_CUDA, std::string("tail stream sync failed: ") + be_last_error());
    if (cudaStreamSynchronize(e->stream2) != cudaSuccess) return fail(SDR_ERR_CUDA, std::string("side stream sync failed: ") + be_last_error());
    if (cudaStreamSynchronize(e->stream_res) != cudaSuccess) return fail(SDR_ERR_CUDA, std::string("result stream sync failed: ") + be_last_error());
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(SDR_ERR_CUDA, std::string("kernel error: ") + cudaGetErrorString(err));
#endif
    {
        std::lock_guard<std::mutex> lk(e->mu);
        e->fetch_sel = e->res_sel;
        time_collect(e);
    }
    return SDR_OK;
}

// Wait only until the results of the block submitted `lag` submits ago (0 = latest, up to kResBuf - 1) are on the host, and
// make the fetch calls read that block.  With lag = 2 the caller keeps two blocks in flight: submit(b+2), wait_block(2),
// fetch b -- the GPU then never idles while the host copies results out.
extern "C" int sdr_engine_wait_block(sdr_engine *e, int lag) {
    if (!e) return fail(SDR_ERR_ARG, "null engine");
    if (lag < 0 || lag >= kResBuf) return fail(SDR_ERR_ARG, "sdr_engine_wait_block: lag must be 0, 1 or 2");
    int q;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        q = (e->res_sel - lag + kResBuf) % kResBuf;
        if (!e->res_on_host[q]) return fail(SDR_ERR_STATE, "sdr_engine_wait_block: that block was not submitted with host results");
    }
#ifndef SDR_EMU
    if (e->res_tail_rec[q] && cudaEventSynchronize(e->ev_res_tail[q]) != cudaSuccess) return fail(SDR_ERR_CUDA, std::string("result wait failed: ") + be_last_error());
    if (e->res_spec_rec[q] && cudaEventSynchronize(e->ev_res_spec[q]) != cudaSuccess) return fail(SDR_ERR_CUDA, std::string("result wait failed: ") + be_last_error());
#endif
    std::lock_guard<std::mutex> lk(e->mu);
    e->fetch_sel = q;
    return SDR_OK;
}

extern "C" int sdr_engine_fetch_block(sdr_engine *e, sdr_cf32 *out, unsigned cap, unsigned *n) {
    if (!e || !out) return fail(SDR_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    unsigned m = std::min(cap, e->last_n);
    be_d2h(out, e->d_blkbuf[e->parity] + e->hist_cap, (size_t)m * sizeof(sdr_cf32), e->stream);
    be_sync(e->stream);
    if (n) *n = m;
    return SDR_OK;
}

extern "C" int sdr_chan_fetch_audio(sdr_chan *c, float *audio, unsigned cap, unsigned *n_audio, double *level_db, float *peak, int *skipped) {
    if (!c) return fail(SDR_ERR_ARG, "null channel");
    sdr_engine *e = c->eng;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->res_on_host[e->fetch_sel] || !e->h_results) return fail(SDR_ERR_STATE, "no host results for the last block");
    const unsigned char *slot = e->hres(e->fetch_sel) + c->cs.res_off;
    const ChanResultHeader *h = reinterpret_cast<const ChanResultHeader *>(slot);
    unsigned m = std::min(cap, std::min(h->n_audio, (unsigned)c->cs.res_audio_cap));
    if (audio) memcpy(audio, slot + kResHdrBytes, (size_t)m * sizeof(float));
    if (n_audio) *n_audio = h->n_audio;
    if (level_db) *level_db = h->level_db;
    if (peak) *peak = h->peak;
    if (skipped) *skipped = h->skipped;
    return SDR_OK;
}

// getSignalLevel / getSignalFloor / getSignalCeil (DemodulatorThread.cpp:372-382) after the fetched block, and the block's
// AudioThreadInput::is_squelch_active (:198, :235-238)
extern "C" int sdr_chan_fetch_meter(sdr_chan *c, float *signal_level, float *signal_floor, float *signal_ceil, int *squelched) {
    if (!c) return fail(SDR_ERR_ARG, "null channel");
    sdr_engine *e = c->eng;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->res_on_host[e->fetch_sel] || !e->h_results) return fail(SDR_ERR_STATE, "no host results for the last block");
    const ChanResultHeader *h = reinterpret_cast<const ChanResultHeader *>(e->hres(e->fetch_sel) + c->cs.res_off);
    if (signal_level) *signal_level = h->sig_level;
    if (signal_floor) *signal_floor = h->sig_floor;
    if (signal_ceil) *signal_ceil = h->sig_ceil;
    if (squelched) *squelched = h->squelched;
    return SDR_OK;
}

extern "C" int sdr_engine_fetch_audio_batch(sdr_engine *e, sdr_chan *const *chans, unsigned n_chan, float *audio, unsigned stride,
                                            unsigned *n_audio, double *level_db, float *peak) {
    if (!e || !chans) return fail(SDR_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->res_on_host[e->fetch_sel] || !e->h_results) return fail(SDR_ERR_STATE, "no host results for the last block");
    for (unsigned i = 0; i < n_chan; ++i) {
        const sdr_chan *c = chans[i];
        if (!c || c->eng != e) return fail(SDR_ERR_ARG, "channel does not belong to this engine");
        const unsigned char *slot = e->hres(e->fetch_sel) + c->cs.res_off;
        const ChanResultHeader *h = reinterpret_cast<const ChanResultHeader *>(slot);
        unsigned m = std::min(stride, std::min(h->n_audio, (unsigned)c->cs.res_audio_cap));
        if (audio) memcpy(audio + (size_t)i * stride, slot + kResHdrBytes, (size_t)m * sizeof(float));
        if (n_audio) n_audio[i] = h->n_audio;
        if (level_db) level_db[i] = h->level_db;
        if (peak) peak[i] = h->peak;
    }
    return SDR_OK;
}

// Zero-copy form of the batch fetch: pointers into the pinned host buffer the block's results were DMA'd to.  They stay valid
// until two further blocks have been submitted (the ring of three result buffers); a host that hands the samples on at once
// (AudioThread mixes them into its own buffer, src/audio/AudioThread.cpp:88-243) needs no intermediate copy.
extern "C" int sdr_engine_view_audio_batch(sdr_engine *e, sdr_chan *const *chans, unsigned n_chan, const float **audio,
                                           unsigned *n_audio, double *level_db, float *peak) {
    if (!e || !chans) return fail(SDR_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->res_on_host[e->fetch_sel] || !e->h_results) return fail(SDR_ERR_STATE, "no host results for the last block");
    for (unsigned i = 0; i < n_chan; ++i) {
        const sdr_chan *c = chans[i];
        if (!c || c->eng != e) return fail(SDR_ERR_ARG, "channel does not belong to this engine");
        const unsigned char *slot = e->hres(e->fetch_sel) + c->cs.res_off;
        const ChanResultHeader *h = reinterpret_cast<const ChanResultHeader *>(slot);
        if (audio) audio[i] = reinterpret_cast<const float *>(slot + kResHdrBytes);
        if (n_audio) n_audio[i] = std::min(h->n_audio, (unsigned)c->cs.res_audio_cap);
        if (level_db) level_db[i] = h->level_db;
        if (peak) peak[i] = h->peak;
    }
    return SDR_OK;
}

extern "C" int sdr_chan_fetch_iq(sdr_chan *c, sdr_cf32 *iq, unsigned cap, unsigned *n_iq) {
    if (!c) return fail(SDR_ERR_ARG, "null channel");
    sdr_engine *e = c->eng;
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->res_on_host[e->fetch_sel] || !e->h_results) return fail(SDR_ERR_STATE, "no host results for the last block");
    if (!c->cs.want_iq) return fail(SDR_ERR_STATE, "sdr_chan_want_iq was not enabled");
    const unsigned char *slot = e->hres(e->fetch_sel) + c->cs.res_off;
    const ChanResultHeader *h = reinterpret_cast<const ChanResultHeader *>(slot);
    unsigned m = std::min(cap, std::min(h->n_iq, (unsigned)c->cs.res_iq_cap));
    if (iq) memcpy(iq, slot + kResHdrBytes + (size_t)c->cs.res_audio_cap * 4, (size_t)m * sizeof(sdr_cf32));
    if (n_iq) *n_iq = h->n_iq;
    return SDR_OK;
}

extern "C" int sdr_spectrum_frames(sdr_spectrum *s, unsigned *n) {
    if (!s || !n) return fail(SDR_ERR_ARG, "null argument");
    *n = s->frames_buf[s->eng->fetch_sel];
    return SDR_OK;
}

extern "C" int sdr_spectrum_fetch(sdr_spectrum *s, unsigned idx, float *points_y, uint8_t *line, double *cf) {
    if (!s) return fail(SDR_ERR_ARG, "null spectrum");
    std::lock_guard<std::mutex> lk(s->eng->mu);
    const int q = s->eng->fetch_sel;
    if (idx >= s->frames_buf[q]) return fail(SDR_ERR_ARG, "frame index out of range");
    if (!s->eng->res_on_host[q]) return fail(SDR_ERR_STATE, "no host results for the last block");
    if (points_y) memcpy(points_y, s->hp(q) + (size_t)idx * s->fft_size, s->fft_size * sizeof(float));
    if (line) memcpy(line, s->hl(q) + (size_t)idx * s->fft_size, s->fft_size);
    if (cf) { cf[0] = s->hc(q)[idx].x; cf[1] = s->hc(q)[idx].y; }
    return SDR_OK;
}

extern "C" int sdr_spectrum_fetch_all(sdr_spectrum *s, unsigned cap_frames, float *points_y, uint8_t *lines, double *cf, unsigned *n_frames) {
    if (!s) return fail(SDR_ERR_ARG, "null spectrum");
    std::lock_guard<std::mutex> lk(s->eng->mu);
    const int q = s->eng->fetch_sel;
    if (!s->eng->res_on_host[q]) return fail(SDR_ERR_STATE, "no host results for the last block");
    const unsigned m = std::min(cap_frames, s->frames_buf[q]);
    if (points_y) memcpy(points_y, s->hp(q), (size_t)m * s->fft_size * sizeof(float));
    if (lines) memcpy(lines, s->hl(q), (size_t)m * s->fft_size);
    if (cf) for (unsigned f = 0; f < m; ++f) { cf[2 * f] = s->hc(q)[f].x; cf[2 * f + 1] = s->hc(q)[f].y; }
    if (n_frames) *n_frames = s->frames_buf[q];
    return SDR_OK;
}

// zero-copy form: the frames of the fetched block where the DMA left them ([n_frames][fft_size] each), same lifetime as above
extern "C" int sdr_spectrum_view(sdr_spectrum *s, const float **points_y, const uint8_t **lines, unsigned *n_frames) {
    if (!s) return fail(SDR_ERR_ARG, "null spectrum");
    std::lock_guard<std::mutex> lk(s->eng->mu);
    const int q = s->eng->fetch_sel;
    if (!s->eng->res_on_host[q]) return fail(SDR_ERR_STATE, "no host results for the last block");
    if (points_y) *points_y = s->hp(q);
    if (lines) *lines = s->hl(q);
    if (n_frames) *n_frames = s->frames_buf[q];
    return SDR_OK;
}

extern "C" int sdr_spectrum_fetch_debug(sdr_spectrum *s, unsigned idx, float *mags, float *maa) {
    if (!s) return fail(SDR_ERR_ARG, "null spectrum");
    std::lock_guard<std::mutex> lk(s->eng->mu);
    use_device(s->eng);
    if (idx >= s->frames_last) return fail(SDR_ERR_ARG, "frame index out of range");
    be_d2h(s->h_dbg, s->d_mags + (size_t)idx * s->N, s->N * sizeof(float), s->eng->stream);
    be_d2h(s->h_dbg + s->N, s->d_maa + (size_t)idx * s->N, s->N * sizeof(float), s->eng->stream);
    be_sync(s->eng->stream);
    if (mags) memcpy(mags, s->h_dbg, s->N * sizeof(float));
    if (maa) memcpy(maa, s->h_dbg + s->N, s->N * sizeof(float));
    return SDR_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// polyphase filterbank channelizer: SDRPostThread::initPFBCH / runPFBCH, SDRPostThread.cpp:401-455
// ---------------------------------------------------------------------------------------------------------------------
struct sdr_pfb {
    sdr_engine *eng = nullptr;
    unsigned M = 0, p = 0, max_frames = 0;
    int osf = 1;                 // 1: firpfbch_crcf, 2: firpfbch2_crcf (one output frame per M/2 inputs)
    long long frame_base = 0;    // output frames produced since reset (osf 2: branch phase)
    size_t hist() const { return osf == 2 ? (size_t)(2 * p - 1) * (M / 2) : (size_t)(p - 1) * M; }
    unsigned step() const { return osf == 2 ? M / 2 : M; }      // input samples per output frame
    float *d_hq = nullptr; float2 *d_tw = nullptr;
    float2 *d_x = nullptr;       // [(p-1) M history | max_frames M new]
    float2 *d_y = nullptr;       // [max_frames M]
    float2 *d_scr = nullptr;     // [(p-1) M] history slide scratch
    float2 *h_io = nullptr;      // pinned staging, max_frames M
};

static void pfb_free(sdr_pfb *p) {
    be_free(p->d_hq); be_free(p->d_tw); be_free(p->d_x); be_free(p->d_y); be_free(p->d_scr); be_host_free(p->h_io);
}

extern "C" int sdr_pfb_optimal_channel_count(long long sample_rate) { return optimal_channel_count(sample_rate); }

static int pfb_create_impl(sdr_engine *e, int osf, unsigned num_channels, unsigned m, float as_db, unsigned max_frames, sdr_pfb **out);
extern "C" int sdr_pfb_create(sdr_engine *e, unsigned num_channels, unsigned m, float as_db, unsigned max_frames, sdr_pfb **out) {
    return pfb_create_impl(e, 1, num_channels, m, as_db, max_frames, out);
}
// firpfbch2_crcf_create_kaiser(LIQUID_ANALYZER, M, m, As) (SDRPostThread::initPFBCH2, SDRPostThread.cpp:457-470): one output
// frame of M samples per M/2 inputs; max_frames counts OUTPUT frames
extern "C" int sdr_pfb2_create(sdr_engine *e, unsigned num_channels, unsigned m, float as_db, unsigned max_frames, sdr_pfb **out) {
    return pfb_create_impl(e, 2, num_channels, m, as_db, max_frames, out);
}
static int pfb_create_impl(sdr_engine *e, int osf, unsigned num_channels, unsigned m, float as_db, unsigned max_frames, sdr_pfb **out) {
    if (!e || !out) return fail(SDR_ERR_ARG, "sdr_pfb_create: null argument");
    if (num_channels < 2 || (num_channels & 1u) || num_channels > (unsigned)kPfbMaxChannels || m < 1 || m > 8 || max_frames < 1)
        return fail(SDR_ERR_ARG, "sdr_pfb_create: num_channels must be even and in [2, 256], m in [1, 8]");
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    auto *p = new sdr_pfb();
    p->eng = e; p->M = num_channels; p->p = 2 * m; p->max_frames = max_frames; p->osf = osf;
    std::vector<float> hq, tw;
    if (osf == 2) pfb2_analysis_tables(num_channels, m, as_db, hq, tw);
    else pfb_analysis_tables(num_channels, m, as_db, hq, tw);
    const size_t hist = p->hist(), cap = (size_t)max_frames * p->M;
    int bad = 0;
    bad |= be_malloc((void **)&p->d_hq, hq.size() * sizeof(float));
    bad |= be_malloc((void **)&p->d_tw, tw.size() * sizeof(float));
    bad |= be_malloc((void **)&p->d_x, (hist + cap) * sizeof(float2));
    bad |= be_malloc((void **)&p->d_y, cap * sizeof(float2));
    bad |= be_malloc((void **)&p->d_scr, hist * sizeof(float2));
    bad |= be_host_alloc((void **)&p->h_io, cap * sizeof(float2));
    if (bad) { pfb_free(p); delete p; return fail(SDR_ERR_NOMEM, "sdr_pfb_create: allocation failed"); }
    be_h2d(p->d_hq, hq.data(), hq.size() * sizeof(float), e->stream);
    be_h2d(p->d_tw, tw.data(), tw.size() * sizeof(float), e->stream);
    be_memset(p->d_x, 0, hist * sizeof(float2), e->stream);      // firpfbch_crcf_reset: empty windows
    be_sync(e->stream);
    *out = p;
    return SDR_OK;
}

extern "C" int sdr_pfb_destroy(sdr_pfb *p) {
    if (!p) return SDR_OK;
    { std::lock_guard<std::mutex> lk(p->eng->mu); use_device(p->eng); be_sync(p->eng->stream); pfb_free(p); }
    delete p;
    return SDR_OK;
}

extern "C" int sdr_pfb_reset(sdr_pfb *p) {
    if (!p) return fail(SDR_ERR_ARG, "null channelizer");
    std::lock_guard<std::mutex> lk(p->eng->mu);
    use_device(p->eng);
    be_memset(p->d_x, 0, p->hist() * sizeof(float2), p->eng->stream);
    be_sync(p->eng->stream);
    p->frame_base = 0;
    return SDR_OK;
}

// device-resident form: d_in / d_out hold n = frames * M samples; history is carried inside the object
static int pfb_run(sdr_pfb *p, const float2 *d_in, unsigned n, float2 *d_out) {
    sdr_engine *e = p->eng;
    const unsigned F = n / p->step();
    const size_t hist = p->hist();
    float2 *xnew = p->d_x + hist;
    be_d2d(xnew, d_in, (size_t)n * sizeof(float2), e->stream);
    PfbArgs a;
    a.x = xnew; a.y = d_out; a.hq = p->d_hq; a.tw = p->d_tw; a.M = (int)p->M; a.p = (int)p->p; a.F = (int)F; a.chan_stride = 0;
    a.osf = p->osf; a.frame_base = p->frame_base;
    p->frame_base += F;
    const int tiles = ((int)F + kPfbTileFrames - 1) / kPfbTileFrames;
#ifndef SDR_EMU
    const size_t smem = pfb_smem_elems(a.M, a.p) * sizeof(float2);
    { int r = pfb_allow_smem(smem); if (r) return r; }
    k_pfb_analyze<<<tiles, kPfbThreads, smem, e->stream>>>(a);
#else
    std::vector<float2> sm(pfb_smem_elems(a.M, a.p));
    for (int t = 0; t < tiles; ++t)
        for (int ph = 0; ph < kPfbPhases; ++ph) for (int th = 0; th < kPfbThreads; ++th) pfb_phase(ph, a, t, th, kPfbThreads, sm.data());
#endif
    e->launches += 1;
    // slide: the last p-1 frames of [history | new] become the history of the next call (through scratch: the ranges overlap
    // when the call was shorter than the history)
    be_d2d(p->d_scr, p->d_x + (size_t)n, hist * sizeof(float2), e->stream);
    be_d2d(p->d_x, p->d_scr, hist * sizeof(float2), e->stream);
    return SDR_OK;
}

extern "C" int sdr_pfb_execute(sdr_pfb *p, const sdr_cf32 *iq, unsigned n, sdr_cf32 *out) {
    if (!p || !iq || !out) return fail(SDR_ERR_ARG, "sdr_pfb_execute: null argument");
    if (n == 0 || n % p->step() != 0 || n / p->step() > p->max_frames)
        return fail(SDR_ERR_ARG, "sdr_pfb_execute: n must be a non-zero multiple of num_channels (half of it for firpfbch2) within max_frames (SoapySDRThread.cpp:668-674)");
    sdr_engine *e = p->eng;
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    memcpy(p->h_io, iq, (size_t)n * sizeof(sdr_cf32));
    be_h2d(p->d_y, p->h_io, (size_t)n * sizeof(float2), e->stream);          // d_y doubles as the input landing buffer
    be_sync(e->stream);
    // input must not alias the output: run from a private copy inside d_x (pfb_run copies d_in there first)
    int r = pfb_run(p, p->d_y, n, p->d_y);
    if (r) return r;
    const size_t n_out = (size_t)n * (size_t)p->osf;                      // firpfbch2 is oversampled by two
    be_d2h(p->h_io, p->d_y, n_out * sizeof(float2), e->stream);
    if (be_sync(e->stream)) return fail(SDR_ERR_CUDA, std::string("sdr_pfb_execute: ") + be_last_error());
#ifndef SDR_EMU
    { cudaError_t err = cudaGetLastError(); if (err != cudaSuccess) return fail(SDR_ERR_CUDA, std::string("sdr_pfb_execute: ") + cudaGetErrorString(err)); }
#endif
    memcpy(out, p->h_io, n_out * sizeof(sdr_cf32));
    return SDR_OK;
}

extern "C" int sdr_pfb_execute_device(sdr_pfb *p, const void *d_iq, unsigned n, void *d_out) {
    if (!p || !d_iq || !d_out) return fail(SDR_ERR_ARG, "sdr_pfb_execute_device: null argument");
    if (n == 0 || n % p->step() != 0 || n / p->step() > p->max_frames) return fail(SDR_ERR_ARG, "sdr_pfb_execute_device: bad length");
    std::lock_guard<std::mutex> lk(p->eng->mu);
    use_device(p->eng);
    return pfb_run(p, (const float2 *)d_iq, n, (float2 *)d_out);
}

// ---------------------------------------------------------------------------------------------------------------------
// diagnostics: measured non-tensor FP32 issue peak of the device (bench.py's roofline denominator for the FP32-bound channel
// front kernel; MEASURED_PEAKS.json only carries HBM and tensor figures).  Eight independent FFMA chains per thread, 16 CTAs
// of 256 threads per SM, best of five launches timed with CUDA events.  Not part of the public header.
// ---------------------------------------------------------------------------------------------------------------------
#ifndef SDR_EMU
__global__ void __launch_bounds__(256) k_fp32_peak(float *out, int iters, float b, float c) {
    float a0 = threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
#endif
extern "C" int sdr_debug_fp32_peak(int device, double *tflops) {
    if (!tflops) return SDR_ERR_ARG;
#ifndef SDR_EMU
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int ctas = prop.multiProcessorCount * 16, iters = 2048;
    float *d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)ctas * 256 * sizeof(float)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0, 0);
        k_fp32_peak<<<ctas, 256>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = (double)ctas * 256.0 * iters * 16.0 * 8.0 * 2.0 / (ms * 1e-3) / 1e12;
        if (r > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops = best;
    return SDR_OK;
#else
    (void)device; *tflops = 0.0;
    return SDR_ERR_UNSUPPORTED;
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// design introspection
// ---------------------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------------------
// audio-side consumers on the device (dsp_sink.h)
// ---------------------------------------------------------------------------------------------------------------------
struct sdr_mixer {
    sdr_engine *eng = nullptr;
    int rate = 0;
    unsigned max_frames = 0;
    int n_src = 0;
    MixSource *d_src = nullptr;
    float *d_rings = nullptr;
    size_t ring_floats = 0, ring_cap_total = 0;
    int *d_map = nullptr;
    double *d_peak = nullptr;
    float *d_out = nullptr, *h_out = nullptr;
    float *d_stage = nullptr, *h_stage = nullptr;
    std::vector<int> src_chan;               // host shadow of MixSource::chan
    std::vector<int> src_active;             // host shadow of MixSource::active
    std::vector<int> order;                  // the controller's boundThreads: active sources in binding order
    int *d_order = nullptr;
#ifndef SDR_EMU
    cudaEvent_t ev_done = nullptr;           // end of the last render's device work
#endif
    // floats per source: the 100 queued blocks + currentInput of a two-channel 48 kHz demodulator at 60 blocks a second
    // (1600 floats each) with room to spare; longer blocks fill the ring before the queue and are dropped from then on
    static const int kRing = 196608;
    int ring_sources = 0;                    // sources the ring pool has room for (grown by doubling, mixer_add)
    static const int kStage = 65536;
};
static MixArgs mixer_args(sdr_mixer *m) {
    MixArgs a;
    a.src = m->d_src; a.n_src = m->n_src; a.order = m->d_order; a.n_order = (int)m->order.size(); a.rings = m->d_rings; a.rate = m->rate; a.cs = m->eng->d_cs; a.results = m->eng->d_results;
    a.n_frames = 0; a.map = m->d_map; a.mix_peak = m->d_peak; a.out = m->d_out;
    return a;
}
static void mixers_feed(sdr_engine *e, sdr_stream_t ts) {
    for (sdr_mixer *m : e->mixers) {
        if (!m->n_src) continue;
        MixArgs a = mixer_args(m);
#ifndef SDR_EMU
        k_mix_feed<<<(m->n_src + 127) / 128, 128, 0, ts>>>(a);
#else
        (void)ts;
        for (int j = 0; j < m->n_src; ++j) mix_feed_thread(a, j);
#endif
        e->launches += 1;
    }
}
static sdr_stream_t sink_stream(sdr_engine *e) {
#ifndef SDR_EMU
    return e->stream_tail;
#else
    return e->stream;
#endif
}
extern "C" int sdr_mixer_create(sdr_engine *e, int sample_rate, unsigned max_frames, sdr_mixer **out) {
    if (!e || !out || sample_rate <= 0 || !max_frames) return fail(SDR_ERR_ARG, "sdr_mixer_create: bad arguments");
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    sdr_mixer *m = new sdr_mixer();
    m->eng = e; m->rate = sample_rate; m->max_frames = max_frames;
    m->ring_sources = 16;
    m->ring_cap_total = (size_t)m->ring_sources * sdr_mixer::kRing;
    int bad = 0;
    bad |= be_malloc((void **)&m->d_src, kMixMaxSources * sizeof(MixSource));
    bad |= be_malloc((void **)&m->d_rings, m->ring_cap_total * sizeof(float));
    bad |= be_malloc((void **)&m->d_map, (size_t)kMixMaxSources * 2 * max_frames * sizeof(int));
    bad |= be_malloc((void **)&m->d_peak, kMixMaxSources * sizeof(double));
    bad |= be_malloc((void **)&m->d_order, kMixMaxSources * sizeof(int));
    bad |= be_malloc((void **)&m->d_out, 2 * (size_t)max_frames * sizeof(float));
    bad |= be_host_alloc((void **)&m->h_out, 2 * (size_t)max_frames * sizeof(float));
    bad |= be_malloc((void **)&m->d_stage, sdr_mixer::kStage * sizeof(float));
    bad |= be_host_alloc((void **)&m->h_stage, sdr_mixer::kStage * sizeof(float));
    if (bad) {
        be_free(m->d_src); be_free(m->d_rings); be_free(m->d_map); be_free(m->d_peak); be_free(m->d_order); be_free(m->d_out); be_host_free(m->h_out);
        be_free(m->d_stage); be_host_free(m->h_stage); delete m;
        return fail(SDR_ERR_NOMEM, "sdr_mixer_create: allocation failed");
    }
    be_memset(m->d_src, 0, kMixMaxSources * sizeof(MixSource), sink_stream(e));
    e->mixers.push_back(m);
    *out = m;
    return SDR_OK;
}
extern "C" int sdr_mixer_destroy(sdr_mixer *m) {
    if (!m) return SDR_OK;
    sdr_engine *e = m->eng;
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    be_sync(sink_stream(e));
    e->mixers.erase(std::remove(e->mixers.begin(), e->mixers.end(), m), e->mixers.end());
#ifndef SDR_EMU
    if (m->ev_done) cudaEventDestroy(m->ev_done);
#endif
    be_free(m->d_src); be_free(m->d_rings); be_free(m->d_map); be_free(m->d_peak); be_free(m->d_order); be_free(m->d_out); be_host_free(m->h_out);
    be_free(m->d_stage); be_host_free(m->h_stage);
    delete m;
    return SDR_OK;
}
// AudioThread::setGain (src/audio/AudioThread.cpp:595-604): clamped to [0, 2]
static float mixer_clamp_gain(float gain) {
    if (gain < 0.0f) gain = 0.0f;
    if (gain > 2.0f) gain = 2.0f;
    return gain;
}
// the binding order changed: upload it (engine lock held; the sink stream is idle afterwards)
static void mixer_upload_order(sdr_mixer *m) {
    sdr_stream_t ss = sink_stream(m->eng);
    be_sync(ss);
    if (!m->order.empty()) be_h2d(m->d_order, m->order.data(), m->order.size() * sizeof(int), ss);
    be_sync(ss);
}
static int mixer_add(sdr_mixer *m, int chan_slot, int feed_rate, float gain, int *index) {
    if (m->n_src >= kMixMaxSources) return fail(SDR_ERR_STATE, "sdr_mixer: too many sources");
    sdr_stream_t ss0 = sink_stream(m->eng);
    if (m->n_src >= m->ring_sources) {                            // grow the ring pool (binding is rare: copy, swap, free)
        const int grown = std::min(2 * m->ring_sources, (int)kMixMaxSources);
        float *d_new = nullptr;
        if (be_malloc((void **)&d_new, (size_t)grown * sdr_mixer::kRing * sizeof(float))) return fail(SDR_ERR_NOMEM, "sdr_mixer: ring pool");
        be_sync(ss0);
        be_d2d(d_new, m->d_rings, m->ring_cap_total * sizeof(float), ss0);
        be_sync(ss0);
        be_free(m->d_rings);
        m->d_rings = d_new; m->ring_sources = grown; m->ring_cap_total = (size_t)grown * sdr_mixer::kRing;
    }
    MixSource s;
    memset(&s, 0, sizeof s);
    // a demodulator's AudioThread: given its gain (setGain), then setActive(true), which binds it to the controller (:577-579)
    s.chan = chan_slot; s.feed_rate = feed_rate; s.gain = mixer_clamp_gain(gain); s.active = 1;
    s.ring_off = (int)((size_t)m->n_src * sdr_mixer::kRing); s.ring_cap = sdr_mixer::kRing;
    sdr_stream_t ss = sink_stream(m->eng);
    be_sync(ss);                                                  // (binding is rare; the copy below reads a stack object)
    be_h2d(m->d_src + m->n_src, &s, sizeof s, ss);
    be_sync(ss);
    if (index) *index = m->n_src;
    m->order.push_back(m->n_src);
    m->n_src += 1;
    m->src_chan.push_back(chan_slot);
    m->src_active.push_back(1);
    mixer_upload_order(m);
    return SDR_OK;
}
// the channel in slot `index` is being destroyed (engine lock held): its slot may be reused, so no source may read it any more
static void mixers_forget_chan(sdr_engine *e, int index) {
    static const int none = -1;
    for (sdr_mixer *m : e->mixers)
        for (size_t j = 0; j < m->src_chan.size(); ++j)
            if (m->src_chan[j] == index) {
                m->src_chan[j] = -1;
                be_h2d(reinterpret_cast<char *>(m->d_src + j) + offsetof(MixSource, chan), &none, sizeof none, sink_stream(e));
            }
}
// AudioThread::bindThread (src/audio/AudioThread.cpp:62-72): the channel's audio of every block joins this source's queue
extern "C" int sdr_mixer_bind(sdr_mixer *m, sdr_chan *c, float gain, int *source_index) {
    if (!m || !c || c->eng != m->eng || c->index < 0) return fail(SDR_ERR_ARG, "sdr_mixer_bind: bad arguments");
    std::lock_guard<std::mutex> lk(m->eng->mu);
    use_device(m->eng);
    return mixer_add(m, c->index, c->audio_rate, gain, source_index);
}
extern "C" int sdr_mixer_add_source(sdr_mixer *m, float gain, int *source_index) {
    if (!m) return fail(SDR_ERR_ARG, "sdr_mixer_add_source: null mixer");
    std::lock_guard<std::mutex> lk(m->eng->mu);
    use_device(m->eng);
    return mixer_add(m, -1, 0, gain, source_index);
}
static int mixer_poke(sdr_mixer *m, int source, size_t field_off, const void *v, size_t bytes) {
    if (!m || source < 0 || source >= m->n_src) return fail(SDR_ERR_ARG, "sdr_mixer: bad source");
    std::lock_guard<std::mutex> lk(m->eng->mu);
    use_device(m->eng);
    sdr_stream_t ss = sink_stream(m->eng);
    be_h2d(reinterpret_cast<char *>(m->d_src + source) + field_off, v, bytes, ss);
    be_sync(ss);
    return SDR_OK;
}
extern "C" int sdr_mixer_set_gain(sdr_mixer *m, int source, float gain) {               // AudioThread::setGain (:595-604)
    gain = mixer_clamp_gain(gain);
    return mixer_poke(m, source, offsetof(MixSource, gain), &gain, sizeof gain);
}
// AudioThread::setActive (:557-589): becoming active binds the thread to the controller -- APPENDED, so a source that was
// inactive for a while is mixed last afterwards (float sums depend on the order); becoming inactive removes it; every call
// flushes the inputQueue ("activity state changing, clear any inputs"); currentInput and audioQueuePtr stay.
extern "C" int sdr_mixer_set_active(sdr_mixer *m, int source, int active) {
    if (!m || source < 0 || source >= m->n_src) return fail(SDR_ERR_ARG, "sdr_mixer: bad source");
    std::lock_guard<std::mutex> lk(m->eng->mu);
    use_device(m->eng);
    const int state = active ? 1 : 0;
    bool order_changed = false;
    if (state && !m->src_active[(size_t)source]) {
        if (std::find(m->order.begin(), m->order.end(), source) == m->order.end()) { m->order.push_back(source); order_changed = true; }
    } else if (!state && m->src_active[(size_t)source]) {
        auto it = std::find(m->order.begin(), m->order.end(), source);
        if (it != m->order.end()) { m->order.erase(it); order_changed = true; }
    }
    m->src_active[(size_t)source] = state;
    sdr_stream_t ss = sink_stream(m->eng);
    be_sync(ss);
    static const int zero2[2] = {0, 0};
    static_assert(offsetof(MixSource, q_count) == offsetof(MixSource, q_head) + sizeof(int), "q_head, q_count adjacent");
    be_h2d(reinterpret_cast<char *>(m->d_src + source) + offsetof(MixSource, q_head), zero2, sizeof zero2, ss);
    be_h2d(reinterpret_cast<char *>(m->d_src + source) + offsetof(MixSource, active), &state, sizeof state, ss);
    be_sync(ss);
    if (order_changed) mixer_upload_order(m);
    return SDR_OK;
}
// one AudioThreadInput pushed by the host (what DemodulatorThread does with audioOutputQueue->try_push, for audio from elsewhere)
extern "C" int sdr_mixer_push(sdr_mixer *m, int source, const float *data, unsigned n, int channels, float peak, int sample_rate) {
    if (!m || source < 0 || source >= m->n_src || (n && !data) || n > (unsigned)sdr_mixer::kStage) return fail(SDR_ERR_ARG, "sdr_mixer_push: bad arguments");
    sdr_engine *e = m->eng;
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    sdr_stream_t ss = sink_stream(e);
    be_sync(ss);
    if (n) memcpy(m->h_stage, data, n * sizeof(float));
    be_h2d(m->d_stage, m->h_stage, n * sizeof(float), ss);
    MixArgs a = mixer_args(m);
#ifndef SDR_EMU
    k_mix_push<<<1, 256, 0, ss>>>(a, source, m->d_stage, (int)n, channels, peak, sample_rate);
#else
    {
        MixSource &s = a.src[source];
        const int w = mix_ring_take(s, (int)n);
        if (w >= 0) {
            for (unsigned i = 0; i < n; ++i) a.rings[s.ring_off + w + i] = m->d_stage[i];
            mix_queue_push(s, w, (int)n, channels, sample_rate, peak);
        }
    }
#endif
    e->launches += 1;
    be_sync(ss);
    return SDR_OK;
}
// audioCallback (src/audio/AudioThread.cpp:88-243): n_frames interleaved stereo frames of the mix into out[2 n_frames]
extern "C" int sdr_mixer_render(sdr_mixer *m, unsigned n_frames, float *out) {
    if (!m || !out || !n_frames || n_frames > m->max_frames) return fail(SDR_ERR_ARG, "sdr_mixer_render: bad arguments");
    sdr_engine *e = m->eng;
    {
        // the engine lock is held only while the kernels are enqueued: the audio thread waits for its own event afterwards, so a
        // callback never stands between SDRPostThread and its next submit
        std::lock_guard<std::mutex> lk(e->mu);
        use_device(e);
        sdr_stream_t ss = sink_stream(e);
        MixArgs a = mixer_args(m);
        a.n_frames = (int)n_frames;
#ifndef SDR_EMU
        if (m->n_src) k_mix_walk<<<(m->n_src + 127) / 128, 128, 0, ss>>>(a);
        k_mix_sum<<<(2 * n_frames + 255) / 256, 256, 0, ss>>>(a);
        k_mix_scale<<<(2 * n_frames + 255) / 256, 256, 0, ss>>>(a);
#else
        for (int j = 0; j < m->n_src; ++j) mix_walk_thread(a, j);
        for (unsigned k = 0; k < 2 * n_frames; ++k) mix_sum_thread(a, (int)k);
        { const float sc = mix_scale(a); if (sc != 1.0f) for (unsigned k = 0; k < 2 * n_frames; ++k) a.out[k] = a.out[k] * sc; }
#endif
        e->launches += 3;
        be_d2h(m->h_out, m->d_out, 2 * (size_t)n_frames * sizeof(float), ss);
#ifndef SDR_EMU
        if (!m->ev_done && cudaEventCreateWithFlags(&m->ev_done, cudaEventDisableTiming) != cudaSuccess) return fail(SDR_ERR_CUDA, "sdr_mixer_render: event");
        CUDA_TRY(cudaEventRecord(m->ev_done, ss));
#endif
    }
#ifndef SDR_EMU
    if (cudaEventSynchronize(m->ev_done) != cudaSuccess) return fail(SDR_ERR_CUDA, "sdr_mixer_render: wait failed");
#endif
    memcpy(out, m->h_out, 2 * (size_t)n_frames * sizeof(float));
    return SDR_OK;
}
// AudioFileWAV::writePayloadToFileStream (src/audio/AudioFileWAV.cpp:143-165): the last block's audio of a channel as the 16-bit words
// a recording receives (clipping prevented through the block's peak); two-channel audio in whole {left, right} pairs
extern "C" int sdr_chan_fetch_pcm16(sdr_chan *c, short *out, unsigned cap, unsigned *n) {
    if (!c || !out || !n || c->index < 0) return fail(SDR_ERR_ARG, "sdr_chan_fetch_pcm16: bad arguments");
    sdr_engine *e = c->eng;
    std::lock_guard<std::mutex> call(e->mu_pcm);                 // one conversion at a time (they share the engine's staging buffers)
    const unsigned room = std::min<unsigned>(cap, (unsigned)c->cs.res_audio_cap);
    {
        std::lock_guard<std::mutex> lk(e->mu);                   // held only while enqueueing (see sdr_mixer_render)
        use_device(e);
        sdr_stream_t ss = sink_stream(e);
        if (room + 2 > e->pcm_cap) {                             // (rare: the first call, or a channel with a larger slot)
            be_sync(ss);
            be_free(e->d_pcm); be_host_free(e->h_pcm);
            e->d_pcm = nullptr; e->h_pcm = nullptr; e->pcm_cap = 0;
            if (be_malloc((void **)&e->d_pcm, ((size_t)room + 2 + 2) * sizeof(short)) || be_host_alloc((void **)&e->h_pcm, ((size_t)room + 2 + 2) * sizeof(short)))
                return fail(SDR_ERR_NOMEM, "sdr_chan_fetch_pcm16");
            e->pcm_cap = room + 2;
        }
        PcmArgs a;
        a.cs = e->d_cs; a.results = e->d_results; a.chan = c->index; a.channels = (c->modem == MODEM_IQ || c->modem == MODEM_FMS) ? 2 : 1;
        a.out = e->d_pcm + 2; a.cap = (int)room; a.n_out = reinterpret_cast<int *>(e->d_pcm);      // [count (one int) | words]
#ifndef SDR_EMU
        k_pcm16<<<(room + 255) / 256 + 1, 256, 0, ss>>>(a);
#else
        for (unsigned i = 0; i < room + 1; ++i) pcm16_thread(a, (int)i);
#endif
        e->launches += 1;
        be_d2h(e->h_pcm, e->d_pcm, ((size_t)room + 2) * sizeof(short), ss);
#ifndef SDR_EMU
        if (!e->ev_pcm && cudaEventCreateWithFlags(&e->ev_pcm, cudaEventDisableTiming) != cudaSuccess) return fail(SDR_ERR_CUDA, "sdr_chan_fetch_pcm16: event");
        CUDA_TRY(cudaEventRecord(e->ev_pcm, ss));
#endif
    }
#ifndef SDR_EMU
    if (cudaEventSynchronize(e->ev_pcm) != cudaSuccess) return fail(SDR_ERR_CUDA, "sdr_chan_fetch_pcm16: wait failed");
#endif
    int hn = 0;
    memcpy(&hn, e->h_pcm, sizeof hn);
    if (hn < 0) hn = 0;
    if ((unsigned)hn > room) hn = (int)room;
    memcpy(out, e->h_pcm + 2, (size_t)hn * sizeof(short));
    *n = (unsigned)hn;
    return SDR_OK;
}
// ScopeVisualProcessor (src/process/ScopeVisualProcessor.cpp:14-217) on the audio of one channel
struct sdr_scope {
    sdr_engine *eng = nullptr; sdr_chan *chan = nullptr;
    int fft_size = 0, max_samples = 0;
    float2 *d_in = nullptr, *d_tw = nullptr; float *d_mags = nullptr; double2 *d_ema = nullptr; double *d_scalars = nullptr;
    float *d_wave = nullptr, *d_spec = nullptr; int *d_counts = nullptr; double *d_fc = nullptr; float2 *d_cf = nullptr; int *d_fs = nullptr;
    unsigned char *h_out = nullptr;          // pinned: [counts (2 int) | floor, ceil (2 double) | waveform | spectrum]
    std::mutex mu_call;
#ifndef SDR_EMU
    cudaEvent_t ev_done = nullptr;
#endif
};
extern "C" int sdr_scope_destroy(sdr_scope *s) {
    if (!s) return SDR_OK;
    std::lock_guard<std::mutex> lk(s->eng->mu);
    use_device(s->eng);
    be_sync(sink_stream(s->eng));
    be_free(s->d_in); be_free(s->d_tw); be_free(s->d_mags); be_free(s->d_ema); be_free(s->d_scalars); be_free(s->d_wave); be_free(s->d_spec);
    be_free(s->d_counts); be_free(s->d_fc); be_free(s->d_cf); be_free(s->d_fs); be_host_free(s->h_out);
#ifndef SDR_EMU
    if (s->ev_done) cudaEventDestroy(s->ev_done);
#endif
    delete s;
    return SDR_OK;
}
extern "C" int sdr_scope_create(sdr_engine *e, sdr_chan *c, unsigned fft_size, unsigned max_samples, sdr_scope **out) {
    if (!e || !c || !out || c->eng != e) return fail(SDR_ERR_ARG, "sdr_scope_create: bad arguments");
    if (fft_size < 64 || (fft_size & (fft_size - 1)) || fft_size > (unsigned)kFft2MaxSub || !max_samples)
        return fail(SDR_ERR_UNSUPPORTED, "sdr_scope_create: fft_size must be a power of two in [64, 8192]");
    std::lock_guard<std::mutex> lk(e->mu);
    use_device(e);
    sdr_scope *s = new sdr_scope();
    s->eng = e; s->chan = c; s->fft_size = (int)fft_size; s->max_samples = (int)max_samples;
    int bad = 0;
    bad |= be_malloc((void **)&s->d_in, fft_size * sizeof(float2));
    bad |= be_malloc((void **)&s->d_tw, fft_size * sizeof(float2));
    bad |= be_malloc((void **)&s->d_mags, fft_size * sizeof(float));
    bad |= be_malloc((void **)&s->d_ema, (fft_size / 2) * sizeof(double2));
    bad |= be_malloc((void **)&s->d_scalars, 4 * sizeof(double));
    bad |= be_malloc((void **)&s->d_wave, 2 * (size_t)max_samples * sizeof(float));
    bad |= be_malloc((void **)&s->d_spec, fft_size * sizeof(float));
    bad |= be_malloc((void **)&s->d_counts, 2 * sizeof(int));
    bad |= be_malloc((void **)&s->d_fc, 2 * sizeof(double));
    bad |= be_malloc((void **)&s->d_cf, sizeof(float2));
    bad |= be_malloc((void **)&s->d_fs, sizeof(int));
    bad |= be_host_alloc((void **)&s->h_out, 32 + 2 * (size_t)max_samples * sizeof(float) + (size_t)fft_size * sizeof(float));
    if (bad) { std::string msg = "sdr_scope_create: allocation failed"; (void)msg; e->mu.unlock(); sdr_scope_destroy(s); e->mu.lock(); return fail(SDR_ERR_NOMEM, "sdr_scope_create: allocation failed"); }
    std::vector<float2> tw(fft_size);
    for (unsigned k = 0; k < fft_size; ++k) { const double ang = -2.0 * 3.14159265358979323846 * (double)k / (double)fft_size; tw[k] = make_float2((float)std::cos(ang), (float)std::sin(ang)); }
    sdr_stream_t ss = sink_stream(e);
    be_h2d(s->d_tw, tw.data(), fft_size * sizeof(float2), ss);
    be_memset(s->d_ema, 0, (fft_size / 2) * sizeof(double2), ss);
    be_memset(s->d_scalars, 0, 4 * sizeof(double), ss);                       // fft_ceil_ma = fft_ceil_maa = fft_floor_ma = fft_floor_maa = 0 (:11-12)
    be_memset(s->d_fs, 0, sizeof(int), ss);
    be_sync(ss);
    *out = s;
    return SDR_OK;
}
// process() on the channel's last block: waveform points {x, y} (n_wave of them) and audio-spectrum points {x, y} (n_spec), floor / ceil
extern "C" int sdr_scope_process(sdr_scope *s, float *wave_xy, unsigned *n_wave, float *spec_xy, unsigned *n_spec, double *floor_ceil) {
    if (!s || !wave_xy || !spec_xy || !n_wave || !n_spec) return fail(SDR_ERR_ARG, "sdr_scope_process: null argument");
    sdr_engine *e = s->eng;
    std::lock_guard<std::mutex> call(s->mu_call);
    std::unique_lock<std::mutex> lk(e->mu);
    use_device(e);
    sdr_stream_t ss = sink_stream(e);
    ScopeArgs a;
    a.cs = e->d_cs; a.results = e->d_results; a.chan = s->chan->index;
    a.channels = (s->chan->modem == MODEM_IQ || s->chan->modem == MODEM_FMS) ? 2 : 1;
    a.sample_rate = s->chan->audio_rate; a.input_rate = s->chan->asked_bandwidth > 0 ? s->chan->asked_bandwidth : s->chan->bandwidth;   // ati->inputRate = inp->sampleRate (DemodulatorThread.cpp:128)
    a.fft_size = s->fft_size; a.max_samples = s->max_samples; a.fft_in = s->d_in; a.mags = s->d_mags; a.ema = s->d_ema; a.scalars = s->d_scalars;
    a.rate = (double)0.65f; a.wave_xy = s->d_wave; a.spec_xy = s->d_spec; a.counts = s->d_counts; a.floor_ceil = s->d_fc; a.red = nullptr;
    SpecArgs f;
    memset(&f, 0, sizeof f);
    f.x = s->d_in; f.frame_start = s->d_fs; f.frame_stride = 0; f.F = 1; f.N = s->fft_size; f.fft_size = s->fft_size / 2; f.P = 1; f.nsub = s->fft_size;
    f.tw = s->d_tw; f.mags = s->d_mags; f.frame_cf = s->d_cf;
    const int work = std::max(s->fft_size, s->max_samples);
#ifndef SDR_EMU
    k_scope_prepare<<<(work + 255) / 256, 256, 0, ss>>>(a);
    {
        const Fft2Plan pl = fft2_plan(f.nsub);
        const size_t smem = (size_t)(f.nsub + f.nsub / pl.r[0]) * sizeof(float2);
        CUDA_TRY(cudaFuncSetAttribute(k_spec_fft2, cudaFuncAttributeMaxDynamicSharedMemorySize, kFft2SmemElems * (int)sizeof(float2)));
        k_spec_fft2<<<1, kFft2Threads, smem, ss>>>(f);
    }
    k_scope_finish<<<1, 1024, 0, ss>>>(a);
#else
    {
        const float pk = scope_wave_peak_serial(a);
        for (int i = 0; i < work; ++i) scope_prepare_thread(a, i, pk);
        const Fft2Plan pl = fft2_plan(f.nsub);
        std::vector<float2> buf((size_t)f.nsub + f.nsub / 2 + 16);
        std::vector<Fft2Regs> regs(kFft2Threads);
        for (int t = 0; t < kFft2Threads; ++t) spec_fft2_load(f, 0, t, kFft2Threads, buf.data());
        for (int hp = 0; hp < 2 * pl.passes; ++hp) for (int t = 0; t < kFft2Threads; ++t) spec_fft2_half(hp, f, 0, t, buf.data(), regs[t]);
        double mx = 0.0, mn = 1.0;
        for (int i = 0; i < a.fft_size / 2; ++i) { const double maa = scope_ema_bin(a, i); if (maa > mx) mx = maa; if (maa < mn) mn = maa; }
        scope_finish_serial(a, mx, mn);
        for (int i = 0; i < a.fft_size / 2; ++i) scope_points_thread(a, i);
    }
#endif
    e->launches += 3;
    // results to pinned host memory; the engine lock is released before the wait (see sdr_mixer_render)
    int *h_counts = reinterpret_cast<int *>(s->h_out);
    double *h_fc = reinterpret_cast<double *>(s->h_out + 16);
    float *h_wave = reinterpret_cast<float *>(s->h_out + 32), *h_spec = h_wave + 2 * (size_t)s->max_samples;
    be_d2h(h_counts, s->d_counts, 2 * sizeof(int), ss);
    be_d2h(h_fc, s->d_fc, 2 * sizeof(double), ss);
    be_d2h(h_wave, s->d_wave, 2 * (size_t)s->max_samples * sizeof(float), ss);
    be_d2h(h_spec, s->d_spec, (size_t)s->fft_size * sizeof(float), ss);
#ifndef SDR_EMU
    if (!s->ev_done && cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming) != cudaSuccess) return fail(SDR_ERR_CUDA, "sdr_scope_process: event");
    CUDA_TRY(cudaEventRecord(s->ev_done, ss));
    lk.unlock();
    if (cudaEventSynchronize(s->ev_done) != cudaSuccess) return fail(SDR_ERR_CUDA, "sdr_scope_process: wait failed");
#endif
    const int nw = h_counts[0] < 0 ? 0 : (h_counts[0] > s->max_samples ? s->max_samples : h_counts[0]);
    const int nsp = h_counts[1] < 0 ? 0 : (h_counts[1] > s->fft_size / 2 ? s->fft_size / 2 : h_counts[1]);
    memcpy(wave_xy, h_wave, 2 * (size_t)nw * sizeof(float));
    memcpy(spec_xy, h_spec, 2 * (size_t)nsp * sizeof(float));
    *n_wave = (unsigned)nw; *n_spec = (unsigned)nsp;
    if (floor_ceil) { floor_ceil[0] = h_fc[0]; floor_ceil[1] = h_fc[1]; }
    return SDR_OK;
}

extern "C" int sdr_design_kaiser(unsigned n, float fc, float as_db, float mu, float *h) { if (!h) return SDR_ERR_ARG; kaiser_lowpass(n, fc, as_db, mu, h); return SDR_OK; }
extern "C" int sdr_design_nco_table(float *t) { if (!t) return SDR_ERR_ARG; nco_sine_table(t); return SDR_OK; }
extern "C" uint32_t sdr_design_nco_step(float d) { return nco_phase_step(d); }
extern "C" int sdr_design_resampler(float rate, float as_db, int *interp, int *S, float *ra, uint32_t *step, float *pfb, int *m) {
    ResamplerPlan p = make_resampler_plan(rate, as_db);
    if (interp) *interp = p.interp; if (S) *S = p.num_halfband; if (ra) *ra = p.rate_arb; if (step) *step = p.arb_step;
    if (pfb) memcpy(pfb, p.pfb.data(), p.pfb.size() * sizeof(float));
    if (m) for (int i = 0; i < p.num_halfband && i < 16; ++i) m[i] = p.hb.m[i];
    return SDR_OK;
}
extern "C" int sdr_design_halfband_branch(int m, float as_db, float *h1) { return (h1 && halfband_branch_taps(m, as_db + 5.0f, h1)) ? SDR_OK : SDR_ERR_UNSUPPORTED; }
extern "C" int sdr_design_am_dc_block(float *h) { if (!h) return SDR_ERR_ARG; am_dc_block_taps(25, 30.0f, h); return SDR_OK; }
extern "C" int sdr_design_hilbert_q(float *hq) { if (!hq) return SDR_ERR_ARG; hilbert_quadrature_taps(5, 90.0f, hq); return SDR_OK; }
extern "C" int sdr_design_butter6_sos(float *o) {
    if (!o) return SDR_ERR_ARG;
    std::vector<Sos> s; butterworth_lowpass_sos(6, 0.25f, s);
    for (int i = 0; i < 3; ++i) { memcpy(o + 6 * i, s[i].b, 12); memcpy(o + 6 * i + 3, s[i].a, 12); }
    return SDR_OK;
}

extern "C" int sdr_design_cheby2_bp5_sos(float fc, float f0, float *o) {
    if (!o) return SDR_ERR_ARG;
    std::vector<Sos> s; cheby2_bandpass5_sos(fc, f0, s);
    for (int i = 0; i < 5; ++i) { memcpy(o + 6 * i, s[i].b, 12); memcpy(o + 6 * i + 3, s[i].a, 12); }
    return SDR_OK;
}
extern "C" int sdr_design_hilbert_q_as(float as_db, float *hq) { if (!hq) return SDR_ERR_ARG; hilbert_quadrature_taps(5, as_db, hq); return SDR_OK; }

// ---------------------------------------------------------------------------------------------------------------------
// several GPUs behind the same boundary
// ---------------------------------------------------------------------------------------------------------------------
#include "group.h"
