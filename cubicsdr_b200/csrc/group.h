// cubicsdr_b200/csrc/group.h -- several GPUs behind the same boundary; included at the end of engine.cu.
//
// Channels are independent given the block (the reference hands the same buffer to every demodulator,
// src/sdr/SDRPostThread.cpp:294-298), so a group places channel c on member c mod N, keeps the spectra on member 0 and moves
// the raw block to every member once per block.  Two forms share this code:
//   local        one process drives N GPUs -- what CubicSDR itself is (one process, src/CubicSDR.cpp:340-397).  The block goes
//                host -> GPU 0 by DMA and from there to the others as ONE ncclBroadcast over NVLink (communicators from
//                ncclCommInitAll; peer copies if no NCCL library can be loaded).
//   distributed  one process per GPU (torch.distributed.run): every process creates the group with its rank and the
//                ncclUniqueId of rank 0; the block enters on rank 0 and the same ncclBroadcast carries it.
// Every member runs its own DC blocker on the raw block: the replay is bit-identical everywhere, and shipping the filtered
// block instead would put the broadcast behind a serial dependency.  There is no reduction step: each member returns the
// audio of its own channels (sdr_chan_fetch_audio works on a group's channels as on any other).
#pragma once
#include <array>

#ifndef SDR_EMU
#include <dlfcn.h>
#include <nccl.h>

namespace {
// NCCL is bound at run time: the product library does not drag a second NCCL into a process that already carries one
// (PyTorch bundles its own libnccl.so.2), and hosts without any NCCL still get the local form over peer copies.
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi *nccl_api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy the process already has (PyTorch's), if any
        if (!a.h) a.h = dlopen("libnccl.so.2", RTLD_NOW);
        if (!a.h) a.h = dlopen("libnccl.so", RTLD_NOW);
        if (!a.h) return;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
        a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.h, "ncclCommInitAll");
        a.Broadcast = (decltype(a.Broadcast))dlsym(a.h, "ncclBroadcast");
        a.GroupStart = (decltype(a.GroupStart))dlsym(a.h, "ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.h, "ncclGroupEnd");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.Broadcast && a.GroupStart && a.GroupEnd && a.CommDestroy;
    });
    return &a;
}
}  // namespace
#define NCCL_TRY(expr)                                                                                          \
    do {                                                                                                        \
        ncclResult_t r_ = (expr);                                                                               \
        if (r_ != ncclSuccess) {                                                                                \
            NcclApi *na_ = nccl_api();                                                                          \
            return fail(SDR_ERR_CUDA, std::string(#expr) + ": " + (na_->GetErrorString ? na_->GetErrorString(r_) : "NCCL error")); \
        }                                                                                                       \
    } while (0)
#endif

struct sdr_group {
    static const int kRing = 3;                 // input blocks in flight
    int world = 1, rank = 0;                    // distributed form: world > 1 and exactly one local member
    std::vector<sdr_engine *> members;          // local engines; members[0] is the group's root iff rank == 0
    unsigned max_block = 0;
    long long next_chan = 0;                    // channels created through the group so far (placement: c mod total)
    int transport = 0;                          // 0 none (one member, or the caller moves the block), 1 NCCL, 2 peer copies
    bool host_results = true;                   // copy every block's results to the host (sdr_group_set_host_results)
    std::mutex mu;
    int sel = 0;
    bool rec[kRing] = {false, false, false};
    sdr_cf32 *h_in[kRing] = {nullptr, nullptr, nullptr};      // pinned staging on the root
    std::vector<std::array<float2 *, kRing>> d_in;            // [member][slot] raw block on that member's GPU
#ifndef SDR_EMU
    std::vector<cudaStream_t> s_in;                           // [member] H2D / broadcast stream
    std::vector<std::array<cudaEvent_t, kRing>> ev_ready, ev_free;
    std::vector<ncclComm_t> comms;                            // [member]
#endif
    int total() const { return world > 1 ? world : (int)members.size(); }
    bool is_root() const { return rank == 0; }
};

static void group_free(sdr_group *g) {
#ifndef SDR_EMU
    NcclApi *na = nccl_api();
    for (size_t i = 0; i < g->members.size(); ++i) {
        sdr_engine *e = g->members[i];
        if (e) { use_device(e); sync_all(e); }
        if (i < g->s_in.size() && g->s_in[i]) cudaStreamSynchronize(g->s_in[i]);
    }
    for (size_t i = 0; i < g->comms.size(); ++i) if (g->comms[i] && na->ok) na->CommDestroy(g->comms[i]);
    for (size_t i = 0; i < g->members.size(); ++i) {
        if (g->members[i]) use_device(g->members[i]);
        if (i < g->d_in.size()) for (auto p : g->d_in[i]) be_free(p);
        if (i < g->ev_ready.size()) for (auto ev : g->ev_ready[i]) if (ev) cudaEventDestroy(ev);
        if (i < g->ev_free.size()) for (auto ev : g->ev_free[i]) if (ev) cudaEventDestroy(ev);
        if (i < g->s_in.size() && g->s_in[i]) cudaStreamDestroy(g->s_in[i]);
    }
#endif
    for (auto &p : g->h_in) be_host_free(p);
    for (auto *e : g->members) if (e) sdr_engine_destroy(e);
    delete g;
}

static int group_make(const int *devices, int n_local, int rank, int world, const void *nccl_id, unsigned max_block, sdr_group **out) {
    if (!out || !devices || n_local < 1 || max_block == 0 || world < 1 || rank < 0 || rank >= world || (world > 1 && n_local != 1))
        return fail(SDR_ERR_ARG, "sdr_group_create: bad arguments");
    sdr_group *g = new sdr_group();
    g->world = world; g->rank = rank; g->max_block = max_block;
    g->members.assign((size_t)n_local, nullptr);
    g->d_in.resize((size_t)n_local);
    for (auto &a : g->d_in) a.fill(nullptr);
#ifndef SDR_EMU
    g->s_in.assign((size_t)n_local, nullptr);
    g->ev_ready.resize((size_t)n_local); g->ev_free.resize((size_t)n_local);
    for (auto &a : g->ev_ready) a.fill(nullptr);
    for (auto &a : g->ev_free) a.fill(nullptr);
    g->comms.assign((size_t)n_local, nullptr);
#endif
    for (int i = 0; i < n_local; ++i) {
        int r = sdr_engine_create(devices[i], max_block, nullptr, &g->members[(size_t)i]);
        if (r) { group_free(g); return r; }
    }
    int bad = 0;
    if (g->is_root()) for (auto &p : g->h_in) bad |= be_host_alloc((void **)&p, (size_t)max_block * sizeof(sdr_cf32));
#ifndef SDR_EMU
    for (int i = 0; i < n_local; ++i) {
        use_device(g->members[(size_t)i]);
        for (int k = 0; k < sdr_group::kRing; ++k) {
            bad |= be_malloc((void **)&g->d_in[(size_t)i][(size_t)k], (size_t)max_block * sizeof(float2));
            bad |= cudaEventCreateWithFlags(&g->ev_ready[(size_t)i][(size_t)k], cudaEventDisableTiming) != cudaSuccess;
            bad |= cudaEventCreateWithFlags(&g->ev_free[(size_t)i][(size_t)k], cudaEventDisableTiming) != cudaSuccess;
        }
        bad |= cudaStreamCreateWithFlags(&g->s_in[(size_t)i], cudaStreamNonBlocking) != cudaSuccess;
    }
    if (bad) { group_free(g); return fail(SDR_ERR_NOMEM, "sdr_group_create: allocation failed"); }
    NcclApi *na = nccl_api();
    if (world > 1) {
        if (nccl_id) {
            if (!na->ok) { group_free(g); return fail(SDR_ERR_UNSUPPORTED, "sdr_group_create_rank: no NCCL library could be loaded"); }
            ncclUniqueId id;
            memcpy(&id, nccl_id, sizeof id);
            use_device(g->members[0]);
            ncclResult_t r = na->CommInitRank(&g->comms[0], world, id, rank);
            if (r != ncclSuccess) { group_free(g); return fail(SDR_ERR_CUDA, "ncclCommInitRank failed"); }
            g->transport = 1;
        }
    } else if (n_local > 1) {
        if (na->ok && na->CommInitAll(g->comms.data(), n_local, devices) == ncclSuccess) g->transport = 1;
        else {
            for (auto &c : g->comms) c = nullptr;
            for (int i = 0; i < n_local; ++i) for (int j = 0; j < n_local; ++j) if (i != j) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
                if (can) { cudaSetDevice(devices[i]); if (cudaDeviceEnablePeerAccess(devices[j], 0) != cudaSuccess) cudaGetLastError(); }
            }
            g->transport = 2;
        }
    }
#else
    (void)nccl_id;
    if (bad) { group_free(g); return fail(SDR_ERR_NOMEM, "sdr_group_create: allocation failed"); }
#endif
    *out = g;
    return SDR_OK;
}

// One process, n GPUs.  devices: CUDA ordinals; devices[0] receives the host block and owns the spectra.
extern "C" int sdr_group_create(const int *devices, int n, unsigned max_block, sdr_group **out) {
    return group_make(devices, n, 0, 1, nullptr, max_block, out);
}
// One process per GPU.  nccl_id: the 128-byte ncclUniqueId that rank 0 obtained from sdr_group_unique_id and passed to the
// other ranks by the launcher's own means; NULL if the caller moves the block itself (then every rank passes it to submit).
extern "C" int sdr_group_create_rank(int device, int rank, int world, const void *nccl_id, unsigned max_block, sdr_group **out) {
    return group_make(&device, 1, rank, world, nccl_id, max_block, out);
}
extern "C" int sdr_group_unique_id(void *id128) {
    if (!id128) return fail(SDR_ERR_ARG, "sdr_group_unique_id: null argument");
#ifndef SDR_EMU
    NcclApi *na = nccl_api();
    if (!na->ok) return fail(SDR_ERR_UNSUPPORTED, "sdr_group_unique_id: no NCCL library could be loaded");
    ncclUniqueId id;
    NCCL_TRY(na->GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return SDR_OK;
#else
    memset(id128, 0, 128);
    return SDR_OK;
#endif
}
extern "C" int sdr_group_destroy(sdr_group *g) {
    if (g) group_free(g);
    return SDR_OK;
}
extern "C" int sdr_group_info(sdr_group *g, int *n_local, int *total, int *transport) {
    if (!g) return fail(SDR_ERR_ARG, "null group");
    if (n_local) *n_local = (int)g->members.size();
    if (total) *total = g->total();
    if (transport) *transport = g->transport;
    return SDR_OK;
}
extern "C" int sdr_group_engine(sdr_group *g, int i, sdr_engine **e) {
    if (!g || !e || i < 0 || i >= (int)g->members.size()) return fail(SDR_ERR_ARG, "sdr_group_engine: bad arguments");
    *e = g->members[(size_t)i];
    return SDR_OK;
}
// Channel c of the group (creation order) lives on member c mod N.  In the distributed form every rank makes the same
// calls in the same order and gets *out = NULL for the channels that live elsewhere.
extern "C" int sdr_group_chan_create(sdr_group *g, const char *modem, long long sample_rate, long long bandwidth, int audio_rate,
                                     sdr_chan **out) {
    if (!g || !out) return fail(SDR_ERR_ARG, "sdr_group_chan_create: null argument");
    *out = nullptr;
    int owner;
    { std::lock_guard<std::mutex> lk(g->mu); owner = (int)(g->next_chan++ % g->total()); }
    if (g->world > 1) {
        if (owner != g->rank) return SDR_OK;
        return sdr_chan_create(g->members[0], modem, sample_rate, bandwidth, audio_rate, out);
    }
    return sdr_chan_create(g->members[(size_t)owner], modem, sample_rate, bandwidth, audio_rate, out);
}
// The group's spectra live on member 0 of rank 0 (*out = NULL elsewhere).
extern "C" int sdr_group_spectrum_create(sdr_group *g, unsigned fft_size, unsigned lps, sdr_spectrum **out) {
    if (!g || !out) return fail(SDR_ERR_ARG, "sdr_group_spectrum_create: null argument");
    *out = nullptr;
    if (!g->is_root()) return SDR_OK;
    return sdr_spectrum_create(g->members[0], fft_size, lps, out);
}

// SDRPostThread hands one block to the group.  iq: host memory on the root (rank 0; a device pointer is accepted too, for
// blocks that are already resident); ignored elsewhere unless the group has no transport.  Like sdr_engine_submit the call only enqueues: DMA to GPU 0, one broadcast, one submit per member.
extern "C" int sdr_group_submit(sdr_group *g, const sdr_cf32 *iq, unsigned n, long long center, long long rate) {
    if (!g) return fail(SDR_ERR_ARG, "sdr_group_submit: null group");
    if (n == 0 || n > g->max_block) return fail(SDR_ERR_ARG, "sdr_group_submit: block size out of range");
    const bool need_iq = g->is_root() || (g->world > 1 && g->transport == 0);
    if (need_iq && !iq) return fail(SDR_ERR_ARG, "sdr_group_submit: null block");
    std::lock_guard<std::mutex> lk(g->mu);
    const int n_local = (int)g->members.size();
#ifdef SDR_EMU
    for (int i = 0; i < n_local; ++i) {
        sdr_engine *e = g->members[(size_t)i];
        std::lock_guard<std::mutex> le(e->mu);
        int r = submit_common(e, nullptr, iq, n, center, rate, g->host_results, 1, nullptr);
        if (r) return r;
    }
    return SDR_OK;
#else
    const int k = g->sel++ % sdr_group::kRing;
    const size_t bytes = (size_t)n * sizeof(sdr_cf32);
    NcclApi *na = nccl_api();
    // ---- the block reaches this process's first GPU
    if (need_iq) {
        sdr_engine *e0 = g->members[0];
        use_device(e0);
        if (g->rec[k]) {
            CUDA_TRY(cudaEventSynchronize(g->ev_ready[0][(size_t)k]));          // the pinned slot's previous DMA has left it (long ago)
            CUDA_TRY(cudaStreamWaitEvent(g->s_in[0], g->ev_free[0][(size_t)k], 0));   // GPU 0's DC stream consumed the device slot
            for (int i = 1; i < n_local; ++i) CUDA_TRY(cudaStreamWaitEvent(g->s_in[0], g->ev_ready[(size_t)i][(size_t)k], 0));   // peers copied it
        }
        const sdr_cf32 *src = iq;
        bool direct = false;                                                  // pinned host memory or device memory: no staging
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, iq) == cudaSuccess) direct = pa.type == cudaMemoryTypeHost || pa.type == cudaMemoryTypeDevice;
        else cudaGetLastError();
        if (!direct) {
            if (!g->h_in[k]) { if (be_host_alloc((void **)&g->h_in[k], (size_t)g->max_block * sizeof(sdr_cf32))) return fail(SDR_ERR_NOMEM, "sdr_group_submit: staging"); }
            memcpy(g->h_in[k], iq, bytes); src = g->h_in[k];
        }
        CUDA_TRY(cudaMemcpyAsync(g->d_in[0][(size_t)k], src, bytes, cudaMemcpyDefault, g->s_in[0]));
    } else if (g->rec[k]) {
        use_device(g->members[0]);
        CUDA_TRY(cudaStreamWaitEvent(g->s_in[0], g->ev_free[0][(size_t)k], 0));
    }
    // ---- ... and every other one: a single broadcast of the raw block
    if (g->transport == 1) {
        for (int i = 1; i < n_local; ++i) if (g->rec[k]) {
            use_device(g->members[(size_t)i]);
            CUDA_TRY(cudaStreamWaitEvent(g->s_in[(size_t)i], g->ev_free[(size_t)i][(size_t)k], 0));
        }
        if (n_local > 1) NCCL_TRY(na->GroupStart());
        for (int i = 0; i < n_local; ++i) {
            use_device(g->members[(size_t)i]);
            NCCL_TRY(na->Broadcast(g->d_in[(size_t)i][(size_t)k], g->d_in[(size_t)i][(size_t)k], (size_t)n * 2, ncclFloat32, 0,
                                   g->comms[(size_t)i], g->s_in[(size_t)i]));
        }
        if (n_local > 1) NCCL_TRY(na->GroupEnd());
    } else if (g->transport == 2) {
        use_device(g->members[0]);
        CUDA_TRY(cudaEventRecord(g->ev_ready[0][(size_t)k], g->s_in[0]));
        for (int i = 1; i < n_local; ++i) {
            use_device(g->members[(size_t)i]);
            CUDA_TRY(cudaStreamWaitEvent(g->s_in[(size_t)i], g->ev_ready[0][(size_t)k], 0));
            if (g->rec[k]) CUDA_TRY(cudaStreamWaitEvent(g->s_in[(size_t)i], g->ev_free[(size_t)i][(size_t)k], 0));
            CUDA_TRY(cudaMemcpyPeerAsync(g->d_in[(size_t)i][(size_t)k], g->members[(size_t)i]->device, g->d_in[0][(size_t)k],
                                         g->members[0]->device, bytes, g->s_in[(size_t)i]));
        }
    }
    // ---- every member runs the block: own DC blocker (bit-identical everywhere), its channels, the spectra on member 0
    for (int i = 0; i < n_local; ++i) {
        sdr_engine *e = g->members[(size_t)i];
        use_device(e);
        CUDA_TRY(cudaEventRecord(g->ev_ready[(size_t)i][(size_t)k], g->s_in[(size_t)i]));
        std::lock_guard<std::mutex> le(e->mu);
        int r = submit_guard(e, submit_common(e, g->d_in[(size_t)i][(size_t)k], nullptr, n, center, rate, g->host_results, 2,
                                              g->ev_ready[(size_t)i][(size_t)k], g->ev_free[(size_t)i][(size_t)k]));
        if (r) return r;
    }
    g->rec[k] = true;
    return SDR_OK;
#endif
}
// Throughput measurement with resident data: leave the results in device memory (the default copies them back every block).
extern "C" int sdr_group_set_host_results(sdr_group *g, int on) {
    if (!g) return fail(SDR_ERR_ARG, "null group");
    std::lock_guard<std::mutex> lk(g->mu);
    g->host_results = on != 0;
    return SDR_OK;
}
extern "C" int sdr_group_wait(sdr_group *g) {
    if (!g) return fail(SDR_ERR_ARG, "null group");
    for (auto *e : g->members) { int r = sdr_engine_wait(e); if (r) return r; }
    return SDR_OK;
}
extern "C" int sdr_group_wait_block(sdr_group *g, int lag) {
    if (!g) return fail(SDR_ERR_ARG, "null group");
    for (auto *e : g->members) { int r = sdr_engine_wait_block(e, lag); if (r) return r; }
    return SDR_OK;
}
