#!/usr/bin/env python3
"""bench.py -- IQ MSamples/s through the spectrum + demodulator chain (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5|cfg4|cfg3|cfg2|cfg1] [--impl ours|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one reference-sized IQ block (fs/60 samples, SoapySDRThread.cpp:668-674) through the whole hot path:
DC blocker -> spectrum FFT / EMA / waterfall line (rank 0) -> NCO mix + multi-stage decimation + demodulation + audio
resampling for every channel of this rank.  Channels are sharded over the ranks (strong scaling: the job is fixed);
when N > 1 the raw block is broadcast from rank 0 over NCCL inside the timed region.

`value`    device-resident inputs (a ring of distinct blocks larger than L2), results left in HBM.  The K steps are timed
           with CUDA events between a barrier + synchronize on both sides; that region is repeated until at least a
           second has been measured and the median repetition is reported (`reps`).
`e2e`      the same through the C-ABI call a CubicSDR thread makes: sdr_engine_submit(host pointer) on pinned host blocks
           (H2D inside), sdr_engine_wait_block, batched audio / spectrum fetch into host arrays (D2H inside).
`roofline` for the dominant kernel (channel front: oscillator mix + half-band cascade; FP32-issue bound with many channels,
           so the bound is "fp32" and the HBM contract figure sits in roofline.hbm), timed with CUDA events on the engine's
           streams inside the timed region.
`cpu_baseline` / `--impl reference`: the reference's own liquid-dsp binary + CubicSDR glue (oracle/_ref) on the host cores
           the process may really use (affinity and cgroup quota), single-channel glue as `value`, the reference's stock
           channelizer path (what SDRPostThread runs above 500 kSps) as `cpu_baseline_channelizer`.
`configs`  (N = 1) the other BASELINE.json configurations, each with its own device / e2e / roofline / CPU figures.
"""
import os
os.environ.setdefault("OMP_NUM_THREADS", "1")   # no spinning OpenMP pool next to the reference's own threads (cpu_baseline / --impl reference)
import argparse
import json
import math
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

CENTER = 100_000_000
WORKLOADS = {
    # BASELINE.json configs[4] (the metric's own config: spectrum + demod chain), run as a fixed job sharded over N GPUs
    "cfg5": dict(fs=61440000, plan="mixed", n_chan=512, fft=16384,
                 desc="cfg5: 512 mixed FM/AM/USB/LSB demod channels + 16384-pt spectrum/waterfall on 61.44 MSps IQ"),
    "cfg4": dict(fs=40000000, plan="mixed", n_chan=256, fft=0,
                 desc="cfg4: 256 mixed FM/AM/USB/LSB demod channels from 40 MSps IQ"),
    "cfg3": dict(fs=20000000, plan="nbfm64", n_chan=64, fft=0, desc="cfg3: 64 NBFM channels @12.5 kHz from 20 MSps IQ"),
    "cfg2": dict(fs=10000000, plan="mixed", n_chan=0, fft=8192, desc="cfg2: 8192-pt spectrum FFT + waterfall on 10 MSps IQ"),
    "cfg1": dict(fs=2400000, plan="wbfm1", n_chan=1, fft=0, desc="cfg1: 1x WBFM demod, 2.4 MSps IQ, 200 kHz -> 48 kHz audio"),
    # SURVEY.md 8(f) row 1 (not a BASELINE config): the reference's channelizer for rates above 500 kSps, on its own
    "pfb": dict(fs=61440000, plan="mixed", n_chan=0, fft=0,
                desc="pfb: firpfbch analysis channelizer (SDRPostThread::runPFBCH), 122 channels on 61.44 MSps IQ"),
}
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12    # non-tensor FP32 issue peak of a B200 at max clock
L2_BYTES = 126e6


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_cores():
    """Host cores this process may really use: scheduler affinity, capped by the cgroup CPU quota (os.cpu_count() sees the
    whole machine even inside a container with a quota)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    quota = None
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(p)
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0 and p > 0:
                quota = q / p
        except Exception:
            pass
    if quota is not None:
        n = max(1, min(n, int(math.floor(quota + 1e-9)) or 1))
    return max(1, n)


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def config_of(wl, N, n_chan):
    """The `config` object: identical keys and values on both arms."""
    return {"workload": wl["desc"], "fs": wl["fs"], "block": N, "channels": n_chan, "fft_size": wl["fft"]}


def torch_wideband(torch, fs, n, chans, t0, device, seed):
    """GPU version of cubicsdr_b200.synth.wideband (same signal model, torch RNG): complex64 tensor [n]."""
    t = (torch.arange(n, device=device, dtype=torch.float64) + float(t0)) / float(fs)
    x = torch.zeros(n, device=device, dtype=torch.complex128)
    amp = 0.5 / max(1.0, np.sqrt(len(chans)) * 1.5) if chans else 0.0
    two_pi = 2.0 * np.pi
    for k, (modem, bw, off) in enumerate(chans):
        tone = 1000.0 + (400.0 if (k & 1) else 0.0)
        if modem in ("FM", "NBFM"):
            dev = 75000.0 if modem == "FM" else 2500.0
            ph = two_pi * off * t + (dev / tone) * torch.sin(two_pi * tone * t)
            x += amp * torch.polar(torch.ones_like(ph), ph)
        elif modem == "AM":
            env = 1.0 + 0.8 * torch.sin(two_pi * tone * t)
            ph = two_pi * off * t
            x += torch.polar(amp * 0.5 * env, ph)
        else:
            sgn = 1.0 if modem == "USB" else -1.0
            x += amp * 0.5 * (torch.polar(torch.ones_like(t), two_pi * (off + sgn * 700.0) * t) +
                              torch.polar(torch.ones_like(t), two_pi * (off + sgn * 1900.0) * t))
    g = torch.Generator(device=device)
    g.manual_seed(seed + t0 % 1000003)
    sigma = 0.01 / np.sqrt(2.0)
    x += sigma * torch.complex(torch.randn(n, device=device, dtype=torch.float64, generator=g),
                               torch.randn(n, device=device, dtype=torch.float64, generator=g))
    x += complex(0.01, -0.02)
    return x.to(torch.complex64)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # the median over the samples taken while the GPU was clocked up (idle samples before / after the loops read low)
        load = [v for v in sm if mx and v >= 0.6 * max(mx)] or sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------------
# CPU reference legs (oracle/_ref: the reference's own liquid-dsp binary + a restatement of the CubicSDR glue around it)
# ---------------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(wl, chans, iq_host, threads, blocks, channelizer=0):
    """Time the reference CPU path on `blocks` blocks after one untimed block.
    channelizer = 1 / 2: through SDRPostThread::runPFBCH / runPFBCH2 instead of runSingleCH."""
    import ctypes as C
    from oracle.refbind import Ref
    ref = Ref()
    n_chan = len(chans)
    names = (C.c_char_p * max(1, n_chan))(*[m.encode() for m, _, _ in chans])
    bws = (C.c_longlong * max(1, n_chan))(*[bw for _, bw, _ in chans])
    shifts = (C.c_longlong * max(1, n_chan))(*[off for _, _, off in chans])
    chk = C.c_double()
    iq = np.ascontiguousarray(iq_host, np.complex64)
    if channelizer:
        freqs = (C.c_longlong * max(1, n_chan))(*[CENTER + off for _, _, off in chans])
        ref.L.ref_bench_pfb_path.restype = C.c_double
        ref.L.ref_bench_pfb_path.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_longlong, C.c_longlong, C.c_int, C.c_uint,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint, C.c_uint, C.c_void_p]
        secs = ref.L.ref_bench_pfb_path(iq.ctypes.data_as(C.c_void_p), iq.size, 1, blocks, wl["fs"], CENTER, int(channelizer), n_chan,
                                        names, bws, freqs, 48000, wl["fft"], threads, C.byref(chk))
        if secs < 0:
            raise RuntimeError("reference channelizer path: block length is not a multiple of the channel count")
        return secs, ref.version()
    secs = ref.L.ref_bench_path(iq.ctypes.data_as(C.c_void_p), iq.size, 1, blocks, wl["fs"], n_chan, names, bws, shifts, 48000,
                                wl["fft"], threads, C.byref(chk))      # 1 untimed block first (page faults, cold filters)
    return secs, ref.version()


def cpu_baselines(wl, chans, iq_host, threads, blocks=1):
    """(cpu_baseline, cpu_baseline_channelizer): the glue that matches the engine's default mode (runSingleCH) and the
    reference's stock path for this rate (runPFBCH above 500 kSps; identical to runSingleCH below)."""
    N = len(iq_host)
    what = "%d block(s) of %d samples (1/60 s each) after 1 untimed block, all %d channels%s" % (
        blocks, N, len(chans), " + spectrum" if wl["fft"] else "")
    out = []
    for mode in (0, 1):
        try:
            if mode == 1 and wl["fs"] <= 500000:
                out.append(None); continue
            secs, ver = cpu_reference_run(wl, chans, iq_host, threads, blocks, channelizer=mode)
            out.append({"value": N * blocks / secs / 1e6, "unit": "MSamples/s", "cores": threads, "kind": "reference",
                        "cpu": cpu_model(),
                        "path": "SDRPostThread::runPFBCH + runDemodChannels (the stock path above 500 kSps)" if mode else
                                "SDRPostThread::runSingleCH (DC blocker + every demodulator at the full rate: the engine's default mode)",
                        "sample": what + ", reference liquid-dsp %s binary + CubicSDR glue" % ver})
        except Exception as ex:        # the oracle is a checker; its absence must not hide the GPU number
            out.append({"value": None, "unit": "MSamples/s", "cores": 0, "kind": "reference", "sample": "unavailable: %r" % (ex,)})
    return out[0], out[1]


def run_reference(args, wl, chans, rank):
    """--impl reference: the reference's own CPU implementation on the host cores, same config / metric / unit."""
    from cubicsdr_b200 import synth
    if rank != 0:
        return
    N = synth.block_size(wl["fs"])
    iq = synth.wideband(wl["fs"], N, chans[:: max(1, len(chans) // 32)] if chans else [], t0=0)   # carriers: a 32-channel subset keeps generation quick
    threads = cpu_cores()
    secs, ver = cpu_reference_run(wl, chans, iq, threads, args.steps, channelizer=1 if args.channelizer else 0)
    val = N * args.steps / secs / 1e6
    stock = None
    if not args.channelizer and wl["fs"] > 500000:
        try:
            s2, _ = cpu_reference_run(wl, chans, iq, threads, max(1, min(args.steps, 4)), channelizer=1)
            stock = {"value": N * max(1, min(args.steps, 4)) / s2 / 1e6, "unit": "MSamples/s", "cores": threads, "kind": "reference",
                     "path": "SDRPostThread::runPFBCH + runDemodChannels (the stock path above 500 kSps)"}
        except Exception as ex:
            stock = {"value": None, "sample": "unavailable: %r" % (ex,)}
    line = {"impl": "reference", "metric": "IQ MSamples/s through spectrum+demod chain", "value": val, "unit": "MSamples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, N, len(chans)),
            "cpu_baseline": {"value": val, "unit": "MSamples/s", "cores": threads, "kind": "reference", "cpu": cpu_model(),
                             "path": "SDRPostThread::runPFBCH" if args.channelizer else "SDRPostThread::runSingleCH (the mode the engine arm runs)",
                             "sample": "1 block (1/60 s of IQ) per step after 1 untimed block, all channels + spectrum, liquid-dsp %s" % ver},
            "cpu_baseline_channelizer": stock,
            "e2e": {"value": val, "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_pfb(args, wl):
    """--workload pfb: the polyphase analysis channelizer alone, device-resident and through host
    buffers, next to the reference binary's firpfbch_crcf on one host core (the reference runs it on one thread)."""
    import torch
    import cubicsdr_b200
    from cubicsdr_b200 import Channelizer, Engine, synth
    fs = wl["fs"]
    N = synth.block_size(fs)
    lib = cubicsdr_b200.lib()
    M = Channelizer.optimal_channel_count(lib, fs)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng = Engine(lib, device=0, max_block=N, stream=stream.cuda_stream)
    ch = Channelizer(eng, M, N // M)
    carriers = synth.channel_plan("mixed", fs, 16)
    ring = [torch_wideband(torch, fs, N, carriers, r * N, dev, 1) for r in range(args.ring)]
    outs = [torch.empty(N, dtype=torch.complex64, device=dev) for _ in range(2)]
    sampler = ClockSampler(0)
    sampler.start()
    for i in range(max(args.warmup, 3) + 200):
        ch.execute_device(ring[i % args.ring].data_ptr(), N, outs[i & 1].data_ptr())
    torch.cuda.synchronize()
    launches0 = eng.launch_count
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        ch.execute_device(ring[i % args.ring].data_ptr(), N, outs[i & 1].data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count - launches0
    value = N * args.steps / (ms * 1e-3) / 1e6
    host_in = ring[0].cpu().numpy()
    for _ in range(3):
        ch.execute(host_in)
    t0 = time.perf_counter()
    for _ in range(max(1, args.steps // 10)):
        ch.execute(host_in)
    t_e2e = (time.perf_counter() - t0) / max(1, args.steps // 10)
    clocks = sampler.stop()
    peak, peak_src = measured_peaks()
    achieved = 16.0 * N / (ms / args.steps * 1e-3) / 1e9
    cpu = None
    if not args.no_cpu_baseline:
        from oracle.refbind import Ref, RefPfb
        rp = RefPfb(Ref(), M)
        rp.execute(host_in)
        t0 = time.perf_counter(); rp.execute(host_in); secs = time.perf_counter() - t0
        cpu = {"value": N / secs / 1e6, "unit": "MSamples/s", "cores": 1, "kind": "reference",
               "sample": "1 block (%d samples) after 1 untimed block, reference liquid-dsp firpfbch_crcf analyzer, one thread as in SDRPostThread" % N}
    line = {"metric": "IQ MSamples/s through the analysis channelizer", "value": value, "unit": "MSamples/s", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "fs": fs, "block": N, "pfb_channels": M,
                       "l2": "%d distinct %.1f MB input blocks rotate" % (args.ring, N * 8 / 1e6)},
            "clocks": clocks,
            "e2e": {"value": N / t_e2e / 1e6, "unit": "MSamples/s", "h2d_bytes_per_step": N * 8, "d2h_bytes_per_step": N * 8,
                    "note": "blocking sdr_pfb_execute with pageable host arrays (staging memcpy + H2D + kernel + D2H + memcpy)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_pfb_analyze", "achieved": achieved, "peak": peak, "peak_source": peak_src,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": None, "alg_bytes_per_sample": 16.0,
                         "kernel_ms_per_block": ms / args.steps},
            "cpu_baseline": cpu}
    ch.close(); eng.close()
    emit(line)


# The contract is ONE JSON line on stdout.  Libraries underneath (NCCL's version banner, for one) also write to fd 1, so
# the real stdout is kept aside and fd 1 is pointed at stderr for the duration of the run.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ---------------------------------------------------------------------------------------------------------------------------
# the engine arm
# ---------------------------------------------------------------------------------------------------------------------------
class Ctx:
    """torch / library / process-group handles shared by the workloads of one bench run"""
    pass


def run_engine_workload(cx, wl, chans, steps, warmup, min_seconds, channelizer=False, want_cpu=True, want_timeline=False):
    """One workload on this rank's GPU (channels sharded over the ranks).  Returns the dict of measured figures on rank 0,
    None elsewhere."""
    torch, lib, dist, dev = cx.torch, cx.lib, cx.dist, cx.dev
    rank, world = cx.rank, cx.world
    from cubicsdr_b200 import Channel, Engine, Spectrum, synth
    fs = wl["fs"]
    N = synth.block_size(fs)
    stream = torch.cuda.Stream(device=dev)          # a dedicated (non-default) stream shared by torch and the engine
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    desc = wl["desc"]
    chan_fs = fs
    grp = None
    if world > 1:
        # several GPUs: the product's group API in its one-process-per-GPU form (sdr_group_create_rank): channel c on rank
        # c mod N, spectrum on rank 0, the raw block moved by the group's own ncclBroadcast.  torch.distributed only carries
        # the 128-byte NCCL id and the timing barriers.
        from cubicsdr_b200 import Group
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(Group.unique_id(lib)), dtype=torch.uint8))
        dist.broadcast(idt, src=0)
        grp = Group(lib, max_block=N, device=cx.local_rank, rank=rank, world=world, nccl_id=bytes(idt.cpu().numpy().tobytes()))
        assert grp.transport == 1 and grp.total == world
        eng = grp.engines[0]
        made = [grp.channel(m, chan_fs, bw, 48000, frequency=CENTER + off) for m, bw, off in chans]
        cobjs = [c for c in made if c is not None]
        mine = [ch for ch, c in zip(chans, made) if c is not None]
        assert mine == chans[rank::world]
        spec = grp.spectrum(wl["fft"], 0) if wl["fft"] else None
    else:
        eng = Engine(lib, device=cx.local_rank, max_block=N, stream=stream.cuda_stream)
        mine = chans
        if channelizer:
            eng.set_channelizer(1)
            n_pfb, chan_fs = eng.channel_plan(fs)
            eng.set_channel_block_max(N // n_pfb)
            desc += " [DIAGNOSTIC: channelizer mode, %d filterbank channels of %d Hz]" % (n_pfb, chan_fs)
        cobjs = [Channel(eng, m, chan_fs, bw, 48000, frequency=CENTER + off) for m, bw, off in mine]
        spec = Spectrum(eng, wl["fft"], 0) if wl["fft"] else None

    def print_timeline(tag):                                    # stderr diagnostic: device timeline of a few steady-state blocks
        try:
            import ctypes as C
            tl = np.zeros((16, 8), np.float32); ntl = C.c_int()
            lib.L.sdr_engine_debug_timeline(eng.h, tl.ctypes.data_as(C.c_void_p), C.byref(ntl))
            print("%s timeline (ms) dc_start dc_end | front_start front_end tail_end | spec_start spec_end tail_start" % tag, file=sys.stderr)
            for r in range(8, min(13, ntl.value)):
                print("  blk %2d " % r + " ".join("%7.3f" % v for v in tl[r]), file=sys.stderr)
        except Exception as ex:
            print("timeline unavailable: %r" % (ex,), file=sys.stderr)

    # ---- inputs: ring of distinct blocks resident in HBM on rank 0, larger than the 126 MB L2 in total ------------------
    n_ring = max(cx.args.ring, int(math.ceil(1.1 * L2_BYTES / (N * 8.0))))
    ring = []
    carriers = chans if chans else synth.channel_plan("mixed", fs, 16)
    if rank == 0:
        if cx.args.noise_inputs:                               # profiling runs: skip the per-carrier synthesis (thousands of launches)
            gq = torch.Generator(device=dev); gq.manual_seed(5)
            base = [(0.05 * torch.view_as_complex(torch.randn(N, 2, device=dev, generator=gq)) + complex(0.01, -0.02)).to(torch.complex64)
                    for r in range(min(n_ring, 16))]
        else:
            base = [torch_wideband(torch, fs, N, carriers, r * N, dev, 1234) for r in range(min(n_ring, 16))]
        ring = list(base)
        g = torch.Generator(device=dev); g.manual_seed(77)
        while len(ring) < n_ring:                             # further distinct blocks: an earlier one plus fresh noise (cheap)
            b = base[len(ring) % len(base)]
            ring.append(b + 0.002 * torch.view_as_complex(torch.randn(N, 2, device=dev, generator=g)))
    # host side of the e2e: four pinned blocks from the C-ABI's own allocator
    host_blocks = []
    if rank == 0:
        for k in range(4):
            a, addr = lib.host_alloc(N)
            a[:] = ring[k % len(ring)].cpu().numpy()
            host_blocks.append((a, addr))
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()

    def step_device(i):
        if grp is not None:
            # the resident block of rank 0 goes to every rank by the group's one exchange step (raw IQ over NVLink)
            grp.submit(ring[i % n_ring].data_ptr() if rank == 0 else None, CENTER, fs, n=N)
            return
        # resident block: the DC blocker of this block overlaps the previous block's channel kernels
        eng.submit_device_async(ring[i % n_ring].data_ptr(), N, CENTER, fs, fetch_to_host=False, ready_event=None)

    # ---- end to end: exactly the host entry a CubicSDR thread calls -- sdr_engine_submit (one GPU) / sdr_group_submit (several)
    # on a pinned host block (DMA to the GPU, the broadcast to the other ranks), sdr_*_wait_block(2), batched fetch.
    def h2d(i):
        pass

    def submit_e2e(i):
        if grp is not None:
            grp.submit(host_blocks[i % 4][1] if rank == 0 else None, CENTER, fs, n=N)
        else:
            eng.submit_ptr(host_blocks[i % 4][1], N, CENTER, fs)

    def fetch_e2e():
        # the block's results were DMA'd to the engine's pinned host buffer; the batch "view" calls hand out pointers, counts,
        # levels and peaks for every channel (what the host's audio sink consumes) without another copy
        nbytes = 0
        if cobjs:
            ptrs, n_audio, _, _ = eng.view_audio_batch(cobjs)
            nbytes += int(n_audio.sum()) * 4
        if spec is not None:
            pts, lines = spec.view()
            nbytes += pts.nbytes + lines.nbytes
        return nbytes

    # Two blocks in flight, like the reference's thread pipeline (queues between its threads): per step, submit block i+2,
    # then wait for and fetch the results of block i.  Every step costs one H2D, one pass of the path and one D2H + fetch.
    host_split = [0.0, 0.0, 0.0]                                # host seconds in submit / wait / fetch (diagnostic, stderr)

    def step_e2e(i):
        t_a = time.perf_counter()
        submit_e2e(i + 2)
        h2d(i + 3)
        t_b = time.perf_counter()
        eng.wait_block(2)
        t_c = time.perf_counter()
        nb = fetch_e2e()
        t_d = time.perf_counter()
        host_split[0] += t_b - t_a; host_split[1] += t_c - t_b; host_split[2] += t_d - t_c
        return nb

    def max_over_ranks(v):
        if dist is None:
            return v
        tt = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- warm-up ------------------------------------------------------------------------------------------------------
    if grp is not None:
        grp.set_host_results(False)                             # device-resident measurement: results stay in HBM
    for i in range(max(3, warmup)):
        step_device(i)
    eng.wait()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(32):                                         # also the estimate that sizes the repetitions
        step_device(i)
    eng.wait(); torch.cuda.synchronize()
    est = max_over_ranks((time.perf_counter() - t0) / 32)
    reps = int(min(400, max(1, math.ceil(min_seconds / max(1e-6, est * steps)))))

    # stage times (CUDA events on the engine's streams, at most 64 blocks per collection): a pass of its own
    eng.set_timing(True)
    for i in range(min(60, max(steps, 16))):
        step_device(i)
    eng.wait()
    stage_ms, stage_blocks = eng.get_timing()
    if rank == 0 and want_timeline:
        print_timeline("device")
    eng.set_timing(False)

    # ---- timed: device-resident.  K steps between barrier + synchronize, repeated `reps` times; the median is reported ----
    launches0 = eng.launch_count
    rep_ms = []
    k = 0
    for r in range(reps):
        barrier(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            step_device(k); k += 1
        eng.wait()                                              # every stream of the engine has drained (device-side work only)
        e1.record(stream)
        torch.cuda.synchronize(); barrier()
        rep_ms.append(max_over_ranks(e0.elapsed_time(e1)))
    launches = (eng.launch_count - launches0) // reps
    ms = float(np.median(rep_ms))
    value = N * steps / (ms * 1e-3) / 1e6

    # ---- timed: end to end through host buffers -------------------------------------------------------------------------
    if grp is not None:
        grp.set_host_results(True)
    h2d(0); submit_e2e(0); h2d(1); submit_e2e(1); h2d(2)
    for i in range(3):
        step_e2e(i)
    eng.wait(); torch.cuda.synchronize()
    e2e_reps = int(min(200, max(1, math.ceil(min_seconds / max(1e-6, 1.3 * est * steps)))))
    rep_s = []
    d2h = 0
    k = 3
    for r in range(e2e_reps):
        # prime two blocks, then K timed steps each of which submits one block and fetches one
        barrier(); torch.cuda.synchronize()
        if r == 0 and rank == 0 and want_timeline:
            eng.set_timing(True)
        t0 = time.perf_counter()
        for i in range(steps):
            d2h = step_e2e(k); k += 1
        torch.cuda.synchronize(); barrier()
        rep_s.append(max_over_ranks(time.perf_counter() - t0))
        if r == 0 and rank == 0 and want_timeline:
            eng.wait()
            e2e_stage_ms, e2e_stage_blocks = eng.get_timing()
            eng.set_timing(False)
            print_timeline("e2e")
            print("e2e device stages per block (ms): %s" % {kk: round(v / max(1, e2e_stage_blocks), 3) for kk, v in e2e_stage_ms.items()}, file=sys.stderr)
    eng.wait()                                                  # drain the blocks still in flight (submitted, not yet fetched)
    t_e2e = float(np.median(rep_s))
    if rank == 0 and want_timeline:
        tot = max(1, e2e_reps * steps + 3)
        print("e2e host split per step (ms): submit %.3f  wait %.3f  fetch %.3f" % tuple(1e3 * v / tot for v in host_split), file=sys.stderr)
    if dist is not None:
        dd = torch.tensor([d2h], device=dev, dtype=torch.float64)
        dist.all_reduce(dd, op=dist.ReduceOp.SUM)
        d2h = int(dd.item())
    e2e_value = N * steps / t_e2e / 1e6

    res = None
    if rank == 0:
        # ---- roofline of the dominant kernel ---------------------------------------------------------------------------------
        peak, peak_src = measured_peaks()
        c_local = len(mine)
        b_alg = 8.0 + c_local * 4.0 * 48000.0 / fs + (2.5 if spec is not None else 0.0)     # bytes / wideband sample / GPU
        per_block = {kk: v / max(1, stage_blocks) for kk, v in stage_ms.items()}
        dom = "front" if c_local else "spectrum"
        if c_local and per_block.get("dc", 0.0) > per_block.get("front", 0.0) and not channelizer:
            dom = "dc"                                            # few channels: the bit-exact DC replay's latency is the longest stage
        t_dom_ms = per_block[dom]
        hbm_gbs = (N * b_alg / (t_dom_ms * 1e-3)) / 1e9 if t_dom_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(cx.args.workload if wl is WORKLOADS.get(cx.args.workload) else "", {}).get(dom)
            except Exception:
                traffic = None
        names = {"front": "k_chan_front (NCO mix + half-band cascade)", "spectrum": "spectrum kernels (k_spec_*)",
                 "dc": "DC blocker stage (double-precision scan + bit-exact float replay: candidate windows k_dc_window / k_dc_compose / k_dc_resolve / k_dc_exact, "
                       "or warmed-up segments k_dc_lanes_tma at >= 256 channels; latency bound)"}
        hbm = {"achieved": hbm_gbs, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": hbm_gbs / peak,
               "alg_bytes_per_sample": b_alg,
               "note": "contract figure: algorithmic bytes per launch / kernel time / measured HBM copy peak"}
        if c_local and dom == "front" and c_local >= 32:
            flops = 30.0 * c_local * N / (per_block["front"] * 1e-3) / 1e12
            fpk = cx.fp32_peak or FP32_NOMINAL_TFLOPS
            roofline = {"bound": "fp32", "kernel": names[dom], "achieved": flops, "peak": fpk,
                        "peak_source": "measured FFMA issue rate (sdr_debug_fp32_peak)" if cx.fp32_peak else "nominal 148 SM x 128 lanes x 2 x 1.965 GHz",
                        "unit": "TFLOP/s", "frac": flops / fpk, "traffic": traffic,
                        "flop_per_sample_per_channel": 30.0,
                        "note": "per-channel mixing replays the reference's table oscillator sample by sample: ~30 flop / sample / channel "
                                "(SURVEY.md 8d); FP32 issue, not HBM, binds with hundreds of channels in single-channel mode",
                        "hbm": hbm}
        else:
            roofline = dict(hbm, bound="hbm", kernel=names[dom], traffic=traffic)
        roofline["kernel_ms_per_block"] = t_dom_ms
        roofline["stage_ms_per_block"] = per_block
        res = {"value": value, "ms_per_step": ms / steps, "reps": reps, "rep_ms_min_max": [min(rep_ms), max(rep_ms)],
               "e2e": {"value": e2e_value, "unit": "MSamples/s", "h2d_bytes_per_step": N * 8, "d2h_bytes_per_step": d2h,
                       "reps": e2e_reps,
                       "api": "sdr_engine_submit(pinned host block) + sdr_engine_wait_block(2) + sdr_engine_view_audio_batch / sdr_spectrum_view (results DMA'd to pinned host memory)"
                              if world == 1 else "sdr_group_submit(pinned host block on rank 0: DMA + one ncclBroadcast inside) + sdr_engine_wait_block(2) + batched fetch"},
               "gpu_launches": int(launches), "roofline": roofline, "N": N, "c_local": c_local, "desc": desc,
               "l2": "%d distinct %.1f MB input blocks rotate (%.0f MB > 126 MB L2)" % (n_ring, N * 8 / 1e6, n_ring * N * 8 / 1e6),
               "realtime_factor": value * 1e6 / fs}
        if want_cpu:
            iq_host = host_blocks[0][0].copy()
            res["cpu_baseline"], res["cpu_baseline_channelizer"] = cpu_baselines(wl, chans, iq_host, cpu_cores())
    for c in cobjs:
        c.close()
    if spec is not None:
        spec.close()
    if grp is not None:
        grp.close()
    else:
        eng.close()
    for _, addr in host_blocks:
        lib.host_free(addr)
    del ring
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ring", type=int, default=16, help="least number of distinct input blocks rotated through (ring * block > L2 is enforced)")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="the timed K-step region is repeated until this much has been measured")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--noise-inputs", action="store_true", help="profiling aid: noise-only input blocks (no per-carrier synthesis); not a contract line")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE.json configurations (configs sub-object)")
    ap.add_argument("--channelizer", action="store_true",
                    help="diagnostic: run the engine in channelizer mode (SDRPostThread::runPFBCH, SURVEY.md 8(f) row 1) -- not a contract line")
    ap.add_argument("--channels", type=int, default=0, help="diagnostic only: override the workload's channel count (not a contract line)")
    ap.add_argument("--fft", type=int, default=-1, help="diagnostic only: override the workload's spectrum size (0 = no spectrum; not a contract line)")
    args = ap.parse_args()
    capture_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    from cubicsdr_b200 import synth
    fs = wl["fs"]
    N = synth.block_size(fs)
    if args.channels > 0 and wl["n_chan"]:
        wl = dict(wl, n_chan=args.channels, desc=wl["desc"] + " [DIAGNOSTIC: %d channels]" % args.channels)
    if args.fft >= 0:
        wl = dict(wl, fft=args.fft, desc=wl["desc"] + " [DIAGNOSTIC: spectrum size %d]" % args.fft)
    chans = synth.channel_plan(wl["plan"], fs, wl["n_chan"]) if wl["n_chan"] else []

    if args.workload == "pfb":
        if rank == 0 and args.impl != "reference":
            run_pfb(args, wl)
        return
    if args.impl == "reference":
        if args.channelizer:
            wl = dict(wl, desc=wl["desc"] + " [DIAGNOSTIC: channelizer mode (SDRPostThread::runPFBCH)]")
        run_reference(args, wl, chans, rank)
        return

    import torch
    import cubicsdr_b200
    if args.warmup < 3:
        args.warmup = 3
    cx = Ctx()
    cx.args, cx.torch, cx.rank, cx.world, cx.local_rank = args, torch, rank, world, local_rank
    torch.cuda.set_device(local_rank)
    cx.dev = torch.device("cuda", local_rank)
    cx.dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=cx.dev)
        cx.dist = dist
    cx.lib = cubicsdr_b200.lib()
    cx.fp32_peak = None
    try:
        import ctypes as C
        tf = C.c_double()
        cx.lib.L.sdr_debug_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        if cx.lib.L.sdr_debug_fp32_peak(local_rank, C.byref(tf)) == 0 and tf.value > 1.0:
            cx.fp32_peak = tf.value
    except Exception:
        pass

    # clocks / throttle reasons are sampled over the whole measured part of the run
    sampler = ClockSampler(local_rank)
    sampler.start()
    main_res = run_engine_workload(cx, wl, chans, args.steps, args.warmup, args.min_seconds, channelizer=args.channelizer,
                                   want_cpu=not args.no_cpu_baseline, want_timeline=True)
    clocks = sampler.stop()

    # the other BASELINE.json configurations (one GPU; shorter runs), reported inside the same line
    configs = None
    if world == 1 and not args.no_configs and args.workload == "cfg5" and not args.channelizer and not args.channels:
        configs = {}
        for name in ("cfg1", "cfg2", "cfg3", "cfg4"):
            w2 = WORKLOADS[name]
            c2 = synth.channel_plan(w2["plan"], w2["fs"], w2["n_chan"]) if w2["n_chan"] else []
            try:
                r2 = run_engine_workload(cx, w2, c2, max(20, min(args.steps, 100)), 3, min(0.5, args.min_seconds), want_cpu=not args.no_cpu_baseline)
                configs[name] = {"config": config_of(w2, r2["N"], len(c2)), "value": r2["value"], "unit": "MSamples/s",
                                 "ms_per_step": r2["ms_per_step"], "reps": r2["reps"], "e2e": r2["e2e"], "roofline": r2["roofline"],
                                 "gpu_launches_per_step": r2["gpu_launches"] / max(1, max(20, min(args.steps, 100))),
                                 "realtime_factor": r2["realtime_factor"], "cpu_baseline": r2.get("cpu_baseline"),
                                 "cpu_baseline_channelizer": r2.get("cpu_baseline_channelizer"),
                                 "note": "cfg4 is specified on 4 GPUs: this is the whole 256-channel job on one" if name == "cfg4" else None}
            except Exception as ex:
                configs[name] = {"error": repr(ex)}
        # the same cfg5 job the way the reference's stock dispatch runs it at this rate (SDRPostThread::runPFBCH: analysis
        # filterbank, demodulators at the channel rate); parity of that mode: tests/test_pfb.py
        try:
            r3 = run_engine_workload(cx, wl, chans, max(20, min(args.steps, 100)), 3, min(0.5, args.min_seconds), channelizer=True, want_cpu=False)
            configs["cfg5_channelizer"] = {"config": config_of(dict(wl, desc=r3["desc"]), r3["N"], len(chans)), "value": r3["value"], "unit": "MSamples/s",
                                           "ms_per_step": r3["ms_per_step"], "reps": r3["reps"], "e2e": r3["e2e"],
                                           "stage_ms_per_block": r3["roofline"]["stage_ms_per_block"],
                                           "gpu_launches_per_step": r3["gpu_launches"] / max(1, max(20, min(args.steps, 100))),
                                           "realtime_factor": r3["realtime_factor"],
                                           "note": "engine in channelizer mode (sdr_engine_set_channelizer(1)); compare with cpu_baseline_channelizer of the main line"}
        except Exception as ex:
            configs["cfg5_channelizer"] = {"error": repr(ex)}

    if rank == 0:
        r = main_res
        line = {"metric": "IQ MSamples/s through spectrum+demod chain", "value": r["value"], "unit": "MSamples/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_of(dict(wl, desc=r["desc"]), r["N"], len(chans)),
                "notes": {"channels_per_gpu": r["c_local"], "l2": r["l2"], "realtime_factor": r["realtime_factor"],
                          "timing": "K steps between barrier+synchronize, CUDA events, max over ranks; repeated %d times (>= %.1f s measured), "
                                    "median repetition reported; repetitions ranged %.3f .. %.3f ms" % (
                                        r["reps"], args.min_seconds, r["rep_ms_min_max"][0], r["rep_ms_min_max"][1]),
                          "host_cores_usable": cpu_cores(), "host_cpu": cpu_model()},
                "reps": r["reps"],
                "clocks": clocks,
                "e2e": r["e2e"],
                "gpu_launches": r["gpu_launches"],
                "roofline": r["roofline"],
                "cpu_baseline": r.get("cpu_baseline"),
                "cpu_baseline_channelizer": r.get("cpu_baseline_channelizer")}
        if configs is not None:
            line["configs"] = configs
        emit(line)
    if cx.dist is not None:
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
