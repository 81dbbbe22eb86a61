#!/usr/bin/env python3
"""bench.py -- IQ MSamples/s through the spectrum + demodulator chain (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5|cfg3|cfg2] [--impl ours|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one reference-sized IQ block (fs/60 samples, SoapySDRThread.cpp:668-674) through the whole hot path:
DC blocker -> spectrum FFT / EMA / waterfall line (rank 0) -> NCO mix + multi-stage decimation + demodulation + audio
resampling for every channel of this rank.  Channels are sharded over the ranks (strong scaling: the job is fixed);
when N > 1 the raw block is broadcast from rank 0 over NCCL inside the timed region.

`value`   device-resident inputs (a ring of distinct blocks larger than L2), results left in HBM.
`e2e`     the same through host buffers: pinned host block -> H2D -> (broadcast) -> kernels -> D2H -> batched audio fetch.
`roofline` for the dominant kernel (channel front: oscillator mix + half-band cascade), timed with CUDA events on the
           engine's stream inside the timed region.
`cpu_baseline` / `--impl reference`: the reference's own liquid-dsp binary + CubicSDR glue (oracle/_ref) on the host cores.
"""
import os
os.environ.setdefault("OMP_NUM_THREADS", "1")   # no spinning OpenMP pool next to the reference's own threads (cpu_baseline / --impl reference)
import argparse
import json
import os
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
    "cfg3": dict(fs=20000000, plan="nbfm64", n_chan=64, fft=0, desc="cfg3: 64 NBFM channels @12.5 kHz from 20 MSps IQ"),
    "cfg2": dict(fs=10000000, plan="mixed", n_chan=0, fft=8192, desc="cfg2: 8192-pt spectrum FFT + waterfall on 10 MSps IQ"),
    # SURVEY.md 8(f) row 1 (not a BASELINE config): the reference's channelizer for rates above 500 kSps, on its own
    "pfb": dict(fs=61440000, plan="mixed", n_chan=0, fft=0,
                desc="pfb: firpfbch analysis channelizer (SDRPostThread::runPFBCH), 122 channels on 61.44 MSps IQ"),
}
FP32_NOMINAL_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12    # non-tensor FP32 issue peak of a B200 at max clock


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(wl, chans, iq_host, threads, blocks, channelizer=0):
    """Time the reference CPU path (oracle/_ref: reference liquid-dsp binary + CubicSDR glue) on `blocks` blocks.
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


def run_reference(args, wl, chans, rank):
    """--impl reference: the reference's own CPU implementation on the host cores, same config / metric / unit."""
    from cubicsdr_b200 import synth
    if rank != 0:
        return
    N = synth.block_size(wl["fs"])
    iq = synth.wideband(wl["fs"], N, chans[:: max(1, len(chans) // 32)] if chans else [], t0=0)   # carriers: a 32-channel subset keeps generation quick
    threads = os.cpu_count() or 1
    secs, ver = cpu_reference_run(wl, chans, iq, threads, args.steps, channelizer=1 if args.channelizer else 0)
    t_tot = secs
    val = N * args.steps / t_tot / 1e6
    line = {"impl": "reference", "metric": "IQ MSamples/s through spectrum+demod chain", "value": val, "unit": "MSamples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "fs": wl["fs"], "block": N, "channels": len(chans), "fft_size": wl["fft"]},
            "cpu_baseline": {"value": val, "unit": "MSamples/s", "cores": threads, "kind": "reference",
                             "sample": "1 block (1/60 s of IQ) per step after 1 untimed block, all channels + spectrum, liquid-dsp %s" % ver},
            "e2e": {"value": val, "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_pfb(args, wl):
    """--workload pfb: the polyphase analysis channelizer alone, device-resident and through host
    buffers, next to the reference binary's firpfbch_crcf on one host core (the reference runs it on one thread)."""
    import ctypes as C
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
        out_host = ch.execute(host_in)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ring", type=int, default=16, help="distinct input blocks rotated through (ring * block > L2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--channelizer", action="store_true",
                    help="diagnostic: run the engine in channelizer mode (SDRPostThread::runPFBCH, SURVEY.md 8(f) row 1) -- not a contract line")
    ap.add_argument("--channels", type=int, default=0, help="diagnostic only: override the workload's channel count (not a contract line)")
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
    from cubicsdr_b200 import Channel, Engine, Spectrum
    if args.warmup < 3:
        args.warmup = 3
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    lib = cubicsdr_b200.lib()

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
    # a dedicated (non-default) stream shared by torch and the engine: torch.cuda.Event timing then sees the kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng = Engine(lib, device=local_rank, max_block=N, stream=stream.cuda_stream)
    mine = chans[rank::world]                                  # channel sharding
    chan_fs = fs
    if args.channelizer:
        eng.set_channelizer(1)
        n_pfb, chan_fs = eng.channel_plan(fs)
        eng.set_channel_block_max(N // n_pfb)
        wl = dict(wl, desc=wl["desc"] + " [DIAGNOSTIC: channelizer mode, %d filterbank channels of %d Hz]" % (n_pfb, chan_fs))
    cobjs = [Channel(eng, m, chan_fs, bw, 48000, frequency=CENTER + off) for m, bw, off in mine]
    spec = Spectrum(eng, wl["fft"], 0) if (wl["fft"] and rank == 0) else None

    # ---- inputs: ring of distinct blocks resident in HBM on rank 0 (ring * block bytes > 126 MB L2) -----------------
    ring = []
    carriers = chans if chans else synth.channel_plan("mixed", fs, 16)
    if rank == 0:
        for r in range(args.ring):
            ring.append(torch_wideband(torch, fs, N, carriers, r * N, dev, 1234))
    recvs = [torch.empty(N, dtype=torch.complex64, device=dev) for _ in range(4)] if world > 1 else None
    host_blocks = None
    if rank == 0:
        host_blocks = [torch.empty(N, dtype=torch.complex64).pin_memory() for _ in range(2)]
        host_blocks[0].copy_(ring[0]); host_blocks[1].copy_(ring[1 % args.ring])
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()

    side = torch.cuda.Stream(device=dev)                       # input production (NCCL broadcast) off the engine's stream
    bc_done = [torch.cuda.Event() for _ in range(4)]

    def step_device(i):
        buf = ring[i % args.ring] if rank == 0 else recvs[i & 3]
        ev = None
        if world > 1:
            with torch.cuda.stream(side):
                dist.broadcast(buf, src=0)                        # the path's one exchange step (raw IQ over NVLink)
                bc_done[i & 3].record(side)
            ev = bc_done[i & 3].cuda_event
        # resident block (or ready at `ev`): the DC blocker of this block overlaps the previous block's channel kernels
        eng.submit_device_async(buf.data_ptr(), N, CENTER, fs, fetch_to_host=False, ready_event=ev)

    # end to end: pinned host block -> H2D (copy stream, overlapped with the previous block's kernels) -> (broadcast) ->
    # kernels -> D2H of every result -> batched fetch into host arrays.  Public API only: submit_device / wait / fetch.
    copy_stream = torch.cuda.Stream(device=dev)
    stages = [torch.empty(N, dtype=torch.complex64, device=dev) for _ in range(4)]
    h2d_done = [torch.cuda.Event() for _ in range(4)]

    def h2d(i):
        if rank == 0:
            with torch.cuda.stream(copy_stream):
                stages[i % 4].copy_(host_blocks[i & 1], non_blocking=True)
                h2d_done[i % 4].record(copy_stream)

    def submit_e2e(i):
        buf = stages[i % 4]
        ev = h2d_done[i % 4]
        if world > 1:
            with torch.cuda.stream(side):
                if rank == 0:
                    side.wait_event(h2d_done[i % 4])
                dist.broadcast(buf, src=0)
                bc_done[i & 3].record(side)
            ev = bc_done[i & 3]
        eng.submit_device_async(buf.data_ptr(), N, CENTER, fs, fetch_to_host=True, ready_event=ev.cuda_event if (rank == 0 or world > 1) else None)

    def fetch_e2e():
        nbytes = 0
        if cobjs:
            audio, n_audio, _, _ = eng.fetch_audio_batch(cobjs, 1024)
            nbytes += int(n_audio.sum()) * 4
        if spec is not None:
            pts, lines, cf = spec.fetch_all(64)
            nbytes += pts.nbytes + lines.nbytes
        return nbytes

    # Two blocks in flight, like the reference's thread pipeline (queues between its threads): per step, submit block i+2,
    # start the H2D of block i+3 (its stage buffer was last read by block i-1, whose results were already waited for), then
    # wait for and fetch the results of block i.  Every step still costs one H2D, one pass of the path and one D2H + fetch.
    host_split = [0.0, 0.0, 0.0]                                # host seconds in submit+h2d / wait / fetch (diagnostic, stderr)

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

    # ---- warm-up ------------------------------------------------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    eng.wait()
    torch.cuda.synchronize()

    # ---- timed: device-resident -----------------------------------------------------------------------------------------
    # clocks / throttle reasons are sampled from just before the timed region until the end of the e2e region
    # (nvidia-smi cannot sample faster than ~10 Hz, and the device-resident region alone lasts a fraction of a second)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(60):                                         # ~50 ms of the same work so the first samples see load
        step_device(i)
    eng.wait()
    eng.set_timing(True)
    launches0 = eng.launch_count
    barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step_device(args.warmup + i)
    e1.record(stream)
    torch.cuda.synchronize(); barrier()
    ms = e0.elapsed_time(e1)
    eng.wait()
    stage_ms, stage_blocks = eng.get_timing()
    if rank == 0:
        print_timeline("device")
    eng.set_timing(False)
    launches = eng.launch_count - launches0
    if dist is not None:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = N * args.steps / (ms * 1e-3) / 1e6

    # ---- timed: end to end through host buffers -------------------------------------------------------------------------
    h2d(0)
    submit_e2e(0)
    h2d(1)
    submit_e2e(1)
    h2d(2)
    for i in range(3):
        step_e2e(i)
    torch.cuda.synchronize()
    barrier(); torch.cuda.synchronize()
    eng.set_timing(True)                                        # stage times of the first 64 e2e blocks (stderr diagnostic)
    t0 = time.perf_counter()
    d2h = 0
    for i in range(args.steps):
        d2h = step_e2e(3 + i)
    eng.wait()                                                  # drain the blocks still in flight (submitted, not yet fetched)
    torch.cuda.synchronize(); barrier()
    t_e2e = time.perf_counter() - t0
    e2e_stage_ms, e2e_stage_blocks = eng.get_timing()
    eng.set_timing(False)
    if rank == 0:
        print_timeline("e2e")
        print("e2e device stages per block (ms): %s" % {k: round(v / max(1, e2e_stage_blocks), 3) for k, v in e2e_stage_ms.items()}, file=sys.stderr)
        print("e2e host split per step (ms): submit+h2d %.3f  wait %.3f  fetch %.3f" % tuple(1e3 * v / (args.steps + 3) for v in host_split), file=sys.stderr)
    if dist is not None:
        tt = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
        dd = torch.tensor([d2h], device=dev, dtype=torch.float64)
        dist.all_reduce(dd, op=dist.ReduceOp.SUM)
        d2h = int(dd.item())
    e2e_value = N * args.steps / t_e2e / 1e6
    clocks = sampler.stop()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------------
    peak, peak_src = measured_peaks()
    c_local = len(mine)
    b_alg = 8.0 + c_local * 4.0 * 48000.0 / fs + (2.5 if spec is not None else 0.0)     # bytes / wideband sample / GPU
    dom = "front" if c_local else "spectrum"
    t_dom_ms = stage_ms[dom] / max(1, stage_blocks)
    achieved = (N * b_alg / (t_dom_ms * 1e-3)) / 1e9 if t_dom_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(args.workload, {}).get(dom)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "note": "contract figure (algorithmic bytes / kernel time / measured HBM peak); with hundreds of channels in "
                                        "single-channel mode the kernel is FP32-pipe bound, see the fp32 object" if c_local else
                                        "contract figure (algorithmic bytes / kernel time / measured HBM peak)", "kernel": "k_chan_front (NCO mix + half-band cascade)" if c_local else "spectrum kernels",
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "alg_bytes_per_sample": b_alg, "kernel_ms_per_block": t_dom_ms,
                "stage_ms_per_block": {k: v / max(1, stage_blocks) for k, v in stage_ms.items()}}
    if c_local:
        flops = 30.0 * c_local * N / (t_dom_ms * 1e-3) / 1e12
        roofline["fp32"] = {"note": "many-channel single-channel-mode is FP32-issue bound (SURVEY.md 8d): ~30 flop / sample / channel",
                            "achieved_tflops": flops, "nominal_peak_tflops": FP32_NOMINAL_TFLOPS, "frac": flops / FP32_NOMINAL_TFLOPS}

    # ---- CPU baseline (bounded sample: one block, all channels) -------------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        try:
            threads = os.cpu_count() or 1
            iq_host = ring[0].cpu().numpy()
            secs, ver = cpu_reference_run(wl, chans, iq_host, threads, 1, channelizer=1 if args.channelizer else 0)
            cpu = {"value": N / secs / 1e6, "unit": "MSamples/s", "cores": threads, "kind": "reference",
                   "sample": "1 block (%d samples = 1/60 s), all %d channels%s, reference liquid-dsp %s binary + CubicSDR glue" % (
                       N, len(chans), " + spectrum" if wl["fft"] else "", ver)}
        except Exception as ex:        # the oracle is a checker; its absence must not hide the GPU number
            cpu = {"value": None, "unit": "MSamples/s", "cores": 0, "kind": "reference", "sample": "unavailable: %r" % (ex,)}

    line = {"metric": "IQ MSamples/s through spectrum+demod chain", "value": value, "unit": "MSamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "fs": fs, "block": N, "channels": len(chans), "channels_per_gpu": c_local,
                       "fft_size": wl["fft"], "l2": "%d distinct %.1f MB input blocks rotate (%.0f MB > 126 MB L2)" % (
                           args.ring, N * 8 / 1e6, args.ring * N * 8 / 1e6),
                       "realtime_factor": value * 1e6 / fs},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "MSamples/s", "h2d_bytes_per_step": N * 8, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
