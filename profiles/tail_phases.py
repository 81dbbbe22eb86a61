#!/usr/bin/env python3
"""Diagnostic: per-phase duration of the channel tail kernel (cfg5 workload) for the first channels of each modem."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cubicsdr_b200
from cubicsdr_b200 import Channel, Engine, Spectrum, synth
import bench
fs = 61440000; N = synth.block_size(fs)
chans = synth.channel_plan("mixed", fs, 512)
lib = cubicsdr_b200.lib(); dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
eng = Engine(lib, max_block=N, stream=st.cuda_stream)
cs = [Channel(eng, m, fs, bw, 48000, frequency=bench.CENTER + off) for m, bw, off in chans]
xs = [bench.torch_wideband(torch, fs, N, chans[::16], r * N, dev, 1) for r in range(2)]
for i in range(3):
    eng.submit_device_async(xs[i % 2].data_ptr(), N, bench.CENTER, fs, fetch_to_host=False)
eng.wait()
lib.L.sdr_engine_debug_front_timeline(eng.h, 1, None, 0)
eng.submit_device_async(xs[1].data_ptr(), N, bench.CENTER, fs, fetch_to_host=False)
eng.wait()
out = np.zeros((64, 64), np.uint64); nph = C.c_int()
lib.L.sdr_engine_debug_tail_phases.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
lib.L.sdr_engine_debug_tail_phases(eng.h, out.ctypes.data_as(C.c_void_p), C.byref(nph))
P = nph.value
names = ["SETUP"] + ["HB%d" % i for i in range(16)] + ["ARB", "DEMOD_A", "DEMOD_B", "AGC_A", "AGC_B", "AGC_C"] + ["AUD%d" % i for i in range(9)] + ["METER_A", "METER_B", "HIST_A", "HIST_B"]
seen = {}
for ch in range(64):
    m = chans[ch][0]
    if m in seen and seen[m] >= 2:
        continue
    seen[m] = seen.get(m, 0) + 1
    t = out[ch].astype(np.int64)
    start = t[63]
    prev = start
    parts = []
    for ph in range(P):
        d = (t[ph] - prev) / 1e3
        prev = t[ph]
        if d >= 0.7:
            parts.append("%s %.1f" % (names[ph] if ph < len(names) else ph, d))
    print("ch %2d %-5s bw %6d total %.1f us :: %s" % (ch, m, chans[ch][1], (t[P - 1] - start) / 1e3, "  ".join(parts)))
t0 = out[:, 63].astype(np.int64); t1 = out[:, P - 1].astype(np.int64)
print("first 64 CTAs: start spread %.1f us, end spread %.1f us" % ((t0.max() - t0.min()) / 1e3, (t1.max() - t1.min()) / 1e3))
