#!/usr/bin/env python3
"""Diagnostic: per-CTA timeline of the channel front kernel (smid, start, after table fill, end) for the cfg5 workload."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import cubicsdr_b200
from cubicsdr_b200 import Channel, Engine, synth
import bench

fs = 61440000; N = synth.block_size(fs)
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
chans = synth.channel_plan("mixed", fs, nch)
# (both mixing directions now share one launch)
lib = cubicsdr_b200.lib()
dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
eng = Engine(lib, max_block=N, stream=st.cuda_stream)
cs = [Channel(eng, m, fs, bw, 48000, frequency=bench.CENTER + off) for m, bw, off in chans]
x = bench.torch_wideband(torch, fs, N, chans[::16], 0, dev, 1)
for _ in range(3):
    eng.submit_device(x.data_ptr(), N, bench.CENTER, fs, fetch_to_host=False)
eng.wait()
lib.L.sdr_engine_debug_front_timeline.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint]
lib.L.sdr_engine_debug_front_timeline(eng.h, 1, None, 0)
eng.submit_device(x.data_ptr(), N, bench.CENTER, fs, fetch_to_host=False)
eng.wait()
out = np.zeros(4 * 1024, np.uint64)
lib.L.sdr_engine_debug_front_timeline(eng.h, 1, out.ctypes.data_as(C.c_void_p), 1024)
d = out.reshape(-1, 4)
d = d[d[:, 1] > 0]
t0 = d[:, 1].min()
print("CTAs recorded (last front launch):", len(d))
start = (d[:, 1] - t0) / 1e3; fill = (d[:, 2] - d[:, 1]) / 1e3; dur = (d[:, 3].astype(np.int64) - d[:, 2].astype(np.int64)) / 1e3
print("start offset us: min %.1f med %.1f max %.1f" % (start.min(), np.median(start), start.max()))
print("table fill  us: min %.1f med %.1f max %.1f" % (fill.min(), np.median(fill), fill.max()))
print("work        us: min %.1f p10 %.1f med %.1f p90 %.1f max %.1f" % (dur.min(), np.percentile(dur, 10), np.median(dur), np.percentile(dur, 90), dur.max()))
order = np.argsort(dur)
print("slowest CTAs (cta, smid, work us):", [(int(i), int(d[i, 0]), round(float(dur[i]), 1)) for i in order[-8:]])
print("fastest CTAs (cta, smid, work us):", [(int(i), int(d[i, 0]), round(float(dur[i]), 1)) for i in order[:8]])
np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "front_timeline.npy"), d)
