#!/usr/bin/env python3
"""Diagnostic: device timeline of consecutive blocks (DC stream vs main stream) for the cfg5 workload, device-resident."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cubicsdr_b200
from cubicsdr_b200 import Channel, Engine, Spectrum, synth
import bench
fs = 61440000; N = synth.block_size(fs)
nch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
chans = synth.channel_plan("mixed", fs, nch)
lib = cubicsdr_b200.lib(); dev = torch.device("cuda", 0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
eng = Engine(lib, max_block=N, stream=st.cuda_stream)
cs = [Channel(eng, m, fs, bw, 48000, frequency=bench.CENTER + off) for m, bw, off in chans]
sp = Spectrum(eng, 16384, 0)
xs = [bench.torch_wideband(torch, fs, N, chans[::16], r * N, dev, 1) for r in range(4)]
for i in range(6):
    eng.submit_device_async(xs[i % 4].data_ptr(), N, bench.CENTER, fs, fetch_to_host=False)
eng.wait()
eng.set_timing(True)
for i in range(8):
    eng.submit_device_async(xs[i % 4].data_ptr(), N, bench.CENTER, fs, fetch_to_host=False)
eng.wait()
out = np.zeros((16, 8), np.float32); n = C.c_int()
lib.L.sdr_engine_debug_timeline(eng.h, out.ctypes.data_as(C.c_void_p), C.byref(n))
print("block: dc_start dc_end | front_start front_end tail_end | spec_start spec_end tail_start   (ms)")
for i in range(n.value):
    print(i, " ".join("%7.3f" % v for v in out[i]))
