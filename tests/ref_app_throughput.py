"""Diagnostic (not a test, not the bench): throughput of the reference's OWN pipeline -- a real SDRPostThread feeding real
DemodulatorInstances on their own threads (oracle/_ref/libdemod_ref.so, built from the unmodified reference sources) -- next to
the restated chain bench.py times as `cpu_baseline` (oracle/ref_chain.cpp: ref_bench_path, one worker per host core), on the
same blocks.  Shows what the baseline's restatement and threading are worth against the real thing.

    python tests/ref_app_throughput.py [cfg1|nbfm16] [blocks]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cubicsdr_b200 import synth                                  # noqa: E402
from oracle import refbind                                       # noqa: E402

CENTER = 100000000


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
    n_blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    if which == "cfg1":
        fs, chans = 2400000, [("FM", 200000, 300000)]
    else:
        fs, chans = 20000000, [("NBFM", 12500, -393750 + 12500 * k) for k in range(16)]
    N = synth.block_size(fs)
    blocks = [synth.wideband(fs, N, chans[:4], seed=1, t0=b * N) for b in range(4)]
    ref, lib = refbind.Ref(), refbind.RefDemodLib()
    # the real pipeline
    app = refbind.SrcApp(lib, 1)
    sds = [refbind.SrcDemod(lib, m, fs, bw, CENTER + off, CENTER) for m, bw, off in chans]
    for d in sds:
        app.add(d)
    for b in range(2):                                           # warm-up
        app.block(blocks[b % 4], CENTER, fs, 1)
        for d in sds:
            d.collect(CENTER)
    t0 = time.perf_counter()
    for b in range(n_blocks):
        app.block(blocks[b % 4], CENTER, fs, 1)
        for d in sds:
            d.collect(CENTER)
    t_app = time.perf_counter() - t0
    for d in sds:
        d.close()
    app.close()
    # the restated chain, the way the parity tests drive it (one thread)
    rcs, ing = [refbind.RefChan(ref, m, fs, bw) for m, bw, _ in chans], refbind.RefIngress(ref)
    t0 = time.perf_counter()
    for b in range(n_blocks):
        y = ing.block(blocks[b % 4])
        for rc, (_, _, off) in zip(rcs, chans):
            rc.process(y, off)
    t_one = time.perf_counter() - t0
    ms = n_blocks * N / 1e6
    print("%s: %d channels, %d blocks of %d samples at %.2f MSps, %d host cores" % (which, len(chans), n_blocks, N, fs / 1e6, len(os.sched_getaffinity(0))))
    print("  reference's own pipeline (its threads)      : %7.2f MS/s" % (ms / t_app))
    print("  restated chain, single thread               : %7.2f MS/s" % (ms / t_one))


if __name__ == "__main__":
    main()
