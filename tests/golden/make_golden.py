#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the reference's own binary (oracle/_ref/libsdr_ref.so).

Run in the build container (where /root/reference exists and oracle/build_ref.py has produced the library):
    python tests/golden/make_golden.py
The vectors pin the oracle itself (tests/test_oracle.py): a transplant that silently broke on another machine would
no longer reproduce them bit-for-bit.  Inputs are regenerated from seeds by cubicsdr_b200.synth, only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cubicsdr_b200 import synth  # noqa: E402
from oracle.refbind import Ref, RefChan, RefDistrib, RefIngress, RefSpectrum  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FIVE = [("FM", 200000, 300000), ("NBFM", 12500, -150000), ("AM", 6000, 450000), ("USB", 5400, -600000), ("LSB", 5400, 750000)]


def main():
    ref = Ref()
    fs, N = 2400000, synth.block_size(2400000)
    out = {}
    ing = RefIngress(ref)
    rcs = [RefChan(ref, m, fs, bw) for m, bw, _ in FIVE]
    sp = RefSpectrum(ref, 256)
    rd = RefDistrib(ref, 512, 100000000)
    for b in range(2):
        x = synth.wideband(fs, N, FIVE, seed=1234, t0=b * N)
        y = ing.block(x)
        out["dc_head_b%d" % b] = y[:64]
        for i, rc in enumerate(rcs):
            r = rc.process(y, FIVE[i][2])
            out["audio_%s_b%d" % (FIVE[i][0], b)] = r["audio"]
            out["iq_%s_b%d" % (FIVE[i][0], b)] = r["iq"][:128]
            out["lvl_%s_b%d" % (FIVE[i][0], b)] = np.array([r["level_db"], r["peak"]])
        frames = rd.push(y, fs)
        for f in range(len(frames)):                   # every frame advances the averages; the first and last are kept
            r = sp.frame(frames[f])
            if f in (0, len(frames) - 1):
                out["spec_points_b%d_f%d" % (b, f)] = r["points"]
                out["spec_line_b%d_f%d" % (b, f)] = r["line"]
    np.savez_compressed(os.path.join(HERE, "chain_2p4msps.npz"), **out)
    print("wrote chain_2p4msps.npz with", len(out), "arrays")
    channelizer_golden(ref)


def channelizer_golden(ref):
    """SDRPostThread::runPFBCH / runPFBCH2: the reference binary's firpfbch / firpfbch2 on a seeded
    block, and the audio of two demodulators fed through the channelizer path (tests/parity.RefChannelizerPath)."""
    from oracle.refbind import RefPfb, RefPfb2
    from tests.parity import CENTER, RefChannelizerPath
    out = {}
    rng = np.random.default_rng(20260922)
    for M in (8, 122):
        x = (rng.standard_normal(M * 24) + 1j * rng.standard_normal(M * 24)).astype(np.complex64)
        out["pfb_in_M%d" % M] = x
        out["pfb_out_M%d" % M] = RefPfb(ref, M).execute(x)
        out["pfb2_out_M%d" % M] = RefPfb2(ref, M).execute(x)
    fs, N = 2400000, synth.block_size(2400000)
    chans = [("FM", 200000, 40000), ("AM", 6000, -650000)]
    for mode in (1, 2):
        path = RefChannelizerPath(ref, fs, chans, mode=mode)
        for b in range(2):
            x = synth.wideband(fs, N, chans, seed=77, t0=b * N)
            res, picks = path.block(x, CENTER, [CENTER + off for _, _, off in chans])
            for (m, _, _), r in zip(chans, res):
                out["mode%d_audio_%s_b%d" % (mode, m, b)] = r["audio"]
        path.close()
    np.savez_compressed(os.path.join(HERE, "channelizer.npz"), **out)
    print("wrote channelizer.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
