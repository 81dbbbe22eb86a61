"""The reference's OWN, unmodified demodulator -- a real DemodulatorInstance with its DemodulatorPreThread, DemodulatorWorkerThread,
DemodulatorThread and modem classes running on their own threads (src/demod/*.cpp, src/modules/modem/*, compiled from where they
lie into oracle/_ref/libdemod_ref.so over the reference's own liquid-dsp binary, oracle/build_ref.py step 8) -- as the pin of
the restated per-channel chain every parity test compares against (oracle/ref_chain.cpp: RefChan): the same wideband blocks in,
the SAME BITS of audio, peak, signal level / floor / ceiling and squelch decision out, block by block."""
import os

import numpy as np
import pytest

from cubicsdr_b200 import synth
from oracle import refbind
from tests.parity import CENTER

FS = 2400000
CASES = [("FM", 200000, 300000, "FM"), ("NBFM", 12500, -150000, "NBFM"), ("AM", 6000, 450000, "AM"), ("USB", 5400, -600000, "USB"),
         ("LSB", 5400, 750000, "LSB"), ("DSB", 5400, 100000, "AM"), ("CW", 500, -200000, "AM"), ("I/Q", 48000, 500000, "AM"),
         ("FMS", 200000, -500000, "FMS")]


@pytest.fixture(scope="module")
def demod_lib():
    if not os.path.exists(refbind.DEMOD_LIB) and not os.path.exists("/root/reference"):
        pytest.skip("oracle/_ref/libdemod_ref.so not built and /root/reference absent")
    return refbind.RefDemodLib()


def _same_block(r, s, what):
    assert r["skipped"] == s["skipped"], what
    assert r["audio"].size == s["audio"].size, "%s: %d samples, the reference's demodulator %d" % (what, r["audio"].size, s["audio"].size)
    assert np.array_equal(r["audio"], s["audio"]), "%s: restated chain differs from the reference's demodulator (max %g)" % (
        what, np.abs(r["audio"] - s["audio"]).max())
    if not r["skipped"]:
        assert np.float32(r["peak"]) == np.float32(s["peak"]), what
        assert r["squelched"] == s["squelched"], what
    for k in ("signal_level", "signal_floor", "signal_ceil"):        # std::atomic<float> members of DemodulatorThread
        assert np.float32(r[k]) == np.float32(s[k]), "%s: %s %r vs %r" % (what, k, r[k], s[k])


def _run(ref, demod_lib, modem, bw, off, on_air, n_blocks, block=None, squelch=None, retune=None, signal=None):
    N = block or synth.block_size(FS)
    rc, sd, ing = refbind.RefChan(ref, modem, FS, bw), refbind.SrcDemod(demod_lib, modem, FS, bw, CENTER + off, CENTER), refbind.RefIngress(ref)
    assert sd.bandwidth == rc.bandwidth
    if squelch is not None:
        rc.set_squelch(True, squelch); sd.set_squelch(True, squelch)
    sig = signal or [(on_air, 6000 if on_air == "AM" else bw, off)]
    flags = []
    for b in range(n_blocks):
        if retune and b in retune:
            off = retune[b]
            sd.set_frequency(CENTER + off)
        y = ing.block(synth.wideband(FS, N, sig, seed=7, t0=b * N))
        r, s = rc.process(y, off), sd.process(y, CENTER)
        _same_block(r, s, "%s block %d" % (modem, b))
        flags.append((r["skipped"], r["squelched"]))
    rc.close(); sd.close(); ing.close()
    return flags


@pytest.mark.parametrize("modem,bw,off,on_air", CASES, ids=[c[0] for c in CASES])
def test_restated_chain_equals_the_reference_demodulator(ref, demod_lib, modem, bw, off, on_air):
    _run(ref, demod_lib, modem, bw, off, on_air, 8)
    _run(ref, demod_lib, modem, bw, off, on_air, 5, block=40003)


def test_retune_zero_shift_and_skip_rule(ref, demod_lib):
    """setFrequency between blocks: up, down, onto the tuner frequency (no mixing), beyond 0.75 fs (the block is dropped,
    DemodulatorPreThread.cpp:161-164) and back."""
    signal = [("FM", 200000, 300000), ("FM", 200000, -620000), ("FM", 200000, 0), ("FM", 200000, 850000)]
    flags = _run(ref, demod_lib, "FM", 200000, 300000, "FM", 12, signal=signal,
                 retune={2: -620000, 4: 0, 6: 2000000, 8: 850000, 10: -1799999})
    assert flags[6][0] and flags[7][0] and not flags[8][0]


def test_squelch_decisions(ref, demod_lib):
    """DemodulatorThread.cpp:165-198: level trackers and the squelch decision for levels above and below the signal."""
    seen = set()
    for modem, bw, off, on_air, level in (("FM", 200000, 300000, "FM", -25.0), ("NBFM", 12500, -150000, "NBFM", -5.0),
                                          ("AM", 6000, 450000, "AM", -40.0), ("USB", 5400, -600000, "USB", -70.0),
                                          ("FM", 200000, -800000, "FM", 10.0)):
        flags = _run(ref, demod_lib, modem, bw, off, on_air, 10, squelch=level)
        seen.update(sq for _, sq in flags)
    assert seen == {True, False}


def test_odd_and_unusual_bandwidths(ref, demod_lib):
    """Modem::checkSampleRate changes the rate the kit and the resampler are built at (odd SSB bandwidths, MIN_BANDWIDTH) while the
    metering keeps dividing by the bandwidth that was asked for (DemodulatorPreThread.cpp:222, DemodulatorThread.cpp:143)."""
    for modem, bw, off, on_air in (("USB", 5401, -600000, "USB"), ("LSB", 2999, 750000, "LSB"), ("AM", 12345, 450000, "AM"),
                                   ("NBFM", 7777, -150000, "NBFM"), ("FM", 96000, 300000, "FM"), ("CW", 701, -200000, "AM")):
        _run(ref, demod_lib, modem, bw, off, on_air, 4)
