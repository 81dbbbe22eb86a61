"""The reference's OWN, unmodified modem classes (src/modules/modem/Modem.cpp, ModemAnalog.cpp, analog/Modem*.cpp -- compiled
from where they lie into oracle/_ref/libmodems_ref.so over the reference's own liquid-dsp binary, oracle/build_ref.py step 6)
as the pin of everything the parity tests compare against:

  * the oracle's restated modem glue (oracle/ref_chain.cpp) must give the SAME BITS as the reference's classes on the same
    decimated IQ, for all nine analog modems;
  * Modem::checkSampleRate of the classes == the restated rule == the C-ABI's rule;
  * the product's Modem plugin seam (sdr_modem_create / sdr_modem_demodulate = buildKit / demodulate, Modem.h:148-153) against
    the classes directly -- emulator here, the CUDA path in the gpu tier.
"""
import ctypes as C

import numpy as np
import pytest

from cubicsdr_b200 import synth
from oracle import refbind
from tests import parity

FS = 2400000
# (modem, bandwidth, offset, what is on the air there)
CASES = [("FM", 200000, 300000, "FM"), ("NBFM", 12500, -150000, "NBFM"), ("AM", 6000, 450000, "AM"), ("USB", 5400, -600000, "USB"),
         ("LSB", 5400, 750000, "LSB"), ("DSB", 5400, 100000, "AM"), ("CW", 500, -200000, "AM"), ("I/Q", 48000, 500000, "AM"),
         ("FMS", 200000, -500000, "FMS")]


@pytest.fixture(scope="module")
def mods():
    import os
    if not os.path.exists(refbind.MODEMS_LIB) and not os.path.exists("/root/reference"):
        pytest.skip("oracle/_ref/libmodems_ref.so not built and /root/reference absent")
    return refbind.RefModems()


def _decimated_iq(ref, modem, bw, off, on_air, n_blocks, block=None, seed=7):
    """Blocks of decimated IQ for one channel + the restated glue's audio for them (oracle/ref_chain.cpp)."""
    N = block or synth.block_size(FS)
    rc, ing = refbind.RefChan(ref, modem, FS, bw), refbind.RefIngress(ref)
    sig = [(on_air, 6000 if on_air == "AM" else bw, off)]
    out = []
    for b in range(n_blocks):
        r = rc.process(ing.block(synth.wideband(FS, N, sig, seed=seed, t0=b * N)), off)
        out.append((r["iq"], r["audio"]))
    rate = rc.bandwidth
    rc.close(); ing.close()
    return rate, out


@pytest.mark.parametrize("modem,bw,off,on_air", CASES, ids=[c[0] for c in CASES])
def test_restated_glue_equals_the_reference_classes_bit_for_bit(ref, mods, modem, bw, off, on_air):
    for block in (None, 40003, 4099):
        rate, blocks = _decimated_iq(ref, modem, bw, off, on_air, 8 if block != 4099 else 24, block)
        rm = refbind.RefModem(mods, modem, bw)
        assert rm.sample_rate == rate
        total = 0
        for b, (iq, audio) in enumerate(blocks):
            a, channels, _ = rm.demodulate(iq)
            assert a.size == audio.size, "%s block %d: %d samples, restated glue %d" % (modem, b, a.size, audio.size)
            assert np.array_equal(a, audio), "%s block %d: restated glue differs from the reference's class (max %g)" % (
                modem, b, np.abs(a - audio).max())
            assert channels == (2 if modem in ("I/Q", "FMS") else 1)
            total += a.size
        assert total > 0
        rm.close()


def test_check_sample_rate_and_registry_equal_the_reference(ref, mods, product_lib):
    defaults = {"FM": 200000, "NBFM": 12500, "FMS": 200000, "AM": 6000, "CW": 500, "LSB": 5400, "USB": 5400, "DSB": 5400, "I/Q": 48000}
    for name, rate in defaults.items():
        assert mods.default_rate(name) == rate                      # CubicSDR.cpp:305-313
    assert not mods.L.refmodem_create(b"nope")
    for name, bw, _, _ in CASES:
        m = refbind.RefModem(mods, name, bw)
        assert m.type == "analog"
        for rate in (1, 499, 500, 501, 1499, 1500, 5399, 5400, 5401, 6000, 12345, 44100, 48000, 100000, 199999, 200000, 250001, 1000000):
            for audio_rate in (44100, 48000):
                want = mods.L.refmodem_check_sample_rate(m.h, rate, audio_rate)
                assert product_lib.L.sdr_modem_check_sample_rate2(name.encode(), rate, audio_rate) == want, (name, rate, audio_rate)
                if audio_rate == 48000 and name != "I/Q":
                    assert ref.L.ref_modem_check_sample_rate(name.encode(), rate) == want, (name, rate)
        m.close()


def _seam(lib, ref, mods, cases, n_blocks=8, tol=parity.TOL):
    """sdr_modem_create + sdr_modem_demodulate against Modem::buildKit + Modem::demodulate of the reference's classes."""
    from cubicsdr_b200 import Engine
    eng = Engine(lib, max_block=synth.block_size(FS))
    L = lib.L
    for modem, bw, off, on_air in cases:
        rate, blocks = _decimated_iq(ref, modem, bw, off, on_air, n_blocks)
        rm = refbind.RefModem(mods, modem, bw)
        h = C.c_void_p()
        assert L.sdr_modem_create(eng.h, modem.encode(), rate, 48000, 65536, C.byref(h)) == 0
        got, want = [], []
        for iq, _ in blocks:
            iq = np.ascontiguousarray(iq, np.complex64)
            au = np.zeros(4 * iq.size + 4096, np.float32)
            n, lvl, pk = C.c_uint(), C.c_double(), C.c_float()
            assert L.sdr_modem_demodulate(h, iq.ctypes.data_as(C.c_void_p), iq.size, au.ctypes.data_as(C.c_void_p), au.size,
                                          C.byref(n), C.byref(lvl), C.byref(pk)) == 0
            a, _, _ = rm.demodulate(iq)
            assert n.value == a.size, "%s: %d samples, the reference's class %d" % (modem, n.value, a.size)
            got.append(au[:n.value].copy()); want.append(a)
            if a.size:                                       # DemodulatorThread.cpp:223-233: max |sample| of what was produced
                assert pk.value == np.abs(au[:n.value]).max()
                if modem != "FMS":
                    assert abs(pk.value - np.abs(a).max()) <= 1e-4 * max(np.abs(a).max(), 1e-3)
        got, want = np.concatenate(got), np.concatenate(want)
        if modem == "FMS":          # the sum channel to the usual bar; the difference channel goes through the chaotic pilot loop
            g, w = got.reshape(-1, 2), want.reshape(-1, 2)
            assert parity.relrms(g.sum(1), w.sum(1)) <= tol
            assert parity.relrms(g[:, 1] - g[:, 0], w[:, 1] - w[:, 0]) <= 2e-2      # (the precise bar for this loop: parity.check_fm_stereo)
        else:
            e = parity.relrms(got, want)
            assert e <= tol, "%s through the plugin seam: %.3g" % (modem, e)
        assert L.sdr_chan_destroy(h) == 0
        rm.close()
    eng.close()


def test_modem_plugin_seam_against_the_reference_classes_emulator(emu_lib, ref, mods):
    _seam(emu_lib, ref, mods, CASES)


@pytest.mark.gpu
def test_modem_plugin_seam_against_the_reference_classes_gpu(gpu_lib, ref, mods):
    _seam(gpu_lib, ref, mods, CASES, n_blocks=12)
