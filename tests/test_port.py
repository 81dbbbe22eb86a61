"""oracle/liquid_port.c (plain-C restatement of the path) pinned against the reference's own binary (oracle/_ref)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cubicsdr_b200 import synth
from oracle.refbind import RefChan, RefDistrib, RefIngress, RefSpectrum
from tests.parity import relrms

FIVE = [("FM", 200000, 300000), ("NBFM", 12500, -150000), ("AM", 6000, 450000), ("USB", 5400, -600000), ("LSB", 5400, 750000)]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def port():
    from oracle import build_port
    L = C.CDLL(build_port.build())
    L.port_chan_create.restype = C.c_void_p
    L.port_chan_create.argtypes = [C.c_char_p, C.c_longlong, C.c_longlong, C.c_int]
    L.port_chan_destroy.argtypes = [C.c_void_p]
    L.port_chan_process.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.port_dc_create.restype = C.c_void_p
    L.port_dc_block.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p]
    L.port_spectrum_create.restype = C.c_void_p
    L.port_spectrum_create.argtypes = [C.c_uint]
    L.port_spectrum_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.port_firdes_kaiser.argtypes = [C.c_uint, C.c_float, C.c_float, C.c_float, C.c_void_p]
    return L


def test_port_halfband_tables_match_reference(port, ref):
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "liquid_port.c")).read()
    for m in (3, 5, 10):
        hf = np.zeros(4 * m + 1, np.float32)
        ref.L.ref_firdespm_halfband_as(m, 65.0, _p(hf))
        body = re.search(r"HB%d\[%d\] = \{([^}]*)\}" % (m, m), src).group(1)
        vals = np.array([float(v.strip().rstrip("f")) for v in body.split(",")], np.float32)
        assert np.array_equal(vals, hf[1:2 * m:2])


def test_port_dc_blocker_bit_exact(port, ref):
    x = synth.wideband(2400000, 40000, FIVE, t0=0)
    y = RefIngress(ref).block(x)
    yp = np.zeros_like(x)
    port.port_dc_block(C.c_void_p(port.port_dc_create()), _p(x), x.size, _p(yp))
    assert np.array_equal(y.view(np.float32), yp.view(np.float32))


def test_port_chain_matches_reference(port, ref):
    fs, N = 2400000, 40000
    ing = RefIngress(ref)
    rcs = [RefChan(ref, m, fs, bw) for m, bw, _ in FIVE]
    pcs = [C.c_void_p(port.port_chan_create(m.encode(), fs, bw, 48000)) for m, bw, _ in FIVE]
    got = [[] for _ in FIVE]; want = [[] for _ in FIVE]
    for b in range(4):
        y = ing.block(synth.wideband(fs, N, FIVE, t0=b * N))
        for i, (m, bw, off) in enumerate(FIVE):
            r = rcs[i].process(y, off)
            iq = np.zeros(8192, np.complex64); au = np.zeros(4096, np.float32)
            niq, na, lvl, pk = C.c_uint(), C.c_uint(), C.c_double(), C.c_float()
            sk = port.port_chan_process(pcs[i], _p(y), N, off, _p(iq), C.byref(niq), _p(au), C.byref(na), C.byref(lvl), C.byref(pk))
            assert sk == 0 and niq.value == len(r["iq"]) and na.value == len(r["audio"]), m
            assert relrms(iq[:niq.value], r["iq"]) <= 1e-5
            assert abs(lvl.value - r["level_db"]) < 2e-3
            got[i].append(au[:na.value].copy()); want[i].append(r["audio"])
    for i, (m, _, _) in enumerate(FIVE):
        assert relrms(np.concatenate(got[i]), np.concatenate(want[i])) <= 1e-5, m


def test_port_spectrum_matches_reference(port, ref):
    fs, N, fft = 2400000, 40000, 1024
    ing = RefIngress(ref); rd = RefDistrib(ref, 2 * fft, 100000000); rs = RefSpectrum(ref, fft)
    ps = C.c_void_p(port.port_spectrum_create(fft))
    for b in range(2):
        for fr in rd.push(ing.block(synth.wideband(fs, N, FIVE, t0=b * N)), fs):
            r = rs.frame(fr)
            pts = np.zeros(fft, np.float32); line = np.zeros(fft, np.uint8); mags = np.zeros(2 * fft, np.float64)
            port.port_spectrum_frame(ps, _p(np.ascontiguousarray(fr)), _p(pts), _p(line), _p(mags))
            assert relrms(mags, r["fft_result"]) <= 1e-6
            assert relrms(pts, r["points"]) <= 1e-5
            assert (line.astype(int) != r["line"].astype(int)).sum() <= 2
