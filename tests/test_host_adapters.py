"""C++ host adapters (cubicsdr_b200/host/cubicsdr_b200.hpp: Modem plugin API, runSingleCH, SpectrumVisualProcessor
mirrors) compiled into a small program and compared with the oracle.  CPU tier links the test-only emulator; the GPU
tier links the product library."""
import os
import subprocess

import numpy as np
import pytest

from cubicsdr_b200 import synth
from oracle.refbind import RefChan, RefDistrib, RefIngress, RefSpectrum
from tests.parity import relrms, TOL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIVE = [("FM", 200000, 300000), ("NBFM", 12500, -150000), ("AM", 6000, 450000), ("USB", 5400, -600000), ("LSB", 5400, 750000)]


def _build(lib_path, out):
    libdir, libname = os.path.dirname(lib_path), os.path.basename(lib_path)[3:-3]
    subprocess.check_call(["g++", "-O1", "-std=c++11", os.path.join(ROOT, "tests", "cpp", "host_adapter_test.cpp"), "-o", out,
                           "-L", libdir, "-l" + libname, "-Wl,-rpath," + libdir])


def _run(lib, ref, tmp):
    fs, N = 2400000, synth.block_size(2400000)
    ing = RefIngress(ref)
    rcs = [RefChan(ref, m, fs, bw) for m, bw, _ in FIVE]
    sp = RefSpectrum(ref, 256); rd = RefDistrib(ref, 512, 100000000)
    want_audio = {m: [] for m, _, _ in FIVE}
    fm_iq, last_pts = [], None
    for b in range(3):
        x = synth.wideband(fs, N, FIVE, t0=b * N)
        y = ing.block(x)
        if b < 2:
            x.tofile(os.path.join(tmp, "wide_%d.bin" % b))
            for fr in rd.push(y, fs):
                last_pts = sp.frame(fr)["points"]
        for i, rc in enumerate(rcs):
            r = rc.process(y, FIVE[i][2])
            r["iq"].tofile(os.path.join(tmp, "iq_%s_%d.bin" % (FIVE[i][0], b)))
            want_audio[FIVE[i][0]].append(r["audio"])
            if i == 0 and b < 2:
                fm_iq.append(r["iq"])
    # channelizer mode: the same two wideband blocks through the reference's runPFBCH path (tests/parity.RefChannelizerPath)
    from tests.parity import CENTER, RefChannelizerPath
    pfb_path = RefChannelizerPath(ref, fs, [FIVE[0]])
    want_pfb = []
    for b in range(2):
        res, _ = pfb_path.block(synth.wideband(fs, N, FIVE, t0=b * N), CENTER, [CENTER + FIVE[0][2]])
        want_pfb.append(res[0]["audio"])
    pfb_path.close()
    exe = os.path.join(tmp, "host_adapter_test")
    _build(lib.path, exe)
    out = subprocess.run([exe, tmp], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0, out.stdout
    for m, _, _ in FIVE:
        got = np.fromfile(os.path.join(tmp, "audio_modem_%s.bin" % m), np.float32)
        want = np.concatenate(want_audio[m])
        assert len(got) == len(want), m
        assert relrms(got, want) <= TOL, (m, relrms(got, want))
    got = np.fromfile(os.path.join(tmp, "audio_chan_FM.bin"), np.float32)
    want = np.concatenate(want_audio["FM"][:2])
    assert len(got) == len(want) and relrms(got, want) <= TOL
    chz = np.fromfile(os.path.join(tmp, "audio_chan_FM_channelizer.bin"), np.float32)
    want_chz = np.concatenate(want_pfb)
    assert len(chz) == len(want_chz) and relrms(chz, want_chz) <= TOL, "channelizer mode through the C++ adapters"
    piped = np.fromfile(os.path.join(tmp, "audio_chan_FM_pipelined.bin"), np.float32)
    assert np.array_equal(piped, got), "pipelined ingress (runSingleCHPipelined) must return the same audio, one block late"
    got = np.fromfile(os.path.join(tmp, "postiq_chan_FM.bin"), np.complex64)
    assert len(got) == sum(len(q) for q in fm_iq) and relrms(got, np.concatenate(fm_iq)) <= TOL
    pts = np.fromfile(os.path.join(tmp, "spec_last_points.bin"), np.float32)
    assert len(pts) == 256 and relrms(pts, last_pts) <= TOL


def test_host_adapters_on_emulator(emu_lib, ref, tmp_path):
    _run(emu_lib, ref, str(tmp_path))


@pytest.mark.gpu
def test_host_adapters_on_gpu(gpu_lib, ref, tmp_path):
    _run(gpu_lib, ref, str(tmp_path))
