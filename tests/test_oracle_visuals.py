"""The reference's OWN, unmodified visual processors (src/process/FFTDataDistributor.cpp, SpectrumVisualProcessor.cpp,
ScopeVisualProcessor.cpp -- compiled from where they lie into oracle/_ref/libvisuals_ref.so over the reference's own liquid-dsp
binary, oracle/build_ref.py step 7) as the pin of the restated glue the parity tests compare against (oracle/ref_chain.cpp):
same frames in, the SAME BITS out -- frames released by the distributor, display points, ceiling and floor of the spectrum in
the full-band and in the zoomed-view branch (zoom, pan, back), waveform and audio spectrum of the scope."""
import os

import numpy as np
import pytest

from cubicsdr_b200 import synth
from oracle import refbind
from tests.parity import CENTER

FS = 2400000
CHANS = synth.channel_plan("mixed", FS, 8)


@pytest.fixture(scope="module")
def vis():
    if not os.path.exists(refbind.VISUALS_LIB) and not os.path.exists("/root/reference"):
        pytest.skip("oracle/_ref/libvisuals_ref.so not built and /root/reference absent")
    return refbind.RefVisuals()


def _blocks(n, block=None, fs=FS, seed=3):
    N = block or synth.block_size(fs)
    return [synth.wideband(fs, N, CHANS, seed=seed, t0=b * N) for b in range(n)]


@pytest.mark.parametrize("lps", [100000000, 30, 400, 1])
def test_distributor_frames_equal_the_reference_class(ref, vis, lps):
    for frame_len, block in ((1024, None), (2048, 40003), (32768, 4099), (600, 257)):
        a, b = refbind.RefDistrib(ref, frame_len, lps), refbind.SrcDistrib(vis, frame_len, lps)
        total = 0
        for k, x in enumerate(_blocks(10 if block != 257 else 40, block)):
            fa, fb = a.push(x, FS, CENTER), b.push(x, FS, CENTER)
            assert fa.shape == fb.shape, (lps, frame_len, k, fa.shape, fb.shape)
            assert np.array_equal(fa, fb)
            total += len(fa)
        assert total > 0 or lps <= 30
        a.close(); b.close()


def test_distributor_tuner_moves_rate_changes_frame_length_changes(ref, vis):
    """inp->frequency or inp->sampleRate changes: the buffered samples are dumped, the line accumulator is not
    (FFTDataDistributor.cpp:41-53); setFFTSize between blocks: only the frame length changes."""
    a, b = refbind.RefDistrib(ref, 1024, 200), refbind.SrcDistrib(vis, 1024, 200)
    plan = {2: dict(center=CENTER + 250000), 4: dict(rate=2000000), 5: dict(frame=4096), 7: dict(center=CENTER, frame=512), 8: dict(rate=FS)}
    center, rate, total = CENTER, FS, 0
    for k, x in enumerate(_blocks(11, 40003)):
        p = plan.get(k, {})
        center, rate = p.get("center", center), p.get("rate", rate)
        if "frame" in p:
            a.set_frame_len(p["frame"]); b.set_frame_len(p["frame"])
        fa, fb = a.push(x, rate, center), b.push(x, rate, center)
        assert fa.shape == fb.shape and np.array_equal(fa, fb), (k, fa.shape, fb.shape)
        total += len(fa)
    assert total > 20
    a.close(); b.close()


def _same_spectrum(ra, rb, what):
    assert rb is not None, what
    assert np.array_equal(ra["points"], rb["points"]), "%s: display points differ from the reference's class (max %g)" % (
        what, np.abs(ra["points"] - rb["points"]).max())
    assert ra["ceil_floor"][0] == rb["ceil_floor"][0] and ra["ceil_floor"][1] == rb["ceil_floor"][1], (what, ra["ceil_floor"], rb["ceil_floor"])


@pytest.mark.parametrize("fft_size,avg_rate", [(512, None), (1024, 0.2), (256, 1.0), (8192, None)])
def test_full_band_spectrum_equals_the_reference_class(ref, vis, fft_size, avg_rate):
    a, b = refbind.RefSpectrum(ref, fft_size), refbind.SrcSpectrum(vis, fft_size, FS)
    b.set_band(CENTER, FS)                                       # AppFrame.cpp:1396-1399
    if avg_rate is not None:
        a.set_average_rate(avg_rate); b.set_average_rate(avg_rate)
    d = refbind.RefDistrib(ref, 2 * fft_size, 100000000)
    n = 0
    for x in _blocks(3 if fft_size < 8192 else 2):
        for f in d.push(x, FS, CENTER):
            _same_spectrum(a.frame(f), b.frame(f, CENTER, FS), "frame %d" % n)
            n += 1
    assert n >= 4
    assert b.desired_input() == 2 * fft_size
    a.close(); b.close(); d.close()


def test_zoomed_view_equals_the_reference_class(ref, vis):
    """The view branch (SpectrumVisualProcessor.cpp:283-386, 454-492) through zoom in, pan, pan across the centre, zoom out, leaving
    the view and coming back -- the sequence tests/test_emu_parity.py and tests/test_gpu_parity.py run the product through."""
    fft_size, C0 = 512, CENTER
    views = {0: (True, C0 + 100000, 1000000), 2: (True, C0 + 100000, 400000), 3: (True, C0 + 160000, 400000),
             4: (True, C0 - 50000, 400000), 5: (True, C0 - 50000, 1100000), 7: (False, C0, FS), 8: (True, C0 - 50000, 1100000),
             10: (True, C0, 1500000)}
    a, b = refbind.RefSpectrum(ref, fft_size), refbind.SrcSpectrum(vis, fft_size, FS)
    d = refbind.RefDistrib(ref, 2 * fft_size, 400)
    view, n, n_view = None, 0, 0
    for k, x in enumerate(_blocks(12)):
        if k in views:
            view = views[k]
            a.set_view(*view); b.set_view(*view)
            want = a.desired_input(FS)
            if want != d.n:
                d.set_frame_len(want)                            # FFTVisualDataThread.cpp:55-61
        for f in d.push(x, FS, C0):
            ra = a.view_frame(f, C0, FS) if view[0] else a.frame(f)
            assert ra is not None
            _same_spectrum(ra, b.frame(f, C0, FS), "block %d frame %d view %r" % (k, n, view))
            assert b.desired_input() == d.n
            n += 1; n_view += 1 if view[0] else 0
    assert n >= 40 and n_view >= 30
    a.close(); b.close(); d.close()


def test_scope_equals_the_reference_class(ref, vis):
    """ScopeVisualProcessor: waveform points (one and two channels) and the audio spectrum with its ceiling / floor, bandwidth
    narrower than the audio rate included (outSize scaled by sampleRate / inputRate, .cpp:194-197).  (sampleRate > inputRate is
    left out: the reference then reads past the end of its averages, .cpp:203.)"""
    rng = np.random.default_rng(11)
    t = np.arange(1024)
    for channels, sample_rate, input_rate in ((1, 48000, 200000), (2, 48000, 200000), (1, 48000, 48000), (1, 5400, 48000)):
        a, b = refbind.RefScope(ref, 1024, 1024), refbind.SrcScope(vis, 1024)
        for k in range(12):
            n = (800, 1600, 1024, 300)[k % 4] * channels
            data = (0.6 * np.sin(2 * np.pi * (0.01 + 0.002 * k) * np.arange(n)) + 0.05 * rng.standard_normal(n)).astype(np.float32)
            if k == 5:
                data *= 3.0                                      # peaks above 1: the waveform is normalised by the peak
            wa, sa, fca = a.process(data, channels, sample_rate, input_rate)
            wb, sb, fcb, mode = b.process(data, channels, sample_rate, input_rate)
            assert wa.size == wb.size and np.array_equal(wa, wb), (channels, k, wa.size, wb.size)
            assert sa.size == sb.size and np.array_equal(sa, sb), (channels, k, sa.size, sb.size)
            assert fca[0] == fcb[0] and fca[1] == fcb[1]
            assert mode == 0                                     # SCOPE_MODE_Y for ordinary demodulator audio (type 0)
        a.close(); b.close()


def test_seeded_fuzz_of_the_spectrum_processor(ref, vis):
    """Eight seeded random sessions: display size, line rate, average rate, and a random walk of views (zoom levels, pans, leaving and
    re-entering the view) -- the restated distributor + processor against the reference's objects, frame by frame."""
    rng = np.random.default_rng(99)
    for session in range(8):
        fft_size = int(rng.choice([256, 512, 1024, 2048]))
        lps = int(rng.choice([100000000, 60, 300, 1000]))
        a, b = refbind.RefSpectrum(ref, fft_size), refbind.SrcSpectrum(vis, fft_size, FS)
        b.set_band(CENTER, FS)
        if rng.random() < 0.5:
            r = float(rng.choice([0.1, 0.35, 0.9]))
            a.set_average_rate(r); b.set_average_rate(r)
        da, db = refbind.RefDistrib(ref, 2 * fft_size, lps), refbind.SrcDistrib(vis, 2 * fft_size, lps)
        view, n = (False, CENTER, FS), 0
        for k, x in enumerate(_blocks(10, seed=100 + session)):
            if rng.random() < 0.45:
                if rng.random() < 0.25:
                    view = (False, CENTER, FS)
                else:
                    bw = int(rng.choice([1500000, 1100000, 700000, 400000, 250000, 120000]))
                    view = (True, CENTER + int(rng.integers(-400000, 400001)), bw)
                a.set_view(*view); b.set_view(*view)
                want = a.desired_input(FS)
                if want != da.n:
                    da.set_frame_len(want); db.set_frame_len(want)
            fa, fb = da.push(x, FS, CENTER), db.push(x, FS, CENTER)
            assert fa.shape == fb.shape and np.array_equal(fa, fb), (session, k)
            for f in fa:
                ra = a.view_frame(f, CENTER, FS) if view[0] else a.frame(f)
                assert ra is not None
                _same_spectrum(ra, b.frame(f, CENTER, FS), "session %d block %d frame %d view %r" % (session, k, n, view))
                n += 1
        assert n > 0
        a.close(); b.close(); da.close(); db.close()


def test_engine_spectrum_against_the_reference_objects_directly(emu_lib, ref, vis):
    """No restatement in between: the engine's spectrum (kernel emulator here) against the reference's own FFTDataDistributor and
    SpectrumVisualProcessor objects on the same DC-blocked blocks -- frame counts, display points, ceiling and floor; full band, then
    a zoomed view, then zoomed out by two octaves."""
    from cubicsdr_b200 import Engine, Spectrum
    from tests.parity import relrms, TOL
    fft_size = 512
    N = synth.block_size(FS)
    eng = Engine(emu_lib, max_block=N)
    sp = Spectrum(eng, fft_size, 0)
    ing = refbind.RefIngress(ref)                                # (the DC blocker; pinned to SDRPostThread in test_oracle_demod.py)
    d, p = refbind.SrcDistrib(vis, 2 * fft_size, 100000000), refbind.SrcSpectrum(vis, fft_size, FS)
    p.set_band(CENTER, FS)
    views = {3: (True, CENTER + 120000, 250000), 6: (True, CENTER + 120000, 1100000)}
    n = 0
    for b, x in enumerate(_blocks(9)):
        if b in views:
            sp.set_view(*views[b]); p.set_view(*views[b])
            d.set_frame_len(sp.desired_input(FS))                # FFTVisualDataThread.cpp:55-61
        eng.submit(x, CENTER, FS); eng.wait()
        frames = d.push(ing.block(x), FS, CENTER)
        assert sp.frames == len(frames), (b, sp.frames, len(frames))
        for f in range(len(frames)):
            r = p.frame(frames[f], CENTER, FS)
            assert r is not None
            pts, line, cf = sp.fetch(f)
            assert relrms(pts, r["points"]) <= TOL, (b, f, relrms(pts, r["points"]))
            assert np.abs(cf - r["ceil_floor"]).max() <= 1e-6 * np.abs(r["ceil_floor"]).max(), (b, f, cf, r["ceil_floor"])
            n += 1
        assert p.desired_input() == d.n
    assert n > 150
    sp.close(); eng.close(); d.close(); p.close(); ing.close()
