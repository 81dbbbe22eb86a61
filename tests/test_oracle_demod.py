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


def test_audio_mixer_equals_the_reference_callback(ref, demod_lib):
    """audioCallback (AudioThread.cpp:88-243) on real AudioThread objects against the restated mixer the device mixer is compared
    with: a seeded random sequence of pushes (one and two channels, empty blocks, a foreign sample rate, peaks above 1), gain
    changes, sources going inactive and coming back, and callbacks of several lengths that never line up with the blocks."""
    rng = np.random.default_rng(2024)
    a, b = refbind.RefMixer(ref, 48000), refbind.SrcMixer(demod_lib, 48000)
    gains = [0.9, 0.6, 1.3, 0.5, 1.0]
    for g in gains:
        assert a.add_source(g) == b.add_source(g)
    n_cb = n_nonzero = 0
    for step in range(1500):
        r = rng.random()
        i = int(rng.integers(len(gains)))
        if r < 0.40:                                             # a demodulator hands over a block
            channels = 2 if i == 3 else 1
            n = int(rng.choice([800, 808, 768, 1600, 0, 33])) * channels if rng.random() < 0.9 else 0
            data = (rng.standard_normal(n) * rng.choice([0.05, 0.3, 1.5])).astype(np.float32)
            rate = 44100 if rng.random() < 0.05 else 48000
            peak = float(np.abs(data).max()) if n else 0.0
            if rng.random() < 0.05:
                channels = 0                                     # AudioThreadInput with no channels: skipped by the callback
            a.push(i, data, channels, peak, rate); b.push(i, data, channels, peak, rate)
        elif r < 0.45:
            g = float(rng.choice([0.0, 0.25, 1.0, 2.0, 3.5, -1.0]))          # setGain clamps to [0, 2]
            a.set_gain(i, g); b.set_gain(i, g)
        elif r < 0.50:
            on = bool(rng.random() < 0.6)
            a.set_active(i, on); b.set_active(i, on)
        else:                                                    # the sound card asks for frames
            n_frames = int(rng.choice([512, 256, 1024, 100]))
            oa, ob = a.callback(n_frames), b.callback(n_frames)
            assert np.array_equal(oa, ob), "callback %d: restated mixer differs from the reference's callback (max %g)" % (
                n_cb, np.abs(oa - ob).max())
            n_cb += 1; n_nonzero += int(np.any(oa != 0))
    assert n_cb > 500 and n_nonzero > 300
    a.close(); b.close()


def test_wav_words_equal_the_reference_writer(ref, demod_lib, tmp_path):
    """AudioFileWAV::writeToFile of the reference (header + 16-bit payload, AudioFileWAV.cpp:122-165) against the restated payload the
    device's PCM words are compared with: one and two channels, peaks below and above 1 (the anti-clipping scale), odd lengths."""
    import struct
    rng = np.random.default_rng(5)
    for channels in (1, 2):
        w = refbind.SrcWav(demod_lib, tmp_path, "rec%d" % channels)
        want = []
        for k in range(9):
            n = int(rng.choice([800, 808, 1600, 33, 1])) * channels + (1 if (channels == 2 and k == 4) else 0)
            data = (rng.standard_normal(n) * rng.choice([0.05, 0.4, 1.7])).astype(np.float32)
            peak = float(np.abs(data).max())
            w.write(data, channels, 48000, peak)
            want.append(refbind.ref_wav_payload(ref, data, channels, peak))
        header, words = w.close()
        want = np.concatenate(want)
        assert words.size == want.size and np.array_equal(words, want)
        assert header[:4] == b"RIFF" and header[8:16] == b"WAVEfmt " and header[36:40] == b"data"
        fmt = struct.unpack("<IHHIIHH", header[16:36])
        assert fmt == (16, 1, channels, 48000, 48000 * 2 * channels, 2 * channels, 16)
        assert struct.unpack("<I", header[40:44])[0] == 2 * words.size and struct.unpack("<I", header[4:8])[0] == 36 + 2 * words.size


# ---- the ingress in front of the demodulators: the reference's own SDRPostThread ------------------------------------------------
def test_ingress_single_channel_equals_the_reference_post_thread(ref, demod_lib):
    """SDRPostThread::runSingleCH (src/sdr/SDRPostThread.cpp:248-299) feeding three real DemodulatorInstances: the DC-blocked block
    it publishes and every demodulator's audio against the restated ingress + chains, with the tuner moving and a retune on the way."""
    chans = [("FM", 200000, 300000), ("AM", 6000, -450000), ("USB", 5400, 600000)]
    N = synth.block_size(FS)
    app = refbind.SrcApp(demod_lib, 1)
    sds = [refbind.SrcDemod(demod_lib, m, FS, bw, CENTER + off, CENTER) for m, bw, off in chans]
    for d in sds:
        app.add(d)
    rcs, ing = [refbind.RefChan(ref, m, FS, bw) for m, bw, _ in chans], refbind.RefIngress(ref)
    freqs = [CENTER + off for _, _, off in chans]
    center = CENTER
    for b in range(10):
        if b == 3:
            center = CENTER + 100000                            # the tuner moves: every shift changes
        if b == 6:
            freqs[0] = CENTER - 350000; sds[0].set_frequency(freqs[0])
        x = synth.wideband(FS, N, chans + [("FM", 200000, -350000)], seed=5, t0=b * N) + np.complex64(0.01 - 0.02j)
        y, yr = app.block(x, center, FS, 1), ing.block(x)
        assert np.array_equal(y.view(np.float32), yr.view(np.float32)), "block %d: DC-blocked block differs" % b
        for d, rc, f in zip(sds, rcs, freqs):
            assert app.route(d)[0] == 0
            _same_block(rc.process(yr, f - center), d.collect(center), "block %d" % b)
    for d in sds:
        d.close()
    for rc in rcs:
        rc.close()
    app.close(); ing.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_ingress_channelizer_equals_the_reference_post_thread(ref, demod_lib, mode):
    """SDRPostThread::runPFBCH / runPFBCH2 + runDemodChannels (:303-511): channel centres, nearest-channel routing, channel 0 through
    the DC blocker, the wrap-around extra channel, the tuner moving and a demodulator hopping to another channel -- the restated
    routing (tests/parity.py: RefChannelizerPath, what the engine's channelizer mode is compared with) against the real thread
    feeding real DemodulatorInstances."""
    from tests.parity import RefChannelizerPath
    fs = 10000000
    chans = [("NBFM", 12500, 10000), ("AM", 6000, 1262000), ("USB", 5400, -2490000), ("NBFM", 12500, 4990000), ("FM", 200000, -3730000)]
    N = synth.block_size(fs)
    oracle = RefChannelizerPath(ref, fs, chans, mode=mode)
    M = oracle.M
    assert N % M == 0
    freqs = [CENTER + off for _, _, off in chans]
    centers = oracle.centers(CENTER)
    picks0 = [oracle.channel_at(centers, f) for f in freqs]
    assert 0 in picks0 and M in picks0                          # channel 0 (DC-blocked) and the extra channel are exercised
    app = refbind.SrcApp(demod_lib, mode)
    sds = [refbind.SrcDemod(demod_lib, m, oracle.demod_rate, bw, f, centers[ci]) for (m, bw, _), f, ci in zip(chans, freqs, picks0)]
    for d in sds:
        app.add(d)
    center, seen, routed = CENTER, set(), 0
    for b in range(8):
        if b == 3:
            center = CENTER + 40000                             # the tuner moves: every channel centre moves
        if b == 5:
            freqs[1] = CENTER + 40000 + 2262000; sds[1].set_frequency(freqs[1])    # hops two channels up
        x = synth.wideband(fs, N, chans + [("AM", 6000, 2302000)], seed=2024, t0=b * N) + np.complex64(0.02 + 0.01j)
        raw = app.block(x, center, fs, M)
        assert np.array_equal(raw.view(np.float32), x.view(np.float32))            # published as it came (getFullSampleRateIqData)
        res, picks = oracle.block(x, center, freqs)
        seen.update(picks)
        for i, (d, r) in enumerate(zip(sds, res)):
            want_center = oracle.centers(center)[picks[i]]
            ch, cc, bw = app.route(d)
            if ch >= 0:             # (-1 right after a refresh: updateActiveDemodulators ran again behind the block, :195-198)
                assert ch == picks[i] and cc == want_center and bw == oracle.chan_bw, (b, i, ch, picks[i])
                routed += 1
            _same_block(r, d.collect(want_center), "mode %d block %d demodulator %d" % (mode, b, i))
    assert len(seen) >= 5 and routed >= 20
    for d in sds:
        d.close()
    oracle.close(); app.close()


def test_seeded_fuzz_of_the_demodulator_chain(ref, demod_lib):
    """Thirty seeded random demodulators (modem, bandwidth, offset, block length, retunes, squelch) through the reference's own
    DemodulatorInstance and through the restated chain: the same bits every time."""
    rng = np.random.default_rng(20260923)
    ranges = {"FM": (60000, 400000), "NBFM": (4000, 40000), "AM": (2000, 20000), "USB": (1500, 12000), "LSB": (1500, 12000),
              "DSB": (1500, 12000), "CW": (500, 2000), "I/Q": (8000, 100000), "FMS": (150000, 300000)}
    on_air = {"FM": "FM", "NBFM": "NBFM", "AM": "AM", "USB": "USB", "LSB": "LSB", "DSB": "AM", "CW": "AM", "I/Q": "AM", "FMS": "FMS"}
    names = list(ranges)
    for case in range(30):
        modem = names[case % len(names)]
        lo, hi = ranges[modem]
        bw = int(rng.integers(lo, hi + 1))
        off = int(rng.integers(-1100000, 1100001))
        block = int(rng.choice([synth.block_size(FS), 40003, 20000, 4099, 65536]))
        n_blocks = 5 if block > 10000 else 12
        retune = {int(b): int(rng.integers(-1150000, 1150001)) for b in rng.choice(np.arange(1, n_blocks), size=2, replace=False)} if rng.random() < 0.6 else None
        if retune and rng.random() < 0.3:
            retune[min(retune)] = int(rng.choice([0, 1900000, -1850000]))          # zero shift / beyond 0.75 fs
        squelch = float(rng.uniform(-60, 0)) if rng.random() < 0.4 else None
        signal = [(on_air[modem], 6000 if on_air[modem] == "AM" else min(bw, 250000), off)]
        _run(ref, demod_lib, modem, bw, off, on_air[modem], n_blocks, block=block, squelch=squelch, retune=retune, signal=signal)


def test_engine_against_the_reference_application_directly(emu_lib, ref, demod_lib):
    """No restatement in between: the same wideband blocks into sdr_engine_submit (kernel emulator here) and into the reference's
    own SDRPostThread feeding its own DemodulatorInstances; the DC-blocked block bit for bit, every demodulator's audio within the
    1e-5 relative RMS the north star grants, identical sample counts, peak and level trackers."""
    from cubicsdr_b200 import Channel, Engine
    from tests.parity import relrms, TOL
    chans = [("FM", 200000, 300000), ("NBFM", 12500, -150000), ("AM", 6000, 450000), ("USB", 5401, -600000), ("LSB", 5400, 750000),
             ("CW", 700, -250000), ("I/Q", 48000, 900000)]
    on_air = [("FM", 200000, 300000), ("NBFM", 12500, -150000), ("AM", 6000, 450000), ("USB", 5400, -600000), ("LSB", 5400, 750000),
              ("AM", 6000, -250000), ("AM", 6000, 900000)]
    N = synth.block_size(FS)
    eng = Engine(emu_lib, max_block=N)
    cs = [Channel(eng, m, FS, bw, 48000, frequency=CENTER + off) for m, bw, off in chans]
    app = refbind.SrcApp(demod_lib, 1)
    sds = [refbind.SrcDemod(demod_lib, m, FS, bw, CENTER + off, CENTER) for m, bw, off in chans]
    for d in sds:
        app.add(d)
    got, want = [[] for _ in chans], [[] for _ in chans]
    for b in range(8):
        x = synth.wideband(FS, N, on_air, seed=11, t0=b * N) + np.complex64(0.01 - 0.02j)
        eng.submit(x, CENTER, FS); eng.wait()
        y = app.block(x, CENTER, FS, 1)
        assert np.array_equal(eng.fetch_block(N).view(np.float32), y.view(np.float32)), "block %d: DC-blocked block" % b
        for i, (c, d) in enumerate(zip(cs, sds)):
            a, lvl, pk, sk = c.fetch_audio()
            s = d.collect(CENTER)
            assert not sk and not s["skipped"] and len(a) == s["audio"].size, (b, chans[i], len(a), s["audio"].size)
            assert abs(pk - s["peak"]) <= 1e-4 * max(s["peak"], 1e-3)
            m = c.fetch_meter()
            assert np.allclose(m[:3], (s["signal_level"], s["signal_floor"], s["signal_ceil"]), rtol=0, atol=2e-4), (b, chans[i], m, s)
            got[i].append(a); want[i].append(s["audio"])
    for i, ch in enumerate(chans):
        e = relrms(np.concatenate(got[i]), np.concatenate(want[i]))
        assert e <= TOL, "%r: audio relative RMS error %.3g against the reference application" % (ch, e)
    for c in cs:
        c.close()
    for d in sds:
        d.close()
    app.close(); eng.close()
