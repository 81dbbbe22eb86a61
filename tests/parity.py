"""Parity harness shared by the emulator tier (no GPU) and the GPU tier.

Both tiers drive the engine through the C-ABI binding (cubicsdr_b200._capi) and compare with the oracle
(oracle/refbind.py: the reference's own liquid-dsp binary + CubicSDR glue) on identical seeded inputs.
"""
import numpy as np

from cubicsdr_b200 import synth
from cubicsdr_b200._capi import Channel, Engine, Spectrum
from oracle.refbind import RefChan, RefDistrib, RefIngress, RefSpectrum

# north_star tolerance: 1e-5 relative RMS on audio and spectrum magnitudes
TOL = 1e-5
CENTER = 100_000_000


def relrms(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    dt = np.complex128 if (a.dtype.kind == "c" or b.dtype.kind == "c") else np.float64
    a = a.astype(dt)
    b = b.astype(dt)
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / nb) if nb > 0 else float(np.linalg.norm(a))


def run_channels(lib, ref, fs, chans, n_blocks, block=None, seed=1234, device=0, want_iq=True, retune=None, signal=None,
                 resident=None, spectrum=0, audio_rate=48000, spec_out=None, squelch=None):
    """Stream n_blocks through engine and oracle.  chans: [(modem, bw, offset_hz)].
    Returns per-channel dict with concatenated audio / iq of both sides and per-block count equality."""
    N = block or synth.block_size(fs)
    eng = Engine(lib, device=device, max_block=N)
    cs = [Channel(eng, m, fs, bw, audio_rate, frequency=CENTER + off, want_iq=want_iq) for m, bw, off in chans]
    # `resident`: further channels (and a spectrum) attached to the same engine but not compared -- they put the engine in
    # the configuration the full-size workload runs in (grid sizes, DC replay placement, stream overlap)
    extra = [Channel(eng, m, fs, bw, 48000, frequency=CENTER + off) for m, bw, off in (resident or [])]
    sp = Spectrum(eng, spectrum, 0) if spectrum else None
    # spec_out (a dict): also compare the attached spectrum, frame by frame, with the reference's (same streams overlapping
    # as in the full workload); filled with the worst relative errors
    rsp = RefSpectrum(ref, spectrum) if (spectrum and spec_out is not None) else None
    rdist = RefDistrib(ref, 2 * spectrum, 100000000) if rsp else None
    if rsp:
        spec_out.update(mags=0.0, points=0.0, frames=0, u8_bad_far=0)
    rcs = [RefChan(ref, m, fs, bw, audio_rate) for m, bw, off in chans]
    ing = RefIngress(ref)
    out = [dict(modem=m, audio=[], ref_audio=[], iq=[], ref_iq=[], counts_ok=True, lvl=[], ref_lvl=[], peak=[], ref_peak=[],
                skipped=[], ref_skipped=[], meter=[], ref_meter=[], squelched=[], ref_squelched=[]) for m, _, _ in chans]
    # squelch: {channel index: level in dB} -- DemodulatorThread::setSquelchLevel on both sides before the first block
    for i, lvl_db in (squelch or {}).items():
        cs[i].set_squelch(True, lvl_db); rcs[i].set_squelch(True, lvl_db)
    dc_err = []
    offs = [off for _, _, off in chans]
    for b in range(n_blocks):
        if retune and b in retune:
            for i, new_off in retune[b].items():
                offs[i] = new_off
                cs[i].set_frequency(CENTER + new_off)
        x = synth.wideband(fs, N, signal if signal is not None else chans, seed=seed, t0=b * N)
        eng.submit(x, CENTER, fs)
        eng.wait()
        y = ing.block(x)
        dc_err.append(relrms(eng.fetch_block(N), y))
        if rsp:
            frames = rdist.push(y, fs, CENTER)
            assert sp.frames == len(frames), "frame count %d vs reference %d in block %d" % (sp.frames, len(frames), b)
            for f in range(sp.frames):
                r = rsp.frame(frames[f])
                pts, line, cf = sp.fetch(f)
                mags, _ = sp.fetch_debug(f)
                spec_out["mags"] = max(spec_out["mags"], relrms(mags, r["fft_result"]))
                spec_out["points"] = max(spec_out["points"], relrms(pts, r["points"]))
                bad = line.astype(np.int32) != r["line"].astype(np.int32)
                v = np.clip(r["points"].astype(np.float64), 0, 0.99) * 255.0
                spec_out["u8_bad_far"] += int((bad & ~(np.abs(v - np.round(v)) < 255.0 * 1e-5)).sum())
                spec_out["frames"] += 1
        for i, (c, rc) in enumerate(zip(cs, rcs)):
            r = rc.process(y, offs[i])
            a, lvl, pk, sk = c.fetch_audio()
            o = out[i]
            o["skipped"].append(sk); o["ref_skipped"].append(r["skipped"])
            if len(a) != len(r["audio"]):
                o["counts_ok"] = False
            o["audio"].append(a); o["ref_audio"].append(r["audio"])
            if want_iq and not sk:
                q = c.fetch_iq()
                if len(q) != len(r["iq"]):
                    o["counts_ok"] = False
                o["iq"].append(q); o["ref_iq"].append(r["iq"])
            o["lvl"].append(lvl); o["ref_lvl"].append(r["level_db"]); o["peak"].append(pk); o["ref_peak"].append(r["peak"])
            m = c.fetch_meter()
            o["meter"].append(m[:3]); o["ref_meter"].append((r["signal_level"], r["signal_floor"], r["signal_ceil"]))
            o["squelched"].append(m[3]); o["ref_squelched"].append(r["squelched"])
    for o in out:
        for k in ("audio", "ref_audio", "iq", "ref_iq"):
            o[k] = np.concatenate(o[k]) if len(o[k]) else np.zeros(0)
    for c in cs + extra:
        c.close()
    if sp is not None:
        sp.close()
    if rsp:
        rsp.close(); rdist.close()
    for rc in rcs:
        rc.close()
    ing.close()
    launches = eng.launch_count
    eng.close()
    return out, dc_err, launches


def check_channels(out, tol=TOL):
    for o in out:
        assert o["counts_ok"], "%s: per-block sample counts differ from the reference" % o["modem"]
        assert o["skipped"] == o["ref_skipped"], "%s: skip rule differs" % o["modem"]
        if len(o["ref_audio"]):
            e = relrms(o["audio"], o["ref_audio"])
            assert e <= tol, "%s: audio relative RMS error %.3g > %.1g" % (o["modem"], e, tol)
        if len(o["ref_iq"]):
            e = relrms(o["iq"], o["ref_iq"])
            assert e <= tol, "%s: decimated IQ relative RMS error %.3g > %.1g" % (o["modem"], e, tol)
        assert np.allclose(o["lvl"], o["ref_lvl"], atol=2e-4), "%s: level metering differs" % o["modem"]
        assert np.allclose(o["peak"], o["ref_peak"], rtol=1e-4, atol=1e-6), "%s: peak differs" % o["modem"]
        # DemodulatorThread's signalLevel / signalFloor / signalCeil trackers (float members) and the squelch decision
        assert np.allclose(o["meter"], o["ref_meter"], rtol=0, atol=2e-4), "%s: squelch trackers differ: %r vs %r" % (
            o["modem"], o["meter"][-1], o["ref_meter"][-1])
        # (a block the 0.75 fs rule drops never reaches DemodulatorThread: there is no AudioThreadInput to carry a flag)
        live = [b for b, sk in enumerate(o["ref_skipped"]) if not sk]
        assert [o["squelched"][b] for b in live] == [o["ref_squelched"][b] for b in live], "%s: squelch decision differs" % o["modem"]


def run_spectrum(lib, ref, fs, fft_size, n_blocks, chans, lps=0, block=None, seed=1234, device=0, avg_rate=None,
                 center_by_block=None, view_by_block=None):
    """Returns worst-case errors over all frames: magnitudes, maa, points, ceil/floor, u8 mismatch fraction.
    center_by_block: {block index: tuner frequency} -- the tuner moves (FFTDataDistributor dumps its buffer, :41-53).
    view_by_block: {block index: (is_view, centre frequency, bandwidth)} -- SpectrumVisualProcessor::setView before that block
    (zoomed view: the distributor then hands out frames of getDesiredInputSize samples, FFTVisualDataThread.cpp:55-61)."""
    N = block or synth.block_size(fs)
    eng = Engine(lib, device=device, max_block=N)
    sp = Spectrum(eng, fft_size, lps)
    ing = RefIngress(ref)
    rsp = RefSpectrum(ref, fft_size)
    rd = RefDistrib(ref, 2 * fft_size, lps if lps else 100000000)
    if avg_rate is not None:                                    # SpectrumVisualProcessor::setFFTAverageRate (:76-78)
        sp.set_average_rate(avg_rate); rsp.set_average_rate(avg_rate)
    worst = dict(mags=0.0, maa=0.0, points=0.0, cf=0.0)
    n_frames = 0
    u8_bad = 0
    u8_bad_far = 0
    center = CENTER
    view = None
    for b in range(n_blocks):
        x = synth.wideband(fs, N, chans, seed=seed, t0=b * N)
        center = (center_by_block or {}).get(b, center) if b else (center_by_block or {}).get(0, CENTER)
        if view_by_block and b in view_by_block:
            view = view_by_block[b]
            sp.set_view(*view); rsp.set_view(*view)
            # the processor's desired input size becomes the distributor's frame length (FFTVisualDataThread.cpp:55-61);
            # what the distributor had buffered stays (its carried remainder), its line accumulator too
            want = rsp.desired_input(fs)
            assert sp.desired_input(fs) == want
            if want != rd.n:
                rd.set_frame_len(want)
        eng.submit(x, center, fs)
        eng.wait()
        frames = rd.push(ing.block(x), fs, center)
        assert sp.frames == len(frames), "frame count %d vs reference %d in block %d" % (sp.frames, len(frames), b)
        for f in range(sp.frames):
            r = rsp.view_frame(frames[f], center, fs) if view and view[0] else rsp.frame(frames[f])
            assert r is not None
            pts, line, cf = sp.fetch(f)
            mags, maa = sp.fetch_debug(f)
            worst["mags"] = max(worst["mags"], relrms(mags, r["fft_result"]))
            worst["maa"] = max(worst["maa"], relrms(maa, r["maa"]))
            worst["points"] = max(worst["points"], relrms(pts, r["points"]))
            worst["cf"] = max(worst["cf"], float(np.abs(cf - r["ceil_floor"]).max() / np.abs(r["ceil_floor"]).max()))
            bad = line.astype(np.int32) != r["line"].astype(np.int32)
            u8_bad += int(bad.sum())
            # a mismatch only counts as "far" if the reference's float is not next to a quantisation edge
            v = np.clip(r["points"].astype(np.float64), 0, 0.99) * 255.0
            near_edge = np.abs(v - np.round(v)) < 255.0 * 1e-5
            u8_bad_far += int((bad & ~near_edge).sum())
            n_frames += 1
    sp.close(); rsp.close(); rd.close(); ing.close(); eng.close()
    worst["frames"] = n_frames
    worst["u8_bad"] = u8_bad
    worst["u8_bad_far"] = u8_bad_far
    worst["u8_total"] = n_frames * fft_size
    return worst


def check_add_remove_mid_stream(lib, ref, fs=2400000, n_blocks=6, device=0):
    """Demodulators come and go while the stream runs (DemodulatorMgr::newThread / deleteThread between blocks): a channel
    attached after block 1 starts from rest with that block's successor, a channel removed after block 3 frees its slot,
    which the next newcomer reuses; everybody else's state must be untouched.  Every channel is compared with its own
    oracle instance over exactly the blocks it lived through."""
    plan = {  # name: (modem, bw, offset, first block, last block exclusive)
        "a": ("FM", 200000, 300000, 0, n_blocks),
        "b": ("AM", 6000, -450000, 0, 4),
        "c": ("NBFM", 12500, 150000, 2, n_blocks),
        "d": ("USB", 5400, -600000, 4, n_blocks),       # takes over b's slot
    }
    signal = [(m, bw, off) for m, bw, off, _, _ in plan.values()]
    N = synth.block_size(fs)
    eng = Engine(lib, device=device, max_block=N)
    ing = RefIngress(ref)
    live, refs = {}, {}
    got = {k: [] for k in plan}
    want = {k: [] for k in plan}
    for b in range(n_blocks):
        for k, (m, bw, off, b0, b1) in plan.items():
            if b == b1 and k in live:
                live.pop(k).close(); refs.pop(k).close()
        for k, (m, bw, off, b0, b1) in plan.items():
            if b == b0:
                live[k] = Channel(eng, m, fs, bw, 48000, frequency=CENTER + off)
                refs[k] = RefChan(ref, m, fs, bw)
        x = synth.wideband(fs, N, signal, seed=4321, t0=b * N)
        eng.submit(x, CENTER, fs)
        eng.wait()
        y = ing.block(x)
        for k in live:
            a, lvl, pk, sk = live[k].fetch_audio()
            r = refs[k].process(y, plan[k][2])
            assert len(a) == len(r["audio"]), "%s block %d: audio count %d vs %d" % (k, b, len(a), len(r["audio"]))
            got[k].append(a); want[k].append(r["audio"])
    for k in list(live):
        live.pop(k).close(); refs.pop(k).close()
    ing.close(); eng.close()
    for k in plan:
        e = relrms(np.concatenate(got[k]), np.concatenate(want[k]))
        assert e <= TOL, "channel %s (%s): audio relative RMS error %.3g" % (k, plan[k][0], e)


def check_pipelined_fetch(lib, fs=2400000, n_blocks=6, fft_size=1024, device=0):
    """submit(b+lag); wait_block(lag); fetch must return exactly block b's results (ring of host result buffers)."""
    chans = [("FM", 200000, 300000), ("NBFM", 12500, -412500), ("AM", 6000, 150000), ("USB", 5400, -250000)]
    N = synth.block_size(fs)
    blocks = [synth.wideband(fs, N, chans, seed=77, t0=b * N) for b in range(n_blocks)]

    def make():
        eng = Engine(lib, device=device, max_block=N)
        cs = [Channel(eng, m, fs, bw, 48000, frequency=CENTER + off) for m, bw, off in chans]
        return eng, cs, Spectrum(eng, fft_size, 0)

    def grab(cs, sp):
        return [c.fetch_audio() for c in cs], [sp.fetch(f) for f in range(sp.frames)]

    eng, cs, sp = make()
    want = []
    for x in blocks:
        eng.submit(x, CENTER, fs)
        eng.wait()
        want.append(grab(cs, sp))
    eng.close()
    for lag in (1, 2):
        eng, cs, sp = make()
        got = []
        for b in range(n_blocks):
            eng.submit(blocks[b], CENTER, fs)
            if b >= lag:
                eng.wait_block(lag)
                got.append(grab(cs, sp))
        for k in range(lag - 1, -1, -1):                        # drain
            eng.wait_block(k)
            got.append(grab(cs, sp))
        eng.close()
        _same_results(want, got, n_blocks)


def _same_results(want, got, n_blocks):
    for b in range(n_blocks):
        (wa, wf), (ga, gf) = want[b], got[b]
        for (a0, l0, p0, s0), (a1, l1, p1, s1) in zip(wa, ga):
            assert np.array_equal(a0, a1) and l0 == l1 and p0 == p1 and s0 == s1, "block %d: pipelined audio differs" % b
        assert len(wf) == len(gf), "block %d: frame count differs" % b
        for (p0, n0, c0), (p1, n1, c1) in zip(wf, gf):
            assert np.array_equal(p0, p1) and np.array_equal(n0, n1) and np.array_equal(c0, c1), "block %d: pipelined spectrum differs" % b


def check_against_golden(lib, device=0):
    """Engine output against the committed golden vectors (tests/golden/chain_2p4msps.npz, produced by the reference binary)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chain_2p4msps.npz"))
    five = [("FM", 200000, 300000), ("NBFM", 12500, -150000), ("AM", 6000, 450000), ("USB", 5400, -600000), ("LSB", 5400, 750000)]
    fs, N = 2400000, synth.block_size(2400000)
    eng = Engine(lib, device=device, max_block=N)
    cs = [Channel(eng, m, fs, bw, 48000, frequency=CENTER + off) for m, bw, off in five]
    sp = Spectrum(eng, 256, 0)
    worst = 0.0
    for b in range(2):
        eng.submit(synth.wideband(fs, N, five, seed=1234, t0=b * N), CENTER, fs)
        eng.wait()
        for i, c in enumerate(cs):
            a = c.fetch_audio()[0]
            ga = g["audio_%s_b%d" % (five[i][0], b)]
            assert len(a) == len(ga), five[i][0]
            worst = max(worst, relrms(a, ga))
        for f in (0, sp.frames - 1):
            pts, line, _ = sp.fetch(f)
            assert relrms(pts, g["spec_points_b%d_f%d" % (b, f)]) <= TOL
    eng.close()
    assert worst <= TOL, worst
    return worst


class RefChannelizerPath:
    """Oracle for channelizer mode: the reference binary's firpfbch_crcf + a restatement of the routing around it.
      SDRPostThread::runPFBCH            src/sdr/SDRPostThread.cpp:416-455  (no wideband DC blocker; one analyzer_execute per frame)
      SDRPostThread::updateChannels      :115-124   channel centre frequencies (integer arithmetic as there)
      SDRPostThread::getChannelAt        :127-139   nearest centre, first minimum wins
      SDRPostThread::runDemodChannels    :303-398   de-interleave, channel 0 through the DC blocker, extra channel = numChannels/2
    Every demodulator then runs DemodulatorPreThread + Modem at the channel rate (RefChan with sample_rate = chanBw)."""

    def __init__(self, ref, fs, chans, audio_rate=48000, mode=1):
        from oracle.refbind import RefPfb, RefPfb2
        self.ref, self.fs, self.mode = ref, fs, mode
        ref.L.ref_optimal_channel_count.argtypes = [__import__("ctypes").c_longlong]
        self.M = ref.L.ref_optimal_channel_count(fs)
        self.chan_bw = fs // self.M
        # runPFBCH2 (:472-511): firpfbch2, and runDemodChannels(chanBw * 2) -- the demodulators see twice the channel rate
        self.pfb = RefPfb2(ref, self.M) if mode == 2 else RefPfb(ref, self.M)
        self.demod_rate = self.chan_bw * (2 if mode == 2 else 1)
        self.dc = RefIngress(ref)                               # dcFilter: the same iirfilt DC blocker object type (:29)
        self.demods = [RefChan(ref, m, self.demod_rate, bw, audio_rate) for m, bw, _ in chans]

    def centers(self, frequency):
        M, c = self.M, [0] * (self.M + 1)
        for i in range(M // 2):
            ofs = int(self.chan_bw * i)
            c[i] = frequency + ofs
            c[i + M // 2] = frequency - (self.fs // 2) + ofs
        c[M] = frequency + (self.fs // 2)
        return c

    def channel_at(self, centers, f):
        chan, min_delta = -1, self.fs
        for i in range(self.M + 1):
            d = abs(f - centers[i])
            if d < min_delta:
                min_delta, chan = d, i
        return chan

    def block(self, x, frequency, demod_freqs):
        """-> list of RefChan.process results (or None when a demodulator got no data), and the chosen channel per demod."""
        y = self.pfb.execute(x)                                 # [frames, M]
        centers = self.centers(frequency)
        picks = [self.channel_at(centers, f) for f in demod_freqs]
        ch0 = self.dc.block(np.ascontiguousarray(y[:, 0])) if 0 in picks else None
        out = []
        for d, f, ci in zip(self.demods, demod_freqs, picks):
            if ci < 0:
                out.append(None); continue
            data = ch0 if ci == 0 else np.ascontiguousarray(y[:, self.M // 2 if ci == self.M else ci])
            out.append(d.process(data, f - centers[ci]))
        return out, picks

    def close(self):
        self.pfb.close(); self.dc.close()
        for d in self.demods:
            d.close()


def check_channelizer_mode(lib, ref, fs, chans, n_blocks, retune=None, device=0, seed=2024, mode=1, center_by_block=None):
    """Engine in channelizer mode against RefChannelizerPath.  chans: [(modem, bw, offset)].  Returns the set of filterbank
    channels that were exercised."""
    N = synth.block_size(fs)
    eng = Engine(lib, device=device, max_block=N)
    eng.set_channelizer(mode)
    M, chan_rate = eng.channel_plan(fs)
    oracle = RefChannelizerPath(ref, fs, chans, mode=mode)
    assert (M, chan_rate) == (oracle.M, oracle.chan_bw) and N % M == 0
    chan_rate = oracle.demod_rate
    eng.set_channel_block_max((2 if mode == 2 else 1) * (N // M))
    cs = [Channel(eng, m, chan_rate, bw, 48000, frequency=CENTER + off) for m, bw, off in chans]
    freqs = [CENTER + off for _, _, off in chans]
    got = [[] for _ in chans]; want = [[] for _ in chans]
    used = set()
    signal = list(chans)
    for b in sorted(retune or {}):
        signal += [(chans[i][0], chans[i][1], off) for i, off in retune[b].items()]
    center = CENTER
    for b in range(n_blocks):
        if retune and b in retune:
            for i, off in retune[b].items():
                freqs[i] = CENTER + off
                cs[i].set_frequency(freqs[i])
        x = synth.wideband(fs, N, signal, seed=seed, t0=b * N) + np.complex64(0.02 + 0.01j)     # a DC offset: channel 0's blocker matters
        center = (center_by_block or {}).get(b, center) if b else (center_by_block or {}).get(0, CENTER)   # the tuner moves: every centre moves
        eng.submit(x, center, fs)
        eng.wait()
        res, picks = oracle.block(x, center, freqs)
        used.update(picks)
        for i, (c, r) in enumerate(zip(cs, res)):
            a, lvl, pk, sk = c.fetch_audio()
            if r is None or r["skipped"]:
                assert sk or len(a) == 0, "demod %d block %d: engine produced audio where the reference had none" % (i, b)
                continue
            assert not sk, "demod %d block %d: engine skipped a block the reference processed" % (i, b)
            assert len(a) == len(r["audio"]), "demod %d block %d: %d audio samples vs %d" % (i, b, len(a), len(r["audio"]))
            got[i].append(a); want[i].append(r["audio"])
    for c in cs:
        c.close()
    oracle.close(); eng.close()
    for i, (m, bw, off) in enumerate(chans):
        e = relrms(np.concatenate(got[i]), np.concatenate(want[i]))
        assert e <= TOL, "demod %d (%s @ %+d Hz): audio relative RMS error %.3g" % (i, m, off, e)
    return used


def dc_window_case(lib, ref, fs, blocks, make_block, modes=(1,), max_block=None):
    """Stream the blocks through the engine's DC blocker with the given replay mode per block and compare every sample, bit for
    bit, with the reference's iirfilt_crcf_execute_block."""
    from cubicsdr_b200 import Engine
    from oracle.refbind import RefIngress
    eng = Engine(lib, max_block=max_block or max(blocks))
    ing = RefIngress(ref)
    t0 = 0
    for b, n in enumerate(blocks):
        eng.set_dc_window(modes[b % len(modes)])
        x = make_block(t0, n).astype(np.complex64)
        t0 += n
        eng.submit(x, CENTER, fs); eng.wait()
        got, want = eng.fetch_block(n).view(np.float32), ing.block(x).view(np.float32)
        assert np.array_equal(got, want), "block %d: %d samples differ" % (b, int(np.count_nonzero(got != want)))
    stats = eng.dc_window_stats()
    eng.close()
    return stats


def check_dc_window(lib, ref, big=False):
    """The candidate-window replay (dsp_dcwin.h) on the inputs that stress it: a DC offset (every boundary enumerable), zero-mean
    noise and a carrier next to 0 Hz (the state keeps crossing zero: runs of boundaries without a window), silence, ragged block
    lengths (partial last tile, blocks below the window path's minimum), and switching between the two replays block by block."""
    rng = np.random.default_rng(7)

    def noise_dc(t0, n):
        return 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n)) + (0.01 - 0.02j)

    def noise_zero_mean(t0, n):
        return 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))

    def carrier_near_dc(t0, n):
        t = np.arange(t0, t0 + n)
        return 0.02 * np.exp(1j * 0.0056 * t) + 0.3 * np.exp(1j * 0.31 * t) + 0.007 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))

    def silence_then_burst(t0, n):
        x = np.zeros(n, np.complex64)
        x[n // 3: n // 3 + 5000] = 0.4 * (rng.standard_normal(min(5000, n - n // 3)) + 0.3)
        return x

    s, blocks = dc_window_case(lib, ref, 2400000, [40000] * 4, noise_dc)
    assert blocks == 4 and s == 0
    s, blocks = dc_window_case(lib, ref, 2400000, [40000] * 5, noise_zero_mean)
    assert blocks == 5
    s, blocks = dc_window_case(lib, ref, 2400000, [40000] * 5, carrier_near_dc)
    assert blocks == 5
    dc_window_case(lib, ref, 2400000, [30000] * 3, silence_then_burst)
    # ragged lengths: partial last tile, exactly one tile more, below the minimum (serial replay), then long again
    s, blocks = dc_window_case(lib, ref, 2400000, [40000, 3073, 33333, 1500, 1024 * 7, 2049, 40000], noise_zero_mean, max_block=40000)
    assert blocks == 6                  # all but the 1500-sample block (two tiles: serial replay)
    # the two replays alternate on one stream (the warm-up replay reads the raw history the window replay does not need)
    s, blocks = dc_window_case(lib, ref, 20000000, [333360] * 4, carrier_near_dc, modes=(1, 0, 0, 1))
    assert blocks == 2
    if big:
        # full-size blocks (61.44 MSps: 1001 tiles, 32 chunks), zero-mean and with an offset
        s, blocks = dc_window_case(lib, ref, 61440000, [1024068] * 3, noise_zero_mean)
        assert blocks == 3
        s, blocks = dc_window_case(lib, ref, 61440000, [1024068] * 2, noise_dc)
        assert blocks == 2 and s == 0


def fm_stereo_floor(ref, fs, chan, signal, n_blocks, eps=1e-7, seed=1234):
    """How far the reference's own FM stereo output moves when its input moves by `eps` (relative, white): the loop's oscillator is
    a 1024-entry table (nco_crcf ignores LIQUID_VCO), so a phase that moves by 1e-7 rad now and then lands in the neighbouring
    table entry, the phase detector sees 6e-3 rad more or less, and the loop (bandwidth 0.25) carries the kick on.  Differences of
    1e-7 at the loop's input therefore surface as ~1e-3 in the difference channel -- for ANY second implementation, the reference
    binary on another libm included.  Returns the relative RMS change of (sum, difference)."""
    m, bw, off = chan
    N = synth.block_size(fs)
    rng = np.random.default_rng(99)
    a, b = RefChan(ref, m, fs, bw), RefChan(ref, m, fs, bw)
    ing = RefIngress(ref)
    A, B = [], []
    for k in range(n_blocks):
        y = ing.block(synth.wideband(fs, N, signal, seed=seed, t0=k * N))
        y2 = (y * (1.0 + eps * rng.standard_normal(N))).astype(np.complex64)
        A.append(a.process(y, off)["audio"]); B.append(b.process(y2, off)["audio"])
    a.close(); b.close(); ing.close()
    A, B = np.concatenate(A).reshape(-1, 2), np.concatenate(B).reshape(-1, 2)
    return relrms(B.sum(1), A.sum(1)), relrms(B[:, 1] - B[:, 0], A[:, 1] - A[:, 0])


def check_fm_stereo(lib, ref, n_blocks=10):
    """ModemFMStereo against the reference: exact per-block counts, the sum channel (discriminator, resampler, de-emphasis, low-pass:
    no feedback) to the usual 1e-5, the difference channel (through the pilot loop) to within a small multiple of the reference's
    own sensitivity to 1e-7 input noise (fm_stereo_floor) -- the tightest bar that means anything for this loop."""
    fs = 2400000
    chans = [("FMS", 200000, 300000), ("FMS", 250000, -500000), ("FMS", 200000, 700000), ("FM", 200000, 700000)]
    signal = [("FMS", 200000, 300000), ("FMS", 250000, -500000), ("FM", 200000, 700000)]
    out, _, _ = run_channels(lib, ref, fs, chans, n_blocks=n_blocks, signal=signal)
    check_channels([o for o in out if o["modem"] != "FMS"])
    res = []
    for o, ch in zip(out, chans):
        if o["modem"] != "FMS":
            continue
        assert o["counts_ok"] and o["skipped"] == o["ref_skipped"]
        a, r = o["audio"].reshape(-1, 2), o["ref_audio"].reshape(-1, 2)
        e_sum, e_diff = relrms(a.sum(1), r.sum(1)), relrms(a[:, 1] - a[:, 0], r[:, 1] - r[:, 0])
        f_sum, f_diff = fm_stereo_floor(ref, fs, ch, signal, n_blocks)
        assert e_sum <= TOL, "FM stereo sum channel: %.3g" % e_sum
        assert e_diff <= max(4.0 * f_diff, 2e-4), "FM stereo difference channel: %.3g (reference's own sensitivity %.3g)" % (e_diff, f_diff)
        assert relrms(o["iq"], o["ref_iq"]) <= TOL
        res.append((e_sum, e_diff, f_sum, f_diff))
    assert len(out[0]["audio"]) == n_blocks * 1600                # 800 frames of {left, right} per block
    a = out[0]["audio"].reshape(-1, 2)[4000:]
    assert np.std(a[:, 0]) > 0.05 and np.std(a[:, 1]) > 0.05 and np.std(a[:, 0] - a[:, 1]) > 0.05    # two different programmes came out
    return res


def check_audio_sinks(lib, ref, n_blocks=8):
    """SURVEY.md 8(f) row 4 on the device against the reference's code restated in the oracle: the audio mix of several demodulators
    (audioCallback: sources summed in binding order, queue / currentInput / audioQueuePtr bookkeeping over callbacks of 512 frames that
    do not line up with the 800-sample blocks, a two-channel source, gains, the peak normalisation), host-fed sources that exercise the
    callback's other branches (empty input, foreign sample rate, inactive source), 16-bit PCM words, scope waveform + audio spectrum."""
    from cubicsdr_b200 import Mixer, Scope, fetch_pcm16
    from oracle.refbind import RefMixer, RefScope, ref_wav_payload
    fs = 2400000
    chans = [("FM", 200000, 300000), ("AM", 6000, -450000), ("USB", 5400, 600000), ("FMS", 200000, 900000), ("NBFM", 12500, -150000)]
    signal = [("FM", 200000, 300000), ("AM", 6000, -450000), ("USB", 5400, 600000), ("FMS", 200000, 900000), ("NBFM", 12500, -150000)]
    gains = [0.9, 0.6, 1.3, 0.5, 0.8]
    N = synth.block_size(fs)
    eng = Engine(lib, max_block=N)
    cs = [Channel(eng, m, fs, bw, 48000, frequency=CENTER + off) for m, bw, off in chans]
    mix, rmix = Mixer(eng, 48000, 2048), RefMixer(ref, 48000)
    for c, g in zip(cs, gains):
        mix.bind(c, g); rmix.add_source(g)
    extra = mix.add_source(0.7); rmix.add_source(0.7)            # a host-fed source for the callback's rarer branches
    scope, rscope = Scope(eng, cs[0], 1024, 1024), RefScope(ref, 1024, 1024)
    scope2, rscope2 = Scope(eng, cs[3], 512, 600), RefScope(ref, 512, 600)        # two-channel input, truncated waveform
    rng = np.random.default_rng(11)
    worst = dict(mix=0.0, wave=0.0, spec=0.0, pcm=0)
    for b in range(n_blocks):
        x = synth.wideband(fs, N, signal, t0=b * N)
        eng.submit(x, CENTER, fs); eng.wait()
        # the reference side queues the ENGINE's audio (the mixer is what is under test here; the demodulators have their own tests)
        for i, c in enumerate(cs):
            a, lvl, pk, sk = c.fetch_audio()
            if len(a):
                rmix.push(i, a, channels=2 if chans[i][0] == "FMS" else 1, peak=pk, sample_rate=48000)
            words = fetch_pcm16(c)
            want = ref_wav_payload(ref, a, 2 if chans[i][0] == "FMS" else 1, pk)
            worst["pcm"] += int(np.count_nonzero(words != want)) if len(words) == len(want) else 10 ** 6
            if i == 0:
                w, s_, fc = scope.process(); rw, rs, rfc = rscope.process(a, 1, 48000, cs[0].bandwidth)
                assert len(w) == len(rw) and len(s_) == len(rs)
                worst["wave"] = max(worst["wave"], float(np.abs(w - rw).max())); worst["spec"] = max(worst["spec"], relrms(s_[1::2], rs[1::2]))
                assert np.allclose(fc, rfc, rtol=1e-5, atol=1e-9)
            if i == 3:
                w, s_, fc = scope2.process(); rw, rs, rfc = rscope2.process(a, 2, 48000, cs[3].bandwidth)
                assert len(w) == len(rw) and len(s_) == len(rs)
                worst["wave"] = max(worst["wave"], float(np.abs(w - rw).max())); worst["spec"] = max(worst["spec"], relrms(s_[1::2], rs[1::2]))
        # host-fed source: ordinary data, an empty input, a foreign rate, a pause
        if b % 4 == 0:
            d = (0.2 * rng.standard_normal(700)).astype(np.float32)
            mix.push(extra, d, 1, 0.6, 48000); rmix.push(extra, d, 1, 0.6, 48000)
        elif b % 4 == 1:
            e0 = np.zeros(0, np.float32)
            mix.push(extra, e0, 1, 0.0, 48000); rmix.push(extra, e0, 1, 0.0, 48000)
            d = (0.2 * rng.standard_normal(300)).astype(np.float32)
            mix.push(extra, d, 2, 0.5, 48000); rmix.push(extra, d, 2, 0.5, 48000)
        elif b % 4 == 2:
            d = (0.2 * rng.standard_normal(400)).astype(np.float32)
            mix.push(extra, d, 1, 0.4, 44100); rmix.push(extra, d, 1, 0.4, 44100)
            mix.push(extra, d, 1, 0.4, 48000); rmix.push(extra, d, 1, 0.4, 48000)
        if b == 5:
            mix.set_gain(1, 2.5); rmix.set_gain(1, 2.5)           # AudioThread::setGain clamps to [0, 2]
        # AudioThread::setActive: a demodulator is muted for a block and comes back -- its queue is flushed both times, it leaves the
        # controller's list and re-enters it at the END (the float sums then run in another order)
        if b == 2:
            mix.set_active(0, False); rmix.set_active(0, False)
        if b == 3:
            mix.set_active(0, True); rmix.set_active(0, True)
            mix.set_active(2, True); rmix.set_active(2, True)     # already active: only the flush happens
        if b == 6:
            mix.set_active(extra, False); rmix.set_active(extra, False)
        for cb in range(2 if b % 3 else 1):                       # 512-frame callbacks, sometimes one, sometimes two per block
            got, want = mix.render(512), rmix.callback(512)
            worst["mix"] = max(worst["mix"], float(np.abs(want).max()))
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "block %d callback %d: %d of %d mixed samples differ (max %.3g)" % (
                b, cb, int(np.count_nonzero(got != want)), got.size, float(np.abs(got - want).max()))
    # a bound demodulator goes away: its source keeps what it has queued and is fed no more (the slot may be reused)
    cs[4].close()
    for b in range(n_blocks, n_blocks + 2):
        x = synth.wideband(fs, N, signal, t0=b * N)
        eng.submit(x, CENTER, fs); eng.wait()
        for i, c in enumerate(cs[:4]):
            a, lvl, pk, sk = c.fetch_audio()
            if len(a):
                rmix.push(i, a, channels=2 if chans[i][0] == "FMS" else 1, peak=pk, sample_rate=48000)
        for cb in range(2):
            got, want = mix.render(512), rmix.callback(512)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "after a channel was destroyed: block %d callback %d" % (b, cb)
    cs = cs[:4]
    # a consumer that lags: more blocks queued than the reference's inputQueue holds (100, DemodulatorInstance.cpp:72-73) -- the
    # surplus is dropped (try_push on a full queue), nothing that is still queued is overwritten, what was kept plays in order
    lag = Mixer(eng, 48000, 1024)
    src = lag.add_source(1.0)
    for k in range(150):
        lag.push(src, np.full(1600, float(k + 1), np.float32), 2, 0.5, 48000)
    played = []
    for k in range(210):
        o = lag.render(800)
        assert np.all(np.isfinite(o))
        played.extend(np.unique(o[o != 0]).tolist())
    assert played == [float(k + 1) for k in range(100)], played[:12]
    # blocks too long for a source's ring are dropped before the queue is full; what was kept still plays in order
    for k in range(40):
        lag.push(src, np.full(30000, float(1000 + k), np.float32), 1, 0.5, 48000)
    played = []
    for k in range(400):
        o = lag.render(1024)
        played.extend(np.unique(o[o != 0]).tolist())
    kept = sorted(set(played))
    assert played == sorted(played) and 3 <= len(kept) < 40 and kept[0] == 1000.0, kept[:12]
    # more sources than the ring pool starts with: it grows, the earlier sources' queued audio survives the move
    many = Mixer(eng, 48000, 512)
    first = many.add_source(1.0)
    many.push(first, np.full(600, 0.25, np.float32), 1, 0.25, 48000)
    others = [many.add_source(1.0) for _ in range(40)]
    many.push(others[-1], np.full(600, 0.5, np.float32), 1, 0.5, 48000)
    many.render(512)                                              # (the first callback only fetches currentInput, :126-134)
    many.push(first, np.zeros(1, np.float32), 1, 0.0, 48000); many.push(others[-1], np.zeros(1, np.float32), 1, 0.0, 48000)
    o = many.render(512)
    assert np.all(o[:1000] == np.float32(0.75)), o[:8]
    many.close()
    lag.close()
    assert worst["mix"] > 0.05                                       # the callbacks did carry audio
    assert worst["pcm"] == 0 and worst["wave"] <= 1e-6 and worst["spec"] <= TOL, worst
    for o in (mix, scope, scope2):
        o.close()
    rmix.close(); rscope.close(); rscope2.close()
    for c in cs:
        c.close()
    eng.close()
    return worst
