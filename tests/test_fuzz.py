"""Seeded random configurations of the single-channel path on the emulator tier against the reference binary: wideband
rate, audio rate, 1-4 demodulators of random modem / bandwidth / offset, 2-6 blocks of mixed lengths, retunes (including
onto 0 Hz and beyond the 0.75 fs skip rule).  FM / NBFM channels are judged on the decimated IQ (an FM discriminator
amplifies 1e-6 IQ differences without bound wherever a retune or start-up transient takes the amplitude through zero);
AM / SSB channels on the audio.  The DC-blocked block must be bit-exact in every block.

Blocks shorter than the front kernel's history window (13 x 2^K <= 832 samples) take the rare paths: around a retune the
history is mixed with every oscillator setting that was in effect (dsp_core.h premix_thread); next to a block skipped by
the 0.75 fs rule the consumed history is not contiguous in the stream and is kept in a rolling private tail
(engine.cu, roll_pending).  The `tiny` variant of the generator exercises both.  CubicSDR's own blocks are 1/60 s
(>= 4167 samples), so the reference never gets there."""
import numpy as np
import pytest

from cubicsdr_b200 import Channel, Engine, synth
from oracle.refbind import RefChan, RefIngress
from tests.parity import CENTER, TOL, relrms

MODEMS = {"FM": [200000, 150000, 120000, 80000, 31000], "NBFM": [12500, 25000, 8000, 16000], "AM": [6000, 10000, 4000, 9000],
          "USB": [5400, 2700, 3000], "LSB": [5400, 2700, 3000]}


def _case(lib, ref, seed, tiny=False):
    rng = np.random.default_rng(seed)
    fs = int(rng.choice([1000000, 1024000, 2000000, 2400000, 2880000, 3200000]))
    audio_rate = int(rng.choice([48000, 48000, 44100, 96000, 32000]))
    chans = []
    for _ in range(int(rng.integers(1, 5))):
        m = str(rng.choice(list(MODEMS)))
        bw = int(rng.choice(MODEMS[m]))
        off = 0 if rng.random() < 0.1 else int(rng.integers(-fs // 2 + bw, fs // 2 - bw))
        chans.append((m, bw, off))
    sizes = [synth.block_size(fs), 4099, 777, 257, 100, 12345] if tiny else [synth.block_size(fs), 4099, 12345, 2048, 40003, 1500]
    blocks = [int(rng.choice(sizes)) for _ in range(int(rng.integers(2, 7)))]
    if tiny:
        blocks = [synth.block_size(fs)] + blocks + [4099]       # a real signal level before and after the short blocks
    retune = {}
    for b in range(1, len(blocks)):
        if rng.random() < 0.4:
            i = int(rng.integers(0, len(chans)))
            r = rng.random()
            if r < 0.15:
                new = int(fs * 0.8) * (1 if rng.random() < 0.5 else -1)          # beyond 0.75 fs: the block is skipped
            elif r < 0.3:
                new = 0
            else:
                new = int(rng.integers(-fs // 2 + chans[i][1], fs // 2 - chans[i][1]))
            retune.setdefault(b, {})[i] = new
    signal = list(chans) + [(chans[i][0], chans[i][1], o if abs(o) < fs // 2 else 0) for b in retune for i, o in retune[b].items()]
    eng = Engine(lib, max_block=max(blocks))
    cs = [Channel(eng, m, fs, bw, audio_rate, frequency=CENTER + off, want_iq=True) for m, bw, off in chans]
    rcs = [RefChan(ref, m, fs, bw, audio_rate) for m, bw, _ in chans]
    ing = RefIngress(ref)
    offs = [o for _, _, o in chans]
    got = [[] for _ in chans]; want = [[] for _ in chans]
    t0 = 0
    for b, n in enumerate(blocks):
        for i, o in retune.get(b, {}).items():
            offs[i] = o
            cs[i].set_frequency(CENTER + o)
        x = synth.wideband(fs, n, signal, seed=seed, t0=t0)
        t0 += n
        eng.submit(x, CENTER, fs)
        eng.wait()
        y = ing.block(x)
        assert np.array_equal(eng.fetch_block(n), y), "seed %d block %d: DC-blocked block not bit-exact" % (seed, b)
        for i, (c, rc) in enumerate(zip(cs, rcs)):
            a, _, _, sk = c.fetch_audio()
            r = rc.process(y, offs[i])
            assert bool(sk) == bool(r["skipped"]), "seed %d demod %d block %d: skip rule" % (seed, i, b)
            assert len(a) == len(r["audio"]), "seed %d demod %d block %d: %d audio samples vs %d" % (seed, i, b, len(a), len(r["audio"]))
            if chans[i][0] in ("FM", "NBFM"):
                if not sk:
                    q = c.fetch_iq()
                    assert len(q) == len(r["iq"])
                    got[i].append(q.view(np.float32)); want[i].append(r["iq"].view(np.float32))
            else:
                got[i].append(a); want[i].append(r["audio"])
    for c in cs:
        c.close()
    for rc in rcs:
        rc.close()
    ing.close(); eng.close()
    for i in range(len(chans)):
        g, w = (np.concatenate(got[i]), np.concatenate(want[i])) if got[i] else (np.zeros(0), np.zeros(0))
        if len(w) and np.linalg.norm(w) > 0:
            e = relrms(g, w)
            assert e <= TOL, "seed %d demod %d %s: relative RMS error %.3g" % (seed, i, chans[i], e)


@pytest.mark.parametrize("seed0", [5000, 5040, 5080])
def test_random_configurations_emulator(emu_lib, ref, seed0):
    for seed in range(seed0, seed0 + 40):
        _case(emu_lib, ref, seed)


def test_random_configurations_with_short_blocks_emulator(emu_lib, ref):
    for seed in range(9000, 9060):
        _case(emu_lib, ref, seed, tiny=True)


CH_MODEMS = {"FM": [200000, 150000, 120000, 80000], "NBFM": [12500, 25000, 16000], "AM": [6000, 10000, 4000], "USB": [5400, 2700],
             "LSB": [5400, 3000]}


def _channelizer_case(lib, ref, seed):
    """Channelizer mode (runPFBCH / runPFBCH2): random rate, 1-4 demodulators anywhere in (and slightly beyond) the band,
    retunes and tuner moves between blocks -- i.e. demodulators hopping between filterbank channels, onto and off channel 0
    (whose DC blocker only advances while it is in use) and onto the wrap channel."""
    from tests.parity import RefChannelizerPath
    rng = np.random.default_rng(seed)
    fs = int(rng.choice([1000000, 2000000, 2400000, 3200000, 5000000, 10000000]))
    mode = int(rng.choice([1, 1, 2]))
    chans = []
    for _ in range(int(rng.integers(1, 5))):
        m = str(rng.choice(list(CH_MODEMS)))
        chans.append((m, int(rng.choice(CH_MODEMS[m])), int(rng.integers(-fs // 2 - 100000, fs // 2 + 100000))))
    nb = int(rng.integers(2, 6))
    N = synth.block_size(fs)
    eng = Engine(lib, max_block=N)
    eng.set_channelizer(mode)
    M, _ = eng.channel_plan(fs)
    oracle = RefChannelizerPath(ref, fs, chans, mode=mode)
    eng.set_channel_block_max(mode * (N // M))
    cs = [Channel(eng, m, oracle.demod_rate, bw, 48000, frequency=CENTER + off, want_iq=True) for m, bw, off in chans]
    freqs = [CENTER + off for _, _, off in chans]
    retune, centers = {}, {}
    for b in range(1, nb):
        if rng.random() < 0.4:
            retune.setdefault(b, {})[int(rng.integers(0, len(chans)))] = int(rng.integers(-fs // 2, fs // 2))
        if rng.random() < 0.2:
            centers[b] = CENTER + int(rng.integers(-500000, 500000))
    signal = list(chans) + [(chans[i][0], chans[i][1], o) for b in retune for i, o in retune[b].items()]
    got = [[] for _ in chans]; want = [[] for _ in chans]
    center = CENTER
    for b in range(nb):
        for i, o in retune.get(b, {}).items():
            freqs[i] = CENTER + o
            cs[i].set_frequency(freqs[i])
        center = centers.get(b, center)
        x = synth.wideband(fs, N, signal, seed=seed, t0=b * N) + np.complex64(0.02 + 0.01j)
        eng.submit(x, center, fs)
        eng.wait()
        res, _ = oracle.block(x, center, freqs)
        for i, (c, r) in enumerate(zip(cs, res)):
            a, _, _, sk = c.fetch_audio()
            if r is None or r["skipped"]:
                assert sk or len(a) == 0, "seed %d demod %d block %d: audio where the reference had none" % (seed, i, b)
                continue
            assert not sk and len(a) == len(r["audio"]), "seed %d demod %d block %d: counts" % (seed, i, b)
            if chans[i][0] in ("FM", "NBFM"):
                q = c.fetch_iq()
                assert len(q) == len(r["iq"])
                got[i].append(q.view(np.float32)); want[i].append(r["iq"].view(np.float32))
            else:
                got[i].append(a); want[i].append(r["audio"])
    for c in cs:
        c.close()
    oracle.close(); eng.close()
    for i in range(len(chans)):
        if got[i]:
            g, w = np.concatenate(got[i]), np.concatenate(want[i])
            if np.linalg.norm(w) > 0:
                e = relrms(g, w)
                assert e <= TOL, "seed %d demod %d %s: relative RMS error %.3g" % (seed, i, chans[i], e)


def test_random_channelizer_configurations_emulator(emu_lib, ref):
    for seed in range(100, 160):
        _channelizer_case(emu_lib, ref, seed)
