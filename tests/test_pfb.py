"""Polyphase filterbank channelizer (SDRPostThread::runPFBCH) against the reference binary's
firpfbch_crcf analyzer.  Emulator tier on CPU, the same checks on the device in the GPU tier."""
import numpy as np
import pytest

from cubicsdr_b200 import Channelizer, Engine, synth
from oracle.refbind import RefPfb, RefPfb2
from tests.parity import TOL, relrms


def _stream(lib, ref, M, frames_per_call, seed=7, device=0, oversampled=False):
    """Several calls of different lengths through engine and oracle: state (the filter windows) must carry.  With
    oversampled=True (firpfbch2) a frame consumes M / 2 samples, and odd frame counts exercise the alternating branch phase."""
    rng = np.random.default_rng(seed)
    eng = Engine(lib, device=device, max_block=4096)
    ch = Channelizer(eng, M, max(frames_per_call), oversampled=oversampled)
    rp = RefPfb2(ref, M) if oversampled else RefPfb(ref, M)
    step = M // 2 if oversampled else M
    gots, wants = [], []
    for F in frames_per_call:
        x = (rng.standard_normal(F * step) + 1j * rng.standard_normal(F * step)).astype(np.complex64)
        x += np.exp(2j * np.pi * 0.123 * np.arange(F * step)).astype(np.complex64) * 3        # a strong tone: one hot channel
        got, want = ch.execute(x), rp.execute(x)
        assert got.shape == want.shape == (F, M)
        gots.append(got); wants.append(want)
    ch.close(); rp.close(); eng.close()
    # over the whole stream: the first frames of a 2-channel bank are numerically zero on both sides (sinc prototype)
    return relrms(np.concatenate(gots), np.concatenate(wants))


CASES = [(2, [5, 1, 9, 40]), (8, [3, 16, 1, 1, 33]), (40, [7, 2, 50]), (122, [20, 3, 41])]


@pytest.mark.parametrize("M,calls", CASES)
def test_channelizer_matches_reference_binary_emulator(emu_lib, ref, M, calls):
    assert _stream(emu_lib, ref, M, calls) <= TOL


def test_optimal_channel_count_rule(emu_lib, ref):
    """SDRThread::getOptimalChannelCount and the block-size rule that makes every block a whole number of frames."""
    ref.L.ref_optimal_channel_count.argtypes = [__import__("ctypes").c_longlong]
    for fs in (250000, 500000, 500001, 1000000, 2400000, 10000000, 20000000, 30720000, 40000000, 61440000):
        want = ref.L.ref_optimal_channel_count(fs)
        assert Channelizer.optimal_channel_count(emu_lib, fs) == want
        if want > 1:
            assert synth.block_size(fs) % want == 0
    assert Channelizer.optimal_channel_count(emu_lib, 61440000) == 122


def test_reset_and_bad_lengths(emu_lib, ref):
    from cubicsdr_b200 import SdrError
    eng = Engine(emu_lib, max_block=4096)
    ch = Channelizer(eng, 8, 16)
    x = (np.arange(64) + 1j).astype(np.complex64)
    a = ch.execute(x)
    ch.reset()
    assert np.array_equal(ch.execute(x), a)                      # reset = empty windows again
    with pytest.raises(SdrError):
        ch.execute(x[:63])                                         # not a whole number of frames
    with pytest.raises(SdrError):
        ch.execute(np.zeros(8 * 17, np.complex64))                 # more frames than the object was sized for
    with pytest.raises(SdrError):
        Channelizer(eng, 7, 16)                                    # the reference only ever makes even counts
    ch.close(); eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("M,calls", CASES + [(122, [8394])])
def test_channelizer_matches_reference_binary_gpu(gpu_lib, ref, M, calls):
    assert _stream(gpu_lib, ref, M, calls) <= TOL


def test_engine_channelizer_mode_emulator(emu_lib, ref):
    """SDRPostThread::runPFBCH end to end at 2.4 MSps (4 channels of 600 kHz; 5 streams with the wrap channel):
    demodulators on channel 0 (DC-corrected), on an ordinary channel, on the wrap channel, and one that hops between
    channels mid-stream."""
    from tests import parity
    fs = 2400000
    chans = [("FM", 200000, 40000), ("AM", 6000, -650000), ("NBFM", 12500, 1150000), ("USB", 5400, 560000), ("FM", 200000, -1100000)]
    used = parity.check_channelizer_mode(emu_lib, ref, fs, chans, n_blocks=5, retune={2: {3: -90000}, 3: {1: 700000}})
    assert {0, 4}.issubset(used) and len(used) >= 4          # channel 0, the extra (wrap) channel and ordinary ones


def test_engine_channelizer_mode_20_channels_emulator(emu_lib, ref):
    """10 MSps: 20 filterbank channels of 500 kHz, demodulators spread over both halves of the band."""
    from tests import parity
    fs = 10000000
    chans = [("FM", 200000, 1480000), ("NBFM", 12500, -2260000), ("AM", 6000, 4890000), ("LSB", 5400, -4990000), ("FM", 150000, -130000)]
    used = parity.check_channelizer_mode(emu_lib, ref, fs, chans, n_blocks=3)
    assert len(used) == 5


@pytest.mark.gpu
def test_engine_channelizer_mode_gpu(gpu_lib, ref):
    from tests import parity
    fs = 2400000
    chans = [("FM", 200000, 40000), ("AM", 6000, -650000), ("NBFM", 12500, 1150000), ("USB", 5400, 560000), ("FM", 200000, -1100000)]
    used = parity.check_channelizer_mode(gpu_lib, ref, fs, chans, n_blocks=5, retune={2: {3: -90000}, 3: {1: 700000}})
    assert {0, 4}.issubset(used)


@pytest.mark.gpu
def test_engine_channelizer_mode_122_channels_gpu(gpu_lib, ref):
    """The BASELINE rate: 61.44 MSps = 122 channels of 503 606 Hz, twelve demodulators across the band, two blocks."""
    from tests import parity
    fs = 61440000
    offs = [-30500000, -21000000, -12345678, -5000000, -250000, 60000, 3000000, 9876543, 15000000, 22222222, 28000000, 30600000]
    modems = ["FM", "AM", "USB", "LSB", "NBFM", "FM"]
    bws = {"FM": 200000, "AM": 6000, "USB": 5400, "LSB": 5400, "NBFM": 12500}
    chans = [(modems[i % 6], bws[modems[i % 6]], o) for i, o in enumerate(offs)]
    used = parity.check_channelizer_mode(gpu_lib, ref, fs, chans, n_blocks=2)
    assert 0 in used and 122 in used and len(used) >= 10


@pytest.mark.parametrize("M,calls", [(2, [5, 1, 9, 40]), (8, [3, 16, 1, 1, 33]), (40, [7, 2, 50]), (122, [21, 3, 40])])
def test_oversampled_channelizer_matches_reference_binary_emulator(emu_lib, ref, M, calls):
    assert _stream(emu_lib, ref, M, calls, oversampled=True) <= TOL


def test_engine_oversampled_channelizer_mode_emulator(emu_lib, ref):
    """SDRPostThread::runPFBCH2: demodulators at twice the channel rate behind firpfbch2."""
    from tests import parity
    chans = [("FM", 200000, 40000), ("AM", 6000, -650000), ("NBFM", 12500, 1150000), ("USB", 5400, 560000)]
    used = parity.check_channelizer_mode(emu_lib, ref, 2400000, chans, n_blocks=4, retune={2: {3: -90000}}, mode=2)
    assert {0, 4}.issubset(used)


@pytest.mark.gpu
@pytest.mark.parametrize("M,calls", [(8, [3, 16, 1, 1, 33]), (122, [21, 3, 40])])
def test_oversampled_channelizer_matches_reference_binary_gpu(gpu_lib, ref, M, calls):
    assert _stream(gpu_lib, ref, M, calls, oversampled=True) <= TOL


@pytest.mark.gpu
def test_engine_oversampled_channelizer_mode_gpu(gpu_lib, ref):
    from tests import parity
    chans = [("FM", 200000, 40000), ("AM", 6000, -650000), ("NBFM", 12500, 1150000), ("USB", 5400, 560000)]
    used = parity.check_channelizer_mode(gpu_lib, ref, 2400000, chans, n_blocks=4, retune={2: {3: -90000}}, mode=2)
    assert {0, 4}.issubset(used)


def test_engine_channelizer_mode_ties_and_band_edges_emulator(emu_lib, ref):
    """getChannelAt corner cases: a demodulator exactly half-way between two channel centres (the first minimum wins), one
    on the lower band edge (channel numChannels/2) and one on the upper edge (the extra channel that wraps onto it)."""
    from tests import parity
    fs = 2400000                                             # 4 channels of 600 kHz: centres 0, +600k, -1200k, -600k, (+1200k)
    chans = [("AM", 6000, 300000), ("AM", 6000, -300000), ("FM", 200000, -1200000), ("FM", 200000, 1200000), ("NBFM", 12500, 900000)]
    used = parity.check_channelizer_mode(emu_lib, ref, fs, chans, n_blocks=3)
    assert {0, 2, 4}.issubset(used)


def test_engine_channelizer_mode_tuner_moves_emulator(emu_lib, ref):
    """The tuner frequency changes between blocks (SDRPostThread::runPFBCH :425-429: updateChannels): every channel centre
    moves, demodulators keep their absolute frequency, so their shifts change and some hop to another filterbank channel."""
    from tests import parity
    fs = 2400000
    chans = [("AM", 6000, 250000), ("USB", 5400, -700000), ("LSB", 5400, 950000), ("AM", 10000, -80000)]
    moves = {2: parity.CENTER + 400000, 4: parity.CENTER - 350000}
    used = parity.check_channelizer_mode(emu_lib, ref, fs, chans, n_blocks=6, center_by_block=moves)
    assert len(used) >= 4


def _check_channelizer_golden(lib, ref, device=0):
    """Committed golden vectors (tests/golden/channelizer.npz, written by the reference binary): they pin the oracle (a
    transplant that broke would stop reproducing them) and the engine."""
    import os
    from cubicsdr_b200 import Channel
    from tests import parity
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "channelizer.npz"))
    eng = Engine(lib, device=device, max_block=4096)
    for M in (8, 122):
        x = g["pfb_in_M%d" % M]
        for key, cls, over in (("pfb_out_M%d", RefPfb, False), ("pfb2_out_M%d", RefPfb2, True)):
            want = g[key % M]
            rp = cls(ref, M)
            assert np.array_equal(rp.execute(x), want), "oracle no longer reproduces its golden vector (%s)" % (key % M)
            rp.close()
            ch = Channelizer(eng, M, want.shape[0], oversampled=over)
            assert relrms(ch.execute(x), want) <= TOL
            ch.close()
    eng.close()
    fs, N = 2400000, synth.block_size(2400000)
    chans = [("FM", 200000, 40000), ("AM", 6000, -650000)]
    for mode in (1, 2):
        eng = Engine(lib, device=device, max_block=N)
        eng.set_channelizer(mode)
        M, chan_rate = eng.channel_plan(fs)
        eng.set_channel_block_max(mode * (N // M))
        cs = [Channel(eng, m, chan_rate * mode, bw, 48000, frequency=parity.CENTER + off) for m, bw, off in chans]
        for b in range(2):
            eng.submit(synth.wideband(fs, N, chans, seed=77, t0=b * N), parity.CENTER, fs)
            eng.wait()
            for (m, _, _), c in zip(chans, cs):
                want = g["mode%d_audio_%s_b%d" % (mode, m, b)]
                a = c.fetch_audio()[0]
                assert len(a) == len(want) and relrms(a, want) <= TOL, (mode, m, b)
        eng.close()


def test_channelizer_golden_vectors_emulator(emu_lib, ref):
    _check_channelizer_golden(emu_lib, ref)


@pytest.mark.gpu
def test_channelizer_golden_vectors_gpu(gpu_lib, ref):
    _check_channelizer_golden(gpu_lib, ref)
