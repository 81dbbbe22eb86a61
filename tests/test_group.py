"""sdr_group: several GPUs behind the same boundary.  Channel c lives on member c mod N, the spectrum on
member 0, the raw block reaches every member by one broadcast; the union of the members' results must equal what a single
engine produces, bit for bit (every member replays the same DC blocker and runs the same kernels on the same bytes)."""
import numpy as np
import pytest

from cubicsdr_b200 import Channel, Engine, Group, Spectrum, synth
from tests.parity import CENTER

CHANS = [("FM", 200000, 300000), ("NBFM", 12500, -150000), ("AM", 6000, 450000), ("USB", 5400, -600000), ("LSB", 5400, 750000),
         ("NBFM", 12500, 100000), ("FM", 200000, -420000)]


def _single(lib, fs, chans, blocks, fft, device=0):
    N = len(blocks[0])
    eng = Engine(lib, device=device, max_block=N)
    cs = [Channel(eng, m, fs, bw, 48000, frequency=CENTER + off) for m, bw, off in chans]
    sp = Spectrum(eng, fft, 0) if fft else None
    audio, pts = [[] for _ in cs], []
    for x in blocks:
        eng.submit(x, CENTER, fs); eng.wait()
        for i, c in enumerate(cs):
            audio[i].append(c.fetch_audio()[0])
        if sp:
            pts.append(np.stack([sp.fetch(f)[0] for f in range(sp.frames)]))
    for c in cs:
        c.close()
    if sp:
        sp.close()
    eng.close()
    return [np.concatenate(a) for a in audio], pts


def _grouped(lib, fs, chans, blocks, fft, devices, pipelined=False):
    N = len(blocks[0])
    g = Group(lib, devices=devices, max_block=N)
    cs = [g.channel(m, fs, bw, 48000, frequency=CENTER + off) for m, bw, off in chans]
    assert all(c is not None for c in cs)
    sp = g.spectrum(fft) if fft else None
    audio, pts = [[] for _ in cs], []

    def grab():
        for i, c in enumerate(cs):
            audio[i].append(c.fetch_audio()[0])
        if sp:
            pts.append(np.stack([sp.fetch(f)[0] for f in range(sp.frames)]))

    for b, x in enumerate(blocks):
        g.submit(x, CENTER, fs)
        if not pipelined:
            g.wait(); grab()
        elif b >= 1:
            g.wait_block(1); grab()
    if pipelined:
        g.wait_block(0); grab()
    info = (g.n_local, g.total, g.transport)
    for c in cs:
        c.close()
    if sp:
        sp.close()
    g.close()
    return [np.concatenate(a) for a in audio], pts, info


def _check(lib, devices, fs=2400000, n_blocks=4, fft=1024, pipelined=False):
    N = synth.block_size(fs)
    blocks = [synth.wideband(fs, N, CHANS, seed=5, t0=b * N) for b in range(n_blocks)]
    want_a, want_p = _single(lib, fs, CHANS, blocks, fft, device=devices[0])
    got_a, got_p, info = _grouped(lib, fs, CHANS, blocks, fft, devices, pipelined)
    assert info[0] == len(devices) and info[1] == len(devices)
    for i, (a, b) in enumerate(zip(want_a, got_a)):
        assert len(a) > 0 and np.array_equal(a, b), CHANS[i]
    assert len(want_p) == len(got_p)
    for a, b in zip(want_p, got_p):
        assert np.array_equal(a, b)
    return info


def test_group_of_three_members_emulator(emu_lib):
    _check(emu_lib, [0, 0, 0])
    _check(emu_lib, [0, 0], pipelined=True)


@pytest.mark.gpu
def test_group_on_one_gpu(gpu_lib):
    """One member: the group is a plain engine.  Two members on the same GPU: the broadcast degenerates to a device copy
    (NCCL refuses a duplicate device, so this exercises the peer-copy transport), the sharding logic is the real one."""
    _check(gpu_lib, [0])
    info = _check(gpu_lib, [0, 0], pipelined=True)
    assert info[2] in (1, 2)


@pytest.mark.gpu
def test_group_two_gpus_matches_single_gpu_bit_for_bit(gpu_lib):
    """Hardware parity of the sharded path: channels on two GPUs, the raw block moved by ncclBroadcast over NVLink."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    info = _check(gpu_lib, [0, 1], fs=10000000, n_blocks=4, fft=8192, pipelined=True)
    assert info[2] == 1, "expected the NCCL transport, got %r" % (info,)
    if torch.cuda.device_count() >= 4:
        _check(gpu_lib, [0, 1, 2, 3], fs=10000000, n_blocks=3, fft=8192)
