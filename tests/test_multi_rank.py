"""world_size-2 test of the sharded path on CPU (gloo): rank 0 owns the raw block, broadcasts it, every rank runs its
channel shard (kernel phases replayed by the test-only emulator), and the union of the shards equals the single-rank
result bit for bit.  Mode 0 goes through the product's group API in its one-process-per-GPU form (sdr_group_create_rank
without an NCCL id: gloo moves the block, the group does placement and submission); the channelizer modes drive the
engines directly."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHANS = [("FM", 200000, 300000), ("NBFM", 12500, -150000), ("AM", 6000, 450000), ("USB", 5400, -600000), ("LSB", 5400, 750000),
         ("NBFM", 12500, 100000)]
FS, NB = 2400000, 3


def _run_rank(rank, world, port, emu_path, out_dir, mode=0):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from cubicsdr_b200 import Channel, Engine, Lib, synth
    from cubicsdr_b200.parallel import broadcast_block, shard_channels
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = Lib(emu_path)
    N = synth.block_size(FS)
    if mode == 0:
        from cubicsdr_b200 import Group
        g = Group(lib, max_block=N, rank=rank, world=world, nccl_id=None)
        assert (g.n_local, g.total, g.transport) == (1, world, 0)
        cs = [(idx, g.channel(m, FS, bw, 48000, frequency=100000000 + off)) for idx, (m, bw, off) in enumerate(CHANS)]
        assert [c is not None for _, c in cs] == [idx % world == rank for idx, _ in cs]      # placement: channel c on rank c mod N
        cs = [(idx, c) for idx, c in cs if c is not None]
        res = {idx: [] for idx, _ in cs}
        for b in range(NB):
            blk = torch.zeros(N, dtype=torch.complex64)
            if rank == 0:
                blk = torch.from_numpy(synth.wideband(FS, N, CHANS, t0=b * N))
            broadcast_block(dist, blk, src=0)
            g.submit(blk.numpy(), 100000000, FS)
            g.wait()
            for idx, c in cs:
                res[idx].append(c.fetch_audio()[0])
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{str(k): np.concatenate(v) for k, v in res.items()})
        dist.barrier()
        dist.destroy_process_group()
        return
    eng = Engine(lib, max_block=N)
    chan_fs = FS
    if mode:                                                     # channelizer mode: every rank runs the filterbank on the block
        eng.set_channelizer(mode)
        n_pfb, chan_fs = eng.channel_plan(FS)
        eng.set_channel_block_max(N // n_pfb)
    mine = shard_channels(list(enumerate(CHANS)), rank, world)
    cs = [(idx, Channel(eng, m, chan_fs, bw, 48000, frequency=100000000 + off)) for idx, (m, bw, off) in mine]
    res = {idx: [] for idx, _ in cs}
    for b in range(NB):
        blk = torch.zeros(N, dtype=torch.complex64)
        if rank == 0:
            blk = torch.from_numpy(synth.wideband(FS, N, CHANS, t0=b * N))
        broadcast_block(dist, blk, src=0)
        eng.submit(blk.numpy(), 100000000, FS)
        eng.wait()
        for idx, c in cs:
            res[idx].append(c.fetch_audio()[0])
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **{str(k): np.concatenate(v) for k, v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", [0, 1])
def test_two_ranks_equal_single_rank(emu_lib, tmp_path, mode):
    import torch.multiprocessing as mp
    from cubicsdr_b200 import Channel, Engine, synth
    from cubicsdr_b200.parallel import owner_of
    port = 29500 + (os.getpid() % 500) + 7 * mode
    mp.spawn(_run_rank, args=(2, port, emu_lib.path, str(tmp_path), mode), nprocs=2, join=True)
    N = synth.block_size(FS)
    eng = Engine(emu_lib, max_block=N)
    chan_fs = FS
    if mode:
        eng.set_channelizer(mode)
        n_pfb, chan_fs = eng.channel_plan(FS)
        eng.set_channel_block_max(N // n_pfb)
    cs = [Channel(eng, m, chan_fs, bw, 48000, frequency=100000000 + off) for m, bw, off in CHANS]
    single = [[] for _ in cs]
    for b in range(NB):
        eng.submit(synth.wideband(FS, N, CHANS, t0=b * N), 100000000, FS)
        eng.wait()
        for i, c in enumerate(cs):
            single[i].append(c.fetch_audio()[0])
    shards = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(2)]
    for i in range(len(CHANS)):
        got = shards[owner_of(i, 2)][str(i)]
        assert np.array_equal(got, np.concatenate(single[i])), CHANS[i]
    assert sorted(int(k) for s in shards for k in s.files) == list(range(len(CHANS)))
