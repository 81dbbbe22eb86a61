"""Multi-GPU plumbing for the one exchange step of the path: channels are sharded over ranks and the raw
IQ block is broadcast from rank 0 (NCCL over NVLink on GPUs; gloo in the CPU tests).  No reduction step exists: every
rank returns the audio of its own channels."""


def shard_channels(chans, rank, world):
    """Channel c of the job goes to rank c mod world (round robin keeps the FM/AM/USB/LSB mix balanced per rank)."""
    return list(chans[rank::world])


def owner_of(channel_index, world):
    return channel_index % world


def broadcast_block(dist, tensor, src=0):
    """Broadcast one raw IQ block (complex64 tensor, viewed as float32 pairs) from `src` to every rank."""
    import torch
    t = torch.view_as_real(tensor) if tensor.is_complex() else tensor
    dist.broadcast(t, src=src)
    return tensor
