"""Multi-GPU plumbing: voices are independent (synth.rs:177-199 touches only `voice.state`), so a bank
shards as contiguous voice ranges, one process per GPU, with no data-path collective.  The only
exchange step is the optional master mix: one reduce (sum) of the per-rank mono buses to rank 0
over NCCL/NVLink, issued once per render, never per block."""


def voice_range(rank: int, world: int, n_voices: int):
    """Contiguous range [lo, hi) owned by `rank`; sizes differ by at most one voice."""
    base, extra = divmod(int(n_voices), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_master_bus(bus, dst: int = 0):
    """Sum the per-rank buses into rank `dst` (torch.distributed: NCCL on GPUs, gloo in CPU tests).

    f32 summation order across ranks is the collective's, not the reference's voice-index order:
    the master bus is compared with a tolerance, never bit-for-bit."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(bus, dst=dst, op=dist.ReduceOp.SUM)
    return bus
