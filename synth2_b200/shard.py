"""Multi-GPU plumbing: voices are independent (synth.rs:177-199 touches only `voice.state`), so a bank
shards as contiguous voice ranges, one process per GPU, with no data-path collective.  The only
exchange step is the optional master mix: one reduce (sum) of the per-rank mono buses to rank 0
over NCCL/NVLink, issued once per render, never per block.

The reduce itself is the library's (`s2_bank_reduce_bus`, include/s2_cuda.h): `MasterBus` builds the NCCL
communicator through the C ABI — the 128-byte unique id travels over whatever the host already has, here
torch.distributed — and calls it.  `reduce_master_bus` keeps the torch.distributed form for CPU tests (gloo)."""
import ctypes as C

import numpy as np


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


class MasterBus:
    """`s2_comm` + `s2_bank_reduce_bus`: the library's own NCCL reduce of the per-GPU buses."""

    def __init__(self, rank: int, world: int, device: int, exchange=None):
        """`exchange(id_bytes_or_None) -> id_bytes`: ships rank 0's 128-byte id to every rank.  Default:
        torch.distributed broadcast (the process group that launched the ranks)."""
        from ._lib import check, lib
        self._lib, self._check = lib(), check
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        uid = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            check(self._lib.s2_comm_unique_id(C.c_void_p(uid.ctypes.data)))
        if exchange is None:
            exchange = self._torch_exchange
        uid = np.frombuffer(exchange(uid.tobytes() if self.rank == 0 else None), dtype=np.uint8).copy()
        self._h = C.c_void_p()
        check(self._lib.s2_comm_create(C.c_void_p(uid.ctypes.data), self.world, self.rank, self.device, C.byref(self._h)))

    def _torch_exchange(self, payload):
        if self.world == 1:
            return payload
        import torch
        import torch.distributed as dist
        t = torch.zeros(128, dtype=torch.uint8, device=torch.device("cuda", self.device))
        if self.rank == 0:
            t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        return bytes(t.cpu().numpy().tobytes())

    def reduce(self, bank, bus_in, bus_out=None, root: int = 0, stream=None):
        """Sum `bus_in` of every rank into `bus_out` on `root` (asynchronous on `stream`)."""
        from ._lib import ptr
        sp = None if stream is None else C.c_void_p(getattr(stream, "cuda_stream", stream))
        self._check(self._lib.s2_bank_reduce_bus(bank._h, self._h, int(root), ptr(bus_in), ptr(bus_out), int(bus_in.numel()), sp))

    def close(self):
        if self._h:
            self._lib.s2_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
