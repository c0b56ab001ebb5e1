"""One process per GPU: observation sharding and the engine's NCCL communicator.

Observations are row-sharded: rank r of W holds the contiguous block
[shard_range(r, W, N)) of the global data set and draws its omega shard from Philox streams
keyed by the GLOBAL observation index, so the chain does not depend on W.  The only exchange
per beta draw is one all-reduce of P*P + P doubles (NCCL over NVLink/NVSwitch, issued inside
the engine on its own stream); torch.distributed is used only to hand the 128-byte NCCL id
from rank 0 to the others.
"""
import ctypes as C


def shard_range(rank, world, N):
    """Contiguous block of observations owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(N), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_comm(rank, world, device=None):
    """Create the engine communicator; call on every rank after torch.distributed is up."""
    import torch
    import torch.distributed as dist
    from . import _lib
    L = _lib.lib()
    if world <= 1:
        return
    buf = (C.c_char * 128)()
    if rank == 0:
        _lib.check(L.bl_comm_unique_id(C.cast(buf, C.c_void_p)))
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if device is not None and dist.get_backend() == "nccl":
        t = t.to(device)
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().numpy().tobytes())
    ident = (C.c_char * 128).from_buffer_copy(raw)
    _lib.check(L.bl_comm_init(C.cast(ident, C.c_void_p), rank, world))
