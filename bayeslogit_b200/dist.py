"""One process per GPU: observation sharding and the engine's NCCL communicator.

Observations are row-sharded: rank r of W holds the contiguous block
[shard_range(r, W, N)) of the global data set and draws its omega shard from Philox streams
keyed by the GLOBAL observation index, so the chain does not depend on W.  The only exchange
per beta draw is one all-reduce of P*P + P doubles (NCCL over NVLink/NVSwitch, issued inside
the engine on its own stream); torch.distributed is used only to hand the 128-byte NCCL id
from rank 0 to the others.
"""
import ctypes as C
import os
import sys


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
    if os.environ.get("BL_PEER_EXCHANGE", "1") != "0" and world <= 8:
        _open_peer_windows(L, rank, world, device)


def init_comm_local(rank, world, device=None):
    """Engine communicator WITHOUT NCCL (one process per GPU, one NVLink node): every exchange, the set-up
    sums included, goes through the CUDA-IPC peer windows.  The window handles travel over whatever
    torch.distributed backend is up."""
    from . import _lib
    L = _lib.lib()
    _lib.check(L.bl_comm_init_local(rank, world))
    if world <= 1:
        return
    _open_peer_windows(L, rank, world, device)
    if not L.bl_comm_peer_active():
        raise _lib.EngineError("bl_comm_init_local: the peer windows could not be mapped (CUDA IPC)")


def _open_peer_windows(L, rank, world, device):
    """All-gather the ranks' CUDA IPC window handles and map them (one node, NVLink): the sharded
    sweeps then exchange their P*P + P sums inside the Gram-reduce / beta-draw kernels instead of
    calling ncclAllReduce.  If any rank cannot map its peers, every rank stays on NCCL."""
    import torch
    import torch.distributed as dist
    on_dev = device is not None and dist.get_backend() == "nccl"
    buf = (C.c_char * 64)()
    ok = L.bl_comm_peer_handle(C.cast(buf, C.c_void_p)) == 0
    mine = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if on_dev:
        mine = mine.to(device)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    raw = b"".join(bytes(g.cpu().numpy().tobytes()) for g in gathered)
    handles = (C.c_char * (64 * world)).from_buffer_copy(raw)
    ok = ok and L.bl_comm_peer_open(C.cast(handles, C.c_void_p)) == 0
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32)
    if on_dev:
        flag = flag.to(device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        L.bl_comm_peer_close()
        L.bl_clear_error()
        if rank == 0:
            print("bayeslogit_b200: peer windows unavailable (CUDA IPC / P2P); using ncclAllReduce", file=sys.stderr)


def peer_exchange_active():
    from . import _lib
    return bool(_lib.lib().bl_comm_peer_active())


def destroy_comm():
    """Barrier, then release the windows and the communicator (every rank)."""
    import torch.distributed as dist
    from . import _lib
    if dist.is_initialized():
        dist.barrier()
    _lib.lib().bl_comm_destroy()
