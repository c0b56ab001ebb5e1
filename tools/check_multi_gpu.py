#!/usr/bin/env python3
"""Multi-GPU check (run under torchrun, one rank per GPU): the sharded logit chain with the
in-stream NCCL all-reduce equals the single-GPU chain, and the sampler's shards equal the
unsharded batch.  Prints MULTI_GPU_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayeslogit_b200 import _lib, dist as bdist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
L = _lib.lib()
_lib.check(L.bl_set_device(local))
dist.init_process_group("nccl", device_id=dev)

rng = np.random.default_rng(0)
N, P, samp, burn = 200_003, 16, 12, 4
X = np.c_[rng.standard_normal((N, P - 1)), np.ones(N)]
bt = np.r_[np.abs(rng.normal(0, 0.3, P - 1)), -0.5]
y = (rng.random(N) < 1 / (1 + np.exp(-X @ bt))).astype(float)
st = torch.cuda.current_stream().cuda_stream


def chain(lo, hi, flags):
    Xd = torch.from_numpy(X[lo:hi].copy()).to(dev); yd = torch.from_numpy(y[lo:hi].copy()).to(dev)
    nd = torch.ones(hi - lo, device=dev, dtype=torch.float64)
    m0 = torch.zeros(P, device=dev, dtype=torch.float64)
    P0 = (0.1 * torch.eye(P, device=dev, dtype=torch.float64)).contiguous()
    beta = torch.zeros(samp, P, device=dev, dtype=torch.float64)
    w = torch.zeros(samp, hi - lo, device=dev, dtype=torch.float64)
    rc = L.bl_logit_gibbs_dev(w.data_ptr(), beta.data_ptr(), yd.data_ptr(), Xd.data_ptr(), nd.data_ptr(),
                              m0.data_ptr(), P0.data_ptr(), hi - lo, P, samp, burn, 4242, flags, lo, st)
    if rc:
        _lib.check(rc)
    torch.cuda.synchronize()
    return beta.cpu().numpy(), w.cpu().numpy()


full = {f: chain(0, N, f) for f in (0, 1)}          # before the communicator exists: single-GPU chains
bdist.init_comm(rank, world, dev)
lo, hi = bdist.shard_range(rank, world, N)
ok = True
for f in (0, 1):
    b, w = chain(lo, hi, f)
    eb = np.max(np.abs(b - full[f][0]) / np.abs(full[f][0]))
    ew = np.max(np.abs(w - full[f][1][:, lo:hi]) / full[f][1][:, lo:hi])
    t = torch.tensor([eb, ew], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"flags={f}: world={world} max rel diff beta {t[0].item():.2e} omega {t[1].item():.2e}")
    ok = ok and t.max().item() < 1e-8


# NB sweep (fixed d): exercises the P extra sums (X'v tail) of the exchange
yc = rng.poisson(np.exp(np.clip(X @ bt * 0.3 + 2.0, None, 4.0))).astype(float)


def nb_chain(lo, hi):
    Xd = torch.from_numpy(X[lo:hi].copy()).to(dev); yd = torch.from_numpy(yc[lo:hi].copy()).to(dev)
    m0 = torch.zeros(P, device=dev, dtype=torch.float64)
    P0 = (0.1 * torch.eye(P, device=dev, dtype=torch.float64)).contiguous()
    beta = torch.zeros(8, P, device=dev, dtype=torch.float64)
    w = torch.zeros(hi - lo, device=dev, dtype=torch.float64)
    rc = L.bl_nb_gibbs_dev(w.data_ptr(), beta.data_ptr(), yd.data_ptr(), Xd.data_ptr(), 5.0,
                           m0.data_ptr(), P0.data_ptr(), hi - lo, P, 8, 777, lo, st)
    if rc:
        _lib.check(rc)
    torch.cuda.synchronize()
    return beta.cpu().numpy(), w.cpu().numpy()


def compare_nb(tag, ref):
    b, w = nb_chain(lo, hi)
    eb = np.max(np.abs(b - ref[0]) / np.abs(ref[0]))
    ew = np.max(np.abs(w - ref[1][lo:hi]) / ref[1][lo:hi])
    t = torch.tensor([eb, ew], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"nb {tag}: max rel diff beta {t[0].item():.2e} omega {t[1].item():.2e}")
    return t.max().item() < 1e-8, b


def nb_df_chain(lo, hi):
    """NB sweep with the dispersion sampled on the device (draw.df): sharded, ymax / the count histogram /
    the log-likelihood sums of every Metropolis step are all-reduced."""
    Xd = torch.from_numpy(X[lo:hi].copy()).to(dev); yd = torch.from_numpy(yc[lo:hi].copy()).to(dev)
    m0 = torch.zeros(P, device=dev, dtype=torch.float64)
    P0 = (0.1 * torch.eye(P, device=dev, dtype=torch.float64)).contiguous()
    beta = torch.zeros(10, P, device=dev, dtype=torch.float64)
    dd = torch.zeros(10, device=dev, dtype=torch.float64)
    rc = L.bl_nb_gibbs_df_dev(None, beta.data_ptr(), dd.data_ptr(), yd.data_ptr(), Xd.data_ptr(), 1.0,
                              m0.data_ptr(), P0.data_ptr(), hi - lo, P, 10, 6, 4711, lo, st)
    if rc:
        _lib.check(rc)
    torch.cuda.synchronize()
    return beta.cpu().numpy(), dd.cpu().numpy()


peer = bdist.peer_exchange_active()
bdist.destroy_comm()
nb_full = nb_chain(0, N)                      # no communicator: single-GPU chain
nbdf_full = nb_df_chain(0, N)
bdist.init_comm(rank, world, dev)
b_df, d_df = nb_df_chain(lo, hi)
e_df = float(np.max(np.abs(b_df - nbdf_full[0]) / np.abs(nbdf_full[0])))
same_d = bool(np.array_equal(d_df, nbdf_full[1]))
t = torch.tensor([e_df, 0.0 if same_d else 1.0], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"nb with dispersion update: max rel diff beta {t[0].item():.2e}; d chain identical on every rank: {t[1].item() == 0.0} "
          f"(d: {d_df[0]:.0f} .. {d_df[-1]:.0f})")
ok = ok and t[0].item() < 1e-8 and t[1].item() == 0.0
ok_nb, b_peer = compare_nb("peer windows" if bdist.peer_exchange_active() else "nccl", nb_full)
ok = ok and ok_nb
# beta must be bit-identical on every rank (replicated draw from identical sums)
g = [torch.empty(b_peer.shape, device=dev, dtype=torch.float64) for _ in range(world)]
dist.all_gather(g, torch.from_numpy(b_peer).to(dev))
same = all(torch.equal(g[0], x) for x in g)
if rank == 0:
    print(f"beta bit-identical across ranks: {same}; peer exchange active: {peer}")
ok = ok and same
if peer:
    # the NCCL path on the same shards
    bdist.destroy_comm()
    os.environ["BL_PEER_EXCHANGE"] = "0"
    bdist.init_comm(rank, world, dev)
    assert not bdist.peer_exchange_active()
    ok_nccl, _ = compare_nb("nccl", nb_full)
    b, w = chain(lo, hi, 1)
    eb = float(np.max(np.abs(b - full[1][0]) / np.abs(full[1][0])))
    if rank == 0:
        print(f"logit nccl: max rel diff beta {eb:.2e}")
    ok = ok and ok_nccl and eb < 1e-8
if rank == 0:
    print("MULTI_GPU_OK" if ok else "MULTI_GPU_MISMATCH")
bdist.destroy_comm()
dist.destroy_process_group()
