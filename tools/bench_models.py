#!/usr/bin/env python3
"""Throughput of the two other sweeps BASELINE.json names: negative-binomial regression (config 4:
N = 10M, P = 256, large counts -> saddle-point-heavy PG(y + d, psi)) and multinomial logit (config
5a: J = 10, N = 1M, P = 32).  Rows are sharded over the ranks (strong scaling, one exchange of
P*P + P sums per beta draw); each rank generates its own shard.

    python tools/bench_models.py [--nb-iters 8 --mlogit-iters 20]
    torchrun --nproc-per-node 2 ... tools/bench_models.py
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _time(fn, dev, world):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, out


def run_nb(N, P, iters, rank, world, local, d=10.0, mean_count=100.0):
    """X[:, :P-1] ~ N(0,1), intercept last, beta scaled so the mean count is ~100
    with d = 10 (b = y + d mostly in (13, 170] with a > 170 tail), y ~ NB(d, mu / (mu + d))."""
    import torch
    from bayeslogit_b200 import _lib, dist as bdist
    dev = torch.device("cuda", local)
    L = _lib.lib()
    lo, hi = bdist.shard_range(rank, world, N)
    n = hi - lo
    g = torch.Generator(device=dev); g.manual_seed(20240004)
    bt = torch.randn(P, generator=g, device=dev, dtype=torch.float64) * 0.03
    sd2 = float((bt[:P - 1] ** 2).sum().item())
    bt[P - 1] = math.log(mean_count) - 0.5 * sd2
    g.manual_seed(20240004 * 1000003 + rank)
    X = torch.empty(n, P, device=dev, dtype=torch.float64)
    blk = 1 << 20
    for r0 in range(0, n, blk):                      # fp64 randn in slices: no 2x temporary of a 20 GB matrix
        X[r0:r0 + blk].normal_(generator=g)
    X[:, P - 1] = 1.0
    mu = torch.exp(X @ bt)
    lam = torch._standard_gamma(torch.full_like(mu, d), generator=g) * (mu / d)
    y = torch.poisson(lam, generator=g)
    del mu, lam
    m0 = torch.zeros(P, device=dev, dtype=torch.float64)
    P0 = (0.01 * torch.eye(P, device=dev, dtype=torch.float64)).contiguous()
    st = torch.cuda.current_stream().cuda_stream
    b = y + d
    shares = {"sp_13_170": float(((b > 13) & (b <= 170)).double().mean().item()),
              "normal_gt_170": float((b > 170).double().mean().item()),
              "alt_le_13": float((b <= 13).double().mean().item())}

    def chain(k):
        beta = torch.zeros(k, P, device=dev, dtype=torch.float64)
        rc = L.bl_nb_gibbs_dev(None, beta.data_ptr(), y.data_ptr(), X.data_ptr(), float(d), m0.data_ptr(),
                               P0.data_ptr(), n, P, k, 20240004, lo, st)
        if rc:
            _lib.check(rc)
        return beta

    chain(2)
    l0 = L.bl_kernel_launches()
    ms, beta = _time(lambda: chain(iters), dev, world)
    launches = L.bl_kernel_launches() - l0
    return {"iters_per_sec": iters / (ms * 1e-3), "ms_per_iter": ms / iters, "iters": iters, "N": N, "P": P, "d": d,
            "n_gpus": world, "mean_count": float(y.mean().item()), "shape_shares": shares,
            "launches_per_iter": launches / iters,
            "gram_tflops": 2.0 * N * P * P / (ms / iters * 1e-3) / 1e12,
            "max_abs_err_vs_truth_last": float((beta[-1] - bt).abs().max().item()),
            "note": "chain from beta = 0 (NBPG-logmean.R:77); the error shrinks as the chain burns in"}


def run_mlogit(N, P, J, iters, rank, world, local):
    """beta ~ N(0, 0.5^2), y one-hot from softmax(X beta) with the last category as baseline."""
    import torch
    from bayeslogit_b200 import _lib, dist as bdist
    dev = torch.device("cuda", local)
    L = _lib.lib()
    lo, hi = bdist.shard_range(rank, world, N)
    n = hi - lo
    U = J - 1
    g = torch.Generator(device=dev); g.manual_seed(20240005)
    B = torch.randn(P, U, generator=g, device=dev, dtype=torch.float64) * 0.5
    g.manual_seed(20240005 * 1000003 + rank)
    X = torch.randn(n, P, generator=g, device=dev, dtype=torch.float64)
    X[:, P - 1] = 1.0
    eta = torch.cat([X @ B, torch.zeros(n, 1, device=dev, dtype=torch.float64)], 1)
    cat = torch.multinomial(torch.softmax(eta, 1), 1, generator=g).squeeze(1)
    ty = torch.nn.functional.one_hot(cat, J)[:, :U].double().contiguous()       # [n][J-1] = (J-1) x n column-major
    nn_ = torch.ones(n, device=dev, dtype=torch.float64)
    m0 = torch.zeros(U, P, device=dev, dtype=torch.float64)
    P0 = (0.01 * torch.eye(P, device=dev, dtype=torch.float64)).repeat(U, 1, 1).contiguous()
    st = torch.cuda.current_stream().cuda_stream

    def chain(k):
        beta = torch.zeros(k, U, P, device=dev, dtype=torch.float64)
        rc = L.bl_mlogit_gibbs_dev(None, beta.data_ptr(), ty.data_ptr(), X.data_ptr(), nn_.data_ptr(), m0.data_ptr(),
                                   P0.data_ptr(), n, P, J, k, 0, 20240005, 2, lo, st)
        if rc:
            _lib.check(rc)
        return beta

    chain(2)
    l0 = L.bl_kernel_launches()
    ms, beta = _time(lambda: chain(iters), dev, world)
    launches = L.bl_kernel_launches() - l0
    post = beta[iters // 2:].mean(0)                                             # [U][P]
    return {"iters_per_sec": iters / (ms * 1e-3), "ms_per_iter": ms / iters, "iters": iters, "N": N, "P": P, "J": J,
            "n_gpus": world, "exchange": ("peer windows" if bdist.peer_exchange_active() else "ncclAllReduce") if world > 1 else None,
            "category_updates_per_sec": iters * U / (ms * 1e-3),
            "launches_per_iter": launches / iters,
            "max_abs_err_vs_truth": float((post - B.t()).abs().max().item())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nb-iters", type=int, default=8)
    ap.add_argument("--nb-N", type=int, default=10_000_000)
    ap.add_argument("--nb-P", type=int, default=256)
    ap.add_argument("--mlogit-iters", type=int, default=20)
    a = ap.parse_args()
    import torch
    from bayeslogit_b200 import _lib, dist as bdist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().bl_set_device(local))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        bdist.init_comm(rank, world, torch.device("cuda", local))
    out = {}
    if a.mlogit_iters > 0:
        out["mlogit_J10_N1M_P32"] = run_mlogit(1_000_000, 32, 10, a.mlogit_iters, rank, world, local)
    if a.nb_iters > 0:
        out["nb_N10M_P256"] = run_nb(a.nb_N, a.nb_P, a.nb_iters, rank, world, local)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        bdist.destroy_comm()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
