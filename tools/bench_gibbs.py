#!/usr/bin/env python3
"""Logit Gibbs iterations/s at BASELINE config C3 (N=1M, P=64), 1..8 GPUs, strong scaling.

    python tools/bench_gibbs.py [--N 1000000 --P 64 --iters 200 --constrained]
    torchrun --nproc-per-node 2 ... tools/bench_gibbs.py

Data: X[:, :P-1] ~ N(0,1), intercept last, beta_true[j] = |N(0,.25^2)|,
beta_true[P-1] = -0.5, y ~ Bernoulli(sigmoid(X beta)), n = 1, m0 = 0, P0 = 0.01 I, seed 20240003.
Each rank generates only its own shard, keyed by the global row index.
Prints one JSON line (rank 0).  Used by bench.py for its `gibbs` extra.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_shard(N, P, lo, hi, device):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(20240003)
    bt = torch.randn(P, generator=g, device=device, dtype=torch.float64).abs() * 0.25
    bt[P - 1] = -0.5
    # rows are generated in blocks of 4096 keyed by the block index so shards agree with the full set
    blk = 4096
    rows = []
    ys = []
    for b0 in range((lo // blk) * blk, hi, blk):
        gb = torch.Generator(device=device)
        gb.manual_seed(20240003 * 1000003 + b0)
        Xb = torch.randn(blk, P, generator=gb, device=device, dtype=torch.float64)
        Xb[:, P - 1] = 1.0
        ub = torch.rand(blk, generator=gb, device=device, dtype=torch.float64)
        yb = (ub < torch.sigmoid(Xb @ bt)).double()
        s, e = max(lo, b0) - b0, min(hi, b0 + blk) - b0
        rows.append(Xb[s:e]); ys.append(yb[s:e])
    return torch.cat(rows).contiguous(), torch.cat(ys).contiguous(), bt


def run(N, P, iters, warm, constrained, rank, world, local, verify=False, unfused=False, one_pass=False, two_pass=False):
    import torch
    import torch.distributed as dist
    from bayeslogit_b200 import _lib, dist as bdist
    dev = torch.device("cuda", local)
    L = _lib.lib()
    lo, hi = bdist.shard_range(rank, world, N)
    X, y, bt = make_shard(N, P, lo, hi, dev)
    n = torch.ones(hi - lo, device=dev, dtype=torch.float64)
    m0 = torch.zeros(P, device=dev, dtype=torch.float64)
    P0 = (0.01 * torch.eye(P, device=dev, dtype=torch.float64)).contiguous()
    flags = 2 | (0 if constrained else 1) | (4 if unfused else 0) | (8 if one_pass else 0) | (16 if two_pass else 0)   # NO_W | PLAIN_BETA | UNFUSED | ONE_PASS | TWO_PASS
    st = torch.cuda.current_stream().cuda_stream

    def chain(k, seed):
        beta = torch.zeros(k, P, device=dev, dtype=torch.float64)
        rc = L.bl_logit_gibbs_dev(None, beta.data_ptr(), y.data_ptr(), X.data_ptr(), n.data_ptr(),
                                  m0.data_ptr(), P0.data_ptr(), hi - lo, P, k, 0, seed, flags, lo, st)
        if rc:
            _lib.check(rc)
        return beta

    chain(warm, 1)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = L.bl_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    beta = chain(iters, 20240003)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = L.bl_kernel_launches() - l0
    post = beta[iters // 2:].mean(0)
    return {"iters_per_sec": iters / (ms * 1e-3), "ms_per_iter": ms / iters, "iters": iters, "N": N, "P": P,
            "n_gpus": world, "psi_and_draw": "two kernels" if unfused
            else "psi + omega + Gram from one TMA-staged read of X (k_logit_sweep)" if (one_pass or (not two_pass and P % 2 == 0 and P <= 64 and hi - lo <= (1 << 18)))
            else "one pass over X (k_logit_psi_draw), Gram in a second",
            "exchange": ("peer windows (fused in the Gram-reduce / beta-draw kernels)"
                                          if bdist.peer_exchange_active() else "ncclAllReduce") if world > 1 else None,
            "beta_draw": "constrained (reference, Logit.hpp:322-400)" if constrained
            else "plain (Logit.hpp:291-320)", "launches_per_iter": launches / iters,
            "max_abs_err_vs_truth": float((post - bt).abs().max().item()),
            "beta_checksum": float(beta.sum().item())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=1_000_000)
    ap.add_argument("--P", type=int, default=64)
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--constrained", action="store_true")
    ap.add_argument("--unfused", action="store_true", help="psi = X beta and the omega draw as two kernels (A/B)")
    ap.add_argument("--one-pass", action="store_true", help="the one-pass sweep kernel (k_logit_sweep)")
    ap.add_argument("--two-pass", action="store_true", help="never the one-pass sweep kernel")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from bayeslogit_b200 import _lib, dist as bdist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().bl_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        bdist.init_comm(rank, world, torch.device("cuda", local))
    out = run(a.N, a.P, a.iters, a.warmup, a.constrained, rank, world, local, unfused=a.unfused, one_pass=a.one_pass, two_pass=a.two_pass)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        bdist.destroy_comm()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
