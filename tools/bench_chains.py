#!/usr/bin/env python3
"""Independent logit chains, BASELINE config 5b: 4096 chains of N = 10k, P = 32 (per GPU block of
the chains when run under torchrun: chains are block-distributed, no communication).

    python tools/bench_chains.py [--chains 4096 --N 10000 --P 32 --iters 20 --constrained]

Prints one JSON line: chain-iterations/s through the batched entry point (bl_logit_chains_dev), and
for comparison the same chains advanced one after the other through bl_logit_gibbs_dev (a sample).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(chains, N, P, iters, constrained, rank, world, local, serial_sample=16, unfused=False):
    """Chains are block-distributed over the ranks (no communication); returns this rank's figures."""
    import torch
    from bayeslogit_b200 import _lib
    dev = torch.device("cuda", local)
    L = _lib.lib()
    C = chains // world
    c0 = rank * C
    g = torch.Generator(device=dev); g.manual_seed(20240006 + c0)
    X = torch.randn(C, N, P, generator=g, device=dev, dtype=torch.float64)
    X[:, :, P - 1] = 1.0
    bt = torch.randn(C, P, generator=g, device=dev, dtype=torch.float64).abs() * 0.25
    bt[:, P - 1] = -0.5
    y = (torch.rand(C, N, generator=g, device=dev, dtype=torch.float64) < torch.sigmoid(torch.einsum("cnp,cp->cn", X, bt))).double()
    n = torch.ones(C, N, device=dev, dtype=torch.float64)
    m0 = torch.zeros(P, device=dev, dtype=torch.float64)
    P0 = (0.01 * torch.eye(P, device=dev, dtype=torch.float64)).contiguous()
    flags = (0 if constrained else 1) | (4 if unfused else 0)
    st = torch.cuda.current_stream().cuda_stream

    def batch(k):
        beta = torch.zeros(C, k, P, device=dev, dtype=torch.float64)
        rc = L.bl_logit_chains_dev(beta.data_ptr(), y.data_ptr(), X.data_ptr(), n.data_ptr(), m0.data_ptr(),
                                   P0.data_ptr(), C, N, P, k, 0, 20240006 + c0, flags, st)
        if rc:
            _lib.check(rc)
        return beta

    batch(3)
    torch.cuda.synchronize()
    l0 = L.bl_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); beta = batch(iters); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = L.bl_kernel_launches() - l0
    # one after the other (the reference's usage: one gibbs() call per chain).  Only without a
    # communicator: with one open, bl_logit_gibbs_dev treats the ranks' rows as shards of ONE chain
    # and exchanges the Gram sums, which is not what independent chains are.
    S = min(serial_sample, C) if world == 1 else 0
    b1 = None
    e0.record()
    for c in range(S):
        b1 = torch.zeros(iters, P, device=dev, dtype=torch.float64)
        rc = L.bl_logit_gibbs_dev(None, b1.data_ptr(), y[c].data_ptr(), X[c].data_ptr(), n[c].data_ptr(),
                                  m0.data_ptr(), P0.data_ptr(), N, P, iters, 0, 20240006 + c0 + c, flags | 2, 0, st)
        if rc:
            _lib.check(rc)
    e1.record(); torch.cuda.synchronize()
    ms1 = e0.elapsed_time(e1)
    same = float((b1 - beta[S - 1]).abs().max().item()) if S else None
    post = beta[:, iters // 2:].mean(1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"chain_iters_per_sec": C * world * iters / (ms * 1e-3), "ms_per_iteration_of_all_chains": ms / iters,
            "chains": C * world, "chains_per_gpu": C, "N": N, "P": P, "iters": iters,
            "launches_per_iteration": launches / iters,
            "beta_draw": "constrained" if constrained else "plain",
            "psi_and_draw": "two kernels" if unfused else "one pass over X",
            "one_after_the_other_chain_iters_per_sec_one_gpu": S * iters / (ms1 * 1e-3) if S else None,
            "serial_sample_chains": S,
            "max_abs_diff_batched_vs_single_entry": same,
            "rms_err_vs_truth": float((post - bt).pow(2).mean().sqrt().item()), "n_gpus": world,
            "x_bytes_per_gpu": C * N * P * 8}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=4096)
    ap.add_argument("--N", type=int, default=10_000)
    ap.add_argument("--P", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--constrained", action="store_true")
    ap.add_argument("--serial-sample", type=int, default=16)
    ap.add_argument("--unfused", action="store_true")
    a = ap.parse_args()
    import torch
    from bayeslogit_b200 import _lib
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _lib.check(_lib.lib().bl_set_device(local))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = run(a.chains, a.N, a.P, a.iters, a.constrained, rank, world, local, a.serial_sample, a.unfused)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
