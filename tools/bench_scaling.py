#!/usr/bin/env python3
"""All multi-GPU figures in one launch (one rank per GPU under torchrun): logit Gibbs N = 1M, P = 64
with the peer-window exchange and with ncclAllReduce, multinomial logit, NB regression (row-sharded,
strong scaling) and the 4096 independent chains (block-distributed).  One JSON line on rank 0.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_scaling.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import torch
    import torch.distributed as dist
    import bench_chains
    import bench_gibbs
    import bench_models
    from bayeslogit_b200 import _lib, dist as bdist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.check(_lib.lib().bl_set_device(local))
    out = {"n_gpus": world}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        bdist.init_comm(rank, world, dev)
    iters = int(os.environ.get("BL_SCALING_ITERS", "300"))
    out["logit_N1M_P64"] = bench_gibbs.run(1_000_000, 64, iters, 5, False, rank, world, local)
    out["mlogit_J10_N1M_P32"] = bench_models.run_mlogit(1_000_000, 32, 10, 20, rank, world, local)
    torch.cuda.empty_cache()
    out["nb_N10M_P256"] = bench_models.run_nb(10_000_000, 256, 8, rank, world, local)
    torch.cuda.empty_cache()
    if world > 1:
        bdist.destroy_comm()
        os.environ["BL_PEER_EXCHANGE"] = "0"
        bdist.init_comm(rank, world, dev)
        out["logit_N1M_P64_nccl"] = bench_gibbs.run(1_000_000, 64, iters, 5, False, rank, world, local)
        out["mlogit_J10_N1M_P32_nccl"] = bench_models.run_mlogit(1_000_000, 32, 10, 20, rank, world, local)
        bdist.destroy_comm()
    torch.cuda.empty_cache()
    out["chains_4096_N10k_P32"] = bench_chains.run(4096, 10_000, 32, 20, False, rank, world, local, serial_sample=2)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
