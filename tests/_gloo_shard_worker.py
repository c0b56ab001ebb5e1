"""Worker for tests/test_gibbs_oracle.py::test_sharded_gram_allreduce_gloo (run under torchrun)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayeslogit_b200.dist import shard_range  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
rng = np.random.default_rng(0)
N, P = 1001, 5
X = rng.standard_normal((N, P)); w = rng.random(N); kappa = rng.standard_normal(N)
lo, hi = shard_range(rank, world, N)
sizes = [shard_range(r, world, N) for r in range(world)]
assert sizes[0][0] == 0 and sizes[-1][1] == N and all(a[1] == b[0] for a, b in zip(sizes, sizes[1:]))
acc = np.concatenate([(X[lo:hi].T * w[lo:hi]) @ X[lo:hi], X[lo:hi].T @ kappa[lo:hi]], axis=None)
t = torch.from_numpy(acc.copy())
dist.all_reduce(t)                       # the P*P + P block of SURVEY.md section 8e
full = np.concatenate([(X.T * w) @ X, X.T @ kappa], axis=None)
assert np.allclose(t.numpy(), full, rtol=1e-12, atol=1e-12), rank
dist.barrier()
if rank == 0:
    print("SHARD_OK")
dist.destroy_process_group()
