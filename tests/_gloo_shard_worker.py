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

# One sharded Gibbs half-step with the CPU oracle as the sampler: rank r draws
# omega for its rows from the streams of the GLOBAL observation index (obs0 = lo), the ranks
# all-reduce the P*P + P sums, and the replicated posterior mean then equals the single-rank one.
from oracle import loader  # noqa: E402
O = loader.Oracle("reference" if loader.available("reference") else "port")
beta = rng.normal(0, 0.3, P)
psi = X @ beta
n1 = np.ones(N, dtype=np.int32)
w_full = O.rpg_devroye(n1, psi, seed=77, obs0=0, call_id=3)
w_mine = O.rpg_devroye(n1[lo:hi], psi[lo:hi], seed=77, obs0=lo, call_id=3)
assert np.array_equal(w_mine, w_full[lo:hi]), "omega shard depends on the sharding"
yk = (rng.random(N) < 0.5) - 0.5
acc = np.concatenate([(X[lo:hi].T * w_mine) @ X[lo:hi], X[lo:hi].T @ yk[lo:hi]], axis=None)
t = torch.from_numpy(acc.copy())
dist.all_reduce(t)
PP = t.numpy()[:P * P].reshape(P, P) + 0.01 * np.eye(P)
m_sharded = np.linalg.solve(PP, t.numpy()[P * P:])
m_single = np.linalg.solve((X.T * w_full) @ X + 0.01 * np.eye(P), X.T @ yk)
assert np.allclose(m_sharded, m_single, rtol=1e-10, atol=1e-12), rank
# every rank must hold the same bits after the all-reduce (replicated beta draw needs no broadcast)
g = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(g, t)
assert all(torch.equal(g[0], x) for x in g), "all-reduced sums differ between ranks"

# Independent chains (config 5b) are block-distributed: rank r owns chains [r*C, (r+1)*C) and calls the
# batched entry with seed + r*C, so chain c always draws from seed + c whatever the world size.
chains, seed = 10, 20240006
C_ = chains // world
mine = [seed + rank * C_ + c for c in range(C_)]
allseeds = [None] * world
dist.all_gather_object(allseeds, mine)
assert sum(allseeds, []) == [seed + c for c in range(chains)]
dist.barrier()
if rank == 0:
    print("SHARD_OK")
dist.destroy_process_group()
