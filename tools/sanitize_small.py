#!/usr/bin/env python3
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck), one tool per call
(tools/_sanitize.sh; tests/test_gpu_parity.py::test_binned_path_is_deterministic is the race check that
runs everywhere): the regime-binned
rpg_hybrid path (set-up kernels, regrouping loop kernels, side stream), PG(1,z) with class binning
off and on, and a short logit Gibbs chain with both beta draws.

    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayeslogit_b200 import api, gibbs_api  # noqa: E402

rng = np.random.default_rng(0)
num = int(os.environ.get("BL_SANITIZE_NUM", "120000"))
z = rng.uniform(-8, 8, num)
h = np.where(rng.random(num) < 0.5, rng.uniform(0.5, 200, num), rng.integers(1, 201, num).astype(float))
x = api.rpg_seeded("hybrid", h, z, seed=1)
print("hybrid", float(x.mean()))
n1 = np.ones(num, dtype=np.int32)
print("devroye", float(api.rpg_seeded("devroye", n1, z, seed=2).mean()))
N, P = 4000, 16
X = np.c_[rng.standard_normal((N, P - 1)), np.ones(N)]
y = (rng.random(N) < 0.5).astype(float)
for flags in (gibbs_api.TWO_PASS, gibbs_api.PLAIN_BETA | gibbs_api.TWO_PASS, gibbs_api.PLAIN_BETA | gibbs_api.ONE_PASS):
    w, b = gibbs_api.logit_gibbs(y, X, np.ones(N), np.zeros(P), np.eye(P), 3, 2, seed=5, flags=flags)
    print("gibbs", flags, float(b.sum()))
# duplicate-row merge on the device
Xd = np.repeat(X[:500], 4, axis=0)
out = gibbs_api.logit_combine(np.tile(y[:500], 4) * 0 + 0.5, Xd, np.ones(2000))
print("combine", out["X"].shape)
