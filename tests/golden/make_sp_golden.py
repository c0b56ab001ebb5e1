#!/usr/bin/env python3
"""Golden values of V(y) = y^{-1}(y) and G(y) = log cos_rt(V(y)) for the saddle-point sampler's
table path (bayeslogit_b200/csrc/pg_sp.cuh), 40-digit mpmath, independent of the table generator's
interpolation nodes.  y: log-uniform on [2^-4, 2^4), plus points hugging the reference's grid
2^(-4 + 0.1 i), y = 1 and the interval edges, where the engine must hand over to the reference's
Newton iteration (InvertY.cpp:57-99).

    python tests/golden/make_sp_golden.py   ->  tests/golden/sp_tables_golden.npz
"""
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "tools"))
import gen_sp_tables as g  # noqa: E402  (y_of_v / v_root / g_of_v in 40-digit arithmetic)

rng = np.random.default_rng(20240007)
y = np.exp2(rng.uniform(-4, 4, 3000))
edges = np.array([2.0 ** e * (1 + j / 16) for e in range(-4, 4) for j in range(16)])
y = np.concatenate([y, edges[1:] * (1 - 1e-15), edges * (1 + 1e-15)])
near = []
for i in range(1, 80):
    gy = 2.0 ** (-4 + 0.1 * i)
    near += [gy * (1 + d) for d in (-3e-5, -1e-6, 1e-6, 3e-5)]
near += [1 - 1e-7, 1 + 1e-7, 1 - 3e-6, 1 + 3e-6]
y_near = np.array(near)
out = {}
for name, ys in (("y", y), ("y_near", y_near)):
    v = [g.v_root(mp.mpf(float(t))) for t in ys]
    out[name] = ys
    out[name + "_v"] = np.array([float(t) for t in v])
    out[name + "_g"] = np.array([float(g.g_of_v(t)) for t in v])
np.savez_compressed(os.path.join(HERE, "sp_tables_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
