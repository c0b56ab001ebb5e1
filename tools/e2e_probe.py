#!/usr/bin/env python3
"""End-to-end rate of rpg_hybrid through the host-pointer C ABI (pinned buffers) for one setting of the
pipeline chunk size (BAYESLOGIT_PIPE_CHUNK_LOG2); one process per setting.

    for c in 22 23 24; do BAYESLOGIT_PIPE_CHUNK_LOG2=$c python tools/e2e_probe.py; done
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from bayeslogit_b200 import _lib
import bench

num = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
L = _lib.lib()
_lib.check(L.bl_set_device(0))
shape_h, z_h = bench.make_inputs_numpy("hybrid", num)
shape_p, z_p = torch.from_numpy(shape_h).pin_memory(), torch.from_numpy(z_h).pin_memory()
x_p = torch.empty(num, dtype=torch.float64).pin_memory()
for w in range(2):
    _lib.check(L.bl_rpg_hybrid_seeded(x_p.data_ptr(), shape_p.data_ptr(), z_p.data_ptr(), num, 1, 100 + w, 0))
torch.cuda.synchronize()
ts = []
for k in range(5):
    t0 = time.perf_counter()
    _lib.check(L.bl_rpg_hybrid_seeded(x_p.data_ptr(), shape_p.data_ptr(), z_p.data_ptr(), num, 1, k, 0))
    ts.append(time.perf_counter() - t0)
print("chunk_log2=%s  e2e draws/s: best %.4g median %.4g  (ms: %s)" % (
    os.environ.get("BAYESLOGIT_PIPE_CHUNK_LOG2", "default"), num / min(ts), num / sorted(ts)[2],
    " ".join("%.1f" % (1e3 * t) for t in ts)))
