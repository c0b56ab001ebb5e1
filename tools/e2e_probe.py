#!/usr/bin/env python3
"""End-to-end rate of the host-pointer rpg_hybrid call: pinned buffers, pageable buffers through the library's
staging ring, pageable buffers through the driver's own staging.

    python tools/e2e_probe.py [draws]           BAYESLOGIT_COPY_THREADS / BAYESLOGIT_PIPE_CHUNK_LOG2 apply
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bayeslogit_b200 import _lib  # noqa: E402

num = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
L = _lib.lib()
_lib.check(L.bl_set_device(0))
shape, z = bench.make_inputs_numpy("hybrid", num)
x = np.zeros(num)
sp, zp, xp = torch.from_numpy(shape).pin_memory(), torch.from_numpy(z).pin_memory(), torch.zeros(num, dtype=torch.float64).pin_memory()


def rate(xa, sa, za, reps=3):
    _lib.check(L.bl_rpg_hybrid_seeded(xa, sa, za, num, 1, 99, 0))
    t = time.perf_counter()
    for k in range(reps):
        _lib.check(L.bl_rpg_hybrid_seeded(xa, sa, za, num, 1, k, 0))
    return num * reps / (time.perf_counter() - t)


print("pinned            %.3e draws/s" % rate(xp.data_ptr(), sp.data_ptr(), zp.data_ptr()))
print("pageable, staged  %.3e draws/s  (copy threads: %s)" % (rate(x.ctypes.data, shape.ctypes.data, z.ctypes.data),
                                                            os.environ.get("BAYESLOGIT_COPY_THREADS", "default")))
os.environ["BAYESLOGIT_NO_STAGING"] = "1"
print("pageable, driver  %.3e draws/s" % rate(x.ctypes.data, shape.ctypes.data, z.ctypes.data, reps=2))
os.environ.pop("BAYESLOGIT_NO_STAGING")
_lib.check(L.bl_rpg_hybrid_seeded(x.ctypes.data, shape.ctypes.data, z.ctypes.data, num, 1, 7, 0))
_lib.check(L.bl_rpg_hybrid_seeded(xp.data_ptr(), sp.data_ptr(), zp.data_ptr(), num, 1, 7, 0))
assert np.array_equal(x, xp.numpy()), "staged pageable call differs from the pinned call"
