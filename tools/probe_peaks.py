#!/usr/bin/env python3
"""Print the pipe-throughput microbenchmarks of the current device (bl_probe_peaks, bl_probe_dmma_scaling)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayeslogit_b200 import _lib  # noqa: E402

L = _lib.lib()
_lib.check(L.bl_set_device(0))
names = ["fp64_fma_tflops", "fp32_fma_tflops", "mufu_gops", "dmma_tflops", "imad_wide_gops", "issue_ginst"]
for _ in range(2):
    b = (C.c_double * 6)()
    _lib.check(L.bl_probe_peaks(C.cast(b, C.c_void_p)))
    print({n: round(v, 2) for n, v in zip(names, b)})
d = (C.c_double * 16)()
L.bl_probe_dmma_scaling.argtypes = [C.c_void_p]
_lib.check(L.bl_probe_dmma_scaling(C.cast(d, C.c_void_p)))
print("dmma TFLOP/s at 1, 2, 4, 8 warps per scheduler (8 accumulator pairs per warp):", [round(v, 2) for v in d[:4]])
print("  ... with A/B operands that change from MMA to MMA:", [round(v, 2) for v in d[4:8]])
print("  ... and a DMUL forming the A operand in front of every 4 MMAs:", [round(v, 2) for v in d[8:12]])
print("  ... in front of every 8 MMAs:", [round(v, 2) for v in d[12:16]])
