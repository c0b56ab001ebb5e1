"""Profiling target: a few launches of the PG(1,z) hot kernel, z~U(-5,5)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayeslogit_b200 import _lib
L = _lib.lib()
num = 1 << 25
zt = torch.rand(num, device="cuda", dtype=torch.float64) * 10 - 5
nt = torch.ones(num, device="cuda", dtype=torch.int32)
xt = torch.empty(num, device="cuda", dtype=torch.float64)
st = torch.cuda.current_stream().cuda_stream
for r in range(4):
    L.bl_rpg_devroye_dev(xt.data_ptr(), nt.data_ptr(), zt.data_ptr(), num, 1, r, 0, st)
torch.cuda.synchronize()
print("ok", xt.mean().item())
