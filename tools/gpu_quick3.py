import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayeslogit_b200 import api, _lib
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
def run(fn, num, zt, nt, call):
    xt = torch.empty(num, device="cuda", dtype=torch.float64)
    fn(xt.data_ptr(), nt.data_ptr(), zt.data_ptr(), num, 7, call, 0, st)
    torch.cuda.synchronize()
    return xt
num = 1 << 22
zt = (torch.rand(num, device="cuda", dtype=torch.float64) * 2 - 1) * 5
nt = torch.ones(num, device="cuda", dtype=torch.int32)
a = run(L.bl_rpg_devroye_dev, num, zt, nt, 1)
b = run(L.bl_rpg_devroye_plain_dev, num, zt, nt, 1)
c = run(L.bl_rpg_devroye_loop_dev, num, zt, nt, 1)
c2 = run(L.bl_rpg_devroye_loop_dev, num, zt, nt, 1)
d = (c != b)
print("ndiff c-b", int(d.sum()), "c==c2", bool(torch.equal(c, c2)))
idx = torch.nonzero(d)[:10, 0]
print(idx.tolist())
print("c", c[idx].tolist()); print("b", b[idx].tolist()); print("z", zt[idx].tolist())
rel = ((c - b).abs() / b)
print("max rel", rel.max().item(), "n rel>1e-12", int((rel > 1e-12).sum()))
