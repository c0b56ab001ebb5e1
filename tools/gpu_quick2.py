"""Dev aid: fast (fp32-filtered) vs plain (all-fp64) Devroye path -- equality and timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bayeslogit_b200 import api, _lib
from oracle.loader import Oracle, make_tape
L = _lib.lib()
O = Oracle("reference")
st = torch.cuda.current_stream().cuda_stream
def run(fn, num, zt, nt, call):
    xt = torch.empty(num, device="cuda", dtype=torch.float64)
    fn(xt.data_ptr(), nt.data_ptr(), zt.data_ptr(), num, 7, call, 0, st)
    torch.cuda.synchronize()
    return xt
for zr, label in ((5.0, "z~U(-5,5)"), (50.0, "z~U(-50,50)"), (0.5, "z~U(-.5,.5)")):
    num = 1 << 26
    zt = (torch.rand(num, device="cuda", dtype=torch.float64) * 2 - 1) * zr
    nt = torch.ones(num, device="cuda", dtype=torch.int32)
    a = run(L.bl_rpg_devroye_dev, num, zt, nt, 1)
    b = run(L.bl_rpg_devroye_plain_dev, num, zt, nt, 1)
    c = run(L.bl_rpg_devroye_loop_dev, num, zt, nt, 1)
    print(label, "refill==plain:", bool(torch.equal(a, b)), "ndiff", int((a != b).sum()), "loop==plain", bool(torch.equal(c, b)), int((c != b).sum()), "mean", a.mean().item())
# tape parity of the fast path vs oracle
M = 200000
rng = np.random.default_rng(1)
z = rng.uniform(-50, 50, M); z[:100000] = rng.uniform(-5, 5, 100000)
tape = make_tape(M, lu=24, le=24, ln=8, seed=3)
n1 = np.ones(M, dtype=np.int32)
xa, ta = api.rpg_tape("devroye", n1, z, tape)
xb, tb = O.rpg_devroye(n1, z, tape=tape, trace=True, nthreads=8)
ok = tb[:, 4] == 0
print("tape trace equal:", (ta == tb).all(axis=1).mean(), "max rel", np.max(np.abs(xa[ok] - xb[ok]) / xb[ok]))
num = 1 << 27
zt = (torch.rand(num, device="cuda", dtype=torch.float64) * 10 - 5)
nt = torch.ones(num, device="cuda", dtype=torch.int32)
xt = torch.empty(num, device="cuda", dtype=torch.float64)
for name, fn in (("refill", L.bl_rpg_devroye_dev), ("loop", L.bl_rpg_devroye_loop_dev), ("plain", L.bl_rpg_devroye_plain_dev)):
    for _ in range(2): fn(xt.data_ptr(), nt.data_ptr(), zt.data_ptr(), num, 1, 0, 0, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(5): fn(xt.data_ptr(), nt.data_ptr(), zt.data_ptr(), num, 1, r, 0, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.3f} ms  {num/ms*1e3/1e9:.2f} Gdraws/s")

# philox parity vs oracle, mixed n
rng = np.random.default_rng(2)
num = 2000000
z = rng.uniform(-6, 6, num); n = rng.integers(0, 5, num).astype(np.int32)
got = api.rpg_seeded("devroye", n, z, seed=99, call_id=4, obs0=(1 << 32) - 500)
want = O.rpg_devroye(n, z, seed=99, call_id=4, obs0=(1 << 32) - 500, nthreads=16)
rel = np.abs(got - want) / np.maximum(want, 1e-300)
print("refill vs oracle: max rel", rel[want > 0].max(), "zeros ok", bool(np.all(got[n == 0] == 0)), "bit equal frac", (got == want).mean())
