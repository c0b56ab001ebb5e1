"""Quick GPU sanity run (development aid): parity vs the oracle + first timings."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bayeslogit_b200 import api, _lib
from oracle.loader import Oracle, make_tape

O = Oracle("reference") if os.path.exists("oracle/_ref/libpg_ref.so") else Oracle("port")
print("oracle kind", O.kind)
rng = np.random.default_rng(5)
N = 200000
z = rng.uniform(-5, 5, N)

def cmp(name, a, b):
    ok = np.isfinite(a) & np.isfinite(b)
    rel = np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), 1e-300)
    print(f"{name}: n={len(a)} finite={ok.sum()} max_rel={rel.max():.3e} n_gt_1e-12={(rel>1e-12).sum()} bit_equal={(a==b).sum()}")

print("philox KAT", [hex(v) for v in api.philox4x32_10([0,0,0,0],[0,0])])
# philox-mode parity
n = rng.integers(1, 4, N).astype(np.int32)
cmp("devroye philox", api.rpg_seeded("devroye", n, z, 11), O.rpg_devroye(n, z, seed=11, nthreads=8))
h = rng.uniform(1, 13, N)
cmp("alt philox", api.rpg_seeded("alt", h, z, 12), O.rpg_alt(h, z, seed=12, nthreads=8))
h = rng.uniform(13, 170, N)
a, ia = api.rpg_seeded("sp", h, z, 13); b, ib = O.rpg_sp(h, z, seed=13, nthreads=8)
cmp("sp philox", a, b); print(" iter equal", (ia == ib).mean())
h = rng.uniform(0.05, 1, 20000)
cmp("gamma philox", api.rpg_seeded("gamma", h, z[:20000], 14), O.rpg_gamma(h, z[:20000], seed=14, nthreads=8))
h = np.where(rng.random(N) < 0.5, rng.uniform(0.5, 200, N), rng.integers(1, 201, N).astype(float))
cmp("hybrid philox", api.rpg_seeded("hybrid", h, z, 15), O.rpg_hybrid(h, z, seed=15, nthreads=8))
# tape parity, devroye
M = 100000
tape = make_tape(M, lu=48, le=48, ln=16, seed=3)
n1 = np.ones(M, dtype=np.int32)
xa, ta = api.rpg_tape("devroye", n1, z[:M], tape)
xb, tb = O.rpg_devroye(n1, z[:M], tape=tape, trace=True)
cmp("devroye tape", xa, xb); print(" trace equal", (ta == tb).all(axis=1).mean(), "exhausted", tb[:,4].sum())

# timing device-resident
L = _lib.lib()
for num in (1 << 20, 1 << 24, 1 << 27):
    zt = (torch.rand(num, device="cuda", dtype=torch.float64) * 10 - 5)
    nt = torch.ones(num, device="cuda", dtype=torch.int32)
    xt = torch.empty(num, device="cuda", dtype=torch.float64)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        L.bl_rpg_devroye_dev(xt.data_ptr(), nt.data_ptr(), zt.data_ptr(), num, 1, 0, 0, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for r in range(reps):
        L.bl_rpg_devroye_dev(xt.data_ptr(), nt.data_ptr(), zt.data_ptr(), num, 1, r, 0, st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"devroye dev num={num}: {ms:.3f} ms  {num/ms/1e6:.1f} Mdraws/s mean={xt.mean().item():.5f}")
num = 1 << 24
zt = (torch.rand(num, device="cuda", dtype=torch.float64) * 10 - 5)
xt = torch.empty(num, device="cuda", dtype=torch.float64)
g = torch.Generator(device="cuda"); g.manual_seed(1)
ht = torch.where(torch.rand(num, device="cuda", generator=g) < 0.5,
                 torch.rand(num, device="cuda", dtype=torch.float64, generator=g) * 199.5 + 0.5,
                 torch.randint(1, 201, (num,), device="cuda", generator=g).double())
st = torch.cuda.current_stream().cuda_stream
L.bl_rpg_hybrid_dev(xt.data_ptr(), ht.data_ptr(), zt.data_ptr(), num, 1, 0, 0, st); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); L.bl_rpg_hybrid_dev(xt.data_ptr(), ht.data_ptr(), zt.data_ptr(), num, 1, 1, 0, st); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"hybrid dev num={num}: {ms:.3f} ms  {num/ms/1e6:.1f} Mdraws/s")
for lo, hi, name in ((1.0001, 13, "alt"), (13.001, 170, "sp")):
    ht = torch.rand(num, device="cuda", dtype=torch.float64) * (hi - lo) + lo
    fn = getattr(L, f"bl_rpg_{name}_dev")
    args = [xt.data_ptr(), ht.data_ptr(), zt.data_ptr(), num] + ([None] if name == "sp" else []) + [1, 0, 0, st]
    fn(*args); torch.cuda.synchronize()
    e0.record(); fn(*args); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name} dev num={num}: {ms:.3f} ms  {num/ms/1e6:.1f} Mdraws/s")
