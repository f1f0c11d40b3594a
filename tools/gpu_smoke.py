"""Quick device-vs-oracle checks (development aid; the real parity tests are tests/test_gpu_*.py)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tne_b200 as T
from oracle import graphs as OG, peps as OP, yao as OY, tebd as OT, timeevolve as OTE, tensorad as OAD, einsum as OE
from helpers import *

ctx = T.context()
if os.environ.get("TNE_DEBUG"): ctx.set_option("debug", 1)
rng = np.random.default_rng(0)
# 1. binary einsum
a = rng.standard_normal((3, 4, 2, 5)) + 1j * rng.standard_normal((3, 4, 2, 5))
b = rng.standard_normal((5, 2, 6, 3)) + 1j * rng.standard_normal((5, 2, 6, 3))
ref = OE.einsum_raw([[1, 2, 3, 4], [4, 3, 5, 1]], [a, b], [5, 2, 1])
out = T.einsum.einsum([[1, 2, 3, 4], [4, 3, 5, 1]], [T.B200Array.from_numpy(a), T.B200Array.from_numpy(b)], [5, 2, 1]).to_numpy()
print("einsum batch", rel(out, ref))
ref = OE.einsum_raw([[1, 2, 3, 4]], [a], [3, 1, 4, 2])
out = T.einsum.einsum([[1, 2, 3, 4]], [T.B200Array.from_numpy(a)], [3, 1, 4, 2]).to_numpy()
print("permute", rel(out, ref))
# 2. 5-site graph: statevec, inner product, expect, fvec
g = OG.test_graph()
for kind in ["simple", "vidal"]:
    op1 = oracle_peps(g, 2, 11, kind, Dmax=4); op2 = oracle_peps(g, 2, 12, kind, Dmax=4)
    dp1 = device_peps_like(T, op1, g, 2, Dmax=4); dp2 = device_peps_like(T, op2, g, 2, Dmax=4)
    print(kind, "statevec", rel(T.vec(dp1), OP.vec(op1)))
    print(kind, "inner", rel(T.inner_product(dp1, dp2).to_numpy(), OP.inner_product(op1, op2)))
    print(kind, "norm", rel(T.norm(dp1).to_numpy(), OP.norm(op1)))
gg = OG.test_graph()
h = OT.rydberg_hamiltonian(gg, 10.0, 0.2, 0.3)
hd = to_device_hamiltonian(T, h)
op1 = oracle_peps(g, 2, 11, Dmax=4); op2 = oracle_peps(g, 2, 12, Dmax=4)
dp1 = device_peps_like(T, op1, g, 2, Dmax=4); dp2 = device_peps_like(T, op2, g, 2, Dmax=4)
print("expect", rel(T.expect(hd, dp1, dp2).to_numpy(), OP.expect(h, op1, op2)))
T.einsum.FUSED_TREES = False
print("expect unfused", rel(T.expect(hd, dp1, dp2).to_numpy(), OP.expect(h, op1, op2)))
T.einsum.FUSED_TREES = True
f_ref = OTE.fvec(op1, h).data
f_dev = T.fvec(dp1, hd).to_numpy()
print("fvec fused", rel(f_dev, f_ref))
T.einsum.FUSED_TREES = False
f_dev = T.fvec(dp1, hd).to_numpy()
print("fvec unfused", rel(f_dev, f_ref))
T.einsum.FUSED_TREES = True
# smatrix
v = OP.variables(op1)
x = (rng.standard_normal(v.size) + 1j * rng.standard_normal(v.size)) / np.sqrt(2)
sx_ref = OTE.apply_smatrix(op1, OAD.DiffTensor(v.copy()), OAD.DiffTensor(v.copy()), OAD.DiffTensor(x, requires_grad=False)).data
sx = T.apply_smatrix(dp1, T.DiffTensor(v.copy()), T.DiffTensor(v.copy()), T.DiffTensor(x, requires_grad=False)).to_numpy()
print("smatrix", rel(sx, sx_ref))
# 3. grids
for (L, D) in [(3, 2), (4, 3)]:
    g = OG.grid([L, L])
    order = sweep_order(L * L)
    op = oracle_peps(g, D, seed_for(L, D), optimizer=order)
    dp = device_peps_like(T, op, g, D, optimizer=order)
    t0 = time.time(); nref = OP.norm(op); t1 = time.time()
    print(L, D, "norm", rel(T.norm(dp).to_numpy(), nref), "cpu s", t1 - t0)
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3); hd = to_device_hamiltonian(T, h)
    t0 = time.time(); eref = OP.expect(h, op, op); t1 = time.time()
    print(L, D, "expect", rel(T.expect(hd, dp, dp).to_numpy(), eref), "cpu s", t1 - t0, ctx.program_stats())
    t0 = time.time(); fref = OTE.fvec(op, h).data; t1 = time.time()
    t2 = time.time(); fd = T.fvec(dp, hd).to_numpy(); t3 = time.time()
    print(L, D, "fvec", rel(fd, fref), "cpu s", t1 - t0, "gpu s", t3 - t2, ctx.program_stats())
    # sliced
    ds = device_peps_like(T, op, g, D, optimizer=order, nslices=2)
    print(L, D, "sliced norm", rel(T.norm(ds).to_numpy(), nref), ds.code_inner_product.slicing)
    fd = T.fvec(ds, hd).to_numpy()
    print(L, D, "sliced fvec", rel(fd, fref))
    ctx.set_option("recompute_threshold_bytes", 4096)
    fd = T.fvec(dp, hd).to_numpy()
    print(L, D, "recompute fvec", rel(fd, fref), ctx.program_stats())
    ctx.set_option("recompute_threshold_bytes", -1)
print("launches", ctx.launch_count())
