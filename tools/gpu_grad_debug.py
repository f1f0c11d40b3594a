import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tne_b200 as T
from oracle import graphs as OG, peps as OP, yao as OY, tebd as OT, timeevolve as OTE, tensorad as OAD, einsum as OE
from helpers import *
TA = T.tensorad
g = OG.test_graph()
h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3); hd = to_device_hamiltonian(T, h)
op1 = oracle_peps(g, 2, 11, Dmax=4); dp1 = device_peps_like(T, op1, g, 2, Dmax=4)
v = OP.variables(op1)
def f1o(x): return OAD.real(OP.inner_product(OP.load_variables(op1, x), OP.load_variables(op1, x)))
def f1d(x): return TA.real(T.inner_product(T.load_variables(dp1, x), T.load_variables(dp1, x)))
def f2o(x): return OAD.real(OP.expect(h, OP.conj(OP.load_variables(op1, x)), OTE.wrap_difftensor(op1)))
def f2d(x): return TA.real(T.expect(hd, T.conj(T.load_variables(dp1, x)), T.wrap_difftensor(dp1)))
def f3o(x): return OAD.real(OP.expect(h, OP.load_variables(op1, x), OTE.wrap_difftensor(op1)))
def f3d(x): return TA.real(T.expect(hd, T.load_variables(dp1, x), T.wrap_difftensor(dp1)))
def f4o(x):
    pl = OP.load_variables(op1, x); return OP.norm(pl)
def f4d(x):
    pl = T.load_variables(dp1, x); return T.norm(pl)
for name, fo, fd in [("ip", f1o, f1d), ("expect conj", f2o, f2d), ("expect", f3o, f3d), ("norm", f4o, f4d)]:
    gref = OAD.gradient(fo, OAD.DiffTensor(v.copy()))[0].data
    for fused in [True, False]:
        T.einsum.FUSED_TREES = fused
        gd = TA.gradient(fd, T.DiffTensor(v.copy()))[0].to_numpy()
        print(name, "fused" if fused else "unfused", rel(gd, gref), np.abs(gref).max())
