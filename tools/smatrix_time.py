"""Times the metric-vector product of the 6x6 D=4 PEPS with and without the sr_step memo (x-independent contractions once)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tne_b200 as T
from helpers import seed_for
ctx = T.context()
L, D = 6, 4
g = T.grid([L, L]); order, segs = T.sweep_plan(g, L)
rng = np.random.default_rng(seed_for(L, D))
p = T.rand_simplepeps(g, D, rng, optimizer=order, segments=segs)
v = T.variables(p).to_numpy()
memo = {}
for k in range(4):
    x = (rng.standard_normal(v.size) + 1j * rng.standard_normal(v.size)) / np.sqrt(2)
    for m in (None, memo):
        ctx.sync(); t0 = time.time()
        sx = T.apply_smatrix(p, T.DiffTensor(v.copy()), T.DiffTensor(v.copy()), T.DiffTensor(x, requires_grad=False), memo=m).to_numpy()
        ctx.sync(); print("product %d %-7s %.3f s  max|Sx| %.6e" % (k, "memo" if m is not None else "plain", time.time() - t0, np.abs(sx).max()), flush=True)
