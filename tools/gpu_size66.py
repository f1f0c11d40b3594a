"""Sizing run: L x L, D PEPS norm / expect / fvec timings with the boundary-sweep tree."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tne_b200 as T
from helpers import sweep_order, seed_for

L = int(sys.argv[1]); D = int(sys.argv[2]); what = sys.argv[3:] or ["norm", "expect", "fvec"]
ctx = T.context(); ctx.set_option("debug", 1)
g = T.grid([L, L])
rng = np.random.default_rng(seed_for(L, D))
t0 = time.time()
p = T.rand_simplepeps(g, D, rng, optimizer=sweep_order(L * L))
h = T.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
print("build", time.time() - t0, flush=True)
def timed(name, f):
    ctx.sync(); t0 = time.time(); r = f(); ctx.sync(); t1 = time.time()
    print(name, "%.3f s" % (t1 - t0), ctx.program_stats(), ctx.mem_info(), flush=True)
    return r
if "norm" in what:
    n = timed("norm", lambda: T.norm(p)); print(n.to_numpy())
    n = timed("norm again", lambda: T.norm(p))
if "expect" in what:
    e = timed("expect", lambda: T.expect(h, p, p)); print(e.to_numpy())
if "fvec" in what:
    f = timed("fvec", lambda: T.fvec(p, h)); print(np.abs(f.to_numpy()).max())
    f = timed("fvec again", lambda: T.fvec(p, h))
