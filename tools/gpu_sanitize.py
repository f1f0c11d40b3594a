"""Small workload touching every kernel family (run under compute-sanitizer --tool memcheck on a B200)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tne_b200 as T
from helpers import seed_for

ctx = T.context()
for (L, D) in [(3, 2), (4, 4)]:
    g = T.grid([L, L])
    order, segs = T.sweep_plan(g, L)
    p = T.rand_simplepeps(g, D, np.random.default_rng(seed_for(L, D)), optimizer=order, segments=segs)
    h = T.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    print(L, D, "norm", T.norm(p).item())
    y, f = T.energy_and_gradient(p, h)
    print(L, D, "energy", complex(y.to_numpy()), "force", float(np.abs(f.to_numpy()).max()), ctx.program_stats()["contractions"], "launches")
q = T.rand_vidalpeps(T.grid([3, 3]), 2, np.random.default_rng(1), Dmax=4)
T.rydberg_tebd_(q, t=0.3, C=10.0, delta=0.2, omega=0.3, nstep=3)
print("tebd norm", T.norm(q).item())
ctx.sync()
print("SANITIZE WORKLOAD DONE", ctx.launch_count(), "launches")
