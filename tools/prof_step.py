"""One <H>+grad evaluation of the benchmark's PEPS (for ncu captures): python tools/prof_step.py [L D [reps]]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import tne_b200 as T
L = int(sys.argv[1]) if len(sys.argv) > 1 else 6
D = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = T.context()
for a in sys.argv[4:]:
    k, v = a.split("=")
    ctx.set_option(k, int(v))
g = T.grid([L, L])
order, segs = T.sweep_plan(g, L)
p = T.rand_simplepeps(g, D, np.random.default_rng(20260000 + 100 * L + D), optimizer=order, segments=segs)
h = T.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
for r in range(reps):
    ctx.sync(); t = time.time()
    ctx.set_option("profile", 1); ctx.profile_collect()
    y, f = T.energy_and_gradient(p, h)
    yv = y.to_numpy()
    dt = time.time() - t
    prof = ctx.profile_collect(fine=bool(os.environ.get("TNE_FINE"))); ctx.set_option("profile", 0)
    print("step %.3f s  y = %r" % (dt, yv))
    for k, v in prof.items():
        if v["launches"]:
            print("  %-22s %6d launches %9.1f ms  %6.2f TFLOP/s  %7.1f GB/s" % (str(k), v["launches"], v["ms"], v["flops"] / v["ms"] / 1e9, v["bytes"] / v["ms"] / 1e6))
