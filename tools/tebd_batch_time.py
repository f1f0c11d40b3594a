"""TEBD throughput with replica batching (one B200): sweeps/s of an 8x8 D=4 PEPS at R = 1, 8, 32, 64, 128 replicas per call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import tne_b200 as T
ctx = T.context()
L, D = 8, 4
g = T.grid([L, L])
for R in (1, 8, 32, 64, 128):
    ps = [T.rand_simplepeps(g, D, np.random.default_rng(1000 + r)) for r in range(R)]
    f = lambda: T.rydberg_tebd_sweep_batch_(ps, g, 10.0, 0.2, 0.3, 0.03)
    f(); f()
    ctx.sync(); t0 = time.time()
    n = 5
    for _ in range(n):
        f()
    ctx.sync(); dt = (time.time() - t0) / n
    print("R = %3d replicas: %8.2f ms per batched sweep, %8.1f PEPS sweeps/s, %9.0f bond updates/s" % (R, dt * 1e3, R / dt, R * g.ne() / dt), flush=True)
