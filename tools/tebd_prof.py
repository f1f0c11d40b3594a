"""Host profile of (replica-batched) TEBD sweeps: python tools/tebd_prof.py [R]"""
import sys, time, cProfile, pstats
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import tne_b200 as T
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = T.context()
g = T.grid([8, 8])
ps = [T.rand_simplepeps(g, 4, np.random.default_rng(1 + r)) for r in range(R)]
args = (g, 10.0, 0.2, 0.3, 0.03)
for _ in range(3): T.rydberg_tebd_sweep_batch_(ps, *args)
ctx.sync(); t0 = time.time()
for _ in range(5): T.rydberg_tebd_sweep_batch_(ps, *args)
ctx.sync(); print("R", R, "batched sweep ms", (time.time() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): T.rydberg_tebd_sweep_batch_(ps, *args)
ctx.sync(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
