import sys, time, cProfile, pstats
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import tne_b200 as T
ctx=T.context()
g=T.grid([8,8])
p=T.rand_simplepeps(g,4,np.random.default_rng(1),optimizer=list(range(128)))
args=dict(C=10.0,delta=0.2,omega=0.3,dt=0.03)
for _ in range(3): T.rydberg_tebd_sweep_(p,g,**args)
ctx.sync(); t0=time.time()
for _ in range(5): T.rydberg_tebd_sweep_(p,g,**args)
ctx.sync(); print("sweep ms", (time.time()-t0)/5*1e3)
ctx.set_option("profile",0)
pr=cProfile.Profile(); pr.enable()
for _ in range(5): T.rydberg_tebd_sweep_(p,g,**args)
ctx.sync(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
