import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import tne_b200 as T
ctx=T.context()
g=T.grid([8,8])
p=T.rand_simplepeps(g,4,np.random.default_rng(1),optimizer=list(range(128)))
args=dict(C=10.0,delta=0.2,omega=0.3,dt=0.03)
ctx.set_option("debug",2)
T.rydberg_tebd_sweep_(p,g,**args)
