import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tne_b200 as T
from oracle import graphs as OG, peps as OP, yao as OY
from helpers import *
g = OG.test_graph(); rng = np.random.default_rng(2)
for kind in ["simple", "vidal"]:
    for (i, j) in [(4, 2), (1, 2), (2, 4), (3, 5)]:
        op = oracle_peps(g, 2, 31, kind, Dmax=4); dp = device_peps_like(T, op, g, 2, Dmax=4)
        m = OY.rand_unitary(4, rng); m4 = np.reshape(m, (2, 2, 2, 2), order="F")
        li, lj = OP.getvlabel(op, i), OP.getvlabel(op, j)
        OP.apply_onbond_(op, i, j, m4); T.apply_onbond_(dp, i, j, m4)
        a, b = dp.vertex_tensors[i - 1].to_numpy(), dp.vertex_tensors[j - 1].to_numpy()
        ra, rb = op.vertex_tensors[i - 1], op.vertex_tensors[j - 1]
        sh = [l for l in li if l in lj][0]
        lo = [l for l in li if l != sh] + [l for l in lj if l != sh]
        from oracle import einsum as OE
        th_d = OE.einsum_raw([li, lj], [a, b], lo); th_r = OE.einsum_raw([li, lj], [ra, rb], lo)
        print(kind, (i, j), "axes", li.index(sh), lj.index(sh), "shapes", a.shape, b.shape, "theta err", rel(th_d, th_r), "state err", rel(T.vec(dp), OP.vec(op)))
