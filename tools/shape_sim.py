"""Host-side simulation of the intermediate layouts/shapes of the boundary-sweep tree (no GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import graphs as OG
L = int(sys.argv[1]); D = int(sys.argv[2])
g = OG.grid([L, L]); n = g.nv()
edges = g.edges(); emap = {e: n + 1 + k for k, e in enumerate(edges)}
maxi = n + len(edges)
vl = []
for i in range(1, n + 1):
    vl.append([i] + [emap.get((i, nb), emap.get((nb, i))) for nb in g.neighbors(i)])
rep = {l: maxi + k + 1 for k, l in enumerate(range(n + 1, maxi + 1))}
ixs = vl + [[rep.get(x, x) for x in l] for l in vl]
size = {}
for ix in ixs:
    for l in ix: size[l] = 2 if l <= n else D
order = []
for i in range(n): order += [i, n + i]
count = {}
for ix in ixs:
    for l in ix: count[l] = count.get(l, 0) + 1
cur = list(ixs[order[0]])
tot = 0
for t in order[1:]:
    b = ixs[t]
    for l in cur: count[l] -= 1
    for l in b: count[l] -= 1
    keep = [l for l in dict.fromkeys(cur + b) if count[l] > 0]
    for l in keep: count[l] += 1
    out = [l for l in cur if l in keep] + [l for l in b if l in keep and l not in cur]
    K = [l for l in cur if l in b and l not in keep]
    M = [l for l in cur if l in keep]
    N = [l for l in b if l in keep and l not in cur]
    pos = [cur.index(l) for l in K]
    strides = np.cumprod([1] + [size[l] for l in cur])[:-1]
    msz = int(np.prod([size[l] for l in M])); nsz = int(np.prod([size[l] for l in N])); ksz = int(np.prod([size[l] for l in K]))
    fl = 8.0 * msz * nsz * ksz; tot += fl
    print("leaf %3d rankA %2d M %9d N %3d K %3d  Kpos %s Kstride %s flops %.2e" % (t, len(cur), msz, nsz, ksz, pos, [int(strides[p]) for p in pos], fl))
    cur = out
print("total %.3e" % tot)
