"""Sizing run: L x L, D PEPS norm / expect / energy+gradient timings with the segmented boundary-sweep plan."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import tne_b200 as T
from helpers import seed_for

L = int(sys.argv[1]); D = int(sys.argv[2]); what = sys.argv[3:] or ["norm", "expect", "grad"]
ctx = T.context(); ctx.set_option("debug", int(os.environ.get("TNE_DEBUG", "2")))
if os.environ.get("TNE_VARIANT"): ctx.set_option("absorb_tma", int(os.environ["TNE_VARIANT"]))
if os.environ.get("TNE_NO_PDL"): ctx.set_option("no_pdl", 1)
if os.environ.get("TNE_PAD"): ctx.set_option("pad", int(os.environ["TNE_PAD"]))
g = T.grid([L, L])
rng = np.random.default_rng(seed_for(L, D))
t0 = time.time()
order, segs = T.sweep_plan(g, L, int(os.environ.get('TNE_NBONDS', '1')))
p = T.rand_simplepeps(g, D, rng, optimizer=order, segments=segs)
h = T.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
print("build", time.time() - t0, flush=True)
def timed(name, f):
    if os.environ.get("TNE_PROFILE") and "again" in name:
        ctx.set_option("profile", 1); ctx.profile_collect(); f()
        for k, v in sorted(ctx.profile_collect(fine=True).items()):
            if v["launches"]: print("   ", k, v["launches"], "launches %.1f ms  %.2f TFLOP/s  %.2f TB/s" % (v["ms"], v["flops"] / v["ms"] / 1e9, v["bytes"] / v["ms"] / 1e9))
        ctx.set_option("profile", 0)
    ctx.sync(); t0 = time.time(); r = f(); t1 = time.time(); ctx.sync(); t2 = time.time()
    st = ctx.program_stats()
    print(name, "%.3f s (host enqueue %.3f s)" % (t2 - t0, t1 - t0), "TFLOP/s %.2f" % ((st["flops"] + st["recompute_flops"]) / (t2 - t0) / 1e12), st, ctx.mem_info(), flush=True)
    return r
if "norm" in what:
    n = timed("norm", lambda: T.norm(p)); print(n.to_numpy())
    n = timed("norm again", lambda: T.norm(p))
if "normgrad" in what:
    TA = T.tensorad
    gfun = lambda: TA.gradient(lambda x: T.norm(T.load_variables(p, x)), T.DiffTensor(T.variables(p)))[0]
    gn = timed("norm gradient", gfun); print(np.abs(gn.to_numpy()).max())
    gn = timed("norm gradient again", gfun)
if "smatrix" in what:
    v = T.variables(p).to_numpy()
    xx = (rng.standard_normal(v.size) + 1j * rng.standard_normal(v.size)) / np.sqrt(2)
    sfun = lambda: T.apply_smatrix(p, T.DiffTensor(v.copy()), T.DiffTensor(v.copy()), T.DiffTensor(xx, requires_grad=False))
    sx = timed("apply_smatrix", sfun); print(np.abs(sx.to_numpy()).max())
    sx = timed("apply_smatrix again", sfun)
if "expect" in what:
    e = timed("expect", lambda: T.expect(h, p, p)); print(e.to_numpy())
if "grad" in what:
    y, f = timed("energy_and_gradient", lambda: T.energy_and_gradient(p, h)); print(y.to_numpy(), np.abs(f.to_numpy()).max())
    y, f = timed("energy_and_gradient again", lambda: T.energy_and_gradient(p, h))
