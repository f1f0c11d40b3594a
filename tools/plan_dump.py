"""Planner tool (runs WITHOUT a GPU): builds the <H>+grad tree program of an L x L, D PEPS in a planner context
(TNE_PLAN_ONLY=1: no kernel ever runs, nothing is computed) and summarises the launch groups the library would issue.

    python tools/plan_dump.py 6 4 [--raw] [--opt no_site2=1 ...]

Cost model for the time column: max(flops / rate, bytes / bw) per launch with the rates measured on B200
(profiles/README_r01.md): absorb 23 TFLOP/s & 4.3 TB/s, site2 25 TFLOP/s, multi 28 TFLOP/s, reduce 22 TFLOP/s & 4.6 TB/s."""
import os, sys, re, collections, subprocess
os.environ["TNE_PLAN_ONLY"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child(L, D, opts, what):
    import numpy as np
    import tne_b200 as T
    from tne_b200 import peps as P, einsum as es
    from tne_b200._lib import TneError
    from helpers import seed_for
    ctx = T.context()
    for k, v in opts:
        ctx.set_option(k, v)
    g = T.grid([L, L])
    order, segs = T.sweep_plan(g, L)
    p = T.rand_simplepeps(g, D, np.random.default_rng(seed_for(L, D)), optimizer=order, segments=segs)
    h = T.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    nt = len(p.alltensors())
    code = p.code_inner_product
    leaves = [P._dev(t) for t in p.alltensors()] * 2
    cj = [1] * nt + [0] * nt
    if what == "norm":
        tlist = []
    else:
        cache = {}
        terms = [P._term_replacements(p, t, cache) for t in h.blocks]
        tlist = [(coef, [(nt + site - 1, t) for site, t in sorted(repl.items())]) for coef, repl in terms]
    try:
        if what == "norm":
            es.expect_terms_device(code, leaves, cj, [(1.0 + 0j, [])], list(range(nt)), 1.0 + 0j)
        else:
            es.expect_terms_device(code, leaves, cj, tlist, list(range(nt)), 1.0 + 0j)
    except TneError as e:
        print("planner returned:", e, file=sys.stderr)


RATES = {  # kind -> (TFLOP/s, TB/s)
    "site2": (25.0, 99.0), "site2b": (25.0, 99.0), "multi": (28.0, 5.0), "absorb": (23.5, 4.4), "reduce": (22.0, 4.6), "generic": (1.0, 1.0),
}


def parse(text):
    ops = []
    for line in text.splitlines():
        if not line.startswith("[tne-final]"):
            continue
        tok = line.split()
        d = {}
        i = 1
        refs = []
        while i < len(tok):
            k = tok[i]
            if k in ("fwd", "bwd"):
                d["dir"] = k
                i += 1
                continue
            v = tok[i + 1]
            if k in ("a", "b", "A", "B1", "B2", "o", "dC", "dB1"):
                vid, n, st = v.split(":")
                refs.append((k, int(vid), int(n), st))
            elif k in ("seg", "phase", "nsl", "first_only", "kernel", "re", "bwdop", "acc", "M", "N", "K", "B", "out"):
                d[k] = int(v)
            elif k in ("flops", "bytes"):
                d[k] = float(v)
            else:
                d[k] = v
            i += 2
        d["refs"] = refs
        ops.append(d)
    return ops


def classify(d):
    if d["kind"] in ("site2", "site2b", "multi", "generic"):
        return d["kind"]
    # dmma kernels: absorb (big M) or reduce (big K)
    return "reduce" if d["K"] >= 8192 else "absorb"


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        L, D, what = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
        opts = [(a.split("=")[0], int(a.split("=")[1])) for a in sys.argv[5:]]
        child(L, D, opts, what)
        return
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    L, D = int(args[0]), int(args[1])
    what = "norm" if "--norm" in sys.argv else "energy"
    opts = [a for a in args[2:] if "=" in a]
    res = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(L), str(D), what] + opts, capture_output=True, text=True)
    if "--raw" in sys.argv:
        sys.stdout.write(res.stderr)
        return
    ops = parse(res.stderr)
    if not ops:
        print(res.stdout[-3000:], res.stderr[-3000:])
        return
    tot = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    per_phase = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0]))
    for d in ops:
        n = 1 if d["first_only"] else d["nsl"]
        c = classify(d)
        r = RATES[c]
        t = max(d["flops"] / (r[0] * 1e12), d["bytes"] / (r[1] * 1e12)) * n
        key = c
        if c == "absorb":
            key = "absorb K=%d N=%d" % (d["K"], d["N"]) if d["M"] >= 4096 else "absorb small"
        for acc in (tot[key], per_phase[(d["phase"], d["seg"], d["dir"])][c]):
            acc[0] += n; acc[1] += d["flops"] * n; acc[2] += d["bytes"] * n; acc[3] += t
    print("%-26s %8s %12s %12s %9s" % ("class", "launches", "flops", "bytes", "model s"))
    T = 0.0
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][3]):
        print("%-26s %8d %12.4e %12.4e %9.3f" % (k, v[0], v[1], v[2], v[3]))
        T += v[3]
    print("total model time %.3f s, flops %.4e" % (T, sum(v[1] for v in tot.values())))
    for ph, cl in sorted(per_phase.items()):
        print("phase %2d seg %d %s:" % ph, "  ".join("%s %d/%.3fs" % (c, v[0], v[3]) for c, v in sorted(cl.items())))


if __name__ == "__main__":
    main()
