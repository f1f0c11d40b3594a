"""Benchmark-size golden fixtures from the ORACLE (CPU only; hours of work, run in the background):

    python tests/golden/make_golden_big.py full 5 4 --procs 4     -> tests/golden/grid_5x5_D4.npz
    python tests/golden/make_golden_big.py full 6 3 --procs 4     -> tests/golden/grid_6x6_D3.npz
    python tests/golden/make_golden_big.py smatrix 5 4            -> adds x / smatrix_x to grid_5x5_D4.npz
    python tests/golden/make_golden_big.py part66                 -> tests/golden/grid_6x6_D4_partial.npz
    python tests/golden/make_golden_big.py norm 8 3               -> tests/golden/grid_8x8_D3_norm.npz  (also 8 2, 10 2)
    python tests/golden/make_golden_big.py energy66 --procs 2     -> tests/golden/grid_6x6_D4_energy.npz (all 96 <psi|h_k|psi>, ~6.5 min each)
    python tests/golden/make_golden_big.py iloss66 --procs 2      -> tests/golden/grid_6x6_D4_iloss.npz  (iloss2's cores psi^T h_k psi, spread over the terms)

The oracle executes the reference's algorithm literally: one full contraction per Hamiltonian term
(``src/peps.jl:469-477``) and one TensorAD tape holding every intermediate (``src/TensorAD/TensorAD.jl:131-164``).
At 5x5 D=4 the tape of a SINGLE term already exceeds this container's 62 GB (measured: OOM), so ``full`` evaluates
``iloss2``/``fvec`` (``src/timeevolve.jl:55-64``) term by term with oracle/lean.py -- the same TTGT contractions and
``einsum_grad`` pullbacks, untaped, one tree alive at a time; checked against the literal tape in
tests/test_oracle_lean.py -- in worker processes, and adds the pieces up.  ``smatrix --method tape`` is the literal
reverse-over-reverse (fits up to 4x4 D=4), ``--method fd`` differentiates the lean gradient numerically.  ``part66`` is forward only (one 6x6 D=4 tree is
2.4e12 flop and ~30 GB in the oracle's TTGT executor): the norm, a representative subset of the Hamiltonian terms, and
directional derivatives of ``iloss2`` restricted to that subset along one-site directions, which a PEPS's
multilinearity turns into plain contractions with the site tensor replaced by the direction.

Per-piece results are cached under ``$TNE_GOLDEN_WORK`` (default /tmp/tne_golden_work) so a killed run resumes."""
import argparse
import os
import sys
import time

os.environ.setdefault("OMP_NUM_THREADS", "1")        # TTGT time is in numpy's single-threaded transposes
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import graphs as OG, peps as OP, tebd as OT, timeevolve as OTE, tensorad as AD, yao as OY, lean as OL  # noqa: E402
from helpers import oracle_peps, seed_for, sweep_order  # noqa: E402

WORK = os.environ.get("TNE_GOLDEN_WORK", "/tmp/tne_golden_work")
# 6x6 D=4 subset (0-based indices into rydberg_hamiltonian's term list: 60 edge terms in edges(g) order, then 36 sites):
# chosen to hit every structural case of the device plan (first/last column, column-crossing and in-column bonds,
# corner / edge / bulk sites).  Resolved to indices in ``subset66``.
SUBSET66_EDGES = [(1, 2), (3, 9), (15, 16), (15, 21), (35, 36)]
SUBSET66_SITES = [15, 36]
DIR_SITES66 = [15, 36]


def setup(L, D):
    g = OG.grid([L, L])
    p = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    return g, p, h


def subset66(g):
    edges = list(g.edges())
    return [edges.index(e) for e in SUBSET66_EDGES] + [len(edges) + s - 1 for s in SUBSET66_SITES]


def _cached(path, fn):
    if os.path.exists(path):
        return dict(np.load(path))
    t = time.time()
    out = fn()
    out["seconds"] = np.asarray(time.time() - t)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez(path + ".tmp.npz", **out)
    os.replace(path + ".tmp.npz", path)
    print("[golden] %s  %.1f s" % (os.path.basename(path), float(out["seconds"])), flush=True)
    return out


def _limit_memory(gb):
    import resource
    resource.setrlimit(resource.RLIMIT_AS, (int(gb * 2 ** 30), int(gb * 2 ** 30)))   # fail this process, not the box


def term_piece(args):
    """One term of iloss2 with the cotangent of the scaled tensors (oracle/lean.py), plus that term's <psi|h_k|psi>."""
    L, D, k, memgb = args
    _limit_memory(memgb)
    g, p, h = setup(L, D)

    def fn():
        e, gpl = OL.iloss2_term_piece(p, h.blocks[k])
        return dict(E=np.asarray(e), gpl=gpl, energy=np.asarray(OP.expect(OY.Add(h.blocks[k]), p, p)))
    return _cached(os.path.join(WORK, "full_%dx%d_D%d" % (L, L, D), "term_%03d.npz" % k), fn)


def make_full(L, D, procs, memgb):
    import multiprocessing as mp
    g, p, h = setup(L, D)
    K = len(h.blocks)
    jobs = [(L, D, k, memgb) for k in range(K)]
    if procs > 1:
        with mp.get_context("fork").Pool(procs, maxtasksperchild=1) as pool:
            pieces = pool.map(term_piece, jobs, chunksize=1)
    else:
        pieces = [term_piece(j) for j in jobs]
    _limit_memory(memgb)
    base = _cached(os.path.join(WORK, "full_%dx%d_D%d" % (L, L, D), "base.npz"),
                   lambda: dict(inner=np.asarray(OP.inner_product(p, p))))
    v = OP.variables(p)
    il, fv = OL.assemble_iloss2_gradient(p, [complex(q["E"]) for q in pieces], np.sum([q["gpl"] for q in pieces], axis=0))
    out = dict(variables=v, inner=base["inner"], norm=np.sqrt(np.abs(base["inner"])),
               iloss2_terms=np.array([complex(q["E"]) for q in pieces]),      # complex E_k; iloss2 = Re sum
               energy_terms=np.array([complex(q["energy"]) for q in pieces]),
               iloss2=np.asarray(il), energy=np.asarray(sum(complex(q["energy"]) for q in pieces)), fvec=fv)
    path = os.path.join(ROOT, "tests", "golden", "grid_%dx%d_D%d.npz" % (L, L, D))
    if os.path.exists(path):                         # keep x / smatrix_x of an earlier ``smatrix`` run
        old = dict(np.load(path))
        for key in ("x", "smatrix_x", "smatrix_method"):
            if key in old:
                out[key] = old[key]
    np.savez_compressed(path, **out)
    print("[golden] wrote", path, flush=True)


def make_smatrix(L, D, method, memgb):
    _limit_memory(memgb)
    g, p, h = setup(L, D)
    v = OP.variables(p)
    r = np.random.default_rng(seed_for(L, D) + 1)
    x = (r.standard_normal(v.size) + 1j * r.standard_normal(v.size)) / np.sqrt(2.0)
    if method == "tape":      # the literal reverse-over-reverse of src/timeevolve.jl:15-52
        fn = lambda: dict(smatrix_x=OTE.apply_smatrix(p, AD.DiffTensor(v.copy()), AD.DiffTensor(v.copy()),
                                                      AD.DiffTensor(x, requires_grad=False)).data)
    else:                     # its tape does not fit: Richardson central differences of the lean gradient (~1e-9)
        fn = lambda: dict(smatrix_x=OL.apply_smatrix_fd(p, v, v.copy(), x))
    sm = _cached(os.path.join(WORK, "full_%dx%d_D%d" % (L, L, D), "smatrix_%s.npz" % method), fn)
    path = os.path.join(ROOT, "tests", "golden", "grid_%dx%d_D%d.npz" % (L, L, D))
    out = dict(np.load(path)) if os.path.exists(path) else dict(variables=v)
    out["x"] = x
    out["smatrix_x"] = sm["smatrix_x"]
    out["smatrix_method"] = np.asarray(method)
    np.savez_compressed(path, **out)
    print("[golden] wrote", path, flush=True)


def make_part66(memgb):
    _limit_memory(memgb)
    L, D = 6, 4
    g, p, h = setup(L, D)
    wd = os.path.join(WORK, "part66")
    ks = subset66(g)
    pc = OP.conj(p)
    out = dict(variables=OP.variables(p), term_index=np.array(ks))
    out["inner"] = _cached(os.path.join(wd, "inner.npz"), lambda: dict(v=np.asarray(OP.inner_product(p, p))))["v"]
    en, il = [], []
    for k in ks:
        hk = OY.Add(h.blocks[k])
        en.append(_cached(os.path.join(wd, "energy_%03d.npz" % k), lambda: dict(v=np.asarray(OP.expect(hk, p, p))))["v"])
        # iloss2's un-normalised core: code(psi..., (h_k psi)...) = expect(h_k, conj(psi), psi)  (src/timeevolve.jl:59)
        il.append(_cached(os.path.join(wd, "iloss_%03d.npz" % k), lambda: dict(v=np.asarray(OP.expect(hk, pc, p))))["v"])
        _write66(out, en, il, [], [], [])
    dirs, dG, dN = [], [], []
    for s in DIR_SITES66:
        t = p.vertex_tensors[s - 1]
        r = np.random.default_rng(seed_for(L, D) + 1000 + s)
        V = ((r.standard_normal(t.size) + 1j * r.standard_normal(t.size)) / np.sqrt(2.0)).reshape(t.shape, order="F")
        pv = OP.replace_tensors(p, [V if j == s - 1 else q for j, q in enumerate(p.vertex_tensors)])
        pvc = OP.conj(pv)
        # d<psi|psi> along V_s = 2 Re <psi[s->V]|psi>;   dG_k along V_s = expect(h_k, conj(psi[s->V]), psi)
        dN.append(_cached(os.path.join(wd, "dnorm_s%02d.npz" % s), lambda: dict(v=np.asarray(OP.inner_product(pv, p))))["v"])
        row = []
        for k in ks:
            hk = OY.Add(h.blocks[k])
            row.append(_cached(os.path.join(wd, "dir_s%02d_%03d.npz" % (s, k)),
                               lambda: dict(v=np.asarray(OP.expect(hk, pvc, p))))["v"])
        dirs.append(np.reshape(V, -1, order="F"))
        dG.append(row)
        _write66(out, en, il, dirs, dG, dN)


def make_norm(L, D, memgb):
    """<psi|psi> of the seeded L x L PEPS through the oracle's inner_product (src/peps.jl:106-110), forward only: the
    validation sizes of BASELINE configs[3] (8x8) and configs[4] (10x10) whose D = 4 boundary does not fit a CPU."""
    _limit_memory(memgb)
    g, p, h = setup(L, D)
    t = time.time()
    inner = np.asarray(OP.inner_product(p, p))
    path = os.path.join(ROOT, "tests", "golden", "grid_%dx%d_D%d_norm.npz" % (L, L, D))
    np.savez_compressed(path, variables=OP.variables(p), inner=inner, norm=np.sqrt(np.abs(inner)))
    print("[golden] wrote %s (%.1f s): inner = %r" % (path, time.time() - t, complex(inner)), flush=True)


def energy66_piece(args):
    """<psi|h_k|psi> of the 6x6 D=4 benchmark PEPS for one Hamiltonian term (one full oracle contraction, cached)."""
    k, memgb = args
    _limit_memory(memgb)
    g, p, h = setup(6, 4)
    hk = OY.Add(h.blocks[k])
    r = _cached(os.path.join(WORK, "part66", "energy_%03d.npz" % k), lambda: dict(v=np.asarray(OP.expect(hk, p, p))))
    return k, complex(r["v"])


def make_energy66(procs, memgb):
    """Every term of the benchmark's Hamiltonian (60 edge terms, 36 site terms) at 6x6 D=4, forward only: the fixture grows as
    the workers finish (term_index lists what is in it), so a killed run still leaves a usable golden."""
    import multiprocessing as mp
    g, p, h = setup(6, 4)
    K = len(h.blocks)
    wd = os.path.join(WORK, "part66")
    inner = _cached(os.path.join(wd, "inner.npz"), lambda: dict(v=np.asarray(OP.inner_product(p, p))))["v"]
    done = {}
    path = os.path.join(ROOT, "tests", "golden", "grid_6x6_D4_energy.npz")

    def write():
        ks = sorted(done)
        np.savez_compressed(path + ".tmp.npz", variables=OP.variables(p), inner=inner, term_index=np.array(ks),
                            energy_terms=np.array([done[k] for k in ks]), n_terms=np.asarray(K))
        os.replace(path + ".tmp.npz", path)
    cached_first = sorted(range(K), key=lambda k: not os.path.exists(os.path.join(wd, "energy_%03d.npz" % k)))
    with mp.get_context("fork").Pool(procs, maxtasksperchild=1) as pool:
        for k, v in pool.imap_unordered(energy66_piece, [(k, memgb) for k in cached_first], chunksize=1):
            done[k] = v
            write()
            print("[golden] energy66: %d of %d terms" % (len(done), K), flush=True)


def iloss66_piece(args):
    """iloss2's un-normalised core of one term, code(psi..., (h_k psi)...) = expect(h_k, conj(psi), psi) (src/timeevolve.jl:59)."""
    k, memgb = args
    _limit_memory(memgb)
    g, p, h = setup(6, 4)
    hk = OY.Add(h.blocks[k])
    pc = OP.conj(p)
    r = _cached(os.path.join(WORK, "part66", "iloss_%03d.npz" % k), lambda: dict(v=np.asarray(OP.expect(hk, pc, p))))
    return k, complex(r["v"])


def make_iloss66(procs, memgb):
    """The cores of iloss2 (the quantity the benchmark differentiates) term by term at 6x6 D=4; cached pieces first, then the terms
    in a stride-7 order so that a run cut short still covers edge and site terms of every lattice region."""
    import multiprocessing as mp
    g, p, h = setup(6, 4)
    K = len(h.blocks)
    wd = os.path.join(WORK, "part66")
    inner = _cached(os.path.join(wd, "inner.npz"), lambda: dict(v=np.asarray(OP.inner_product(p, p))))["v"]
    done = {}
    path = os.path.join(ROOT, "tests", "golden", "grid_6x6_D4_iloss.npz")

    def write():
        ks = sorted(done)
        np.savez_compressed(path + ".tmp.npz", variables=OP.variables(p), inner=inner, term_index=np.array(ks),
                            iloss_core_terms=np.array([done[k] for k in ks]), n_terms=np.asarray(K))
        os.replace(path + ".tmp.npz", path)
    spread = [(7 * q) % K for q in range(K)]            # 7 and 96 are coprime: a permutation of the terms
    order = sorted(spread, key=lambda k: not os.path.exists(os.path.join(wd, "iloss_%03d.npz" % k)))
    with mp.get_context("fork").Pool(procs, maxtasksperchild=1) as pool:
        for k, v in pool.imap_unordered(iloss66_piece, [(k, memgb) for k in order], chunksize=1):
            done[k] = v
            write()
            print("[golden] iloss66: %d of %d terms" % (len(done), K), flush=True)


def _write66(out, en, il, dirs, dG, dN):
    o = dict(out)
    o["energy_terms"] = np.array([complex(x) for x in en])
    o["iloss_core_terms"] = np.array([complex(x) for x in il])
    o["dir_sites"] = np.array(DIR_SITES66[:len(dirs)])
    for s, V, row, dn in zip(DIR_SITES66, dirs, dG, dN):
        o["dir_V_s%02d" % s] = V
        o["dir_dcore_s%02d" % s] = np.array([complex(x) for x in row])
        o["dir_dinner_s%02d" % s] = np.asarray(complex(dn))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "grid_6x6_D4_partial.npz"), **o)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["full", "smatrix", "part66", "norm", "energy66", "iloss66"])
    ap.add_argument("L", type=int, nargs="?", default=5)
    ap.add_argument("D", type=int, nargs="?", default=4)
    ap.add_argument("--procs", type=int, default=1)
    ap.add_argument("--memgb", type=float, default=14.0, help="address-space limit per process")
    ap.add_argument("--method", choices=["tape", "fd"], default="tape")
    a = ap.parse_args()
    if a.what == "full":
        make_full(a.L, a.D, a.procs, a.memgb)
    elif a.what == "smatrix":
        make_smatrix(a.L, a.D, a.method, a.memgb)
    elif a.what == "norm":
        make_norm(a.L, a.D, a.memgb)
    elif a.what == "energy66":
        make_energy66(a.procs, a.memgb)
    elif a.what == "iloss66":
        make_iloss66(a.procs, a.memgb)
    else:
        make_part66(a.memgb)
