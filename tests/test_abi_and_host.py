"""CPU-side checks of the drop-in boundary and the host logic: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/tne.h declares (no compute call is made without a GPU); the host mirror of the
OMEinsum objects produces valid trees with the flop counts SURVEY.md 8(d) quotes."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "tne.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tne_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_binding_surface():
    import tne_b200
    assert set(header_functions()) == set(tne_b200._lib.SYMBOLS)


def test_library_builds_loads_and_exports_every_symbol():
    import tne_b200
    path = tne_b200.build()
    L = ctypes.CDLL(path)
    for name in header_functions():
        assert hasattr(L, name), name
    assert tne_b200._lib.lib().tne_version() >= 100


def test_julia_shim_matches_the_header():
    """julia/TensorNetworkEvolveB200.jl cannot be executed here (no Julia runtime): every ccall signature and struct
    mirror is diffed mechanically against include/tne.h, and the PEPS dispatch uses the reference's three type parameters."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import check_julia_ccalls as chk
    assert chk.problems() == []
    jl = open(os.path.join(ROOT, "julia", "TensorNetworkEvolveB200.jl")).read()
    assert len(chk.julia_ccalls(jl)) >= 15
    assert not re.search(r"PEPS\{[^}]*,[^}]*,[^}]*,[^}]*\}", jl)      # abstract type PEPS{T,NF,LT}: src/peps.jl:4
    bad = jl.replace("(Ptr{Cvoid}, Int64, Ptr{Int64}), CTX[], x.handle, h))", "(Ptr{Cvoid}, Cint, Ptr{Int64}), CTX[], x.handle, h))", 1)
    assert bad != jl                                                   # the checker does catch a wrong argument type
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "julia")); os.makedirs(os.path.join(d, "include"))
        open(os.path.join(d, "julia", "TensorNetworkEvolveB200.jl"), "w").write(bad)
        open(os.path.join(d, "include", "tne.h"), "w").write(open(os.path.join(ROOT, "include", "tne.h")).read())
        root, chk.ROOT = chk.ROOT, d
        try:
            assert len(chk.problems()) == 1
        finally:
            chk.ROOT = root


def test_no_cpu_fallback_without_a_device():
    import torch
    import tne_b200
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(tne_b200.TneError):
        tne_b200.array.Context(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "tensornetworkevolve.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_sweep_tree_flops_match_survey():
    from tne_b200 import einsum as es, graphs
    from helpers import sweep_order
    for L, D, flops in [(3, 2, 7.2e4), (4, 3, 8.5e7), (6, 4, 2.38e12)]:
        g = graphs.grid([L, L])
        n, edges = g.nv(), g.edges()
        emap = {e: n + 1 + k for k, e in enumerate(edges)}
        vl = [[i] + [emap.get((i, nb), emap.get((nb, i))) for nb in g.neighbors(i)] for i in range(1, n + 1)]
        mx = n + len(edges)
        rep = {l: mx + k + 1 for k, l in enumerate(range(n + 1, mx + 1))}
        ixs = vl + [[rep.get(x, x) for x in l] for l in vl]
        sd = {l: (2 if l <= n else D) for ix in ixs for l in ix}
        code = es.optimize_code(ixs, [], sd, sweep_order(n))
        f, b, big = es.tree_cost(code, sd)
        assert abs(f - flops) / flops < 0.02, (L, D, f)
        ser = es.SerializedTree(code, [])
        assert ser.struct.n_leaves == 2 * n and ser.struct.n_nodes == 2 * n - 1


def test_greedy_tree_is_a_valid_binary_tree():
    from tne_b200 import einsum as es
    from oracle import einsum as OE
    rng = np.random.default_rng(0)
    ixs = [[1, 2, 3], [3, 4], [4, 5, 1], [5, 2, 6]]
    sd = {l: 3 for l in range(1, 7)}
    code = es.optimize_code(ixs, [6], sd, "greedy")
    assert sorted(range(4)) == sorted(i for i in range(len(code.leaf_ixs())))
    # the same tree evaluated by the oracle's executor equals a direct numpy einsum
    xs = [rng.standard_normal([3] * len(ix)) + 0j for ix in ixs]

    def ev(nd):
        if nd.isleaf():
            return xs[nd.tensorindex]
        a, b = [ev(c) for c in nd.args]
        return OE.einsum_raw(nd.eins.ixs, [a, b], nd.eins.iy)
    assert np.allclose(ev(code), np.einsum("abc,cd,dea,ebf->f", *xs))
