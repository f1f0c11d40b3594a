"""The memory-lean oracle evaluation (oracle/lean.py, used for the benchmark-size golden vectors) against the literal
tape restatement and the committed small golden files."""
import os

import numpy as np
import pytest

from oracle import graphs as OG, peps as OP, tebd as OT, timeevolve as OTE, tensorad as AD, lean as OL
from helpers import oracle_peps, seed_for, sweep_order, rel

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("L,D", [(3, 2), (4, 3)])
def test_lean_fvec_matches_golden(L, D):
    g = OG.grid([L, L])
    p = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    gold = np.load(os.path.join(GOLD, "grid_%dx%d_D%d.npz" % (L, L, D)))
    val, f = OL.fvec(p, h)
    assert abs(val - float(gold["iloss2"])) <= 1e-13 * abs(float(gold["iloss2"]))
    assert rel(f, gold["fvec"]) < 1e-13


def test_lean_fvec_matches_literal_tape_on_test_graph():
    g = OG.SimpleGraph(5)
    for e in [(1, 2), (1, 3), (2, 4), (2, 5), (3, 4), (3, 5)]:   # test/peps.jl:6
        g.add_edge(*e)
    p = oracle_peps(g, 2, 7)
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    val, f = OL.fvec(p, h)
    assert rel(f, OTE.fvec(p, h).data) < 1e-13
    assert abs(val - float(OTE.iloss2(h, p, OP.variables(p)))) < 1e-13


def test_lean_overlap_gradient_and_fd_smatrix():
    L, D = 3, 2
    g = OG.grid([L, L])
    p = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))
    gold = np.load(os.path.join(GOLD, "grid_3x3_D2.npz"))
    v = OP.variables(p)
    r = np.random.default_rng(5)
    v2 = v + 0.1 * (r.standard_normal(v.size) + 1j * r.standard_normal(v.size))
    y, g1 = OL.normalized_overlap_grad1(p, v, v2)
    lit = AD.gradient(lambda a: OTE.normalized_overlap(p, a, AD.DiffTensor(v2, requires_grad=False)), AD.DiffTensor(v))[0].data
    assert rel(g1, lit) < 1e-12
    sx = OL.apply_smatrix_fd(p, v, v.copy(), gold["x"])
    assert rel(sx, gold["smatrix_x"]) < 1e-8
