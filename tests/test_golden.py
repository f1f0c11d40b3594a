"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle):
CPU: the oracle still reproduces them; GPU: the device path reproduces them through the C ABI."""
import os

import numpy as np
import pytest

from oracle import graphs as OG, peps as OP, tebd as OT, timeevolve as OTE
from helpers import device_peps_like, oracle_peps, rel, seed_for, sweep_order, to_device_hamiltonian

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(L, D):
    return np.load(os.path.join(GOLD, "grid_%dx%d_D%d.npz" % (L, L, D)))


@pytest.mark.parametrize("L,D", [(3, 2), (4, 3)])
def test_oracle_reproduces_golden(L, D):
    gold = load(L, D)
    g = OG.grid([L, L])
    p = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    assert rel(OP.variables(p), gold["variables"]) == 0.0          # the synthetic-input rule is bit-stable
    assert rel(OP.norm(p), gold["norm"]) < 1e-13
    assert rel(OP.expect(h, p, p), gold["energy"]) < 1e-12
    if L == 3:
        assert rel(OTE.fvec(p, h).data, gold["fvec"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("L,D", [(3, 2), (4, 3)])
def test_device_reproduces_golden(tne, L, D):
    gold = load(L, D)
    g = OG.grid([L, L])
    order, segs = tne.sweep_plan(tne.grid([L, L]), L)
    op = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))
    dp = device_peps_like(tne, op, g, D, optimizer=order, segments=segs)
    hd = to_device_hamiltonian(tne, OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3))
    assert rel(tne.variables(dp).to_numpy(), gold["variables"]) == 0.0
    assert rel(tne.norm(dp).to_numpy(), gold["norm"]) < 1e-10
    assert rel(tne.inner_product(dp, dp).to_numpy(), gold["inner"]) < 1e-10
    assert rel(tne.expect(hd, dp, dp).to_numpy(), gold["energy"]) < 1e-10
    y, f = tne.energy_and_gradient(dp, hd)
    assert rel(y.to_numpy(), gold["iloss2"]) < 1e-10 and rel(f.to_numpy(), gold["fvec"]) < 1e-10
    v = gold["variables"]
    sx = tne.apply_smatrix(dp, tne.DiffTensor(v.copy()), tne.DiffTensor(v.copy()), tne.DiffTensor(gold["x"], requires_grad=False))
    assert rel(sx.to_numpy(), gold["smatrix_x"]) < 1e-9             # tree-level metric-vector product vs reverse-over-reverse
    if L == 3:
        dq = device_peps_like(tne, oracle_peps(g, D, seed_for(L, D)), g, D)
        tne.rydberg_tebd_(dq, tne.grid([L, L]), t=0.3, C=10.0, delta=0.2, omega=0.3, nstep=10)
        assert rel(tne.vec(dq), gold["tebd_state"]) < 1e-8          # gauge-free comparison of the evolved state
