"""Pins the oracle's PEPS layer on the reference's own tests (test/peps.jl), restated one by one.
The reference stores no golden vectors: every check is a known answer or an identity against a dense
2^n statevector, which is what is done here (oracle.yao.ArrayReg is the dense simulator)."""
import numpy as np
import pytest

from oracle import einsum as OE, graphs as OG, peps as OP, yao as OY


def e(n, k):
    x = np.zeros(1 << n, dtype=np.complex128)
    x[k - 1] = 1  # Julia's 1-based x[k] = 1
    return x


@pytest.fixture
def g():
    return OG.test_graph()  # test/peps.jl:5-8


def test_initial_state_known_answers(g):  # test/peps.jl:4-24
    peps = OP.zero_vidalpeps(g, 2, nslices=3)  # TreeSA(nslices=3) -> a SlicedEinsum (test/peps.jl:9)
    assert isinstance(peps.code_inner_product, OE.SlicedEinsum)
    assert OP.virtualbonds(peps) == [(1, 2), (1, 3), (2, 4), (2, 5), (3, 4), (3, 5)]
    assert np.allclose(OP.vec(peps), e(5, 1))
    assert OP.newlabel(peps, 2) == 11 + 2
    assert OP.newlabel(peps, 4) == 11 + 4
    assert OP.getphysicallabel(peps, 2) == 2
    assert len(OP.getvlabel(peps, 2)) == 4
    OP.apply_onsite_(peps, 1, OY.X)
    assert np.allclose(OP.vec(peps), e(5, 2))
    OP.apply_onbond_(peps, 1, 2, np.reshape(OY.CNOT, (2, 2, 2, 2), order="F"))
    assert np.allclose(OP.vec(peps), e(5, 4))
    peps = OP.rand_vidalpeps(g, 2, np.random.default_rng(1))
    OP.normalize_(peps)
    assert np.isclose(OP.norm(peps), 1)


def test_random_gate(g):  # test/peps.jl:26-44
    rng = np.random.default_rng(2)
    peps = OP.rand_vidalpeps(g, 2, rng, Dmax=4)
    reg = OY.ArrayReg(OP.vec(peps))
    m = OY.rand_unitary(2, rng)
    OP.apply_onsite_(peps, 1, m)
    reg.apply_(OY.put(5, 1, m))
    assert np.allclose(OP.vec(peps), reg.state)
    m = OY.rand_unitary(4, rng)
    reg.apply_(OY.put(5, (4, 2), m))
    OP.apply_onbond_(peps, 4, 2, np.reshape(m, (2, 2, 2, 2), order="F"))
    assert np.allclose(OP.vec(peps), reg.state)


def test_conj_variables_load(g):  # test/peps.jl:47-63
    rng = np.random.default_rng(3)
    peps = OP.rand_simplepeps(g, 2, rng, Dmax=4)
    assert np.allclose(OP.statevec(OP.conj(peps)), np.conj(OP.statevec(peps)))
    ref = OP.statevec(peps).copy()
    v = OP.variables(peps).copy()
    OP.randn_(peps, rng)
    assert not np.allclose(ref, OP.statevec(peps))
    OP.load_variables_(peps, v)
    assert np.allclose(ref, OP.statevec(peps))
    p2 = OP.randn_(peps, rng)
    p2 = OP.load_variables(p2, v)
    assert np.allclose(ref, OP.statevec(p2))


@pytest.mark.parametrize("kind", ["simple", "vidal"])
def test_inner_product_and_normalize(g, kind):  # test/peps.jl:65-86
    rng = np.random.default_rng(4)
    mk = OP.rand_vidalpeps if kind == "vidal" else OP.rand_simplepeps
    p1, p2 = mk(g, 2, rng, Dmax=4), mk(g, 2, rng, Dmax=4)
    assert np.isclose(OP.inner_product(p1, p1), np.vdot(OP.statevec(p1), OP.statevec(p1)))
    assert np.isclose(OP.inner_product(p1, p2), np.vdot(OP.statevec(p1), OP.statevec(p2)))
    assert np.isclose(np.linalg.norm(OP.vec(p1)), OP.norm(p1))
    OP.normalize_(p1)
    assert np.isclose(OP.norm(p1), 1)


def rand_hamiltonian(g, rng):  # test/peps.jl:89-100
    n = g.nv()
    blocks = []
    for (i, j) in g.edges():
        blocks.append(OY.put(n, (i, j), rng.random() * OY.kron_mat(OY.X, OY.X))
                      + rng.random() * OY.kron(n, (i, OY.Y), (j, OY.Y))
                      + OY.control(n, i, j, rng.random() * OY.Z))
    for i in g.vertices():
        blocks.append(OY.put(n, i, rng.random() * OY.X))
    h = blocks[0]
    for b in blocks[1:]:
        h = h + b
    return h


def test_expectation_values(g):  # test/peps.jl:88-110
    rng = np.random.default_rng(5)
    h = rand_hamiltonian(g, rng)
    p1, p2 = OP.rand_simplepeps(g, 2, rng, Dmax=4), OP.rand_simplepeps(g, 2, rng, Dmax=4)
    H = OY.mat(h)
    assert np.isclose(OP.expect(h, p1, p2), np.vdot(OP.statevec(p1), H @ OP.statevec(p2)))
    assert np.isclose(OP.expect(h, p1, p1), np.vdot(OP.statevec(p1), H @ OP.statevec(p1)))


# -- beyond the reference's tests: grids (the bench configs) against the dense statevector ----------
@pytest.mark.parametrize("L,D", [(3, 2), (4, 2)])
def test_grid_against_dense(L, D):
    from helpers import dense_apply, oracle_peps, seed_for, sweep_order
    from oracle import tebd as OT
    g = OG.grid([L, L])
    p = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))
    q = oracle_peps(g, D, seed_for(L, D), optimizer="greedy")
    psi = OP.statevec(p)
    assert np.isclose(OP.inner_product(p, p), np.vdot(psi, psi))
    assert np.isclose(OP.inner_product(q, q), np.vdot(psi, psi))  # any contraction order, same number
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    assert np.isclose(OP.expect(h, p, p), np.vdot(psi, dense_apply(h, psi)))


def test_sliced_equals_unsliced():
    from helpers import oracle_peps
    g = OG.grid([3, 3])
    a = oracle_peps(g, 2, 7, optimizer="greedy")
    b = oracle_peps(g, 2, 7, optimizer="greedy", nslices=2)
    assert isinstance(b.code_inner_product, OE.SlicedEinsum)
    assert np.isclose(OP.inner_product(a, a), OP.inner_product(b, b))


def test_einsum_ttgt_against_numpy():
    rng = np.random.default_rng(6)
    a = rng.standard_normal((3, 4, 2, 5)) + 1j * rng.standard_normal((3, 4, 2, 5))
    b = rng.standard_normal((5, 2, 6, 3)) + 1j * rng.standard_normal((5, 2, 6, 3))
    out = OE.einsum_raw([[1, 2, 3, 4], [4, 3, 5, 1]], [a, b], [5, 2, 1])
    assert np.allclose(out, np.einsum("abcd,dcea->eba", a, b))
    assert np.allclose(OE.einsum_raw([[1, 2, 3, 4]], [a], [3, 1]), np.einsum("abcd->ca", a))
