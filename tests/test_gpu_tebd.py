"""Parity of the device bond update / TEBD (tne_apply_onbond_batched, tne_svd_trunc) against the oracle and the
dense statevector, up to the SVD gauge: states, norms, singular values and truncation errors are compared, never
the raw U / V factors (BASELINE.json north_star: "post-TEBD tensors up to SVD gauge")."""
import numpy as np
import pytest

from oracle import graphs as OG, peps as OP, tebd as OT, yao as OY
from helpers import device_peps_like, oracle_peps, rel, seed_for

pytestmark = pytest.mark.gpu


def test_svd_trunc_matches_lapack(tne):
    rng = np.random.default_rng(0)
    for (m, n) in [(16, 8), (128, 16), (8, 8), (32, 4)]:
        a = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
        u, s, v, kept, err = tne.svd_trunc(tne.B200Array.from_numpy(a), Dmax=n, eps=0.0)
        sref = np.linalg.svd(a, compute_uv=False)
        assert kept == n
        assert rel(s.to_numpy(), sref) < 1e-12
        U, S, V = u.to_numpy(), s.to_numpy(), v.to_numpy()
        assert rel(U @ np.diag(S) @ V.conj().T, a) < 1e-12            # A = U S V'
        assert rel(U.conj().T @ U, np.eye(n)) < 1e-12 and rel(V.conj().T @ V, np.eye(n)) < 1e-12
        u, s, v, kept, err = tne.svd_trunc(tne.B200Array.from_numpy(a), Dmax=n // 2, eps=0.0)
        assert kept == n // 2 and abs(err - sref[n // 2:].sum()) < 1e-11 * sref.sum()
    # rank-deficient input and the absolute eps cut (src/peps.jl:383-385)
    a = np.outer(rng.standard_normal(16), rng.standard_normal(8)) + 0j
    u, s, v, kept, err = tne.svd_trunc(tne.B200Array.from_numpy(a), Dmax=8, eps=1e-10)
    assert kept == 1 and rel(u.to_numpy() @ np.diag(s.to_numpy()) @ v.to_numpy().conj().T, a) < 1e-12


def test_known_answer_cnot(tne):  # test/peps.jl:18-19
    g = tne.graph_from_edges(5, OG.test_graph().edges())
    p = tne.zero_vidalpeps(g, 2)
    tne.apply_onsite_(p, 1, OY.X)
    tne.apply_onbond_(p, 1, 2, np.reshape(OY.CNOT, (2, 2, 2, 2), order="F"))
    e4 = np.zeros(32, complex); e4[3] = 1
    assert np.allclose(tne.vec(p), e4)


@pytest.mark.parametrize("kind", ["simple", "vidal"])
def test_random_two_site_gate_vs_dense_and_oracle(tne, kind):  # test/peps.jl:26-44
    g = OG.test_graph()
    rng = np.random.default_rng(2)
    op = oracle_peps(g, 2, 31, kind, Dmax=4)
    dp = device_peps_like(tne, op, g, 2, Dmax=4)
    reg = OY.ArrayReg(OP.vec(op))
    for n, (i, j) in enumerate([(4, 2), (1, 3), (3, 5), (2, 1)]):
        m = OY.rand_unitary(4, rng)
        reg.apply_(OY.put(5, (i, j), m))
        m4 = np.reshape(m, (2, 2, 2, 2), order="F")
        OP.apply_onbond_(op, i, j, m4)
        tne.apply_onbond_(dp, i, j, m4)
        if n < 3:  # exact while the new bond dimension fits Dmax (the last gate needs 8 > Dmax = 4 and truncates)
            assert rel(tne.vec(dp), reg.state) < 1e-10
        assert rel(tne.vec(dp), OP.vec(op)) < 1e-10
        assert [t.shape for t in dp.vertex_tensors] == [t.shape for t in op.vertex_tensors]
        if kind == "vidal":  # the stored bond vector = singular values (gauge-free)
            b = op.virtual_labels.index([l for l in OP.getvlabel(op, i) if l in OP.getvlabel(op, j)][0])
            assert rel(dp.bond_tensors[b].to_numpy(), op.bond_tensors[b]) < 1e-10


def test_truncating_gate_matches_oracle(tne):
    """Dmax = D: every generic gate truncates; the truncated state is unique (gauge-free) and must match."""
    g = OG.grid([3, 3])
    rng = np.random.default_rng(3)
    op = oracle_peps(g, 2, seed_for(3, 2))
    dp = device_peps_like(tne, op, g, 2)
    for (i, j) in [(1, 2), (5, 6), (4, 5), (2, 5)]:
        m4 = np.reshape(OY.rand_unitary(4, rng), (2, 2, 2, 2), order="F")
        OP.LOG.clear(); tne.peps.LOG.clear()
        OP.apply_onbond_(op, i, j, m4)
        tne.apply_onbond_(dp, i, j, m4)
        assert rel(tne.vec(dp), OP.vec(op)) < 1e-10
        assert len(tne.peps.LOG) == len(OP.LOG) == 1
        e_dev = float(tne.peps.LOG[0].split()[-1]); e_ref = float(OP.LOG[0].split()[-1])
        assert abs(e_dev - e_ref) < 1e-10 * max(1.0, e_ref)


@pytest.mark.parametrize("kind", ["simple", "vidal"])
def test_rydberg_tebd_test_graph(tne, kind):  # test/tebd.jl:4-26
    g = OG.test_graph()
    kw = dict(Dmax=10) if kind == "simple" else dict(Dmax=10, eps=1e-10)
    op = oracle_peps(g, 2, 41, kind, **kw)
    dp = device_peps_like(tne, op, g, 2, **kw)
    reg = OY.ArrayReg(OP.vec(op))
    args = dict(t=0.3, C=10.0, delta=0.2, omega=0.3, nstep=10)
    gd = tne.graph_from_edges(5, g.edges())
    if kind == "simple":
        tne.rydberg_tebd_(dp, gd, **args)
    else:
        tne.rydberg_tebd_(dp, **args)
    OT.rydberg_tebd_(op, g, **args)
    assert not np.allclose(tne.vec(dp), reg.state, atol=1e-3)
    OT.rydberg_tebd_(reg, g, **args)
    assert np.allclose(tne.vec(dp), reg.state, atol=1e-3)      # the reference's own assertion
    assert rel(tne.vec(dp), OP.vec(op)) < 1e-8                  # and parity with the oracle's tensors (gauge-free)


def test_tebd_grid_config1(tne):
    """BASELINE.json configs[0]: 3x3, D = 2, 10 TEBD steps (9 sweeps), wavefront-batched on the device."""
    L, D = 3, 2
    g = OG.grid([L, L])
    op = oracle_peps(g, D, seed_for(L, D))
    dp = device_peps_like(tne, op, g, D)
    args = dict(t=0.3, C=10.0, delta=0.2, omega=0.3, nstep=10)
    OT.rydberg_tebd_(op, g, **args)
    tne.rydberg_tebd_(dp, tne.grid([L, L]), **args)
    assert rel(tne.vec(dp), OP.vec(op)) < 1e-8
    assert rel(tne.norm(dp).to_numpy(), OP.norm(op)) < 1e-8
    lv = tne.tebd.wavefronts(g.edges())
    assert sum(len(b) for b in lv) == g.ne() and all(len({s for e in b for s in e}) == 2 * len(b) for b in lv)


@pytest.mark.parametrize("kind", ["simple", "vidal"])
def test_tebd_sweep_from_product_state_exact_ranks(tne, kind):
    """TEBD from |0...0> (test/tebd.jl starts from random states; this is the data-dependent-rank case): the first
    sweeps keep FEWER singular values than the cap (absolute eps cut, src/peps.jl:383-390), so tne_tebd_sweep's
    speculative pass is rejected and the sweep is redone with exact ranks.  Bond dimensions, bond vectors and the state
    must match the oracle's strictly sequential loop."""
    L, D = 3, 2
    g = OG.grid([L, L])
    kw = dict(Dmax=4, eps=1e-10)
    f = OP.zero_vidalpeps if kind == "vidal" else OP.zero_simplepeps
    op = f(g, D, **kw)
    dp = device_peps_like(tne, op, g, D, **kw)
    args = dict(t=0.4, C=10.0, delta=0.2, omega=0.3, nstep=4)
    OP.LOG.clear(); tne.peps.LOG.clear()
    OT.rydberg_tebd_(op, g, **args)
    tne.rydberg_tebd_(dp, tne.grid([L, L]), **args)
    assert [t.shape for t in dp.vertex_tensors] == [np.shape(t) for t in op.vertex_tensors]
    assert rel(tne.vec(dp), OP.vec(op)) < 1e-8
    if kind == "vidal":
        for b in range(g.ne()):
            assert dp.bond_tensors[b].shape == np.shape(op.bond_tensors[b])
            assert rel(dp.bond_tensors[b].to_numpy(), op.bond_tensors[b]) < 1e-8
    assert len(tne.peps.LOG) == len(OP.LOG)


@pytest.mark.parametrize("kind", ["simple", "vidal"])
def test_replica_batched_sweeps_match_single_sweeps(tne, kind):
    """R replica PEPS through ONE tne_tebd_sweep call per sweep (rydberg_tebd_batch_): the R copies of a gate share a
    wavefront launch, every replica ends with exactly the tensors of its own rydberg_tebd_ run (same kernel, same
    inputs), and a batched sweep issues as many launches as a single one."""
    L, D, R = 4, 3, 5
    g = tne.grid([L, L])
    mk = tne.rand_simplepeps if kind == "simple" else tne.rand_vidalpeps
    par = dict(t=0.09, C=10.0, delta=0.2, omega=0.3, nstep=4)
    singles = [mk(g, D, np.random.default_rng(100 + r)) for r in range(R)]
    batch = [mk(g, D, np.random.default_rng(100 + r)) for r in range(R)]
    ctx = tne.context()
    for q in singles:
        tne.rydberg_tebd_(q, g, **par)
    l0 = ctx.launch_count()
    tne.rydberg_tebd_(mk(g, D, np.random.default_rng(7)), g, **par)
    l_single = ctx.launch_count() - l0
    l0 = ctx.launch_count()
    tne.rydberg_tebd_batch_(batch, g, **par)
    l_batch = ctx.launch_count() - l0
    assert l_batch == l_single
    for a, b in zip(singles, batch):
        for x, y in zip(a.alltensors(), b.alltensors()):
            assert x.shape == y.shape and np.array_equal(x.to_numpy(), y.to_numpy())
    # product-state start: some bonds keep fewer singular values than the cap, the whole batch is redone with exact ranks
    zs = [(tne.zero_simplepeps if kind == "simple" else tne.zero_vidalpeps)(g, D) for _ in range(2)]
    z1 = (tne.zero_simplepeps if kind == "simple" else tne.zero_vidalpeps)(g, D)
    tne.rydberg_tebd_batch_(zs, g, **par)
    tne.rydberg_tebd_(z1, g, **par)
    for q in zs:
        for x, y in zip(q.alltensors(), z1.alltensors()):
            assert x.shape == y.shape and np.array_equal(x.to_numpy(), y.to_numpy())


@pytest.mark.parametrize("kind", ["simple", "vidal"])
def test_multi_sweep_call_matches_sweep_by_sweep(tne, kind):
    """tne_tebd_sweeps: rydberg_tebd!'s nstep - 1 sweeps enqueued back to back in ONE call (no synchronisation between sweeps,
    ranks verified once) give bit for bit the tensors, and the same truncation log, as one call per sweep -- for a random
    state with growing bonds (every sweep reaches its cap) and from a product state (early sweeps keep fewer singular values:
    the speculative chain is cut, that sweep redone exactly and the rest enqueued again)."""
    L, D = 3, 2
    g = tne.grid([L, L])
    par = dict(C=10.0, delta=0.2, omega=0.3)
    for start in ("random", "product"):
        kw = dict(Dmax=6, eps=1e-10)
        if start == "random":
            mk = lambda: (tne.rand_simplepeps if kind == "simple" else tne.rand_vidalpeps)(g, D, np.random.default_rng(5), **kw)
        else:
            mk = lambda: (tne.zero_simplepeps if kind == "simple" else tne.zero_vidalpeps)(g, D, **kw)
        a, b = mk(), mk()
        tne.peps.LOG.clear()
        for _ in range(5):
            tne.rydberg_tebd_sweep_(a, g, dt=0.3 / 6, **par)
        log_a = list(tne.peps.LOG)
        tne.peps.LOG.clear()
        tne.rydberg_tebd_(b, g, t=0.3, nstep=6, **par)          # the same 5 sweeps of dt = 0.3 / 6 in one library call
        log_b = list(tne.peps.LOG)
        assert [t.shape for t in a.alltensors()] == [t.shape for t in b.alltensors()], start
        for x, y in zip(a.alltensors(), b.alltensors()):
            assert np.array_equal(x.to_numpy(), y.to_numpy()), start
        assert log_a == log_b, start
