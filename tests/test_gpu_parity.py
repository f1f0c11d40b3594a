"""Parity of the CUDA path (through the C ABI, via the ctypes host mirror) against the oracle on identical
seeded inputs.  Tolerance: 1e-10 relative (BASELINE.json north_star) for energies, norms, gradients."""
import numpy as np
import pytest

from oracle import einsum as OE, graphs as OG, peps as OP, tebd as OT, tensorad as OAD, timeevolve as OTE, yao as OY
from helpers import (device_peps_like, oracle_peps, rel, seed_for, sweep_order, to_device_hamiltonian)

pytestmark = pytest.mark.gpu
TOL = 1e-10


def crand(rng, shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


# -- einsum (OMEinsum.einsum dispatch point) -------------------------------------------------------
@pytest.mark.parametrize("force_generic", [0, 1])
def test_binary_einsum_shapes(tne, force_generic):
    ctx = tne.context()
    ctx.set_option("force_generic_kernel", force_generic)
    rng = np.random.default_rng(0)
    cases = [
        ([[1, 2, 3, 4], [4, 3, 5, 1]], [(3, 4, 2, 5), (5, 2, 6, 3)], [5, 2, 1]),      # batch label 1
        ([[1, 2], [2, 3]], [(7, 5), (5, 9)], [1, 3]),                                   # plain matmul
        ([[1, 2, 3], [3, 2, 1]], [(4, 3, 5), (5, 3, 4)], []),                            # full contraction
        ([[1, 2], [3, 4]], [(3, 4), (2, 5)], [4, 1, 3, 2]),                              # outer product + permute
        ([[1, 2, 3, 4, 5, 6], [2, 7, 5]], [(4, 4, 4, 4, 4, 4), (4, 2, 4)], [1, 3, 4, 6, 7]),  # absorb shape
        ([[1, 2, 3], [1, 2, 4]], [(64, 64, 3), (64, 64, 5)], [3, 4]),                   # reduce shape (long K)
        # DMMA absorb kernel: big M, K legs at the fastest and a slow position, N appended (forward sweep step)
        ([[1, 2, 3, 4, 5, 6, 7, 8], [1, 9, 7, 10]], [(4, 4, 4, 4, 4, 4, 4, 2), (4, 2, 4, 4)], [2, 3, 4, 5, 6, 8, 9, 10]),
        # ... its reverse: K slow, N lands on the fastest output position
        ([[2, 3, 4, 5, 6, 8, 9, 10], [1, 9, 7, 10]], [(4, 4, 4, 4, 4, 2, 2, 4), (4, 2, 4, 4)], [1, 2, 3, 4, 5, 6, 7, 8]),
        # odd dimensions (D = 3): padded k-steps and n-tiles
        ([[1, 2, 3, 4, 5, 6, 7, 8], [1, 9, 7, 10]], [(3, 3, 3, 3, 3, 3, 3, 2), (3, 2, 3, 3)], [2, 3, 4, 5, 6, 8, 9, 10]),
        ([[1, 2, 3, 4], [3, 5]], [(16, 16, 3, 16), (3, 5)], [1, 2, 5, 4]),
        # DMMA reduce kernel: leaf gradient shape (tiny output, long reduction)
        ([[1, 2, 3, 4, 5], [1, 2, 3, 6, 7]], [(32, 32, 16, 4, 2), (32, 32, 16, 4, 4)], [4, 5, 6, 7]),
        ([[4, 1, 2, 3, 5], [1, 6, 2, 3, 7]], [(4, 32, 32, 16, 2), (32, 3, 32, 16, 3)], [6, 7, 5, 4]),
        ([[1, 2, 3], [1, 2, 4]], [(128, 80, 3), (128, 80, 5)], [3, 4]),
        ([[1, 2, 3]], [(3, 4, 5)], [3, 1, 2]),                                           # unary permute
        ([[1, 2, 3]], [(3, 4, 5)], [2]),                                                 # unary sum
    ]
    for ixs, shapes, iy in cases:
        xs = [crand(rng, s) for s in shapes]
        ref = OE.einsum_raw(ixs, xs, iy)
        for conj in ([0] * len(xs), [1] * len(xs)):
            got = tne.einsum.einsum(ixs, [tne.B200Array.from_numpy(x) for x in xs], iy, conj=conj).to_numpy()
            want = ref if not conj[0] else OE.einsum_raw(ixs, [np.conj(x) for x in xs], iy)
            assert rel(got, want) < 1e-13, (ixs, iy, conj)
    ctx.set_option("force_generic_kernel", 0)


def test_elementwise_ops(tne):
    rng = np.random.default_rng(1)
    a = crand(rng, (5, 7)) + 2
    x = tne.B200Array.from_numpy(a)
    assert rel(x.conj().to_numpy(), np.conj(a)) < 1e-15
    assert rel(x.abs().to_numpy(), np.abs(a)) < 1e-15
    assert rel(x.sqrt().to_numpy(), np.sqrt(a)) < 1e-14
    assert rel(x.inv().to_numpy(), 1 / a) < 1e-14
    assert rel(x.pow(1 / 9).to_numpy(), a ** (1 / 9)) < 1e-14
    assert rel((x * (2 - 1j)).to_numpy(), a * (2 - 1j)) < 1e-15
    assert rel((x + x).to_numpy(), 2 * a) < 1e-15
    assert rel(x.slice(3, 10).to_numpy(), a.reshape(-1, order="F")[3:13]) < 1e-15


# -- the reference's test graph (test/peps.jl) ------------------------------------------------------
@pytest.mark.parametrize("kind", ["simple", "vidal"])
def test_statevec_inner_norm(tne, kind):
    g = OG.test_graph()
    op1, op2 = oracle_peps(g, 2, 11, kind, Dmax=4), oracle_peps(g, 2, 12, kind, Dmax=4)
    dp1, dp2 = device_peps_like(tne, op1, g, 2, Dmax=4), device_peps_like(tne, op2, g, 2, Dmax=4)
    assert rel(tne.vec(dp1), OP.vec(op1)) < TOL
    assert rel(tne.inner_product(dp1, dp2).to_numpy(), OP.inner_product(op1, op2)) < TOL
    assert rel(tne.norm(dp1).to_numpy(), OP.norm(op1)) < TOL
    tne.normalize_(dp1)
    assert abs(tne.norm(dp1).item() - 1) < 1e-12


def test_known_answers_on_device(tne):  # test/peps.jl:11,17
    g = tne.graph_from_edges(5, OG.test_graph().edges())
    p = tne.zero_vidalpeps(g, 2, nslices=3)
    e1 = np.zeros(32, complex); e1[0] = 1
    assert np.allclose(tne.vec(p), e1)
    tne.apply_onsite_(p, 1, OY.X)
    e2 = np.zeros(32, complex); e2[1] = 1
    assert np.allclose(tne.vec(p), e2)


def test_expect_generic_blocks(tne):  # test/peps.jl:88-110 on the device (one-site-product terms fused)
    g = OG.test_graph()
    rng = np.random.default_rng(5)
    n = 5
    blocks = []
    for (i, j) in g.edges():
        blocks.append(rng.random() * OY.kron(n, (i, OY.Y), (j, OY.Y)))
    for i in g.vertices():
        blocks.append(OY.put(n, i, rng.random() * OY.X))
    h = OY.Add(*blocks)
    op1, op2 = oracle_peps(g, 2, 21, Dmax=4), oracle_peps(g, 2, 22, Dmax=4)
    dp1, dp2 = device_peps_like(tne, op1, g, 2, Dmax=4), device_peps_like(tne, op2, g, 2, Dmax=4)
    hd = to_device_hamiltonian(tne, h)
    ref = OP.expect(h, op1, op2)
    assert rel(tne.expect(hd, dp1, dp2).to_numpy(), ref) < TOL
    tne.einsum.FUSED_TREES = False
    try:
        assert rel(tne.expect(hd, dp1, dp2).to_numpy(), ref) < TOL
    finally:
        tne.einsum.FUSED_TREES = True


@pytest.mark.parametrize("fused", [True, False])
def test_fvec_test_graph(tne, fused):  # test/timeevolve.jl:5-35
    g = OG.test_graph()
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    hd = to_device_hamiltonian(tne, h)
    op1 = oracle_peps(g, 2, 11, Dmax=4)
    dp1 = device_peps_like(tne, op1, g, 2, Dmax=4)
    ref = OTE.fvec(op1, h).data
    tne.einsum.FUSED_TREES = fused
    try:
        got = tne.fvec(dp1, hd).to_numpy()
    finally:
        tne.einsum.FUSED_TREES = True
    assert rel(got, ref) < TOL


def test_apply_smatrix(tne):  # test/timeevolve.jl:50-67
    g = OG.test_graph()
    rng = np.random.default_rng(3)
    op1 = oracle_peps(g, 2, 11, Dmax=4)
    dp1 = device_peps_like(tne, op1, g, 2, Dmax=4)
    v = OP.variables(op1)
    x = crand(rng, v.size) / np.sqrt(2)
    ref = OTE.apply_smatrix(op1, OAD.DiffTensor(v.copy()), OAD.DiffTensor(v.copy()), OAD.DiffTensor(x, requires_grad=False)).data
    got = tne.apply_smatrix(dp1, tne.DiffTensor(v.copy()), tne.DiffTensor(v.copy()), tne.DiffTensor(x, requires_grad=False)).to_numpy()
    assert rel(got, ref) < 1e-9


def test_apply_smatrix_tree_level_vs_taped_and_oracle(tne):
    """The explicit tree-level metric-vector product (fused trees) equals the reference's reverse-over-reverse
    (op-by-op tape, einsum.FUSED_TREES = False) and the oracle, also for v1 != v2 and on a grid with segments."""
    rng = np.random.default_rng(4)
    g = OG.test_graph()
    op1, op2 = oracle_peps(g, 2, 11, Dmax=4), oracle_peps(g, 2, 12, Dmax=4)
    dp1 = device_peps_like(tne, op1, g, 2, Dmax=4)
    v1, v2 = OP.variables(op1), OP.variables(op2)
    x = crand(rng, v1.size) / np.sqrt(2)
    ref = OTE.apply_smatrix(op1, OAD.DiffTensor(v1.copy()), OAD.DiffTensor(v2.copy()), OAD.DiffTensor(x, requires_grad=False)).data
    got = tne.apply_smatrix(dp1, tne.DiffTensor(v1.copy()), tne.DiffTensor(v2.copy()), tne.DiffTensor(x, requires_grad=False)).to_numpy()
    assert rel(got, ref) < 1e-9
    tne.einsum.FUSED_TREES = False
    try:
        taped = tne.apply_smatrix(dp1, tne.DiffTensor(v1.copy()), tne.DiffTensor(v2.copy()), tne.DiffTensor(x, requires_grad=False)).to_numpy()
    finally:
        tne.einsum.FUSED_TREES = True
    assert rel(got, taped) < 1e-9
    L, D = 3, 2
    gg = OG.grid([L, L])
    order, segs = tne.sweep_plan(tne.grid([L, L]), L)
    op = oracle_peps(gg, D, seed_for(L, D), optimizer=sweep_order(L * L))
    dp = device_peps_like(tne, op, gg, D, optimizer=order, segments=segs)
    v = OP.variables(op)
    x = crand(rng, v.size) / np.sqrt(2)
    ref = OTE.apply_smatrix(op, OAD.DiffTensor(v.copy()), OAD.DiffTensor(v.copy()), OAD.DiffTensor(x, requires_grad=False)).data
    got = tne.apply_smatrix(dp, tne.DiffTensor(v.copy()), tne.DiffTensor(v.copy()), tne.DiffTensor(x, requires_grad=False)).to_numpy()
    assert rel(got, ref) < 1e-9
    # the memo of sr_step's GMRES loop (same v1 = v2 for every product): x-independent contractions run once
    memo = {}
    launches = []
    for xk in (x, crand(rng, v.size) / np.sqrt(2)):
        refk = OTE.apply_smatrix(op, OAD.DiffTensor(v.copy()), OAD.DiffTensor(v.copy()), OAD.DiffTensor(xk, requires_grad=False)).data
        l0 = tne.context().launch_count()
        gotk = tne.apply_smatrix(dp, tne.DiffTensor(v.copy()), tne.DiffTensor(v.copy()), tne.DiffTensor(xk, requires_grad=False), memo=memo).to_numpy()
        launches.append(tne.context().launch_count() - l0)
        assert rel(gotk, refk) < 1e-9
    assert set(memo) == {"n2sq", "norm"} and launches[1] < launches[0]


def test_normalized_overlap_gradient_zero(tne):  # test/timeevolve.jl:37-48
    g = OG.test_graph()
    op1 = oracle_peps(g, 2, 11, Dmax=4)
    dp1 = device_peps_like(tne, op1, g, 2, Dmax=4)
    v = OP.variables(op1)
    v1 = tne.DiffTensor(v.copy(), requires_grad=False)
    g1 = tne.gradient(lambda x: tne.normalized_overlap(dp1, v1, x), tne.DiffTensor(v.copy()))[0]
    assert np.abs(g1.to_numpy()).max() < 1e-8


# -- BASELINE.json configs 1 and 2: grids with the boundary-sweep tree -------------------------------
@pytest.mark.parametrize("L,D", [(3, 2), (4, 3)])
def test_grid_norm_energy_gradient(tne, L, D):
    ctx = tne.context()
    g = OG.grid([L, L])
    order = sweep_order(L * L)
    op = oracle_peps(g, D, seed_for(L, D), optimizer=order)
    dp = device_peps_like(tne, op, g, D, optimizer=order)
    assert rel(tne.norm(dp).to_numpy(), OP.norm(op)) < TOL
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    hd = to_device_hamiltonian(tne, h)
    assert rel(tne.expect(hd, dp, dp).to_numpy(), OP.expect(h, op, op)) < TOL
    fref = OTE.fvec(op, h).data
    assert rel(tne.fvec(dp, hd).to_numpy(), fref) < TOL
    y, f = tne.energy_and_gradient(dp, hd)
    assert rel(f.to_numpy(), fref) < TOL
    assert rel(y.to_numpy(), OTE.iloss2(h, op, OP.variables(op))) < TOL
    # globally sliced tree (SlicedEinsum, test/peps.jl:9): same numbers
    ds = device_peps_like(tne, op, g, D, optimizer=order, nslices=2)
    assert len(ds.code_inner_product.slicing) == 2
    assert rel(tne.norm(ds).to_numpy(), OP.norm(op)) < TOL
    assert rel(tne.fvec(ds, hd).to_numpy(), fref) < TOL
    # term sharding: partial energies / forces of 2 ranks add up
    parts = [tne.energy_and_gradient(dp, hd, shard=(r, 2)) for r in range(2)]
    assert rel(sum(p[1].to_numpy() for p in parts), fref) < TOL
    assert ctx.launch_count() > 0


@pytest.mark.parametrize("nbonds", [2, 3])
def test_multi_bond_segments(tne, nbonds):
    """sweep_plan with several sliced bonds per column (row-wise sub-segments; checkpoints that carry sliced labels
    are written and read through strided views): same numbers as the oracle."""
    L, D = 4, 3
    g = OG.grid([L, L])
    order, segs = tne.sweep_plan(tne.grid([L, L]), L, nbonds)
    assert len(segs) == 1 + (L - 1) * nbonds
    op = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))
    dp = device_peps_like(tne, op, g, D, optimizer=order, segments=segs)
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    hd = to_device_hamiltonian(tne, h)
    assert rel(tne.norm(dp).to_numpy(), OP.norm(op)) < TOL
    y, f = tne.energy_and_gradient(dp, hd)
    assert rel(f.to_numpy(), OTE.fvec(op, h).data) < TOL
    assert rel(y.to_numpy(), OTE.iloss2(h, op, OP.variables(op))) < TOL


def test_linearity_and_scaling_properties(tne):
    """size-independent properties: <a p|p> = conj(a) <p|p>; norm scales; gradient of a real loss is
    orthogonal to the radial direction after normalisation (iloss2 is scale-invariant in x)."""
    L, D = 4, 3
    g = tne.grid([L, L])
    rng = np.random.default_rng(9)
    p = tne.rand_simplepeps(g, D, rng, optimizer=sweep_order(L * L))
    h = tne.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    n0 = tne.norm(p).item().real
    q = tne.peps.mul(p, 2.0)
    assert abs(tne.norm(q).item().real / n0 - 2.0) < 1e-11
    f = tne.fvec(p, h).to_numpy()
    v = tne.variables(p).to_numpy()
    # L(s x) = L(x) for real s > 0  =>  Re <x, grad> = 0 with grad = i * fvec
    grad = 1j * f
    assert abs(np.real(np.vdot(v, grad))) < 1e-9 * np.linalg.norm(v) * np.linalg.norm(grad)


# -- checkpointed segments with segment-local slicing (the 6x6 D=4 execution plan) at oracle-checkable size --
@pytest.mark.parametrize("L,D", [(3, 2), (4, 3)])
def test_segmented_program_matches_oracle(tne, L, D):
    ctx = tne.context()
    g = OG.grid([L, L])
    gd = tne.grid([L, L])
    order, segs = tne.sweep_plan(gd, L)
    assert len(segs) == L and sum(2 if isinstance(t, tuple) else 1 for t in order) == 2 * L * L
    op = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))   # any order gives the same numbers
    dp = device_peps_like(tne, op, g, D, optimizer=order, segments=segs)
    assert dp.code_inner_product.segments is not None
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    hd = to_device_hamiltonian(tne, h)
    assert rel(tne.norm(dp).to_numpy(), OP.norm(op)) < TOL
    assert rel(tne.expect(hd, dp, dp).to_numpy(), OP.expect(h, op, op)) < TOL
    fref = OTE.fvec(op, h).data
    y, f = tne.energy_and_gradient(dp, hd)
    st = ctx.program_stats()
    assert rel(f.to_numpy(), fref) < TOL
    assert rel(y.to_numpy(), OTE.iloss2(h, op, OP.variables(op))) < TOL
    assert st["recompute_flops"] > 0  # the reverse pass did re-run segment forwards
    # segments + a globally sliced label + term sharding together
    ds = device_peps_like(tne, op, g, D, optimizer=order, segments=segs, nslices=1)
    parts = [tne.energy_and_gradient(ds, hd, shard=(r, 3)) for r in range(3)]
    assert rel(sum(p[1].to_numpy() for p in parts), fref) < TOL
    # gradient of the plain norm w.r.t. every leaf through the segmented reverse pass
    gref = OAD.gradient(lambda x: OP.norm(OP.load_variables(op, x)), OAD.DiffTensor(OP.variables(op)))[0].data
    gdev = tne.gradient(lambda x: tne.norm(tne.load_variables(dp, x)), tne.DiffTensor(tne.variables(dp)))[0].to_numpy()
    assert rel(gdev, gref) < TOL


def test_sr_step_matches_oracle(tne):
    """src/timeevolve.jl:1-4 as intended (GMRES on S x = -i f with the device metric-vector product): the same
    Krylov iteration on the host around device matvecs gives the oracle's update."""
    g = OG.test_graph()
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    hd = to_device_hamiltonian(tne, h)
    op1 = oracle_peps(g, 2, 11, Dmax=4)
    dp1 = device_peps_like(tne, op1, g, 2, Dmax=4)
    ref = OTE.sr_step(op1, h, tol=1e-8, maxiter=3)
    got = tne.sr_step(dp1, hd, tol=1e-8, maxiter=3)
    assert rel(tne.variables(got).to_numpy(), OP.variables(ref)) < 1e-6


def test_fused_site_absorption_d4(tne):
    """D = 4 grid: the bra and ket step of a bulk site run as one two-stage DMMA kernel (tne_site2.cu).  Same numbers
    as the step-by-step path (option no_site2) and as the oracle."""
    ctx = tne.context()
    L, D = 4, 4
    g = OG.grid([L, L])
    order, segs = tne.sweep_plan(tne.grid([L, L]), L)
    op = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    hd = to_device_hamiltonian(tne, h)
    nref, eref = OP.norm(op), OP.expect(h, op, op)
    for segments in (None, segs):
        dp = device_peps_like(tne, op, g, D, optimizer=order, segments=segments)
        ctx.set_option("profile", 1); ctx.profile_collect()
        n1 = tne.norm(dp).to_numpy()
        y1, f1 = tne.energy_and_gradient(dp, hd)
        used = ctx.profile_collect(fine=True)
        ctx.set_option("profile", 0)
        ctx.set_option("no_site2", 1)
        try:
            n0 = tne.norm(dp).to_numpy()
            y0, f0 = tne.energy_and_gradient(dp, hd)
        finally:
            ctx.set_option("no_site2", 0)
        assert rel(n1, nref) < TOL and rel(n0, nref) < TOL
        assert rel(tne.expect(hd, dp, dp).to_numpy(), eref) < TOL
        assert rel(y1.to_numpy(), y0.to_numpy()) < TOL and rel(f1.to_numpy(), f0.to_numpy()) < TOL
        if segments is None:
            assert any(k[0] == "gett_site2" for k in used), used   # the fused kernel did run
