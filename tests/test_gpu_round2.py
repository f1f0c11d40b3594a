"""Round-2 parity additions (all through the C ABI): the block types of ``expect`` the reference tests at
test/peps.jl:94-110, the rest of the TensorAD surface (trace / diag / repeat einsums, ``jacobian`` / ``hessian`` / ``cat``:
test/TensorAD/TensorAD.jl:22-39, test/TensorAD/rules.jl:13), a converged ``sr_step``, the error status of the bond
kernel, and the benchmark-size golden vectors of tests/golden/make_golden_big.py at 1e-10 through the BENCHMARK's plan
(boundary sweep, one checkpointed segment per column, segment-local slices, fused kernels on)."""
import os

import numpy as np
import pytest

from oracle import einsum as OE, graphs as OG, peps as OP, tebd as OT, tensorad as OAD, timeevolve as OTE, yao as OY
from helpers import device_peps_like, oracle_peps, rel, seed_for, sweep_order, to_device_hamiltonian

pytestmark = pytest.mark.gpu
TOL = 1e-10
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def crand(rng, shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


# -- expect with two-site put / control terms (test/peps.jl:88-110) ------------------------------------------------
def test_expect_put2_kron_control_terms(tne):
    g = OG.test_graph()
    n = 5
    rng = np.random.default_rng(31)
    blocks = []
    for (i, j) in g.edges():      # test/peps.jl:94: rand() * put(n, (i, j) => kron(Y, Y)) / kron / control, alternating
        c = rng.random()
        kind = len(blocks) % 3
        if kind == 0:
            blocks.append(c * OY.put(n, (i, j), OY.kron_mat(OY.Y, OY.Y)))
        elif kind == 1:
            blocks.append(c * OY.kron(n, (i, OY.Y), (j, OY.Y)))
        else:
            blocks.append(c * OY.control(n, i, j, OY.Y))
    for i in g.vertices():
        blocks.append(OY.put(n, i, rng.random() * OY.X))
    h = OY.Add(*blocks)
    for kind in ("simple", "vidal"):
        op1, op2 = oracle_peps(g, 2, 41, kind, Dmax=4), oracle_peps(g, 2, 42, kind, Dmax=4)
        dp1, dp2 = device_peps_like(tne, op1, g, 2, Dmax=4), device_peps_like(tne, op2, g, 2, Dmax=4)
        ref = OP.expect(h, op1, op2)
        got = tne.expect(to_device_hamiltonian(tne, h), dp1, dp2).to_numpy()
        assert rel(got, ref) < 1e-9, kind             # two-site terms go through apply! + SVD (src/peps.jl:456-460)
        # ... and against the dense state vector, like the reference does (test/peps.jl:109-110)
        from helpers import dense_apply
        dense = np.vdot(OP.vec(op1), dense_apply(h, OP.vec(op2)))
        assert abs(complex(got) - dense) < 1e-8 * abs(dense)


def test_expect_gradient_wrt_ket_is_not_dropped(tne):
    """<psi| H |psi(x)> differentiated w.r.t. the KET: the fused programme only has bra pullbacks, so this must take the
    term-by-term taped path (src/peps.jl:474-477 tapes apply -> inner_product for both arguments)."""
    g = OG.test_graph()
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    hd = to_device_hamiltonian(tne, h)
    op1 = oracle_peps(g, 2, 11, Dmax=4)
    dp1 = device_peps_like(tne, op1, g, 2, Dmax=4)
    v = OP.variables(op1)
    obra = OP.replace_tensors(op1, [OAD.DiffTensor(t, requires_grad=False) for t in op1.alltensors()])
    dbra = tne.replace_tensors(dp1, [tne.DiffTensor(t, requires_grad=False) for t in dp1.alltensors()])
    ref = OAD.gradient(lambda x: OAD.real(OP.expect(h, obra, OP.load_variables(op1, x))), OAD.DiffTensor(v.copy()))[0].data
    got = tne.gradient(lambda x: tne.tensorad.real(tne.expect(hd, dbra, tne.load_variables(dp1, x))), tne.DiffTensor(v.copy()))[0].to_numpy()
    assert np.abs(ref).max() > 0 and rel(got, ref) < TOL


def test_term_sharding_with_an_empty_share(tne):
    g = OG.test_graph()
    op1 = oracle_peps(g, 2, 11, Dmax=4)
    dp1 = device_peps_like(tne, op1, g, 2, Dmax=4)
    h = OY.Add(OY.put(5, 2, OY.X))
    hd = to_device_hamiltonian(tne, h)
    parts = [tne.expect(hd, dp1, dp1, shard=(r, 3)).to_numpy() for r in range(3)]
    assert parts[1] == 0 and parts[2] == 0            # NOT the bare overlap <p|p>
    assert rel(sum(parts), OP.expect(h, op1, op1)) < TOL


# -- OMEinsum unary rules with repeated labels and output-only labels ----------------------------------------------
def test_einsum_trace_diag_repeat(tne):
    rng = np.random.default_rng(2)
    a = crand(rng, (6, 6))
    b = crand(rng, (4, 5, 4))
    v = crand(rng, (5,))
    dev = tne.B200Array.from_numpy
    E = tne.einsum.einsum
    assert rel(E([[1, 1]], [dev(a)], []).to_numpy(), np.trace(a)) < 1e-14                        # ein"ii->"
    assert rel(E([[1, 1]], [dev(a)], [1]).to_numpy(), np.diag(a)) < 1e-15                        # ein"ii->i"
    assert rel(E([[1, 2, 1]], [dev(b)], [2]).to_numpy(), np.einsum("iji->j", b)) < 1e-14         # partial trace
    assert rel(E([[1]], [dev(v)], [1, 1]).to_numpy(), np.diag(v)) < 1e-15                        # ein"i->ii" (zeros off the diagonal)
    assert rel(E([[1, 2, 1], [2]], [dev(b), dev(v)], []).to_numpy(), np.einsum("iji,j->", b, v)) < 1e-14
    s = crand(rng, ())
    got = E([[]], [dev(s)], [3, 3], size_dict={3: 4}).to_numpy()                                 # ein"->ii": pullback of a trace
    assert rel(got, s * np.eye(4)) < 1e-15
    got = E([[1]], [dev(v)], [1, 7], size_dict={7: 3}).to_numpy()                                # repeat along a new label
    assert rel(got, np.repeat(v[:, None], 3, axis=1)) < 1e-15
    with pytest.raises(ValueError):
        E([[1]], [dev(v)], [1, 7])


# -- TensorAD: jacobian / hessian / cat (test/TensorAD/TensorAD.jl:22-39, test/TensorAD/rules.jl:13) ------------------
def _both(tne):
    """(module, einsum, DiffTensor) pairs for the oracle and the device so that one test body defines f once."""
    class O:
        ad = OAD
        ein = staticmethod(lambda ixs, xs, iy: OE.einsum(ixs, xs, iy))
        wrap = staticmethod(lambda a: OAD.DiffTensor(np.array(a)))
        out = staticmethod(lambda t: np.asarray(t.data if hasattr(t, "data") else t))

    class Dv:
        ad = tne.tensorad
        ein = staticmethod(lambda ixs, xs, iy: tne.einsum.einsum(ixs, xs, iy))
        wrap = staticmethod(lambda a: tne.DiffTensor(np.array(a)))
        out = staticmethod(lambda t: t.to_numpy())
    return O, Dv


@pytest.mark.parametrize("dtype", ["real", "complex"])
def test_hessians_match_the_oracle(tne, dtype):
    rng = np.random.default_rng(8)
    O, Dv = _both(tne)
    mk = (lambda *s: rng.standard_normal(s)) if dtype == "real" else (lambda *s: crand(rng, s))

    def f1(M):      # ein"ii->"(ein"ij,jk->ik"(x, y)) on the halves of one flat vector (ADTest.jl builds the same wrapper)
        return lambda xy: M.ein([[1, 1]], [M.ein([[1, 2], [2, 3]], [M.ad.reshape(xy[0:25], (5, 5)), M.ad.reshape(xy[25:50], (5, 5))], [1, 3])], [])

    def f2(M):      # sin.(sin.(sin.(xy))) on a 0-dim tensor
        return lambda xy: M.ad.bsin(M.ad.bsin(M.ad.bsin(xy)))

    def f3(M):      # ein"i->"((x .* y) .* z) with x = y = z = xy[1:1]
        return lambda xy: M.ein([[1]], [M.ad.bmul(M.ad.bmul(xy[0:1], xy[0:1]), xy[0:1])], [])
    for f, x0 in ((f1, mk(50)), (f2, np.asarray(mk())), (f3, mk(3))):
        if dtype == "complex" and f is f2:
            x0 = np.asarray(0.5 + 0.2j)
        ref = O.out(O.ad.hessian(f(O), O.wrap(x0)))
        got = Dv.out(Dv.ad.hessian(f(Dv), Dv.wrap(x0)))
        assert got.shape == ref.shape and rel(got, ref) < 1e-12, f.__name__


def test_jacobian_of_cat_and_hcat(tne):
    rng = np.random.default_rng(9)
    O, Dv = _both(tne)
    x0 = crand(rng, (24,))

    def f(M):   # cat(reshape(x[1:16], 4, 4), reshape(x[17:24], 4, 2); dims = 2), then a contraction that mixes the columns
        def g(x):
            c = M.ad.cat([M.ad.reshape(x[0:16], (4, 4)), M.ad.reshape(x[16:24], (4, 2))], 2)
            return M.ein([[1, 2], [3, 2]], [c, c], [1, 3])
        return g
    ref = O.out(O.ad.jacobian(f(O), O.wrap(x0)))
    got = Dv.out(Dv.ad.jacobian(f(Dv), Dv.wrap(x0)))
    assert rel(got, ref) < 1e-12
    # cat along the FIRST axis of matrices (the rotated path) and hcat / vcat of vectors
    def f_rows(M):
        return lambda x: M.ad.cat([M.ad.reshape(x[0:8], (2, 4)), M.ad.reshape(x[8:24], (4, 4))], 1)
    assert rel(Dv.out(f_rows(Dv)(Dv.wrap(x0))), np.concatenate([x0[:8].reshape(2, 4, order="F"), x0[8:].reshape(4, 4, order="F")], 0)) < 1e-15
    assert rel(Dv.out(Dv.ad.jacobian(f_rows(Dv), Dv.wrap(x0))), O.out(O.ad.jacobian(f_rows(O), O.wrap(x0)))) < 1e-12
    h = Dv.ad.hcat(Dv.wrap(x0[:6]), Dv.wrap(x0[6:12]))
    assert h.shape == (6, 2) and rel(h.to_numpy(), np.stack([x0[:6], x0[6:12]], 1)) < 1e-15
    assert rel(Dv.ad.vcat(Dv.wrap(x0[:6]), Dv.wrap(x0[6:9])).to_numpy(), x0[:9]) < 1e-15


# -- a converged TDVP step (src/timeevolve.jl:1-4) on a grid ---------------------------------------------------------
def test_sr_step_converged_on_a_grid(tne):
    L, D = 3, 2
    g = OG.grid([L, L])
    order, segs = tne.sweep_plan(tne.grid([L, L]), L)
    op = oracle_peps(g, D, seed_for(L, D), optimizer=sweep_order(L * L))
    dp = device_peps_like(tne, op, g, D, optimizer=order, segments=segs)
    h = OT.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    ref = OTE.sr_step(op, h, tol=1e-12, maxiter=40)
    got = tne.sr_step(dp, to_device_hamiltonian(tne, h), tol=1e-12, maxiter=40)
    assert rel(tne.variables(got).to_numpy(), OP.variables(ref)) < 1e-8


# -- bond kernel: non-finite input is an error status, not a fault ----------------------------------------------------
def test_bond_update_reports_non_finite_input(tne):
    g = OG.test_graph()
    op = oracle_peps(g, 2, 5, Dmax=4)
    bad = [np.array(t) for t in op.alltensors()]
    bad[0][(0,) * bad[0].ndim] = np.nan
    dp = tne.peps.from_numpy_tensors(device_peps_like(tne, op, g, 2, Dmax=4), bad)
    gate = OY.time_evolve(OT.bond_hamiltonian(10.0, 0.2, 0.3, 2, 3), 0.03).reshape(2, 2, 2, 2, order="F")
    with pytest.raises(tne.TneError, match="non-finite"):
        tne.apply_onbond_(dp, 1, 2, gate)
    # the context is still alive and correct afterwards
    dq = device_peps_like(tne, op, g, 2, Dmax=4)
    tne.apply_onbond_(dq, 1, 2, gate)
    OP.apply_onbond_(op, 1, 2, gate)
    assert rel(tne.vec(dq), OP.vec(op)) < 1e-10


# -- benchmark-size golden vectors ------------------------------------------------------------------------------------
def _gold(name):
    path = os.path.join(GOLD, name)
    if not os.path.exists(path):
        pytest.skip("%s not generated (tests/golden/make_golden_big.py)" % name)
    return np.load(path)


def _bench_peps(tne, L, D):
    """The benchmark's PEPS and plan for an L x L lattice (bench.py): same seed rule, sweep_plan segments."""
    g = tne.grid([L, L])
    order, segs = tne.sweep_plan(g, L)
    p = tne.rand_simplepeps(g, D, np.random.default_rng(seed_for(L, D)), optimizer=order, segments=segs)
    return g, p


@pytest.mark.parametrize("L,D", [(4, 4), (5, 4), (6, 3)])
def test_full_energy_and_force_golden(tne, L, D):
    gold = _gold("grid_%dx%d_D%d.npz" % (L, L, D))
    if "fvec" not in gold:
        pytest.skip("force not generated yet")
    g, p = _bench_peps(tne, L, D)
    h = tne.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    assert rel(tne.variables(p).to_numpy(), gold["variables"]) == 0.0
    assert rel(tne.inner_product(p, p).to_numpy(), gold["inner"]) < TOL
    assert rel(tne.expect(h, p, p).to_numpy(), gold["energy"]) < TOL
    y, f = tne.energy_and_gradient(p, h)
    assert rel(y.to_numpy(), gold["iloss2"]) < TOL
    assert rel(f.to_numpy(), gold["fvec"]) < TOL
    assert rel(tne.fvec(p, h).to_numpy(), gold["fvec"]) < TOL


@pytest.mark.parametrize("L,D", [(4, 4), (5, 4)])
def test_smatrix_golden(tne, L, D):
    gold = _gold("grid_%dx%d_D%d.npz" % (L, L, D))
    if "smatrix_x" not in gold:
        pytest.skip("S.x not generated yet")
    g, p = _bench_peps(tne, L, D)
    v = gold["variables"]
    sx = tne.apply_smatrix(p, tne.DiffTensor(v.copy()), tne.DiffTensor(v.copy()), tne.DiffTensor(gold["x"], requires_grad=False)).to_numpy()
    # "tape": the oracle's literal reverse-over-reverse; "fd": Richardson central differences of the oracle gradient
    tol = 1e-9 if str(gold["smatrix_method"]) == "tape" else 2e-8
    assert rel(sx, gold["smatrix_x"]) < tol


def test_6x6_d4_against_the_oracle(tne):
    """The headline configuration itself.  Oracle values (forward only: one 6x6 D=4 tree is 2.4e12 flop on the CPU): the norm,
    a structurally representative subset of the 96 terms -- energies <psi|h_k|psi> and iloss2's cores psi^T h_k psi -- and
    directional derivatives of iloss2 restricted to that subset along one-site directions, all through the benchmark's
    plan with the same term-DP / segments / slices / fused kernels."""
    gold = _gold("grid_6x6_D4_partial.npz")
    L, D = 6, 4
    g, p = _bench_peps(tne, L, D)
    assert rel(tne.variables(p).to_numpy(), gold["variables"]) == 0.0
    h = tne.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    inner = complex(tne.inner_product(p, p).to_numpy())
    assert abs(inner - complex(gold["inner"])) < TOL * abs(complex(gold["inner"]))
    ks = [int(k) for k in gold["term_index"]]
    nE = len(gold["energy_terms"])
    if nE == 0:
        pytest.skip("no term energies generated yet")
    sub = tne.yao.Add(*[h.blocks[k] for k in ks[:nE]])
    e_sub = complex(tne.expect(sub, p, p).to_numpy())
    ref = complex(np.sum(gold["energy_terms"]))
    assert abs(e_sub - ref) < TOL * np.sum(np.abs(gold["energy_terms"]))
    for k, want in zip(ks[:nE], gold["energy_terms"]):      # every term alone as well (different DP patterns)
        got = complex(tne.expect(tne.yao.Add(h.blocks[k]), p, p).to_numpy())
        assert abs(got - complex(want)) < TOL * abs(complex(want)), k
    # iloss2 = Re sum_k core_k / <psi|psi> at x = variables(psi)   (src/timeevolve.jl:55-60)
    y, f = tne.energy_and_gradient(p, sub)
    n2 = abs(complex(gold["inner"]))
    want_y = float(np.real(np.sum(gold["iloss_core_terms"][:nE]))) / n2
    assert abs(float(np.real(y.to_numpy())) - want_y) < TOL * np.sum(np.abs(gold["iloss_core_terms"][:nE])) / n2
    # directional derivatives: dL = dG / n2 - G <dN> / (2 n2^2),  G = Re sum core, dN = 2 Re <psi[s -> V]|psi>
    if nE < len(ks):
        return
    grad = 1j * f.to_numpy()                                 # fvec = -i grad
    sizes = [t.size for t in p.alltensors()]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    G = float(np.real(np.sum(gold["iloss_core_terms"])))
    for s in [int(s) for s in gold["dir_sites"]]:
        V = gold["dir_V_s%02d" % s]
        dG = float(np.real(np.sum(gold["dir_dcore_s%02d" % s])))
        dN = 2.0 * float(np.real(complex(gold["dir_dinner_s%02d" % s])))
        want = dG / n2 - G * dN / (2.0 * n2 * n2)
        got = float(np.real(np.vdot(grad[offs[s - 1]:offs[s]], V)))
        scale = np.linalg.norm(grad[offs[s - 1]:offs[s]]) * np.linalg.norm(V)
        assert abs(got - want) < 1e-9 * scale, (s, got, want)


def test_6x6_d4_energy_of_every_generated_term(tne):
    """<psi|h_k|psi> of the benchmark PEPS for every term the oracle has contracted so far (grid_6x6_D4_energy.npz grows up to all
    96: 60 edge terms C P1 P1 and 36 site terms): their sum in one expect (one term-DP programme over the benchmark's plan) and
    eight of them alone."""
    gold = _gold("grid_6x6_D4_energy.npz")
    L, D = 6, 4
    g, p = _bench_peps(tne, L, D)
    assert rel(tne.variables(p).to_numpy(), gold["variables"]) == 0.0
    h = tne.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    ks = [int(k) for k in gold["term_index"]]
    assert len(h.blocks) == int(gold["n_terms"]) and len(ks) >= 7
    e_sub = complex(tne.expect(tne.yao.Add(*[h.blocks[k] for k in ks]), p, p).to_numpy())
    assert abs(e_sub - complex(np.sum(gold["energy_terms"]))) < TOL * np.sum(np.abs(gold["energy_terms"]))
    for q in sorted(set(np.linspace(0, len(ks) - 1, 8).astype(int).tolist())):
        got = complex(tne.expect(tne.yao.Add(h.blocks[ks[q]]), p, p).to_numpy())
        assert abs(got - complex(gold["energy_terms"][q])) < TOL * abs(complex(gold["energy_terms"][q])), ks[q]


def test_6x6_d4_iloss2_of_every_generated_core(tne):
    """iloss2 -- the quantity the benchmark evaluates and differentiates (src/timeevolve.jl:55-60: psi^T H psi_0 / <psi|psi>, the
    reference's double conjugation) -- restricted to the terms whose cores the oracle has contracted (grid_6x6_D4_iloss.npz, spread
    over edge and site terms of every lattice region), through energy_and_gradient on the benchmark's plan."""
    gold = _gold("grid_6x6_D4_iloss.npz")
    L, D = 6, 4
    g, p = _bench_peps(tne, L, D)
    assert rel(tne.variables(p).to_numpy(), gold["variables"]) == 0.0
    h = tne.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    ks = [int(k) for k in gold["term_index"]]
    assert len(ks) >= 7
    y, f = tne.energy_and_gradient(p, tne.yao.Add(*[h.blocks[k] for k in ks]))
    n2 = abs(complex(gold["inner"]))
    want = float(np.real(np.sum(gold["iloss_core_terms"]))) / n2
    assert abs(float(np.real(y.to_numpy())) - want) < TOL * np.sum(np.abs(gold["iloss_core_terms"])) / n2
    # the force of a normalised state is orthogonal to the radial direction (size-independent property of fvec)
    v = tne.variables(p).to_numpy()
    assert abs(np.vdot(v, 1j * f.to_numpy()).real) < 1e-9 * np.linalg.norm(v) * np.linalg.norm(f.to_numpy())


# -- fused reverse pass of a site (tne_site2b.cu) -------------------------------------------------------------------------
def test_fused_reverse_site_kernel(tne):
    """D = 4 grid with the full Rydberg H: the three reverse ops of a bulk site run as ONE kernel.  Same energy and force as
    the step-by-step reverse pass (option no_site2b) to rounding, the kernel did launch, and two runs agree bit for bit in
    the leaf gradients it produced (fixed-order reduction, no floating-point atomics in this kernel)."""
    ctx = tne.context()
    L, D = 4, 4
    g, p = _bench_peps(tne, L, D)
    h = tne.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    ctx.set_option("profile", 1); ctx.profile_collect()
    y1, f1 = tne.energy_and_gradient(p, h)
    used = ctx.profile_collect()
    ctx.set_option("profile", 0)
    assert used["gett_site2_bwd"]["launches"] > 0, used
    ctx.set_option("no_site2b", 1)
    try:
        y0, f0 = tne.energy_and_gradient(p, h)
    finally:
        ctx.set_option("no_site2b", 0)
    assert rel(y1.to_numpy(), y0.to_numpy()) < 1e-13
    assert rel(f1.to_numpy(), f0.to_numpy()) < 1e-12
    # the norm programme (gradient of <psi|psi> w.r.t. every bra leaf) through the same kernel, against the oracle
    op = oracle_peps(OG.grid([L, L]), D, seed_for(L, D), optimizer=sweep_order(L * L))
    gref = OAD.gradient(lambda x: OP.norm(OP.load_variables(op, x)), OAD.DiffTensor(OP.variables(op)))[0].data
    gdev = tne.gradient(lambda x: tne.norm(tne.load_variables(p, x)), tne.DiffTensor(tne.variables(p)))[0].to_numpy()
    assert rel(gdev, gref) < TOL


def test_gradients_are_bit_reproducible(tne):
    """No floating-point atomics on the gradient path (fixed-order two-pass reductions in the reduce kernels and in the
    fused reverse kernel): two evaluations of the same programme give bit-identical energies and forces."""
    for L, D in ((4, 4), (4, 3)):
        g, p = _bench_peps(tne, L, D)
        h = tne.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
        y1, f1 = tne.energy_and_gradient(p, h)
        y2, f2 = tne.energy_and_gradient(p, h)
        assert np.array_equal(y1.to_numpy(), y2.to_numpy())
        assert np.array_equal(f1.to_numpy(), f2.to_numpy())


# -- general dense nodes: tiled complex128 DMMA GEMM (tne_gemm.cu) -----------------------------------------------------------
def test_gemm_kernel_random_shapes(tne):
    """Contractions with M, N and K all large and / or batch legs (greedy / TreeSA tree nodes, the paired boundary tree) run on
    the tiled DMMA GEMM: random label orders, odd sizes, batch legs, conjugations, against numpy at 1e-13."""
    ctx = tne.context()
    rng = np.random.default_rng(17)
    cases = [
        ([[1, 2], [2, 3]], [(300, 70), (70, 50)], [1, 3]),                                         # plain, ragged tiles
        ([[1, 2], [2, 3]], [(1024, 256), (256, 256)], [3, 1]),                                     # paired-tree shape, transposed output
        ([[2, 1], [3, 2]], [(96, 200), (40, 96)], [1, 3]),                                         # both operands transposed
        ([[1, 5, 2, 6], [6, 3, 5, 4]], [(16, 6, 9, 10), (10, 8, 6, 5)], [4, 1, 2, 3]),             # two K legs, split M / N, permuted output
        ([[7, 1, 2], [2, 7, 3]], [(5, 64, 48), (48, 5, 33)], [1, 7, 3]),                           # batch leg 7 in the middle of the output
        ([[1, 2, 3, 4], [4, 3, 5, 1]], [(6, 40, 9, 11), (11, 9, 36, 6)], [5, 2, 1]),               # batch label 1 slowest / fastest
        ([[1, 2, 3, 4, 5], [3, 6, 5, 7]], [(4, 4, 4, 4, 4), (4, 4, 4, 4)], [1, 2, 4, 6, 7]),       # D = 4 legs only (M 64, N 16, K 16)
        ([[1, 2, 3], [3, 4]], [(27, 27, 81), (81, 27)], [4, 2, 1]),                                # D = 3 powers
    ]
    ctx.set_option("profile", 1); ctx.profile_collect()
    for ixs, shapes, iy in cases:
        xs = [crand(rng, s) for s in shapes]
        for conj in ([0, 0], [1, 0], [0, 1], [1, 1]):
            want = OE.einsum_raw(ixs, [np.conj(x) if c else x for x, c in zip(xs, conj)], iy)
            got = tne.einsum.einsum(ixs, [tne.B200Array.from_numpy(x) for x in xs], iy, conj=conj).to_numpy()
            assert rel(got, want) < 1e-13, (ixs, iy, conj)
    used = ctx.profile_collect()
    ctx.set_option("profile", 0)
    assert used["gett_gemm"]["launches"] >= 16, used       # the tiny cases may stay on the FMA kernel (latency-bound anyway)
    # the same contractions on the FMA kernel (option no_gemm): the two kernels agree to rounding
    ctx.set_option("no_gemm", 1)
    try:
        ixs, shapes, iy = cases[3]
        xs = [crand(rng, s) for s in shapes]
        ref = tne.einsum.einsum(ixs, [tne.B200Array.from_numpy(x) for x in xs], iy).to_numpy()
    finally:
        ctx.set_option("no_gemm", 0)
    assert rel(tne.einsum.einsum(ixs, [tne.B200Array.from_numpy(x) for x in xs], iy).to_numpy(), ref) < 1e-13


def test_greedy_and_paired_trees_run_on_dmma_kernels(tne):
    """Trees other than the hand-built sweep on a 5x5 D=4 PEPS, against the ORACLE golden <psi|psi>:
    * the reference's DEFAULT contraction order (GreedyMethod, src/peps.jl:244-245): 1.1e12 flop, nodes like
      4096 x 65536 x 256 and 256 x 4096 x 65536;
    * the fully paired boundary tree (bra . ket per site first, then N = K = 256 absorptions).
    Both match at 1e-10 and the FMA kernel gets < 5 % of the time: the heavy nodes run on the DMMA GEMM / absorb / reduce kernels."""
    gold = _gold("grid_5x5_D4.npz")
    ctx = tne.context()
    L, D = 5, 4
    g = tne.grid([L, L])
    n = L * L
    base = tne.rand_simplepeps(g, D, np.random.default_rng(seed_for(L, D)), optimizer=sweep_order(n))
    for name, opt in (("greedy", "greedy"), ("paired", [(i, n + i) for i in range(n)])):
        p = tne.replace_tensors(tne.zero_simplepeps(g, D, optimizer=opt), base.alltensors())
        tne.inner_product(p, p)                          # warm-up (workspace)
        ctx.set_option("profile", 1); ctx.profile_collect()
        got = tne.inner_product(p, p).to_numpy()
        used = ctx.profile_collect()
        ctx.set_option("profile", 0)
        assert rel(got, gold["inner"]) < TOL, name
        total = sum(v["ms"] for v in used.values())
        shares = {k: (v["launches"], round(v["ms"], 2), round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 1)) for k, v in used.items() if v["launches"]}
        assert used["gett_gemm"]["launches"] > 0, (name, shares)
        assert used["gett_generic"]["ms"] < 0.05 * total, (name, shares)     # ~30 latency-bound launches of tiny operands (15 us each)


# -- tensor permutation kernel (tne_permute.cu) -----------------------------------------------------------------------------
def test_permutation_kernel(tne):
    ctx = tne.context()
    rng = np.random.default_rng(23)
    cases = [((3, 4, 5), [3, 1, 2]), ((64, 64), [2, 1]), ((4,) * 8, [8, 3, 1, 6, 2, 7, 4, 5]), ((2, 4, 4, 4, 4, 4), [6, 5, 4, 3, 2, 1]),
             ((3, 9, 27, 3), [2, 4, 1, 3]), ((100, 7, 33), [1, 3, 2]), ((81, 50), [2, 1]), ((5,), [1]), ((16384, 6), [2, 1]),
             ((2, 3, 2, 3, 2, 3, 2), [7, 1, 6, 2, 5, 3, 4])]
    l0 = ctx.launch_count()
    for shape, perm in cases:
        x = crand(rng, shape)
        labels = list(range(1, len(shape) + 1))
        for conj in (0, 1):
            got = tne.einsum.einsum([labels], [tne.B200Array.from_numpy(x)], perm, conj=[conj]).to_numpy()
            want = np.transpose(np.conj(x) if conj else x, [p - 1 for p in perm])
            assert got.shape == want.shape and np.array_equal(got, want), (shape, perm)
    assert ctx.launch_count() - l0 == 2 * len(cases)       # one launch each: no zero-fill, no contraction


# -- big lattices: two-bond sliced sweep (BASELINE configs[3], the exact 8x8 D=4 norm) and globally sliced sums (configs[4]) -----
def _row_bonds(tne, L, rows):
    g = tne.grid([L, L])
    n = L * L
    emap = {e: n + 1 + k for k, e in enumerate(g.edges())}
    return [emap[(r + (y - 2) * L, r + (y - 1) * L)] for r in rows for y in range(2, L + 1)]


@pytest.mark.parametrize("L,D,name", [(4, 3, "grid_4x4_D3.npz"), (5, 4, "grid_5x5_D4.npz"), (8, 2, "grid_8x8_D2_norm.npz"),
                                      (8, 3, "grid_8x8_D3_norm.npz"), (10, 2, "grid_10x10_D2_norm.npz")])
def test_two_bond_sliced_sweep_norm_golden(tne, L, D, name):
    """sweep_plan(slice_first_row=True) -- the plan that fits the exact 8x8 D=4 norm (64 GiB boundaries) into one B200 --
    against the oracle's <psi|psi> at the validation sizes: every column loops over the bond entering its last site AND the
    bond leaving its first site; the next boundary is written slice by slice through strided views."""
    gold = _gold(name)
    g = tne.grid([L, L])
    order, segs = tne.sweep_plan(g, L, slice_first_row=True)
    p = tne.rand_simplepeps(g, D, np.random.default_rng(seed_for(L, D)), optimizer=order, segments=segs)
    assert np.array_equal(tne.variables(p).to_numpy(), gold["variables"])
    assert rel(tne.inner_product(p, p).to_numpy(), gold["inner"]) < TOL


def test_globally_sliced_inner_product_sums_to_the_golden(tne):
    """inner_product_sliced (BASELINE configs[4]: sliced bond indices, partial sums per rank): the sum over ALL values of the
    horizontal bonds of one lattice row equals the oracle's <psi|psi>; dealing the terms to 3 'ranks' partitions the sum."""
    gold = _gold("grid_3x3_D2.npz")
    L, D = 3, 2
    g = tne.grid([L, L])
    order, segs = tne.sweep_plan(g, L)
    p = tne.rand_simplepeps(g, D, np.random.default_rng(seed_for(L, D)), optimizer=order, segments=segs)
    want = complex(tne.inner_product(p, p).to_numpy())
    if "inner" in gold:
        assert rel(want, gold["inner"]) < TOL
    bonds = _row_bonds(tne, L, [2])
    total, n = tne.inner_product_sliced(p, p, bonds)
    assert n == (D * D) ** len(bonds) and rel(total, want) < TOL
    parts = [tne.inner_product_sliced(p, p, bonds, shard=(r, 3)) for r in range(3)]
    assert sum(q[1] for q in parts) == n and rel(sum(q[0] for q in parts), want) < TOL
    # a slice is a PEPS with dimension-1 legs; two different bra / ket slices give one off-diagonal term
    one = tne.slice_bonds(p, bonds, (1, 0))
    assert sum(1 in t.shape for t in one.vertex_tensors) == L


def test_plan_cache_rebinds_buffers(tne):
    """tne_expect_terms caches compiled programs: a second PEPS of the same shape (different values, different buffers)
    is served from the cache and matches a freshly planned evaluation bit for bit; a different Hamiltonian misses."""
    ctx = tne.context()
    L, D = 4, 3
    g = tne.grid([L, L])
    order, segs = tne.sweep_plan(g, L)
    h = tne.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    pa = tne.rand_simplepeps(g, D, np.random.default_rng(1), optimizer=order, segments=segs)
    pb = tne.rand_simplepeps(g, D, np.random.default_rng(2), optimizer=order, segments=segs)
    ctx.set_option("no_plan_cache", 1)
    try:
        ya, fa = tne.energy_and_gradient(pa, h)
        yb, fb = tne.energy_and_gradient(pb, h)
        ya, fa, yb, fb = ya.to_numpy(), fa.to_numpy(), yb.to_numpy(), fb.to_numpy()
        assert ctx.plan_cache_info()[0] == 0
    finally:
        ctx.set_option("no_plan_cache", 0)
    _, h0 = ctx.plan_cache_info()
    y1, f1 = tne.energy_and_gradient(pa, h)         # plans (norm programme + energy programme)
    n1, h1 = ctx.plan_cache_info()
    y2, f2 = tne.energy_and_gradient(pb, h)         # same shapes, other buffers: served from the cache
    n2, h2 = ctx.plan_cache_info()
    y3, f3 = tne.energy_and_gradient(pa, h)
    assert n1 >= 2 and n2 == n1 and h1 == h0 and h2 - h1 >= 2
    assert np.array_equal(y1.to_numpy(), ya) and np.array_equal(f1.to_numpy(), fa)
    assert np.array_equal(y2.to_numpy(), yb) and np.array_equal(f2.to_numpy(), fb)
    assert np.array_equal(y3.to_numpy(), ya) and np.array_equal(f3.to_numpy(), fa)
    assert not np.array_equal(fa, fb)
    h2_ = tne.rydberg_hamiltonian(g, 7.0, 0.2, 0.3)   # other coefficients: a different programme
    tne.energy_and_gradient(pa, h2_)
    assert ctx.plan_cache_info()[0] > n2


def test_tdvp_drivers_reject_vidal_like_the_reference(tne):
    """``load_variables`` (the AD version) exists for SimplePEPS only (src/peps.jl:212): ``iloss2`` / ``fvec`` /
    ``normalized_overlap`` / ``apply_smatrix`` on a VidalPEPS are MethodErrors in the reference and TypeErrors here -- there is
    no Vidal metric-vector product to accelerate."""
    g = OG.test_graph()
    op1 = oracle_peps(g, 2, 21, kind="vidal", Dmax=4)
    dp1 = device_peps_like(tne, op1, g, 2, Dmax=4)
    assert dp1.kind == "vidal"
    v = OP.variables(op1)
    with pytest.raises(TypeError):
        tne.apply_smatrix(dp1, tne.DiffTensor(v.copy()), tne.DiffTensor(v.copy()), tne.DiffTensor(v.copy(), requires_grad=False))
    with pytest.raises(TypeError):
        tne.normalized_overlap(dp1, tne.DiffTensor(v.copy()), tne.DiffTensor(v.copy()))


def test_treesa_sliced_codes_on_device(tne):
    """The reference's ``optimizer = TreeSA(nslices = 3)`` (test/peps.jl:9): annealed contraction orders with a SlicedEinsum
    around them, for the state-tensor code and the inner-product code.  The known answers of test/peps.jl:9-19 on the test
    graph, and <psi|psi> of the 4x4 D=3 golden PEPS through a TreeSA tree with 2 sliced labels."""
    gt = tne.graph_from_edges(5, OG.test_graph().edges())
    p = tne.zero_vidalpeps(gt, 2, optimizer="treesa", nslices=3)
    assert isinstance(p.code_inner_product, tne.SlicedEinsum) and len(p.code_inner_product.slicing) == 3
    x = np.zeros(32, dtype=complex); x[0] = 1
    assert np.allclose(tne.vec(p), x)
    tne.apply_onsite_(p, 1, np.array([[0, 1], [1, 0]], dtype=complex))
    x = np.zeros(32, dtype=complex); x[1] = 1
    assert np.allclose(tne.vec(p), x)
    q = tne.rand_vidalpeps(gt, 2, np.random.default_rng(3), optimizer="treesa", nslices=3)
    tne.normalize_(q)
    assert abs(tne.norm(q).item() - 1) < 1e-12
    gold = _gold("grid_4x4_D3.npz")
    L, D = 4, 3
    g = tne.grid([L, L])
    ps = tne.rand_simplepeps(g, D, np.random.default_rng(seed_for(L, D)), optimizer="treesa", nslices=2)
    assert np.array_equal(tne.variables(ps).to_numpy(), gold["variables"])
    assert rel(tne.inner_product(ps, ps).to_numpy(), gold["inner"]) < TOL
    h = tne.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    assert rel(tne.expect(h, ps, ps).to_numpy(), gold["energy"]) < TOL


def test_tape_retention_is_bit_identical(tne):
    """Forward tapes of as many (segment, slice) units as fit in memory are kept for the reverse pass instead of being
    recomputed (option tape_gib: -1 auto, 0 off): the same kernels on the same inputs, so energy and force do not change
    by a bit, and fewer contractions run."""
    ctx = tne.context()
    L, D = 4, 4
    g = tne.grid([L, L])
    order, segs = tne.sweep_plan(g, L)
    p = tne.rand_simplepeps(g, D, np.random.default_rng(seed_for(L, D)), optimizer=order, segments=segs)
    h = tne.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    out = {}
    for gib in (0, -1, 1):
        ctx.set_option("tape_gib", gib)
        try:
            y, f = tne.energy_and_gradient(p, h)
            out[gib] = (y.to_numpy(), f.to_numpy(), ctx.program_stats())
        finally:
            ctx.set_option("tape_gib", -1)
    for gib in (-1, 1):
        assert np.array_equal(out[gib][0], out[0][0]) and np.array_equal(out[gib][1], out[0][1])
    # everything fits at this size: only the slice-independent crumbs (tiny leaf-level ops) are still recomputed
    assert out[-1][2]["recompute_flops"] < 1e-3 * out[0][2]["recompute_flops"] and out[0][2]["recompute_flops"] > 0.2 * out[0][2]["flops"]
    assert out[-1][2]["contractions"] < out[0][2]["contractions"]
    assert out[-1][2]["flops"] == out[0][2]["flops"]


def test_automatic_sweep_plan_on_device(tne):
    """optimizer="sweep" with segments="auto" / "auto2" (order, checkpointed segments and free slices derived from the network,
    no hand-built plan): energy + force of the 4x4 D=3 golden PEPS and <psi|psi> of the 8x8 D=2 / 5x5 D=4 goldens; a
    rectangular 3x5 lattice against the reference-default greedy tree."""
    gold = _gold("grid_4x4_D3.npz")
    L, D = 4, 3
    g = tne.grid([L, L])
    p = tne.rand_simplepeps(g, D, np.random.default_rng(seed_for(L, D)), optimizer="sweep", segments="auto")
    assert len(p.code_inner_product.segments) == L
    h = tne.rydberg_hamiltonian(g, 10.0, 0.2, 0.3)
    y, f = tne.energy_and_gradient(p, h)
    assert rel(y.to_numpy(), gold["iloss2"]) < TOL and rel(f.to_numpy(), gold["fvec"]) < TOL
    for (L2, D2, name) in ((8, 2, "grid_8x8_D2_norm.npz"), (5, 4, "grid_5x5_D4.npz")):
        g2 = tne.grid([L2, L2])
        q = tne.rand_simplepeps(g2, D2, np.random.default_rng(seed_for(L2, D2)), optimizer="sweep", segments="auto2")
        assert rel(tne.inner_product(q, q).to_numpy(), _gold(name)["inner"]) < TOL
    # VidalPEPS: the bond vectors are leaves of the network too (their bra / ket copies form groups of their own)
    g4 = tne.grid([4, 4])
    va = tne.rand_vidalpeps(g4, 2, np.random.default_rng(11), optimizer="sweep", segments="auto")
    vb = tne.replace_tensors(tne.zero_vidalpeps(g4, 2, optimizer="greedy"), va.alltensors())
    assert rel(tne.inner_product(va, va).to_numpy(), tne.inner_product(vb, vb).to_numpy()) < TOL
    assert rel(tne.vec(va), tne.vec(vb)) < TOL
    g3 = tne.grid([3, 5])
    a = tne.rand_simplepeps(g3, 3, np.random.default_rng(9), optimizer="sweep", segments="auto")
    b = tne.replace_tensors(tne.zero_simplepeps(g3, 3, optimizer="greedy"), a.alltensors())
    assert rel(tne.inner_product(a, a).to_numpy(), tne.inner_product(b, b).to_numpy()) < TOL
    assert rel(tne.expect(tne.rydberg_hamiltonian(g3, 10.0, 0.2, 0.3), a, a).to_numpy(),
               tne.expect(tne.rydberg_hamiltonian(g3, 10.0, 0.2, 0.3), b, b).to_numpy()) < TOL
