"""Pins the oracle's TEBD / TDVP drivers and TensorAD restatement on the reference's own tests
(test/tebd.jl, test/timeevolve.jl, test/TensorAD/*).  ForwardDiff / FiniteDifferences are replaced by an
independent torch-CPU autograd restatement on dense statevectors (same complex-gradient convention
g = dL/dre + i dL/dim, checked below) and by central finite differences."""
import string

import numpy as np
import pytest
import torch

from oracle import graphs as OG, peps as OP, tebd as OT, tensorad as AD, timeevolve as OTE, yao as OY
from helpers import dense_apply


@pytest.fixture
def g():
    return OG.test_graph()


# ---------------------------------------------------------------------------------------------
# TEBD against dense evolution (test/tebd.jl:4-26)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["simple", "vidal"])
def test_rydberg_tebd_matches_dense(g, kind):
    rng = np.random.default_rng(10)
    if kind == "simple":
        peps = OP.rand_simplepeps(g, 2, rng, Dmax=10)
    else:
        peps = OP.rand_vidalpeps(g, 2, rng, Dmax=10, eps=1e-10)
    reg = OY.ArrayReg(OP.vec(peps))
    kw = dict(t=0.3, C=10.0, delta=0.2, omega=0.3, nstep=10)
    if kind == "simple":
        OT.rydberg_tebd_(peps, g, **kw)
    else:
        OT.rydberg_tebd_(peps, **kw)  # builds the graph from the PEPS's own bonds (tebd.jl:33-38)
    assert not np.allclose(OP.statevec(peps), reg.state, atol=1e-3)
    OT.rydberg_tebd_(reg, g, **kw)
    assert np.allclose(OP.statevec(peps), reg.state, atol=1e-3)


def test_tebd_runs_nstep_minus_one_sweeps(g):  # tebd.jl:27 quirk, observable behaviour
    rng = np.random.default_rng(11)
    p = OP.rand_simplepeps(g, 2, rng, Dmax=10)
    reg = OY.ArrayReg(OP.vec(p))
    OT.rydberg_tebd_(reg, g, t=0.3, C=10.0, delta=0.2, omega=0.3, nstep=3)
    ref = OY.ArrayReg(OP.vec(p))
    for _ in range(2):
        OT.rydberg_tebd_sweep_(ref, g, 10.0, 0.2, 0.3, 0.1)
    assert np.allclose(reg.state, ref.state)


def test_truncation_keeps_dmax_and_reports():
    g = OG.grid([2, 2])
    rng = np.random.default_rng(12)
    p = OP.rand_simplepeps(g, 2, rng)  # Dmax = D = 2: a generic two-site gate must truncate
    m = OY.rand_unitary(4, rng)
    OP.apply_onbond_(p, 1, 2, np.reshape(m, (2, 2, 2, 2), order="F"))
    assert all(max(t.shape[1:]) <= 2 for t in p.vertex_tensors)


# ---------------------------------------------------------------------------------------------
# torch restatement of the dense quantities
# ---------------------------------------------------------------------------------------------
def torch_statevec(peps, tensors):
    """einsum of all vertex tensors -> state tensor with site 1 fastest, flattened column-major."""
    labels = sorted({l for vl in peps.vertex_labels for l in vl})
    sym = {l: string.ascii_letters[i] for i, l in enumerate(labels)}
    ins = ",".join("".join(sym[l] for l in vl) for vl in peps.vertex_labels)
    out = "".join(sym[l] for l in reversed(peps.physical_labels))  # C-order flatten => site 1 fastest
    return torch.einsum(ins + "->" + out, *tensors).reshape(-1)


def torch_tensors(peps, v):
    ts, k = [], 0
    for t in peps.vertex_tensors:
        n = t.size
        ts.append(v[k:k + n].reshape(tuple(reversed(t.shape))).permute(*reversed(range(t.ndim))))  # column-major
        k += n
    return ts


def torch_iloss2(h, peps, v):
    psi = torch_statevec(peps, torch_tensors(peps, v))
    psi = psi / torch.sqrt(torch.abs(torch.vdot(psi, psi)))
    psi0 = OP.statevec(peps)
    psi0 = psi0 / np.linalg.norm(psi0)
    hpsi0 = torch.tensor(dense_apply(h, psi0))
    return torch.real(torch.dot(psi, hpsi0))  # un-conjugated: timeevolve.jl:59 + peps.jl:107 cancel


def rand_hamiltonian_sr(g, rng):  # test/timeevolve.jl:6-17
    n = g.nv()
    blocks = []
    for (i, j) in g.edges():
        for P in (OY.X, OY.Y, OY.Z):
            blocks.append(rng.standard_normal() * OY.kron(n, (i, P), (j, P)))
    for i in g.vertices():
        blocks.append(OY.put(n, i, rng.random() * OY.X))
    return OY.Add(*blocks)


def test_statevec_restatement(g):
    p = OP.rand_simplepeps(g, 2, np.random.default_rng(13), Dmax=4)
    v = torch.tensor(OP.variables(p))
    assert np.allclose(torch_statevec(p, torch_tensors(p, v)).numpy(), OP.statevec(p))


def test_fvec_against_autograd(g):  # test/timeevolve.jl:5-35
    rng = np.random.default_rng(14)
    h = rand_hamiltonian_sr(g, rng)
    p1 = OP.rand_simplepeps(g, 2, rng, Dmax=4)
    gv = AD.gradient(lambda x: OP.norm(OP.load_variables(p1, x)), AD.DiffTensor(OP.variables(p1)))[0]
    assert isinstance(gv, AD.DiffTensor)
    assert np.ndim(OTE.iloss2(h, p1, OP.variables(p1))) == 0
    f = OTE.fvec(p1, h)
    assert isinstance(f, AD.DiffTensor)
    v = torch.tensor(OP.variables(p1), requires_grad=True)
    loss = torch_iloss2(h, p1, v)
    assert np.isclose(loss.item(), OTE.iloss2(h, p1, OP.variables(p1)))
    loss.backward()
    # torch: grad = dL/dre + i dL/dim for a real loss -- the reference's convention (test/timeevolve.jl:32-34)
    assert np.allclose(f.data, -1j * v.grad.numpy(), rtol=1e-9, atol=1e-12)


def test_gradient_convention_finite_difference(g):
    rng = np.random.default_rng(15)
    h = rand_hamiltonian_sr(g, rng)
    p1 = OP.rand_simplepeps(g, 2, rng, Dmax=4)
    v0 = OP.variables(p1)
    f = OTE.fvec(p1, h).data
    for idx in rng.integers(0, v0.size, 6):
        eps = 1e-6
        d = np.zeros_like(v0); d[idx] = eps
        dre = (OTE.iloss2(h, p1, v0 + d) - OTE.iloss2(h, p1, v0 - d)) / (2 * eps)
        dim = (OTE.iloss2(h, p1, v0 + 1j * d) - OTE.iloss2(h, p1, v0 - 1j * d)) / (2 * eps)
        assert np.isclose(f[idx], -1j * (dre + 1j * dim), rtol=1e-5, atol=1e-8)


def test_normalized_overlap_gradient_zero(g):  # test/timeevolve.jl:37-48
    p1 = OP.rand_simplepeps(g, 2, np.random.default_rng(16), Dmax=4)
    v1 = AD.DiffTensor(OP.variables(p1).copy(), requires_grad=False)
    v2 = AD.DiffTensor(OP.variables(p1))
    g1 = AD.gradient(lambda x: OTE.normalized_overlap(p1, v1, x), v2)[0]
    assert np.allclose(g1.data, 0, atol=1e-8)


def test_smatrix_apply_against_double_backward(g):  # test/timeevolve.jl:50-67
    rng = np.random.default_rng(17)
    p1 = OP.rand_simplepeps(g, 2, rng, Dmax=4)
    v = OP.variables(p1)
    n = v.size
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2)
    sx = OTE.apply_smatrix(p1, AD.DiffTensor(v.copy()), AD.DiffTensor(v.copy()), AD.DiffTensor(x, requires_grad=False)).data

    def overlap(a, b):
        pa = torch_statevec(p1, torch_tensors(p1, a)); pa = pa / torch.sqrt(torch.abs(torch.vdot(pa, pa)))
        pb = torch_statevec(p1, torch_tensors(p1, b)); pb = pb / torch.sqrt(torch.abs(torch.vdot(pb, pb)))
        return torch.real(torch.vdot(pa, pb))
    # H12 * x_real with interleaved reals: real Hessian block d2y / dv1 dv2 applied to x
    a = torch.tensor(np.stack([v.real, v.imag], 1), requires_grad=True)
    b = torch.tensor(np.stack([v.real, v.imag], 1), requires_grad=True)
    y = overlap(torch.view_as_complex(a), torch.view_as_complex(b))
    (gb,) = torch.autograd.grad(y, b, create_graph=True)
    z = (gb * torch.tensor(np.stack([x.real, x.imag], 1))).sum()
    (ga,) = torch.autograd.grad(z, a)
    ref = ga.numpy()[:, 0] + 1j * ga.numpy()[:, 1]
    assert np.allclose(sx, ref, rtol=1e-8, atol=1e-11)


# ---------------------------------------------------------------------------------------------
# TensorAD rules: jacobians against finite differences (test/TensorAD/rules.jl via ADTest.jl:53-75)
# ---------------------------------------------------------------------------------------------
def _fd_vjp(f, x, cot, eps=1e-6):
    """<cot, J dx> convention check: gradient of Re<cot, f(x)> by central differences."""
    g = np.zeros_like(x, dtype=np.complex128)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        for part, unit in ((0, 1.0), (1, 1j)):
            if part == 1 and not np.iscomplexobj(x):
                continue
            d = np.zeros_like(x); d[i] = eps * unit
            val = (np.real(np.vdot(cot, f(x + d))) - np.real(np.vdot(cot, f(x - d)))) / (2 * eps)
            g[i] += val * (1.0 if part == 0 else 1j)
    return g


RULES = [
    ("neg", lambda t: AD.neg(t), lambda a: -a),
    ("inv", lambda t: AD.binv(t), lambda a: 1 / a),
    ("pow3", lambda t: AD.bpow(t, 3), lambda a: a ** 3),
    ("sqrt", lambda t: AD.bsqrt(t), lambda a: np.sqrt(a)),
    ("copy", lambda t: AD.copy(t), lambda a: a.copy()),
    ("transpose", lambda t: AD.transpose(t), lambda a: a.T),
    ("reshape", lambda t: AD.reshape(t, (t.data.size,)), lambda a: a.reshape(-1, order="F")),
    ("scalar*", lambda t: AD.mul(2.5 - 1j, t), lambda a: (2.5 - 1j) * a),
    ("sum", lambda t: AD.sum_(t), lambda a: np.asarray(a.sum())),
    ("sin", lambda t: AD.bsin(t), lambda a: np.sin(a)),
    ("cos", lambda t: AD.bcos(t), lambda a: np.cos(a)),
    ("real", lambda t: AD.real(t), lambda a: np.real(a).astype(np.complex128)),
    ("imag", lambda t: AD.imag(t), lambda a: np.imag(a).astype(np.complex128)),
    ("conj", lambda t: AD.conj(t), lambda a: np.conj(a)),
    ("abs", lambda t: AD.babs(t), lambda a: np.abs(a).astype(np.complex128)),
    ("self*", lambda t: AD.bmul(t, t), lambda a: a * a),
    ("matmul", lambda t: AD.mul(t, AD.transpose(t)), lambda a: a @ a.T),
]


@pytest.mark.parametrize("name,fad,fnp", RULES, ids=[r[0] for r in RULES])
def test_rule_vjp(name, fad, fnp):
    rng = np.random.default_rng(abs(hash(name)) % 1000)
    x = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3)) + 2.0
    y0 = fnp(x)
    cot = rng.standard_normal(np.shape(y0)) + 1j * rng.standard_normal(np.shape(y0))
    if name in ("real", "imag", "abs"):
        cot = cot.real.astype(np.complex128)  # real outputs take real cotangents
    t = AD.DiffTensor(x.copy())
    AD.GLOBAL_TAPE.instructs.clear()
    y = fad(t)
    assert np.allclose(y.data, y0)
    storage = {}
    seed = cot if np.iscomplexobj(y.data) else np.real(cot)
    AD.accumulate_gradient_(storage, AD.getid(y), AD.DiffTensor(np.asarray(seed), requires_grad=False))
    AD.back_(AD.GLOBAL_TAPE, storage)
    got = AD.getgrad(storage, t).data
    assert np.allclose(got, _fd_vjp(fnp, x, cot), rtol=1e-5, atol=1e-6)


def test_gradient_and_requires_grad_pruning():  # test/TensorAD/TensorAD.jl:10-20, 41-81
    rng = np.random.default_rng(20)
    a = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    x = AD.DiffTensor(a.copy())
    c = AD.DiffTensor(rng.standard_normal((4, 4)) + 0j, requires_grad=False)
    g = AD.gradient(lambda t: AD.real(AD.sum_(AD.bmul(AD.mul(t, c), t))), x)[0]
    ref = _fd_vjp(lambda z: np.asarray(np.real(np.sum((z @ c.data) * z))).astype(np.complex128), a, np.asarray(1.0 + 0j))
    assert np.allclose(g.data, ref, rtol=1e-5, atol=1e-6)
    # inputs that do not require grad get no pullback recorded (TensorAD.jl:41-47)
    AD.GLOBAL_TAPE.instructs.clear()
    AD.mul(x, c)
    assert len(AD.GLOBAL_TAPE.instructs[-1].backs) == 1
