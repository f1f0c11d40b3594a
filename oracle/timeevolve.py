"""numpy restatement of ``src/timeevolve.jl``.  TEST INFRASTRUCTURE ONLY (see oracle/__init__)."""
import numpy as np

from . import einsum as es
from . import peps as P
from . import tensorad as ad
from .tensorad import DiffTensor


def _inv(x):
    return ad.binv(x) if isinstance(x, DiffTensor) else 1.0 / x


def _real(x):
    return ad.real(x) if isinstance(x, DiffTensor) else np.real(x)


def normalized_overlap(peps, v1, v2):  # timeevolve.jl:6-12
    pl = P.load_variables(peps, v1)
    pl = P.mul(pl, _inv(P.norm(pl)))
    pr = P.load_variables(peps, v2)
    pr = P.mul(pr, _inv(P.norm(pr)))
    return _real(P.inner_product(pl, pr))


def apply_smatrix(peps, v1, v2, x):  # timeevolve.jl:15-52
    ad.GLOBAL_TAPE.instructs.clear()
    ad.requires_grad_(v1, True)
    ad.requires_grad_(v2, True)
    pl = P.load_variables(peps, v1)
    pl = P.mul(pl, _inv(P.norm(pl)))
    pr = P.load_variables(peps, v2)
    pr = P.mul(pr, _inv(P.norm(pr)))
    y = _real(P.inner_product(pl, pr))
    storage = ad.init_storage_(y, requires_grad=True)
    ad.back_(ad.GLOBAL_TAPE, storage)
    g2 = ad.getgrad(storage, v2)
    assert ad.requires_grad(g2)
    z = ad.real(es.einsum([[1], [1]], [ad.conj(x), g2], []))
    assert ad.requires_grad(z)
    storage = ad.init_storage_(z)
    ad.back_(ad.GLOBAL_TAPE, storage)
    return ad.getgrad(storage, v1)


def iloss2(h, peps, variables):  # timeevolve.jl:55-60
    pl = P.load_variables(peps, variables)
    pl = P.mul(pl, _inv(P.norm(pl)))
    pr = P.mul(peps, _inv(P.norm(peps)))
    return _real(P.expect(h, P.conj(pl), pr))


def wrap_difftensor(peps):  # timeevolve.jl:65-67
    return P.replace_tensors(peps, [DiffTensor(t) for t in peps.alltensors()])


def fvec(peps, h):  # timeevolve.jl:61-64
    variables = P.variables(peps)
    g = ad.gradient(lambda x: iloss2(h, wrap_difftensor(peps), x), DiffTensor(variables))[0]
    return ad.mul(-1j, g)


def sr_step(peps, h, tol=1e-10, maxiter=200):
    """timeevolve.jl:1-4 as intended (the reference calls a 2-argument ``apply_smatrix`` that does
    not exist; restated with ``v1 = v2 = variables(peps)``).  GMRES via scipy stands in for KrylovKit."""
    import scipy.sparse.linalg as sla
    v = P.variables(peps)
    n = v.size

    def matvec(xr):
        xc = xr[:n] + 1j * xr[n:]
        sx = apply_smatrix(peps, DiffTensor(v.copy()), DiffTensor(v.copy()), DiffTensor(xc, requires_grad=False)).data
        return np.concatenate([sx.real, sx.imag])
    b = -1j * fvec(peps, h).data
    op = sla.LinearOperator((2 * n, 2 * n), matvec=matvec, dtype=np.float64)
    sol, _ = sla.gmres(op, np.concatenate([b.real, b.imag]), x0=np.concatenate([v.real, v.imag]), rtol=tol,
                       maxiter=maxiter)
    return P.load_variables_(peps.copy(), sol[:n] + 1j * sol[n:])
