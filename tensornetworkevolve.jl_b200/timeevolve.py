"""TDVP drivers on the device: src/timeevolve.jl with unchanged signatures.

``fvec`` runs the fused path (one device program for the K-term expectation and its reverse pass,
plus the two norm contractions); ``apply_smatrix`` needs reverse-over-reverse and therefore tapes
every binary einsum like the reference does (``einsum.FUSED_TREES`` is switched off inside it)."""
import numpy as np

from . import einsum as es
from . import peps as P
from . import tensorad as ad
from .array import B200Array
from .tensorad import DiffTensor


def _inv(x):
    return ad.binv(x) if isinstance(x, DiffTensor) else x.inv()


def _real(x):
    return ad.real(x) if isinstance(x, DiffTensor) else x.real()


def normalized_overlap(peps, v1, v2):  # src/timeevolve.jl:6-12
    pl = P.load_variables(peps, v1)
    pl = P.mul(pl, _inv(P.norm(pl)))
    pr = P.load_variables(peps, v2)
    pr = P.mul(pr, _inv(P.norm(pr)))
    return _real(P.inner_product(pl, pr))


def apply_smatrix(peps, v1, v2, x):  # src/timeevolve.jl:15-52
    fused = es.FUSED_TREES
    es.FUSED_TREES = False
    try:
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
    finally:
        es.FUSED_TREES = fused


def iloss2(h, peps, variables, shard=(0, 1)):  # src/timeevolve.jl:55-60
    pl = P.load_variables(peps, variables)
    pl = P.mul(pl, _inv(P.norm(pl)))
    pr = P.mul(peps, _inv(P.norm(peps)))
    return _real(P.expect(h, P.conj(pl), pr, shard=shard))


def wrap_difftensor(peps, requires_grad=True):  # src/timeevolve.jl:65-67
    return P.replace_tensors(peps, [t if isinstance(t, DiffTensor) else DiffTensor(t, requires_grad=requires_grad)
                                    for t in peps.alltensors()])


def fvec(peps, h, shard=(0, 1)):  # src/timeevolve.jl:61-64
    """The reference wraps the constant ``peps`` with ``requires_grad = true`` and then discards those
    gradients; here the constant side is wrapped with ``requires_grad = false`` (same result, no wasted
    reverse pass through ``norm(peps)``)."""
    variables = P.variables(peps)
    g = ad.gradient(lambda x: iloss2(h, wrap_difftensor(peps, False), x, shard=shard), DiffTensor(variables))[0]
    return ad.mul(-1j, g)


def energy_and_gradient(peps, h, shard=(0, 1)):
    """One ``<H> + grad`` evaluation (the benchmark's unit of work): ``(iloss2 value, fvec)`` from a
    single taped pass.  With ``shard = (rank, world)`` both are this rank's partial sums over its
    Hamiltonian terms (the normalisation chain is replicated; its gradient contribution is
    proportional to the rank's partial energy, so partial results add up exactly)."""
    variables = DiffTensor(P.variables(peps))
    ad.GLOBAL_TAPE.instructs.clear()
    y = iloss2(h, wrap_difftensor(peps, False), variables, shard=shard)
    storage = ad.init_storage_(y)
    ad.back_(ad.GLOBAL_TAPE, storage)
    g = ad.getgrad(storage, variables)
    return y, ad.mul(-1j, g)


def sr_step(peps, h, tol=1e-10, maxiter=200):
    """src/timeevolve.jl:1-4 as intended: GMRES on ``S x = -i f`` with the device metric-vector product
    (the reference calls a 2-argument ``apply_smatrix`` that does not exist; ``v1 = v2 = variables``).
    The Krylov loop is host code (KrylovKit in the reference); only its matvec runs on the device."""
    import scipy.sparse.linalg as sla
    v = P.variables(peps).to_numpy()
    n = v.size

    def matvec(xr):
        xc = xr[:n] + 1j * xr[n:]
        sx = apply_smatrix(peps, DiffTensor(v.copy()), DiffTensor(v.copy()), DiffTensor(xc, requires_grad=False)).to_numpy()
        return np.concatenate([sx.real, sx.imag])
    b = -1j * fvec(peps, h).to_numpy()
    op = sla.LinearOperator((2 * n, 2 * n), matvec=matvec, dtype=np.float64)
    sol, _ = sla.gmres(op, np.concatenate([b.real, b.imag]), x0=np.concatenate([v.real, v.imag]), rtol=tol, maxiter=maxiter)
    return P.load_variables_(peps.copy(), B200Array.from_numpy(sol[:n] + 1j * sol[n:]))
