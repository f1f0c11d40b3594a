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


def apply_smatrix(peps, v1, v2, x, memo=None):  # src/timeevolve.jl:15-52
    """Metric-vector product ``sx = d/dv1 Re(conj(x) . dy/dv2)`` of ``y = Re<psi^(v1)|psi^(v2)>``.

    The reference differentiates the taped backward pass a second time (reverse over reverse).  With fused
    trees the same quantity is evaluated by an explicit tree-level algorithm: ``Re(conj(x) . dy/dv2)`` is the
    directional derivative of ``y`` along ``x`` in ``v2``; because the PEPS is multilinear in its tensors that
    derivative is ``z(v1) = Re<psi^(v1)|phi>`` with the fixed ket
    ``phi = (sum_s psi2[s -> x_s] - c psi2) / n2``, ``c = Re<psi2|sum_s psi2[s -> x_s]> / n2^2``, i.e. a sum of
    one-site replacement terms -- exactly what ``tne_expect_terms`` contracts in one device programme
    (shared sub-contractions, checkpointed reverse pass).  ``sx`` is then the ordinary first-order gradient
    of ``z``.  Works at sizes where the reference's tape cannot exist (6x6, D=4).

    ``memo``: a dict the caller keeps across calls with the SAME ``v1`` and ``v2`` (the matvecs of one GMRES solve in
    ``sr_step``): the x-independent pieces -- ``<psi2|psi2>`` and the norm programme of ``v1`` (value + pullbacks) -- are then
    contracted once instead of once per product."""
    if es.FUSED_TREES:   # SimplePEPS and VidalPEPS alike: the state is multilinear in the bond vectors too
        return _apply_smatrix_fused(peps, v1, v2, x, memo)
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


def _norm_memo(pl, memo):
    """``norm(pl)`` for a PEPS of DiffTensors; with ``memo`` the fused norm programme (value and pullbacks for a unit
    cotangent) runs once and is re-attached to the current leaf objects afterwards."""
    if memo is None:
        return P.norm(pl)
    xs = P.conj(pl).alltensors() + pl.alltensors()
    leaves = [j for j in range(len(xs)) if ad.requires_grad(xs[j])]
    if "norm" not in memo:
        memo["norm"] = es.expect_terms_device(pl.code_inner_product, [t.data for t in xs], None, [], leaves, 1.0 + 0j)
    y, grads = memo["norm"]
    return ad.bsqrt(ad.babs(ad.wrap_tree_result(xs, leaves, y, grads)))


def _apply_smatrix_fused(peps, v1, v2, x, memo=None):
    ad.requires_grad_(v1, True)
    nt = len(peps.alltensors())
    p2 = P.load_variables(peps, v2.data if isinstance(v2, DiffTensor) else v2)          # plain device tensors
    xs = P.load_variables(peps, x.data if isinstance(x, DiffTensor) else x).alltensors()
    code = peps.code_inner_product
    ket = [P._dev(t) for t in p2.alltensors()]
    if memo is not None and "n2sq" in memo:
        n2sq = memo["n2sq"]
    else:
        n2sq = es.contract_tree_device(code, ket + ket, [1] * nt + [0] * nt).item()   # <psi2|psi2>
        if memo is not None:
            memo["n2sq"] = n2sq
    n2 = float(np.sqrt(abs(n2sq)))
    site_terms = [(1.0 + 0j, [(nt + s, P._dev(xs[s]))]) for s in range(nt)]
    dval, _ = es.expect_terms_device(code, ket + ket, [1] * nt + [0] * nt, site_terms)  # <psi2|dpsi2>
    c = float(np.real(dval.item())) / (n2 * n2)
    terms = [(1.0 / n2, {s + 1: P._dev(xs[s])}) for s in range(nt)] + [(-c / n2, {})]

    def z_of(v):
        pl = P.load_variables(peps, v)
        pl = P.mul(pl, _inv(_norm_memo(pl, memo)))
        return _real(P._expect_fused(terms, pl, p2, (0, 1)))
    return ad.gradient(z_of, v1)[0]


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
    # iloss2 (src/timeevolve.jl:55-60) evaluated at x = variables(peps): ||psi(x)|| and ||psi|| are the same number, so
    # the norm network is contracted once (value + pullbacks in one device program) and its value is reused, detached,
    # for the constant side
    pl = P.load_variables(wrap_difftensor(peps, False), variables)
    nl = P.norm(pl)
    pl = P.mul(pl, _inv(nl))
    pr = P.mul(wrap_difftensor(peps, False), _inv(DiffTensor(nl.data, requires_grad=False)))
    y = _real(P.expect(h, P.conj(pl), pr, shard=shard))
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

    memo = {}   # v1 = v2 = variables(peps) for every matvec: the x-independent contractions run once per solve

    def matvec(xr):
        xc = xr[:n] + 1j * xr[n:]
        sx = apply_smatrix(peps, DiffTensor(v.copy()), DiffTensor(v.copy()), DiffTensor(xc, requires_grad=False), memo=memo).to_numpy()
        return np.concatenate([sx.real, sx.imag])
    b = -1j * fvec(peps, h).to_numpy()
    op = sla.LinearOperator((2 * n, 2 * n), matvec=matvec, dtype=np.float64)
    sol, _ = sla.gmres(op, np.concatenate([b.real, b.imag]), x0=np.concatenate([v.real, v.imag]), rtol=tol, maxiter=maxiter)
    return P.load_variables_(peps.copy(), B200Array.from_numpy(sol[:n] + 1j * sol[n:]))
