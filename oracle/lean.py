"""Memory-lean evaluation of ``iloss2`` / ``fvec`` / ``normalized_overlap`` for benchmark-size golden vectors.

TEST INFRASTRUCTURE ONLY (see oracle/__init__).

The literal restatement (:mod:`oracle.timeevolve` on :mod:`oracle.tensorad`) keeps, like the reference
(``src/TensorAD/TensorAD.jl:25,41-47``), every intermediate of every Hamiltonian term AND every array of the
(itself taped) backward pass alive on one global tape: a single term of a 5x5 D=4 PEPS already needs more than
this container's 62 GB.  This module computes the SAME numbers with bounded memory:

* the heavy arithmetic is unchanged -- every contraction is ``oracle.einsum.einsum_raw`` (TTGT) on the PEPS's own
  ``code_inner_product`` tree and every pullback is ``oracle.einsum.einsum_grad`` exactly as the einsum rule calls it
  (``src/TensorAD/rules.jl:1-8``: ``conj(einsum_grad(..., conj(dy), i))``) -- but only one tree's forward values are
  held, the backward pass is not taped, and Hamiltonian terms are processed one at a time (``expect`` over an ``Add``
  is a plain sum, ``src/peps.jl:469-473``);
* the scalar normalisation chain of ``iloss2`` (``src/timeevolve.jl:55-60``: ``pl = load(x) * inv.(norm(load(x)))``,
  ``peps*c`` scaling every tensor by ``c^(1/nsite)``, ``src/peps.jl:292-295,302``) is differentiated in closed form:
  with ``c = nrm^(-1/n)``, ``nrm = sqrt|<x|x>|``, ``g_pl`` the cotangent of the scaled tensors and ``E_s`` the
  environment ``d<x|x>/d conj(x_s)``,
      ``g_x[s] = c g_pl[s] + A E_s``,  ``A = (sum_t Re<g_pl[t], x_t>) (-1/n) nrm^(-1/n-2)``
  (gradient convention ``g = dL/dre + i dL/dim``, ``src/TensorAD/rules.jl:71-75``).

``tests/test_oracle_lean.py`` checks it against the literal tape at the sizes where that runs (1e-13)."""
import numpy as np

from . import einsum as es
from . import peps as P
from . import yao


def _leafset(node, memo):
    k = id(node)
    if k not in memo:
        memo[k] = frozenset([node.tensorindex]) if node.isleaf() else frozenset().union(*[_leafset(a, memo) for a in node.args])
    return memo[k]


def tree_value_and_grads(nested, xs, want):
    """Value of ``nested(*xs)`` and, for every leaf index in ``want``, the cotangent the einsum rule's pullbacks
    deliver for seed ``dy = 1`` (``src/TensorAD/rules.jl:1-8`` applied node by node, untaped)."""
    if isinstance(nested, es.SlicedEinsum):
        assert not nested.slicing
        nested = nested.eins
    want = frozenset(want)
    memo, store = {}, {}

    def fwd(nd):
        if nd.isleaf():
            return xs[nd.tensorindex]
        ys = [fwd(a) for a in nd.args]
        if _leafset(nd, memo) & want:
            store[id(nd)] = ys
        return es.einsum_raw(nd.eins.ixs, ys, nd.eins.iy)

    val = fwd(nested)
    grads = {}

    def bwd(nd, dy):
        if nd.isleaf():
            i = nd.tensorindex
            grads[i] = dy if i not in grads else grads[i] + dy
            return
        ys = store.pop(id(nd))
        sd = es._size_dict(nd.eins.ixs, ys)
        cdy = np.conj(dy)
        del dy
        todo = [i for i, a in enumerate(nd.args) if _leafset(a, memo) & want]
        gs = [es.einsum_grad(nd.eins.ixs, ys, nd.eins.iy, sd, cdy, i) for i in todo]
        del ys, cdy
        for i, g in zip(todo, gs):
            bwd(nd.args[i], g)

    if want:
        bwd(nested, np.ones((), dtype=np.complex128))
    return val, grads


def _scale_factor(peps):
    """``inv.(norm(peps)) .^ (1/nsite)`` (``src/peps.jl:292-295,302``) and the norm itself."""
    nrm = float(np.sqrt(np.abs(P.inner_product(peps, peps))))
    return (1.0 / nrm) ** (1.0 / P.nsite(peps)), nrm


def _norm_environments(peps):
    """``E_s = d<x|x>/d conj(x_s)`` for every site (vertex tensors only; SimplePEPS)."""
    n = P.nsite(peps)
    ts = peps.alltensors()
    _, g = tree_value_and_grads(peps.code_inner_product, [np.conj(t) for t in ts] + list(ts), range(n))
    return [np.conj(g[s]) for s in range(n)]            # pullback delivers conj(dN/db_s), b_s = conj(x_s)


def _flat(ts):
    return np.concatenate([np.reshape(t, -1, order="F") for t in ts])


def iloss2_term_piece(peps, term, c=None, c0=None):
    """One Hamiltonian term of ``iloss2`` at ``x = variables(peps)``: ``E_k = code(pl..., (h_k pr)...)`` (complex; the
    two conjugations of ``src/timeevolve.jl:59`` and ``src/peps.jl:107`` cancel) and the cotangent of ``pl`` for seed 1."""
    assert peps.kind == "simple"
    n = P.nsite(peps)
    if c is None:
        c, _ = _scale_factor(peps)
        c0 = c
    pl = [t * c for t in peps.vertex_tensors]
    pr = P.replace_tensors(peps, [t * c0 for t in peps.vertex_tensors])
    opb = P.apply_block(pr, term)
    val, g = tree_value_and_grads(peps.code_inner_product, pl + list(opb.alltensors()), range(n))
    return complex(val), _flat([g[s] for s in range(n)])


def assemble_iloss2_gradient(peps, E_terms, gpl_sum):
    """``iloss2`` and ``fvec`` (``src/timeevolve.jl:55-64``) from the per-term pieces."""
    n = P.nsite(peps)
    c, nrm = _scale_factor(peps)
    x = P.variables(peps)
    envs = _flat(_norm_environments(peps))
    A = float(np.sum(np.real(np.conj(gpl_sum) * x))) * (-1.0 / n) * nrm ** (-1.0 / n - 2.0)
    gx = c * gpl_sum + A * envs
    return float(np.real(sum(E_terms))), -1j * gx


def fvec(peps, h):
    """Lean ``(iloss2, fvec)``; ``h`` an ``Add`` of terms or a single block."""
    terms = h.blocks if isinstance(h, yao.Add) else [h]
    c, _ = _scale_factor(peps)
    Es, gsum = [], 0
    for t in terms:
        e, g = iloss2_term_piece(peps, t, c, c)
        Es.append(e)
        gsum = gsum + g
    return assemble_iloss2_gradient(peps, Es, gsum)


def normalized_overlap_grad1(peps, v1, v2):
    """``y = normalized_overlap(peps, v1, v2)`` (``src/timeevolve.jl:6-12``) and its gradient w.r.t. ``v1``."""
    assert peps.kind == "simple"
    n = P.nsite(peps)
    p1 = P.load_variables(peps, v1)
    p2 = P.load_variables(peps, v2)
    c1, nrm1 = _scale_factor(p1)
    c2, _ = _scale_factor(p2)
    pl = [t * c1 for t in p1.vertex_tensors]
    pr = [t * c2 for t in p2.vertex_tensors]
    val, g = tree_value_and_grads(peps.code_inner_product, [np.conj(t) for t in pl] + pr, range(n))
    gpl = _flat([np.conj(g[s]) for s in range(n)])          # conj rule: b_s = conj(pl_s)
    envs = _flat(_norm_environments(p1))
    A = float(np.sum(np.real(np.conj(gpl) * v1))) * (-1.0 / n) * nrm1 ** (-1.0 / n - 2.0)
    return float(np.real(val)), c1 * gpl + A * envs


def apply_smatrix_fd(peps, v1, v2, x, eps=1e-3):
    """``apply_smatrix`` (``src/timeevolve.jl:15-52``) by Richardson-extrapolated central differences of the lean
    first-order gradient: ``sx = d/de grad_{v1} y(v1, v2 + e x)``.  Accurate to ~1e-9 relative, not to rounding."""
    def d(e):
        return (normalized_overlap_grad1(peps, v1, v2 + e * x)[1] - normalized_overlap_grad1(peps, v1, v2 - e * x)[1]) / (2 * e)
    return (4.0 * d(eps) - d(2 * eps)) / 3.0
