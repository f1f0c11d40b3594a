"""numpy restatement of ``src/peps.jl`` (PEPS state layer + Yao glue).

TEST INFRASTRUCTURE ONLY (see oracle/__init__).  Sites, labels and bond numbering are
1-based exactly like the reference; arrays carry the Julia ``size`` as numpy ``shape``
and every ``vec``/``reshape`` is column-major (``order='F'``).  The code is generic over
plain ``numpy.ndarray`` and :class:`oracle.tensorad.DiffTensor` like the Julia source is
generic over ``AbstractArray``.
"""
import numpy as np

from . import einsum as es
from . import tensorad as ad
from . import yao
from .tensorad import DiffTensor

LOG = []  # truncation messages (@info at peps.jl:386)


# -- small dispatch helpers (Julia multiple dispatch on the array type) -----------------
def _isdiff(x):
    return isinstance(x, DiffTensor)


def _conj(x):
    return ad.conj(x) if _isdiff(x) else np.conj(x)


def _reshape(x, shape):
    return ad.reshape(x, tuple(shape)) if _isdiff(x) else np.reshape(x, shape, order="F")


def _vec(x):
    return np.reshape(ad.getdata(x), -1, order="F")


class PEPS:
    """Common behaviour of ``SimplePEPS`` / ``VidalPEPS`` (``peps.jl:4``)."""

    kind = "simple"

    def __init__(self, physical_labels, virtual_labels, vertex_labels, vertex_tensors, max_index,
                 code_statetensor, code_inner_product, Dmax, eps, nflavor=2, bond_tensors=None):
        self.physical_labels = list(physical_labels)
        self.virtual_labels = list(virtual_labels)
        self.vertex_labels = [list(l) for l in vertex_labels]
        self.vertex_tensors = list(vertex_tensors)
        self.max_index = max_index
        self.code_statetensor = code_statetensor
        self.code_inner_product = code_inner_product
        self.Dmax = Dmax
        self.eps = eps
        self.nflavor = nflavor
        if bond_tensors is not None:
            self.bond_tensors = list(bond_tensors)

    # peps.jl:96-98, 171-172
    def alllabels(self):
        if self.kind == "vidal":
            return self.vertex_labels + [[l] for l in self.virtual_labels]
        return self.vertex_labels

    def alltensors(self):
        if self.kind == "vidal":
            return self.vertex_tensors + self.bond_tensors
        return self.vertex_tensors

    def copy(self):  # peps.jl:99-101 (the Vidal method at :175-177 is broken upstream; restated as intended)
        cp = lambda t: t.copy() if not _isdiff(t) else ad.copy(t)
        return replace_tensors(self, [cp(t) for t in self.alltensors()])


class SimplePEPS(PEPS):
    kind = "simple"


class VidalPEPS(PEPS):
    kind = "vidal"


def _optimized_code(alllabels, physical_labels, virtual_labels, max_ind, nflavor, D, optimizer, nslices):
    """peps.jl:84-93."""
    size_dict = {l: nflavor for l in physical_labels}
    size_dict.update({l: D for l in virtual_labels})
    st_opt = "greedy" if isinstance(optimizer, (list, tuple)) else optimizer  # an explicit order names inner-product leaves
    code_st = es.optimize_code(alllabels, physical_labels, size_dict, st_opt, nslices)
    rep = {l: max_ind + i + 1 for i, l in enumerate(virtual_labels)}
    for l in rep.values():
        size_dict[l] = D
    ixs = [list(l) for l in alllabels] + [[rep.get(x, x) for x in l] for l in alllabels]
    code_ip = es.optimize_code(ixs, [], size_dict, optimizer, nslices)
    return code_st, code_ip


def make_peps(kind, vertex_labels, vertex_tensors, virtual_labels, Dmax, eps, nflavor=2, bond_tensors=None,
              optimizer="greedy", nslices=0):
    """The label-deriving constructors, peps.jl:69-82 and :142-151."""
    physical_labels = []
    for vl in vertex_labels:
        ph = [l for l in vl if l not in virtual_labels]
        assert len(ph) == 1
        physical_labels.append(ph[0])
    for t, vl in zip(vertex_tensors, vertex_labels):  # peps.jl:73-75
        for s, l in zip(t.shape, vl):
            if (l in physical_labels and s != nflavor) or (l not in physical_labels and s > Dmax):
                raise ValueError("Input tensor shape error, either: \n 1. physical degrees of freedom do not match, "
                                 "\n 2. or input tensor virtual degrees of freedom larger than `Dmax = %d`." % Dmax)
    max_ind = max(max(physical_labels), max(virtual_labels))
    if kind == "vidal":
        alll = [list(l) for l in vertex_labels] + [[l] for l in virtual_labels]
    else:
        alll = [list(l) for l in vertex_labels]
    cst, cip = _optimized_code(alll, physical_labels, virtual_labels, max_ind, nflavor, Dmax, optimizer, nslices)
    cls = VidalPEPS if kind == "vidal" else SimplePEPS
    return cls(physical_labels, virtual_labels, vertex_labels, vertex_tensors, max_ind, cst, cip, Dmax, eps,
               nflavor, bond_tensors)


def replace_tensors(peps, tensors):  # peps.jl:222-234
    tensors = list(tensors)
    if peps.kind == "vidal":
        nv = nsite(peps)
        return VidalPEPS(peps.physical_labels, peps.virtual_labels, peps.vertex_labels, tensors[:nv], peps.max_index,
                         peps.code_statetensor, peps.code_inner_product, peps.Dmax, peps.eps, peps.nflavor,
                         tensors[nv:])
    return SimplePEPS(peps.physical_labels, peps.virtual_labels, peps.vertex_labels, tensors, peps.max_index,
                      peps.code_statetensor, peps.code_inner_product, peps.Dmax, peps.eps, peps.nflavor)


def conj(peps):  # peps.jl:178-180
    return replace_tensors(peps, [_conj(t) for t in peps.alltensors()])


# -- label API, peps.jl:183-199 ---------------------------------------------------------
def getvlabel(peps, i):
    return peps.vertex_labels[i - 1]


def getphysicallabel(peps, i):
    return peps.physical_labels[i - 1]


def newlabel(peps, offset):
    return peps.max_index + offset


def virtualbonds(peps):
    bs = []
    for b in peps.virtual_labels:
        i, j = [k + 1 for k, l in enumerate(peps.vertex_labels) if b in l]
        bs.append((i, j))
    return bs


def findbondtensor(peps, b):
    return peps.bond_tensors[peps.virtual_labels.index(b)]


def nsite(peps):
    return len(peps.physical_labels)


# -- variables, peps.jl:202-220 ---------------------------------------------------------
def variables(peps):
    return np.concatenate([_vec(t).astype(np.complex128) for t in peps.alltensors()])


def load_variables_(peps, variables):
    k = 0
    for t in peps.alltensors():
        t[...] = np.reshape(variables[k:k + t.size], t.shape, order="F")
        k += t.size
    return peps


def load_variables(peps, variables):
    ats = peps.alltensors()
    ks = np.cumsum([t.size for t in ats])
    tensors = []
    for i, t in enumerate(ats):
        lo = 0 if i == 0 else int(ks[i - 1])
        tensors.append(_reshape(variables[lo:int(ks[i])], t.shape))
    return replace_tensors(peps, tensors)


# -- factories, peps.jl:244-280 ---------------------------------------------------------
def _peps_zero_state(kind, g, D, nflavor=2, Dmax=None, eps=1e-12, optimizer="greedy", nslices=0,
                     dtype=np.complex128):
    Dmax = D if Dmax is None else Dmax
    nv, ne = g.nv(), g.ne()
    virtual_labels = list(range(nv + 1, nv + ne + 1))
    edge_map = dict(zip(g.edges(), virtual_labels))
    vertex_labels, vertex_tensors = [], []
    for i in range(1, nv + 1):
        labs = [i] + [edge_map.get((i, nb), edge_map.get((nb, i), 0)) for nb in g.neighbors(i)]
        t = np.zeros([nflavor] + [D] * g.degree(i), dtype=dtype)
        t[(0,) * t.ndim] = 1
        vertex_labels.append(labs)
        vertex_tensors.append(t)
    if any(0 in vl for vl in vertex_labels):
        raise ValueError("incorrect input labels1")
    bonds = [np.ones(D, dtype=dtype) for _ in range(ne)] if kind == "vidal" else None
    return make_peps(kind, vertex_labels, vertex_tensors, virtual_labels, Dmax, eps, nflavor, bonds, optimizer, nslices)


def zero_simplepeps(g, D, **kw):
    return _peps_zero_state("simple", g, D, **kw)


def zero_vidalpeps(g, D, **kw):
    return _peps_zero_state("vidal", g, D, **kw)


def randn_(peps, rng):
    """``randn!(peps)`` (peps.jl:275-280): iid circular complex normal, unit total variance,
    tensors filled in ``alltensors`` order, column-major (SURVEY.md 8(d) synthetic-input rule)."""
    for t in peps.alltensors():
        n = t.size
        z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2.0)
        t[...] = z.reshape(t.shape, order="F")
    return peps


def rand_simplepeps(g, D, rng, **kw):
    return randn_(zero_simplepeps(g, D, **kw), rng)


def rand_vidalpeps(g, D, rng, **kw):
    return randn_(zero_vidalpeps(g, D, **kw), rng)


# -- scalar ops and norm, peps.jl:283-306 -----------------------------------------------
def rmul_(peps, c):
    f = c ** (1.0 / nsite(peps))
    peps.vertex_tensors = [t * f for t in peps.vertex_tensors]
    return peps


def mul(peps, c):
    """``peps * c`` for a number (peps.jl:288-291) or a 0-dim array (peps.jl:292-295)."""
    n = nsite(peps)
    if _isdiff(c):
        f = ad.bpow(c, 1.0 / n)
        tensors = [t * f for t in peps.vertex_tensors]
    elif isinstance(c, np.ndarray):
        f = c ** (1.0 / n)
        tensors = [t * f[()] if not _isdiff(t) else t * DiffTensor(f, requires_grad=False) for t in peps.vertex_tensors]
    else:
        f = c ** (1.0 / n)
        tensors = [t * f for t in peps.vertex_tensors]
    rest = peps.alltensors()[len(peps.vertex_tensors):]
    return replace_tensors(peps, tensors + list(rest))


def inner_product(p1, p2):  # peps.jl:106-110
    p1c = conj(p1)
    return p1.code_inner_product(*p1c.alltensors(), *p2.alltensors())


def norm(peps):  # peps.jl:302
    ip = inner_product(peps, peps)
    if _isdiff(ip):
        return ad.bsqrt(ad.babs(ip))
    return np.sqrt(np.abs(ip))


def normalize_(peps):  # peps.jl:303-306
    nm = np.sqrt(np.abs(inner_product(peps, peps)))
    return rmul_(peps, 1.0 / nm[()])


def statetensor(peps):  # peps.jl:320-322
    return peps.code_statetensor(*peps.alltensors())


def vec(peps):  # peps.jl:319
    return _vec(statetensor(peps))


statevec = vec


# -- gates, peps.jl:329-430 -------------------------------------------------------------
def _onsite_code(peps, i):
    old = getvlabel(peps, i)
    mlabel = [newlabel(peps, 1), getphysicallabel(peps, i)]
    return es.EinCode([old, mlabel], [mlabel[0] if l == mlabel[1] else l for l in old])


def apply_onsite_(peps, i, mat):  # peps.jl:329-336
    assert mat.shape[0] == mat.shape[1]
    peps.vertex_tensors[i - 1] = _onsite_code(peps, i)(peps.vertex_tensors[i - 1], mat)
    return peps


def apply_onsite(peps, i, mat):  # peps.jl:338-347
    assert mat.shape[0] == mat.shape[1]
    code = _onsite_code(peps, i)
    tensors = [code(t, mat) if j == i - 1 else t for j, t in enumerate(peps.vertex_tensors)]
    rest = peps.alltensors()[len(peps.vertex_tensors):]
    return replace_tensors(peps, tensors + list(rest))


def safe_inv(y, eps=1e-10):  # peps.jl:419-426
    y = np.asarray(y)
    small = np.abs(y) < eps
    repl = np.where(np.real(y) >= 0, eps, -eps).astype(y.dtype)
    return 1.0 / np.where(small, repl, y)


def _apply_vec(t, l, v, b):  # peps.jl:428-430
    return es.EinCode([l, [b]], l)(t, v)


def apply_onbond_(peps, i, j, mat):  # peps.jl:353-417
    ti, tj = peps.vertex_tensors[i - 1], peps.vertex_tensors[j - 1]
    li, lj = getvlabel(peps, i), getvlabel(peps, j)
    shared = [l for l in li if l in lj]
    assert len(shared) == 1
    only_left = [l for l in li if l not in lj]
    only_right = [l for l in lj if l not in li]
    lij = only_left + only_right
    vidal = peps.kind == "vidal"
    if vidal:
        for b in li:
            if b == getphysicallabel(peps, i):
                continue
            ti = _apply_vec(ti, li, np.sqrt(findbondtensor(peps, b)), b)
        for b in lj:
            if b == getphysicallabel(peps, j):
                continue
            tj = _apply_vec(tj, lj, np.sqrt(findbondtensor(peps, b)), b)
    tij = es.EinCode([li, lj], lij)(ti, tj)
    lijkl = [newlabel(peps, 1), newlabel(peps, 2), getphysicallabel(peps, i), getphysicallabel(peps, j)]
    lkl = [lijkl[0] if l == lijkl[2] else (lijkl[1] if l == lijkl[3] else l) for l in lij]
    tkl = es.EinCode([lij, lijkl], lkl)(tij, mat)
    sl = [ti.shape[li.index(l)] for l in only_left]
    sr = [tj.shape[lj.index(l)] for l in only_right]
    A = np.reshape(tkl, (int(np.prod(sl)), int(np.prod(sr))), order="F")
    U, S, Vh = np.linalg.svd(A, full_matrices=False)      # Julia: U, S, V = svd(A); A = U*Diagonal(S)*V'
    Vt = np.conj(Vh.conj().T)                               # ``Vt = conj.(V)`` (peps.jl:382)
    below = np.nonzero(S < peps.eps)[0]
    D = min(peps.Dmax, U.shape[1]) if below.size == 0 else min(int(below[0]), peps.Dmax)
    if D < U.shape[1]:
        LOG.append("truncation error is %r" % float(np.sum(S[D:])))
        S, U, Vt = S[:D], U[:, :D], Vt[:, :D]
    ti_ = es.EinCode([only_left + shared], li)(np.reshape(U, sl + [U.shape[1]], order="F"))
    tj_ = es.EinCode([only_right + shared], lj)(np.reshape(Vt, sr + [Vt.shape[1]], order="F"))
    if vidal:
        for b in only_left:
            if b == getphysicallabel(peps, i):
                continue
            ti_ = _apply_vec(ti_, li, safe_inv(np.sqrt(findbondtensor(peps, b))), b)
        for b in only_right:
            if b == getphysicallabel(peps, j):
                continue
            tj_ = _apply_vec(tj_, lj, safe_inv(np.sqrt(findbondtensor(peps, b))), b)
        peps.bond_tensors[peps.virtual_labels.index(shared[0])] = S
    else:
        sq = np.sqrt(S)
        ti_ = _apply_vec(ti_, li, sq, shared[0])
        tj_ = _apply_vec(tj_, lj, sq, shared[0])
    peps.vertex_tensors[i - 1] = ti_
    peps.vertex_tensors[j - 1] = tj_
    return peps


# -- Yao glue, peps.jl:432-477 ----------------------------------------------------------
def _match(peps, m):  # array_match_type, peps.jl:198-199
    m = np.asarray(m, dtype=np.complex128)
    return DiffTensor(m) if _isdiff(peps.vertex_tensors[0]) else m


def apply_block_(peps, block):
    """``YaoBlocks.unsafe_apply!`` methods, peps.jl:437-460."""
    if isinstance(block, yao.Put) and len(block.locs) == 1:
        return apply_onsite_(peps, block.locs[0], _match(peps, block.content))
    if isinstance(block, yao.Kron):
        for loc, m in zip(block.locs, block.subblocks):
            peps = apply_onsite_(peps, loc, _match(peps, m))
        return peps
    if isinstance(block, yao.Control):
        return apply_block_(peps, yao.Put(block.n, (block.ctrl, block.loc), block.two_site_matrix()))
    if isinstance(block, yao.Scale):
        return rmul_(apply_block_(peps, block.content), (1.0 + 0j) * block.factor)
    if isinstance(block, yao.Put) and len(block.locs) == 2:
        return apply_onbond_(peps, block.locs[0], block.locs[1],
                             np.reshape(block.content, (2, 2, 2, 2), order="F"))
    if isinstance(block, yao.Add):
        raise TypeError("Add is not a gate")
    raise TypeError(type(block))


def apply_block(peps, block):
    """Non-mutating ``YaoBlocks.apply`` methods (peps.jl:437-455); two-site ``put`` falls back to
    ``apply!(copy(peps), block)`` like Yao's generic method (peps.jl:456)."""
    if isinstance(block, yao.Put) and len(block.locs) == 1:
        return apply_onsite(peps, block.locs[0], _match(peps, block.content))
    if isinstance(block, yao.Kron):
        for loc, m in zip(block.locs, block.subblocks):
            peps = apply_onsite(peps, loc, _match(peps, m))
        return peps
    if isinstance(block, yao.Control):
        return apply_block(peps, yao.Put(block.n, (block.ctrl, block.loc), block.two_site_matrix()))
    if isinstance(block, yao.Scale):
        return mul(apply_block(peps, block.content), (1.0 + 0j) * block.factor)
    if isinstance(block, yao.Put) and len(block.locs) == 2:
        return apply_block_(peps.copy(), block)
    raise TypeError(type(block))


def expect(op, pa, pb):  # peps.jl:469-477
    if isinstance(op, yao.Add):
        out = None
        for term in op.blocks:
            e = expect(term, pa, pb)
            out = e if out is None else out + e
        return out
    opb = apply_block(pb, op)
    return inner_product(pa, opb)
