"""PEPS state layer on the device: same names, argument meaning and error behaviour as
``src/peps.jl`` (Julia's ``f!`` is spelled ``f_``), with ``vertex_tensors`` holding
:class:`B200Array` (or ``DiffTensor`` of them).  Sites, labels and bond numbering are 1-based
exactly like the reference.

Every tensor operation is a call into ``libtne_b200.so``:
``inner_product`` / ``statetensor`` run the PEPS's optimised code as one device tree program,
``expect`` over an ``Add`` is one ``tne_expect_terms`` call (terms share sub-contractions),
``apply_onbond_`` is the fused device bond update (gate + SVD truncation).
"""
import ctypes as C

import numpy as np

from . import _lib
from . import array as _arr
from . import einsum as es
from . import tensorad as ad
from . import yao
from .array import B200Array
from .tensorad import DiffTensor

LOG = []  # truncation messages (@info at src/peps.jl:386)


def _isdiff(x):
    return isinstance(x, DiffTensor)


def _dev(x):
    """numpy / B200Array / DiffTensor -> B200Array data."""
    if isinstance(x, DiffTensor):
        return x.data
    if isinstance(x, B200Array):
        return x
    return B200Array.from_numpy(np.asarray(x))


def _conj(x):
    return ad.conj(x) if _isdiff(x) else x.conj()


def _reshape(x, shape):
    return ad.reshape(x, tuple(shape)) if _isdiff(x) else x.reshape(tuple(shape))


class PEPS:
    """``PEPS{T,NF,LT}`` (src/peps.jl:4); fields as in ``SimplePEPS`` (src/peps.jl:36-50)."""

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

    def alllabels(self):  # src/peps.jl:96, 171
        if self.kind == "vidal":
            return self.vertex_labels + [[l] for l in self.virtual_labels]
        return self.vertex_labels

    def alltensors(self):  # src/peps.jl:98, 172
        if self.kind == "vidal":
            return self.vertex_tensors + self.bond_tensors
        return self.vertex_tensors

    def copy(self):  # src/peps.jl:99-101 (the Vidal method at :175-177 reads a missing field upstream)
        return replace_tensors(self, [ad.copy(t) if _isdiff(t) else t.copy() for t in self.alltensors()])

    def __repr__(self):  # src/peps.jl:236-241
        return "%s (Dmax = %d, ϵ = %g)\n # of spins = %d\n # of virtual degrees = %d" % (
            type(self).__name__, self.Dmax, self.eps, nsite(self), len(self.virtual_labels))


class SimplePEPS(PEPS):
    kind = "simple"


class VidalPEPS(PEPS):
    kind = "vidal"


def _optimized_code(alllabels, physical_labels, virtual_labels, max_ind, nflavor, D, optimizer, nslices, segments=None):
    """src/peps.jl:84-93.  ``segments``: ``[(site, [bond labels]), ...]`` for an explicit sweep order -- a
    checkpointed segment of the inner-product tree starts where ``site``'s bra tensor is absorbed and loops
    over the given bonds (bra and ket copy of each)."""
    size_dict = {l: nflavor for l in physical_labels}
    size_dict.update({l: D for l in virtual_labels})
    st_opt = "greedy" if isinstance(optimizer, (list, tuple)) else optimizer
    code_st = es.optimize_code(alllabels, physical_labels, size_dict, st_opt, nslices) if len(alllabels) > 1 else None
    rep = {l: max_ind + i + 1 for i, l in enumerate(virtual_labels)}
    for l in rep.values():
        size_dict[l] = D
    ixs = [list(l) for l in alllabels] + [[rep.get(x, x) for x in l] for l in alllabels]
    segs = None
    if segments and isinstance(optimizer, (list, tuple)):
        def pos_of(site):   # position of the item that holds the site's bra tensor (a leaf index or a (bra, ket) pair)
            for k, it in enumerate(optimizer):
                if it == site - 1 or (isinstance(it, (tuple, list)) and it[0] == site - 1):
                    return k
            raise ValueError("site %d is not in the contraction order" % site)
        segs = [(pos_of(site), [l for b in bonds for l in (b, rep[b])]) for site, bonds in segments]
    # optimizer = "sweep": the bra and the ket copy of a tensor are absorbed one right after the other; segments = "auto" /
    # "auto2": checkpointed segments and segment-local slices derived from the network (einsum.sweep_segments)
    nl = len(alllabels)
    if isinstance(segments, str):
        segs = segments
    code_ip = es.optimize_code(ixs, [], size_dict, optimizer, nslices, segments=segs, groups=[[i, nl + i] for i in range(nl)])
    return code_st, code_ip


def make_peps(kind, vertex_labels, vertex_tensors, virtual_labels, Dmax, eps, nflavor=2, bond_tensors=None,
              optimizer="greedy", nslices=0, segments=None):
    """The label-deriving constructors, src/peps.jl:69-82 and :142-151."""
    physical_labels = []
    for vl in vertex_labels:
        ph = [l for l in vl if l not in virtual_labels]
        if len(ph) != 1:
            raise ValueError("each vertex needs exactly one physical label")
        physical_labels.append(ph[0])
    for t, vl in zip(vertex_tensors, vertex_labels):  # src/peps.jl:73-75
        for s, l in zip(t.shape, vl):
            if (l in physical_labels and s != nflavor) or (l not in physical_labels and s > Dmax):
                raise ValueError("Input tensor shape error, either: \n 1. physical degrees of freedom do not match, "
                                 "\n 2. or input tensor virtual degrees of freedom larger than `Dmax = %d`." % Dmax)
    max_ind = max(max(physical_labels), max(virtual_labels))
    alll = [list(l) for l in vertex_labels]
    if kind == "vidal":
        alll = alll + [[l] for l in virtual_labels]
    cst, cip = _optimized_code(alll, physical_labels, virtual_labels, max_ind, nflavor, Dmax, optimizer, nslices, segments)
    cls = VidalPEPS if kind == "vidal" else SimplePEPS
    return cls(physical_labels, virtual_labels, vertex_labels, vertex_tensors, max_ind, cst, cip, Dmax, eps,
               nflavor, bond_tensors)


def replace_tensors(peps, tensors):  # src/peps.jl:222-234
    tensors = list(tensors)
    if peps.kind == "vidal":
        nv = nsite(peps)
        return VidalPEPS(peps.physical_labels, peps.virtual_labels, peps.vertex_labels, tensors[:nv], peps.max_index,
                         peps.code_statetensor, peps.code_inner_product, peps.Dmax, peps.eps, peps.nflavor,
                         tensors[nv:])
    return SimplePEPS(peps.physical_labels, peps.virtual_labels, peps.vertex_labels, tensors, peps.max_index,
                      peps.code_statetensor, peps.code_inner_product, peps.Dmax, peps.eps, peps.nflavor)


def conj(peps):  # src/peps.jl:178-180
    return replace_tensors(peps, [_conj(t) for t in peps.alltensors()])


# -- label API, src/peps.jl:183-199 -----------------------------------------------------
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


def nflavor(peps):
    return peps.nflavor


def Dmax(peps):
    return peps.Dmax


# -- variables, src/peps.jl:202-220 -----------------------------------------------------
def variables(peps):
    """``vcat(vec.(alltensors(peps))...)`` as a device vector."""
    return _arr.concat([_dev(t) for t in peps.alltensors()])


def load_variables_(peps, variables):
    variables = _dev(variables)
    k = 0
    nv = len(peps.vertex_tensors)
    for idx, t in enumerate(peps.alltensors()):
        new = variables.slice(k, t.size).reshape(t.shape).copy()
        if idx < nv:
            peps.vertex_tensors[idx] = new
        else:
            peps.bond_tensors[idx - nv] = new
        k += t.size
    return peps


def load_variables(peps, variables):
    if peps.kind != "simple":
        raise TypeError("load_variables is only defined for SimplePEPS (src/peps.jl:212)")
    if not isinstance(variables, (DiffTensor, B200Array)):
        variables = B200Array.from_numpy(np.asarray(variables))
    ats = peps.alltensors()
    tensors, k = [], 0
    for t in ats:
        if _isdiff(variables):
            tensors.append(ad.reshape(variables[k:k + t.size], t.shape))
        else:
            tensors.append(variables.slice(k, t.size).reshape(t.shape))
        k += t.size
    return replace_tensors(peps, tensors)


# -- factories, src/peps.jl:244-280 -----------------------------------------------------
def sweep_plan(g, L, nbonds=1, pair_last_row=True, slice_first_row=False):
    """Boundary-sweep contraction plan for an ``L x L`` ``grid`` PEPS (vertex ``v = x + (y-1) L``): the leaf
    order ``bra_1, ket_1, bra_2, ...`` and checkpointed segments with segment-local slices.

    ``nbonds = 1``: one segment per lattice column ``y``; column ``y >= 2`` loops over the bond between the last
    sites of columns ``y-1`` and ``y``: it is open during the whole column, so slicing it costs no flops and divides
    the column's tape by ``D^2``.
    ``pair_last_row``: the bra and ket tensor of the last site of every sliced column are contracted with each other
    first (a 16 x 16 matrix at D = 4) and absorbed in one step.
    ``nbonds = k > 1`` (bigger lattices): the bonds of the last ``k`` rows between columns ``y-1`` and ``y`` are all
    looped over (``D^(2k)`` slices).  The bond of row ``r`` closes at site ``(r, y)``, so the column is cut into
    ``k`` segments -- rows ``1..L-k+1``, then one row each -- and every segment only loops over the bonds that are
    still open in it: again no flops are repeated.
    ``slice_first_row`` (big lattices, forward only: the exact 8 x 8, D = 4 norm): every column also loops over the bond
    that leaves its FIRST site towards column ``y+1``.  That leg is created by the column's very first absorption, so
    every intermediate of the column carries one fixed value of it -- ``D^2`` times smaller again -- and nothing but the
    first (tiny) bra step is repeated; the column's output is written slice by slice through a strided view of the
    next boundary.  With both bonds looped over the workspace of 8 x 8, D = 4 is two 64 GiB boundaries plus 12 GiB."""
    n = g.nv()
    order = []
    for i in range(n):
        x, y = i % L + 1, i // L + 1
        # last-row sites of the sliced columns have neither down legs nor (inside a slice) left legs: contracting
        # bra with ket first costs no extra flops and the boundary is read / written once instead of twice
        if pair_last_row and x == L and y >= 2 and L > 1:
            order.append((i, n + i))
        else:
            order += [i, n + i]
    emap = {e: n + 1 + k for k, e in enumerate(g.edges())}
    site = lambda x, y: x + (y - 1) * L
    hbond = lambda x, y: emap[(site(x, y - 1), site(x, y))]    # bond between (x, y-1) and (x, y)
    k = max(1, min(int(nbonds), L - 1))
    if slice_first_row:
        if k != 1:
            raise ValueError("slice_first_row needs nbonds = 1")
        return order, [(site(1, y), ([hbond(L, y)] if y >= 2 else []) + ([hbond(1, y + 1)] if y < L else [])) for y in range(1, L + 1)]
    segments = [(1, [])]
    for y in range(2, L + 1):
        # segment 0 of the column: rows 1 .. L-k+1, all k bonds open; then one segment per remaining row
        first_rows = [1] + list(range(L - k + 2, L + 1))
        for j, x0 in enumerate(first_rows):
            open_rows = range(L - k + 1 + j, L + 1) if j > 0 else range(L - k + 1, L + 1)
            segments.append((site(x0, y), [hbond(x, y) for x in open_rows]))
    return order, segments


def _peps_zero_state(kind, g, D, nflavor=2, Dmax=None, eps=1e-12, optimizer="greedy", nslices=0, segments=None):
    Dmax = D if Dmax is None else Dmax
    nv, ne = g.nv(), g.ne()
    virtual_labels = list(range(nv + 1, nv + ne + 1))
    edge_map = dict(zip(g.edges(), virtual_labels))
    vertex_labels, vertex_tensors = [], []
    for i in range(1, nv + 1):
        labs = [i] + [edge_map.get((i, nb), edge_map.get((nb, i), 0)) for nb in g.neighbors(i)]
        t = np.zeros([nflavor] + [D] * g.degree(i), dtype=np.complex128)
        t[(0,) * t.ndim] = 1
        vertex_labels.append(labs)
        vertex_tensors.append(B200Array.from_numpy(t))
    if any(0 in vl for vl in vertex_labels):
        raise ValueError("incorrect input labels1")
    bonds = [B200Array.from_numpy(np.ones(D, dtype=np.complex128)) for _ in range(ne)] if kind == "vidal" else None
    return make_peps(kind, vertex_labels, vertex_tensors, virtual_labels, Dmax, eps, nflavor, bonds, optimizer, nslices, segments)


def zero_simplepeps(g, D, **kw):
    return _peps_zero_state("simple", g, D, **kw)


def zero_vidalpeps(g, D, **kw):
    return _peps_zero_state("vidal", g, D, **kw)


def randn_(peps, rng):
    """``randn!(peps)`` (src/peps.jl:275-280): iid circular complex normal with unit variance, filled
    in ``alltensors`` order, column-major, from a numpy Generator."""
    new = []
    for t in peps.alltensors():
        n = t.size
        z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2.0)
        new.append(B200Array.from_numpy(z.reshape(t.shape, order="F")))
    nv = len(peps.vertex_tensors)
    peps.vertex_tensors = new[:nv]
    if peps.kind == "vidal":
        peps.bond_tensors = new[nv:]
    return peps


def rand_simplepeps(g, D, rng, **kw):
    return randn_(zero_simplepeps(g, D, **kw), rng)


def rand_vidalpeps(g, D, rng, **kw):
    return randn_(zero_vidalpeps(g, D, **kw), rng)


def from_numpy_tensors(template, tensors):
    """Load host arrays (Julia shapes) into a PEPS with ``template``'s structure."""
    return replace_tensors(template, [B200Array.from_numpy(np.asarray(t)) for t in tensors])


# -- scalar ops and norm, src/peps.jl:283-306 -------------------------------------------
def rmul_(peps, c):
    f = complex(c) ** (1.0 / nsite(peps)) if np.iscomplexobj(c) else c ** (1.0 / nsite(peps))
    peps.vertex_tensors = [t * f for t in peps.vertex_tensors]
    return peps


def mul(peps, c):
    """``peps * c`` for a number (src/peps.jl:288-291) or a 0-dim array (src/peps.jl:292-295)."""
    n = nsite(peps)
    if _isdiff(c):
        f = ad.bpow(c, 1.0 / n)
        tensors = [t * f for t in peps.vertex_tensors]
    elif isinstance(c, B200Array):
        f = c.pow(1.0 / n)
        tensors = [t * (DiffTensor(f, requires_grad=False) if _isdiff(t) else f) for t in peps.vertex_tensors]
    else:
        f = c ** (1.0 / n)
        tensors = [t * f for t in peps.vertex_tensors]
    rest = peps.alltensors()[len(peps.vertex_tensors):]
    return replace_tensors(peps, tensors + list(rest))


def inner_product(p1, p2):  # src/peps.jl:106-110
    p1c = conj(p1)
    return p1.code_inner_product(*p1c.alltensors(), *p2.alltensors())


def slice_bonds(peps, bonds, index, host=None):
    """One term of a SlicedEinsum over virtual bonds (the reference's ``TreeSA(nslices = k)`` codes, test/peps.jl:9, loop
    over the values of the sliced labels): the PEPS with bond ``bonds[k]`` fixed at ``index[k]`` (0-based), kept as a leg of
    dimension 1 so that labels, contraction order and segments stay valid.  ``host``: cached numpy copies of the tensors."""
    host = host if host is not None else [t.to_numpy() for t in peps.alltensors()]
    fix = dict(zip(bonds, index))
    out = []
    for vl, a in zip(peps.vertex_labels, host):
        if any(l in fix for l in vl):
            a = a[tuple(slice(fix[l], fix[l] + 1) if l in fix else slice(None) for l in vl)]
            out.append(B200Array.from_numpy(np.asfortranarray(a)))
        else:
            out.append(None)
    tens = [o if o is not None else t for o, t in zip(out, peps.alltensors())]
    return replace_tensors(peps, tens)


def inner_product_sliced(p1, p2, bonds, subset=None, shard=(0, 1)):
    """``<p1|p2>`` as the sum over the values of the sliced ``bonds`` (bra and ket copy of each bond are sliced
    independently: ``D^(2 len(bonds))`` terms) -- BASELINE configs[4]: every term is an ordinary contraction with
    dimension-1 legs, the terms are dealt round-robin to the ranks (``shard = (rank, world)``) and the caller all-reduces
    the partial sums.  ``subset``: a list of ``(bra index tuple, ket index tuple)`` to evaluate instead of all of them
    (timing a fixed subset of the slices).  Returns ``(partial sum as a Python complex, number of terms evaluated)``."""
    import itertools
    dims = []
    for b in bonds:
        for vl, t in zip(p1.vertex_labels, p1.vertex_tensors):
            if b in vl:
                dims.append(t.shape[vl.index(b)])
                break
    if subset is None:
        one = list(itertools.product(*[range(d) for d in dims]))
        subset = [(a, b) for a in one for b in one]
    h1 = [t.to_numpy() for t in p1.alltensors()]
    h2 = h1 if p2 is p1 else [t.to_numpy() for t in p2.alltensors()]
    total, n = 0.0 + 0.0j, 0
    rank, world = shard
    for k, (ia, ib) in enumerate(subset):
        if k % world != rank:
            continue
        total += complex(inner_product(slice_bonds(p1, bonds, ia, h1), slice_bonds(p2, bonds, ib, h2)).to_numpy())
        n += 1
    return total, n


def norm(peps):  # src/peps.jl:302
    ip = inner_product(peps, peps)
    if _isdiff(ip):
        return ad.bsqrt(ad.babs(ip))
    return ip.abs().sqrt()


def normalize_(peps):  # src/peps.jl:303-306
    nm = inner_product(peps, peps).abs().sqrt()
    return rmul_(peps, 1.0 / nm.item())


def statetensor(peps):  # src/peps.jl:320-322
    if peps.code_statetensor is None:
        return peps.alltensors()[0]
    return peps.code_statetensor(*peps.alltensors())


def vec(peps):  # src/peps.jl:319
    return np.reshape(_dev(statetensor(peps)).to_numpy(), -1, order="F")


statevec = vec


# -- gates, src/peps.jl:329-430 ---------------------------------------------------------
def _onsite_code(peps, i):
    old = getvlabel(peps, i)
    mlabel = [newlabel(peps, 1), getphysicallabel(peps, i)]
    return es.EinCode([old, mlabel], [mlabel[0] if l == mlabel[1] else l for l in old])


def _as_operand(peps, m):  # array_match_type, src/peps.jl:198-199
    if isinstance(m, (DiffTensor, B200Array)):
        return m
    m = B200Array.from_numpy(np.asarray(m, dtype=np.complex128))
    return DiffTensor(m) if _isdiff(peps.vertex_tensors[0]) else m


def apply_onsite_(peps, i, mat):  # src/peps.jl:329-336
    if mat.shape[0] != mat.shape[1]:
        raise AssertionError("size(mat, 1) == size(mat, 2)")
    peps.vertex_tensors[i - 1] = _onsite_code(peps, i)(peps.vertex_tensors[i - 1], _as_operand(peps, mat))
    return peps


def apply_onsite(peps, i, mat):  # src/peps.jl:338-347
    if mat.shape[0] != mat.shape[1]:
        raise AssertionError("size(mat, 1) == size(mat, 2)")
    code = _onsite_code(peps, i)
    m = _as_operand(peps, mat)
    tensors = [code(t, m) if j == i - 1 else t for j, t in enumerate(peps.vertex_tensors)]
    rest = peps.alltensors()[len(peps.vertex_tensors):]
    return replace_tensors(peps, tensors + list(rest))


def bond_job(peps, i, j, mat):
    """Describe ``apply_onbond!(peps, i, j, mat)`` for ``tne_apply_onbond_batched``."""
    li, lj = getvlabel(peps, i), getvlabel(peps, j)
    shared = [l for l in li if l in lj]
    if len(shared) != 1:
        raise AssertionError("length(shared_label) == 1")
    job = _lib.TneBondJob()
    job.ti = _dev(peps.vertex_tensors[i - 1]).handle
    job.tj = _dev(peps.vertex_tensors[j - 1]).handle
    job.axis_i = li.index(shared[0])
    job.axis_j = lj.index(shared[0])
    g = np.asarray(mat, dtype=np.complex128).reshape(4, 4, order="F")  # (o_i + 2 o_j, i_i + 2 i_j)
    flat = np.asfortranarray(g).reshape(-1, order="F").view(np.float64)
    for q in range(32):
        job.gate[q] = flat[q]
    job.Dmax = peps.Dmax
    job.eps = peps.eps
    job.vidal = 1 if peps.kind == "vidal" else 0
    if peps.kind == "vidal":
        for q, b in enumerate(li):
            job.lambda_i[q] = 0 if q == 0 else _dev(findbondtensor(peps, b)).handle
        for q, b in enumerate(lj):
            job.lambda_j[q] = 0 if q == 0 else _dev(findbondtensor(peps, b)).handle
    return job, shared[0]


def apply_onbond_(peps, i, j, mat):  # src/peps.jl:353-417
    return apply_onbond_batch_(peps, [(i, j, mat)])


def apply_onbond_batch_(peps, gates):
    """``apply_onbond!`` for a list of ``(i, j, mat)`` on pairwise disjoint sites in ONE device launch (one CTA
    per bond: contract -> gate -> QR-reduced one-sided Jacobi SVD -> truncate -> split / rescale)."""
    sites = [s for i, j, _ in gates for s in (i, j)]
    if len(set(sites)) != len(sites):
        raise ValueError("bond updates of one batch must act on disjoint sites")
    for i, j, _ in gates:
        if getphysicallabel(peps, i) != getvlabel(peps, i)[0] or getphysicallabel(peps, j) != getvlabel(peps, j)[0]:
            raise ValueError("the device bond update expects the physical leg first (src/peps.jl:253)")
    built = [bond_job(peps, i, j, mat) for i, j, mat in gates]
    ctx = _dev(peps.vertex_tensors[gates[0][0] - 1]).ctx
    jobs = (_lib.TneBondJob * len(built))(*[b[0] for b in built])
    ctx.check(ctx.L.tne_apply_onbond_batched(ctx.h, len(built), jobs))
    for (i, j, _), (_, shared), job in zip(gates, built, jobs):
        si = list(_dev(peps.vertex_tensors[i - 1]).shape)
        sj = list(_dev(peps.vertex_tensors[j - 1]).shape)
        si[job.axis_i] = job.kept
        sj[job.axis_j] = job.kept
        if job.trunc_err >= 0:
            LOG.append("truncation error is %r" % job.trunc_err)
        peps.vertex_tensors[i - 1] = B200Array(job.ti_out, si, False, ctx)
        peps.vertex_tensors[j - 1] = B200Array(job.tj_out, sj, False, ctx)
        s = B200Array(job.s_out, (job.kept,), True, ctx)
        if peps.kind == "vidal":
            peps.bond_tensors[peps.virtual_labels.index(shared)] = s
    return peps


class SweepPlan:
    """The label bookkeeping of a gate sequence, done once: the ``tne_sweep_gate`` array of ``tne_tebd_sweep`` (site and bond
    indices, shared-leg positions, gate matrices).  Valid for every PEPS with the same labels; ``kept`` / ``trunc_err`` are
    overwritten by each sweep."""

    def __init__(self, peps, gates):
        for i, j, _ in gates:
            if getphysicallabel(peps, i) != getvlabel(peps, i)[0] or getphysicallabel(peps, j) != getvlabel(peps, j)[0]:
                raise ValueError("the device bond update expects the physical leg first (src/peps.jl:253)")
        self.vidal = peps.kind == "vidal"
        self.sites = [(i, j) for i, j, _ in gates]
        bond_index = {l: k for k, l in enumerate(peps.virtual_labels)}
        n = len(gates)
        # filled through a numpy view of the ctypes array (one vectorised store per field instead of ~70 per gate)
        self.arr = (_lib.TneSweepGate * n)()
        view = np.frombuffer(self.arr, dtype=np.dtype(_lib.TneSweepGate))
        lam_i = np.full((n, _lib.MAX_RANK), -1, dtype=np.int32)
        lam_j = np.full((n, _lib.MAX_RANK), -1, dtype=np.int32)
        gmat = np.empty((n, 32), dtype=np.float64)
        head = np.empty((n, 5), dtype=np.int32)
        for q, (i, j, mat) in enumerate(gates):
            li, lj = getvlabel(peps, i), getvlabel(peps, j)
            shared = [l for l in li if l in lj]
            if len(shared) != 1:
                raise AssertionError("length(shared_label) == 1")
            head[q] = (i - 1, j - 1, li.index(shared[0]), lj.index(shared[0]), bond_index[shared[0]])
            g = np.asarray(mat, dtype=np.complex128).reshape(4, 4, order="F")  # (o_i + 2 o_j, i_i + 2 i_j)
            gmat[q] = np.asfortranarray(g).reshape(-1, order="F").view(np.float64)
            if self.vidal:
                lam_i[q, 1:len(li)] = [bond_index[l] for l in li[1:]]
                lam_j[q, 1:len(lj)] = [bond_index[l] for l in lj[1:]]
        for k, name in enumerate(("site_i", "site_j", "axis_i", "axis_j", "bond")):
            view[name] = head[:, k]
        view["gate"] = gmat
        view["lambda_i"] = lam_i
        view["lambda_j"] = lam_j
        self.view = view


def apply_sweep_(peps, gates, n_sweeps=1):
    """``apply_onbond!`` for a whole sequence ``[(i, j, mat), ...]`` (one TEBD sweep, src/tebd.jl:16-21) in ONE library
    call: ``tne_tebd_sweep`` levels the sequence into wavefronts of site-disjoint gates and enqueues them without any
    host synchronisation in between.  Results and the ``LOG`` order are those of the sequential loop.  ``gates`` may be a
    prepared :class:`SweepPlan` (the drivers reuse one plan for all sweeps).  ``n_sweeps > 1``: the same sequence that many
    times back to back (``rydberg_tebd!``'s loop, src/tebd.jl:27-29) in one ``tne_tebd_sweeps`` call: no host round trip and
    no synchronisation between the sweeps either."""
    plan = gates if isinstance(gates, SweepPlan) else SweepPlan(peps, gates)
    n = len(plan.sites)
    if n == 0:
        return peps
    vidal = peps.kind == "vidal"
    if vidal != plan.vidal:
        raise ValueError("the sweep plan was prepared for the other PEPS kind")
    ctx = _dev(peps.vertex_tensors[0]).ctx
    nv, nb = len(peps.vertex_tensors), len(peps.virtual_labels)
    old_sites = [_dev(t) for t in peps.vertex_tensors]
    sites = (C.c_int64 * nv)(*[t.handle for t in old_sites])
    old_bonds = [_dev(t) for t in peps.bond_tensors] if vidal else []
    bonds = (C.c_int64 * max(1, nb))(*([t.handle for t in old_bonds] if vidal else [0]))
    n_sweeps = int(n_sweeps)
    if n_sweeps < 1:
        return peps
    kept_all = (C.c_int32 * (n * n_sweeps))()
    err_all = (C.c_double * (n * n_sweeps))()
    ctx.check(ctx.L.tne_tebd_sweeps(ctx.h, nv, sites, nb if vidal else 0, bonds, n, plan.arr, int(peps.Dmax), float(peps.eps),
                                    1 if vidal else 0, n_sweeps, kept_all, err_all))
    # shapes: the bond a gate rewrote now has the rank it kept (the last gate on a bond wins)
    shapes = [list(t.shape) for t in old_sites]
    blen = {}
    kept, err = plan.view["kept"].tolist(), plan.view["trunc_err"].tolist()
    ax_i, ax_j, bnd = plan.view["axis_i"].tolist(), plan.view["axis_j"].tolist(), plan.view["bond"].tolist()
    for e in list(err_all)[:n * (n_sweeps - 1)]:      # the @info lines of the earlier sweeps, in the sequential order
        if e >= 0:
            LOG.append("truncation error is %r" % e)
    for q, (i, j) in enumerate(plan.sites):
        shapes[i - 1][ax_i[q]] = kept[q]
        shapes[j - 1][ax_j[q]] = kept[q]
        blen[bnd[q]] = kept[q]
        if err[q] >= 0:
            LOG.append("truncation error is %r" % err[q])
    for k in range(nv):
        if sites[k] != old_sites[k].handle:
            peps.vertex_tensors[k] = B200Array(sites[k], shapes[k], False, ctx)
    if vidal:
        for k in range(nb):
            if bonds[k] != old_bonds[k].handle:
                peps.bond_tensors[k] = B200Array(bonds[k], (blen[k],), True, ctx)
    return peps


def apply_sweep_batch_(pepses, gates, n_sweeps=1):
    """One sweep ``[(i, j, mat), ...]`` applied to R independent PEPS of the SAME structure (replicas: an ensemble of
    initial states / parameter points) in ONE ``tne_tebd_sweep`` call.  The site and bond tables of the replicas are
    concatenated and the gate list is tiled with offset indices; the library's wavefront levelling then puts the R copies
    of a gate into the same launch, so a wavefront runs ``R x`` more CTAs instead of 4-6 on 148 SMs.  TEBD does not shard
    inside one PEPS (replicas only) -- this is how it fills a GPU, and R / n_gpus replicas per rank is its
    multi-GPU mode (no communication).  Every replica gets exactly the tensors ``apply_sweep_`` would give it."""
    pepses = list(pepses)
    if not pepses:
        return pepses
    p0 = pepses[0]
    plan = gates if isinstance(gates, SweepPlan) else SweepPlan(p0, gates)
    n, R = len(plan.sites), len(pepses)
    if n == 0:
        return pepses
    vidal = p0.kind == "vidal"
    nv, nb = len(p0.vertex_tensors), len(p0.virtual_labels)
    for q in pepses:
        if q.kind != p0.kind or q.vertex_labels != p0.vertex_labels or q.Dmax != p0.Dmax or q.eps != p0.eps:
            raise ValueError("replicas must share the PEPS structure, Dmax and eps")
    if vidal != plan.vidal:
        raise ValueError("the sweep plan was prepared for the other PEPS kind")
    tiled = plan.__dict__.setdefault("_tiled", {})
    if R not in tiled:
        arr = (_lib.TneSweepGate * (n * R))()
        view = np.frombuffer(arr, dtype=np.dtype(_lib.TneSweepGate))
        for r in range(R):
            blk = view[r * n:(r + 1) * n]
            blk[:] = plan.view
            blk["site_i"] += r * nv
            blk["site_j"] += r * nv
            blk["bond"] += r * nb
            for name in ("lambda_i", "lambda_j"):
                lam = blk[name]
                lam[lam >= 0] += r * nb
        tiled[R] = (arr, view)
    arr, view = tiled[R]
    ctx = _dev(p0.vertex_tensors[0]).ctx
    old_sites = [_dev(t) for q in pepses for t in q.vertex_tensors]
    sites = (C.c_int64 * (nv * R))(*[t.handle for t in old_sites])
    old_bonds = [_dev(t) for q in pepses for t in q.bond_tensors] if vidal else []
    bonds = (C.c_int64 * max(1, nb * R))(*([t.handle for t in old_bonds] if vidal else [0]))
    n_sweeps = int(n_sweeps)
    if n_sweeps < 1:
        return pepses
    err_all = (C.c_double * (n * R * n_sweeps))()
    ctx.check(ctx.L.tne_tebd_sweeps(ctx.h, nv * R, sites, nb * R if vidal else 0, bonds, n * R, arr, int(p0.Dmax), float(p0.eps),
                                    1 if vidal else 0, n_sweeps, None, err_all))
    for e in list(err_all)[:n * R * (n_sweeps - 1)]:
        if e >= 0:
            LOG.append("truncation error is %r" % e)
    kept, err = view["kept"].tolist(), view["trunc_err"].tolist()
    ax_i, ax_j, bnd = plan.view["axis_i"].tolist(), plan.view["axis_j"].tolist(), plan.view["bond"].tolist()
    for r, q in enumerate(pepses):
        shapes = [list(t.shape) for t in old_sites[r * nv:(r + 1) * nv]]
        blen = {}
        for g_, (i, j) in enumerate(plan.sites):
            kq = kept[r * n + g_]
            shapes[i - 1][ax_i[g_]] = kq
            shapes[j - 1][ax_j[g_]] = kq
            blen[bnd[g_]] = kq
            if err[r * n + g_] >= 0:
                LOG.append("truncation error is %r" % err[r * n + g_])
        for k in range(nv):
            if sites[r * nv + k] != old_sites[r * nv + k].handle:
                q.vertex_tensors[k] = B200Array(sites[r * nv + k], shapes[k], False, ctx)
        if vidal:
            for k in range(nb):
                if bonds[r * nb + k] != old_bonds[r * nb + k].handle:
                    q.bond_tensors[k] = B200Array(bonds[r * nb + k], (blen[k],), True, ctx)
    return pepses


# -- Yao glue, src/peps.jl:432-477 ------------------------------------------------------
def apply_block_(peps, block):
    """``YaoBlocks.unsafe_apply!`` methods, src/peps.jl:437-460."""
    if isinstance(block, yao.Put) and len(block.locs) == 1:
        return apply_onsite_(peps, block.locs[0], block.content)
    if isinstance(block, yao.Kron):
        for loc, m in zip(block.locs, block.subblocks):
            peps = apply_onsite_(peps, loc, m)
        return peps
    if isinstance(block, yao.Control):
        return apply_block_(peps, yao.Put(block.n, (block.ctrl, block.loc), block.two_site_matrix()))
    if isinstance(block, yao.Scale):
        return rmul_(apply_block_(peps, block.content), (1.0 + 0j) * block.factor)
    if isinstance(block, yao.Put) and len(block.locs) == 2:
        return apply_onbond_(peps, block.locs[0], block.locs[1], np.reshape(block.content, (2, 2, 2, 2), order="F"))
    raise TypeError("no apply! method for %s" % type(block).__name__)


def apply_block(peps, block):
    """Non-mutating ``YaoBlocks.apply`` (src/peps.jl:437-455); a two-site ``put`` falls back to
    ``apply!(copy(peps), block)`` like Yao's generic method (src/peps.jl:456)."""
    if isinstance(block, yao.Put) and len(block.locs) == 1:
        return apply_onsite(peps, block.locs[0], block.content)
    if isinstance(block, yao.Kron):
        for loc, m in zip(block.locs, block.subblocks):
            peps = apply_onsite(peps, loc, m)
        return peps
    if isinstance(block, yao.Control):
        return apply_block(peps, yao.Put(block.n, (block.ctrl, block.loc), block.two_site_matrix()))
    if isinstance(block, yao.Scale):
        return mul(apply_block(peps, block.content), (1.0 + 0j) * block.factor)
    if isinstance(block, yao.Put) and len(block.locs) == 2:
        return apply_block_(peps.copy(), block)
    raise TypeError("no apply method for %s" % type(block).__name__)


def _term_replacements(pb, block, cache=None):
    """(coef, {site: replacement tensor}) of one Hamiltonian term applied to the ket, or None if the
    term is not a product of one-site operators (then the generic path is used).  Scalar factors of the
    one-site matrices are moved into ``coef`` and equal (site, matrix) pairs share ONE device tensor
    (``cache``), so that terms which agree on a site share their partial contractions in the device
    programme (C P1_i P1_j and C P1_i P1_k open the same pattern at site i)."""
    cache = {} if cache is None else cache
    coef = 1.0 + 0j
    while isinstance(block, yao.Scale):
        coef *= block.factor
        block = block.content
    if isinstance(block, yao.Put) and len(block.locs) == 1:
        pairs = [(block.locs[0], block.content)]
    elif isinstance(block, yao.Kron):
        pairs = list(zip(block.locs, block.subblocks))
    else:
        return None
    merged = {}
    for loc, m in pairs:   # several factors on one site multiply (later factors act last)
        m = np.asarray(m, dtype=np.complex128)
        merged[loc] = m @ merged[loc] if loc in merged else m
    repl = {}
    for loc, m in merged.items():
        flat = m.reshape(-1, order="F")
        nz = np.flatnonzero(np.abs(flat) > 0)
        if nz.size == 0:
            return 0.0 + 0j, {}
        nu = flat[nz[0]]
        coef *= nu
        m = m / nu
        key = (loc, m.tobytes())
        if key not in cache:
            old = getvlabel(pb, loc)
            mlabel = [newlabel(pb, 1), getphysicallabel(pb, loc)]
            cache[key] = es.einsum([old, mlabel], [_dev(pb.vertex_tensors[loc - 1]), B200Array.from_numpy(m)],
                                   [mlabel[0] if l == mlabel[1] else l for l in old])
        repl[loc] = cache[key]
    return coef, repl


def expect(op, pa, pb, shard=(0, 1)):  # src/peps.jl:469-477
    """``expect(op, pa, pb)``.  An ``Add`` of one-site-product terms runs as ONE device program in which
    the terms share sub-contractions; anything else follows the reference term by term.  ``shard =
    (rank, world)`` keeps only this rank's terms (the caller sums the partial values)."""
    if isinstance(op, yao.Add):
        fused = es.FUSED_TREES and len(pa.alltensors()) == len(pb.alltensors())
        cache = {}
        terms = [_term_replacements(pb, t, cache) for t in op.blocks] if fused else None
        # the fused programme returns pullbacks for the bra leaves only: a ket that requires gradients (the reference
        # tapes apply -> inner_product for both arguments, src/peps.jl:474-477) takes the term-by-term taped path
        ket_grad = any(_isdiff(t) and ad.requires_grad(t) for t in pb.alltensors())
        if fused and not ket_grad and all(t is not None for t in terms) and \
                all(_isdiff(t) for t in pa.alltensors()) == any(_isdiff(t) for t in pa.alltensors()):
            return _expect_fused(terms, pa, pb, shard)
        out = None
        for k, term in enumerate(op.blocks):
            if k % shard[1] != shard[0]:
                continue
            e = expect(term, pa, pb)
            out = e if out is None else out + e
        return _zero_scalar(pa) if out is None else out
    opb = apply_block(pb, op)
    return inner_product(pa, opb)


def _zero_scalar(pa):
    """The value of an empty sum of terms (a rank whose share of the Hamiltonian is empty): an explicit 0-dim zero
    that carries no gradient -- NOT the plain overlap <pa|pb> that ``tne_expect_terms`` evaluates for ``n_terms == 0``."""
    t0 = pa.alltensors()[0]
    z = B200Array.from_numpy(np.zeros((), dtype=np.complex128), _dev(t0).ctx)
    return DiffTensor(z, requires_grad=False) if _isdiff(t0) else z


def _expect_fused(terms, pa, pb, shard):
    code = pa.code_inner_product
    nt = len(pa.alltensors())
    bra = pa.alltensors()
    ket = pb.alltensors()
    diff = any(_isdiff(t) for t in bra)
    leaves = [_dev(t) for t in bra] + [_dev(t) for t in ket]
    cj = [1] * nt + [0] * nt
    tlist = []
    for k, (coef, repl) in enumerate(terms):
        if k % shard[1] != shard[0]:
            continue
        tlist.append((coef, [(nt + site - 1, t) for site, t in sorted(repl.items())]))
    if not tlist:
        return _zero_scalar(pa)
    if not diff:
        val, _ = es.expect_terms_device(code, leaves, cj, tlist)
        return val
    # value and the pullbacks for a unit cotangent in ONE device program (forward + reverse sweep); the
    # pullback is linear in the cotangent, so back(dy) only rescales the stored gradients
    gl = [i for i in range(nt) if ad.requires_grad(bra[i])]
    val, grads = es.expect_terms_device(code, leaves, cj, tlist, gl, 1.0 + 0j)
    gmap = dict(zip(gl, grads))

    def mk(i):
        def back(dy):
            # bra leaves enter conjugated: the pullback is antilinear in the cotangent
            return DiffTensor(gmap[i].bmul(dy.data.to_complex().conj()), requires_grad=False)
        return back
    return ad.difftensor(val, "expect", *[(bra[i], mk(i)) for i in range(nt)])
