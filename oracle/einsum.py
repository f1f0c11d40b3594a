"""OMEinsum 0.7 restatement (not vendored in the reference; ``Project.toml:19``).

TEST INFRASTRUCTURE ONLY (see oracle/__init__).  Restates the published algorithm the
reference's call sites rely on (``src/peps.jl:85-92,109,321,334,344,373,376,392-393,429``;
``src/TensorAD/rules.jl:1-8``):

* ``EinCode(ixs, iy)`` is callable on tensors; a binary contraction is executed as
  TTGT: classify labels (batch / outer-left / contracted / outer-right), ``permutedims``
  -> ``reshape`` to 3-D -> batched ``gemm`` -> ``reshape`` -> ``permutedims``.
* ``NestedEinsum`` is a binary tree of ``EinCode`` s; ``SlicedEinsum`` loops serially over
  the assignments of the sliced labels and accumulates.
* ``optimize_code(code, size_dict, GreedyMethod())`` = repeatedly contract the pair with the
  smallest ``size(out) - size(a) - size(b)`` (deterministic tie-break here; the reference's
  is randomised, any order is mathematically identical).
* ``einsum_grad(ixs, xs, iy, size_dict, cdy, i)`` = ``conj(einsum(ixs[i->iy] -> ixs[i]; xs[i->cdy]))``
  projected to the element type of ``xs[i]``.

Arrays are numpy arrays whose ``shape`` equals the Julia ``size``; einsum semantics do
not depend on memory order.  ``vec``/``reshape`` (column-major in Julia) live in
:mod:`oracle.peps` and always pass ``order='F'``.
"""
import itertools

import numpy as np

# hook installed by oracle.tensorad: called when every operand is a DiffTensor
_DIFF_HOOK = None
FLOP_COUNTER = {"flops": 0, "bytes": 0, "enabled": False}


def _size_dict(ixs, xs, size_dict=None):
    sd = dict(size_dict) if size_dict else {}
    for ix, x in zip(ixs, xs):
        for l, s in zip(ix, x.shape):
            if l in sd and sd[l] != s and False:
                raise ValueError("inconsistent size for label %r" % (l,))
            sd[l] = s
    return sd


def _unary(ix, x, iy, sd):
    ix = list(ix)
    iy = list(iy)
    if len(set(ix)) == len(ix) and sorted(map(repr, ix)) == sorted(map(repr, iy)) and len(set(iy)) == len(iy):
        return np.transpose(x, [ix.index(l) for l in iy])  # pure permutation
    # general unary rule (sum / trace / diag / repeat) through numpy's einsum
    labs = {l: k for k, l in enumerate(dict.fromkeys(list(ix) + list(iy)))}
    missing = [l for l in iy if l not in ix]
    src_iy = [l for l in iy if l in ix]
    y = np.einsum(x, [labs[l] for l in ix], [labs[l] for l in dict.fromkeys(src_iy)])
    cur = list(dict.fromkeys(src_iy))
    if missing or cur != iy:
        # broadcast missing labels / duplicate output labels
        full = list(dict.fromkeys(iy))
        for l in full:
            if l not in cur:
                y = np.broadcast_to(y[..., None], y.shape + (sd[l],))
                cur.append(l)
        y = np.transpose(y, [cur.index(l) for l in full])
        if full != iy:
            out = np.zeros([sd[l] for l in iy], dtype=y.dtype)
            for idx in itertools.product(*[range(sd[l]) for l in full]):
                m = dict(zip(full, idx))
                out[tuple(m[l] for l in iy)] = y[idx]
            y = out
    return np.array(y)


def _binary_ttgt(ia, a, ib, b, iy, sd):
    """permutedims -> reshape -> batched gemm -> reshape -> permutedims."""
    ia, ib, iy = list(ia), list(ib), list(iy)
    ok = len(set(ia)) == len(ia) and len(set(ib)) == len(ib) and len(set(iy)) == len(iy) \
        and all(l in ia or l in ib for l in iy)
    if not ok:
        labs = {l: k for k, l in enumerate(dict.fromkeys(ia + ib + iy))}
        return np.array(np.einsum(a, [labs[l] for l in ia], b, [labs[l] for l in ib], [labs[l] for l in iy]))
    # labels only in one operand and not in the output are summed first (OMEinsum simplifies them away)
    for l in [l for l in ia if l not in ib and l not in iy]:
        a = a.sum(axis=ia.index(l))
        ia.remove(l)
    for l in [l for l in ib if l not in ia and l not in iy]:
        b = b.sum(axis=ib.index(l))
        ib.remove(l)
    batch = [l for l in ia if l in ib and l in iy]
    contr = [l for l in ia if l in ib and l not in iy]
    outa = [l for l in ia if l not in ib]
    outb = [l for l in ib if l not in ia]
    prod = lambda ls: int(np.prod([sd[l] for l in ls], dtype=np.int64)) if ls else 1
    nb, nk, nm, nn = prod(batch), prod(contr), prod(outa), prod(outb)
    a3 = np.transpose(a, [ia.index(l) for l in batch + outa + contr]).reshape(nb, nm, nk)
    b3 = np.transpose(b, [ib.index(l) for l in batch + contr + outb]).reshape(nb, nk, nn)
    if FLOP_COUNTER["enabled"]:
        cplx = np.iscomplexobj(a3) or np.iscomplexobj(b3)
        FLOP_COUNTER["flops"] += (8 if cplx else 2) * nb * nm * nk * nn
        FLOP_COUNTER["bytes"] += 16 * (a.size + b.size + nb * nm * nn)
    c3 = np.matmul(np.ascontiguousarray(a3), np.ascontiguousarray(b3))
    cur = batch + outa + outb
    c = c3.reshape([sd[l] for l in cur])
    return np.array(np.transpose(c, [cur.index(l) for l in iy]), order='C')


def einsum_raw(ixs, xs, iy, size_dict=None):
    """``OMEinsum.einsum(EinCode(ixs, iy), xs, size_dict)`` on plain numpy arrays."""
    xs = [np.asarray(x) for x in xs]
    sd = _size_dict(ixs, xs, size_dict)
    if len(xs) == 1:
        return _unary(ixs[0], xs[0], iy, sd)
    if len(xs) == 2:
        return _binary_ttgt(ixs[0], xs[0], ixs[1], xs[1], iy, sd)
    labs = {l: k for k, l in enumerate(dict.fromkeys([l for ix in ixs for l in ix] + list(iy)))}
    args = []
    for ix, x in zip(ixs, xs):
        args += [x, [labs[l] for l in ix]]
    return np.array(np.einsum(*args, [labs[l] for l in iy], optimize=True))


def einsum(ixs, xs, iy, size_dict=None):
    """Dispatch point both in-tree array plug-ins hook (``src/TensorAD/rules.jl:1-8``,
    ``src/cracker.jl:4-8``): all-DiffTensor operands go to the AD rule."""
    if _DIFF_HOOK is not None and len(xs) > 0 and all(_DIFF_HOOK[0](x) for x in xs):
        return _DIFF_HOOK[1](ixs, xs, iy, size_dict)
    return einsum_raw(ixs, xs, iy, size_dict)


class EinCode:
    """``EinCode(ixs, iy)``; calling it contracts directly (no order optimisation)."""

    def __init__(self, ixs, iy):
        self.ixs = [list(ix) for ix in ixs]
        self.iy = list(iy)

    def __call__(self, *xs, size_dict=None):
        return einsum(self.ixs, list(xs), self.iy, size_dict)


class NestedEinsum:
    """Binary contraction tree.  A leaf has ``tensorindex >= 0`` (0-based)."""

    def __init__(self, args=None, eins=None, tensorindex=-1):
        self.args = args or []
        self.eins = eins
        self.tensorindex = tensorindex

    def isleaf(self):
        return self.tensorindex >= 0

    def __call__(self, *xs, size_dict=None):
        sd = size_dict
        return _run_nested(self, list(xs), sd)

    def leaves(self):
        if self.isleaf():
            return [self.tensorindex]
        return [i for a in self.args for i in a.leaves()]


def _run_nested(node, xs, sd):
    if node.isleaf():
        return xs[node.tensorindex]
    ys = [_run_nested(a, xs, sd) for a in node.args]
    return einsum(node.eins.ixs, ys, node.eins.iy, sd)


class SlicedEinsum:
    """``SlicedEinsum(slicing, nested)``: serial loop over the sliced labels' values."""

    def __init__(self, slicing, eins, ixs, iy):
        self.slicing = list(slicing)
        self.eins = eins
        self.ixs = [list(ix) for ix in ixs]
        self.iy = list(iy)

    def __call__(self, *xs, size_dict=None):
        xs = list(xs)
        sd = {}
        for ix, x in zip(self.ixs, xs):
            for l, s in zip(ix, x.shape):
                sd[l] = s
        if not self.slicing:
            return self.eins(*xs)
        out = None
        for vals in itertools.product(*[range(sd[l]) for l in self.slicing]):
            m = dict(zip(self.slicing, vals))
            sl = []
            for ix, x in zip(self.ixs, xs):
                idx = tuple(slice(m[l], m[l] + 1) if l in m else slice(None) for l in ix)
                sl.append(_getslice(x, idx))
            y = self.eins(*sl)
            if any(l in self.iy for l in self.slicing):
                raise NotImplementedError("sliced labels in the output are not needed on this path")
            out = y if out is None else out + y
        return out


def _getslice(x, idx):
    if _DIFF_HOOK is not None and _DIFF_HOOK[0](x):
        return _DIFF_HOOK[2](x, idx)
    return x[idx]


# ---------------------------------------------------------------------------------------
# contraction-order optimisation (GreedyMethod) and slicing
# ---------------------------------------------------------------------------------------
def optimize_greedy(ixs, iy, size_dict):
    """Greedy pairwise order, minimising ``size(out) - size(a) - size(b)``.  Deterministic."""
    ixs = [list(ix) for ix in ixs]
    iy = list(iy)
    n = len(ixs)
    if n == 1:
        return NestedEinsum([NestedEinsum(tensorindex=0)], EinCode([ixs[0]], iy))
    nodes = {i: NestedEinsum(tensorindex=i) for i in range(n)}
    labels = {i: list(ixs[i]) for i in range(n)}
    count = {}
    for ix in ixs:
        for l in set(ix):
            count[l] = count.get(l, 0) + 1
    for l in iy:
        count[l] = count.get(l, 0) + 1
    size = lambda ls: float(np.prod([size_dict[l] for l in ls])) if ls else 1.0
    nxt = n
    while len(nodes) > 1:
        best = None
        keys = sorted(nodes)
        lsets = {k: set(labels[k]) for k in keys}
        pairs = [(a, b) for ai, a in enumerate(keys) for b in keys[ai + 1:] if lsets[a] & lsets[b]]
        if not pairs:
            pairs = [(a, b) for ai, a in enumerate(keys) for b in keys[ai + 1:]]
        for a, b in pairs:
            la, lb = labels[a], labels[b]
            out = [l for l in dict.fromkeys(la + lb)
                   if count[l] - (l in lsets[a]) - (l in lsets[b]) > 0]
            cost = size(out) - size(la) - size(lb)
            if best is None or cost < best[0]:
                best = (cost, a, b, out)
        _, a, b, out = best
        if len(nodes) == 2:
            out = iy
        for l in set(labels[a]):
            count[l] -= 1
        for l in set(labels[b]):
            count[l] -= 1
        for l in set(out):
            count[l] += 1
        nodes[nxt] = NestedEinsum([nodes[a], nodes[b]], EinCode([labels[a], labels[b]], out))
        labels[nxt] = list(out)
        del nodes[a], nodes[b], labels[a], labels[b]
        nxt += 1
    return nodes[max(nodes)]


def optimize_sequence(ixs, iy, order):
    """Left-deep ("boundary sweep") tree absorbing tensors in ``order`` (0-based indices)."""
    ixs = [list(ix) for ix in ixs]
    count = {}
    for ix in ixs:
        for l in set(ix):
            count[l] = count.get(l, 0) + 1
    for l in iy:
        count[l] = count.get(l, 0) + 1
    cur = NestedEinsum(tensorindex=order[0])
    lab = list(ixs[order[0]])
    for step, t in enumerate(order[1:]):
        lb = ixs[t]
        for l in set(lab):
            count[l] -= 1
        for l in set(lb):
            count[l] -= 1
        out = [l for l in dict.fromkeys(lab + lb) if count[l] > 0]
        if step == len(order) - 2:
            out = list(iy)
        for l in set(out):
            count[l] += 1
        cur = NestedEinsum([cur, NestedEinsum(tensorindex=t)], EinCode([lab, lb], out))
        lab = out
    return cur


def slice_code(nested, ixs, iy, size_dict, nslices):
    """Pick ``nslices`` labels of the largest intermediate and strip them from the tree
    (the role ``TreeSA(nslices=k)`` plays at ``test/peps.jl:9``)."""
    if nslices <= 0:
        return SlicedEinsum([], nested, ixs, iy)
    inter = []

    def walk(nd):
        if nd.isleaf():
            return
        for a in nd.args:
            walk(a)
        inter.append(nd.eins.iy)
    walk(nested)
    inter.sort(key=lambda ls: -float(np.prod([size_dict[l] for l in ls])) if ls else 0)
    chosen = []
    for ls in inter:
        for l in ls:
            if l not in chosen and l not in iy and len(chosen) < nslices:
                chosen.append(l)
    if len(chosen) < nslices:
        for ix in ixs:
            for l in ix:
                if l not in chosen and l not in iy and len(chosen) < nslices:
                    chosen.append(l)

    def strip(nd):
        if nd.isleaf():
            return NestedEinsum(tensorindex=nd.tensorindex)
        return NestedEinsum([strip(a) for a in nd.args],
                            EinCode([[l for l in ix if l not in chosen] for ix in nd.eins.ixs],
                                    [l for l in nd.eins.iy if l not in chosen]))
    return SlicedEinsum(chosen, _SqueezeNested(strip(nested), ixs, chosen), ixs, iy)


class _SqueezeNested:
    """Runs a label-stripped tree on tensors whose sliced axes have length 1."""

    def __init__(self, nested, ixs, chosen):
        self.nested = nested
        self.ixs = ixs
        self.chosen = chosen

    def __call__(self, *xs):
        sq = []
        for ix, x in zip(self.ixs, xs):
            shape = tuple(s for l, s in zip(ix, x.shape) if l not in self.chosen)
            sq.append(_reshape(x, shape))
        return self.nested(*sq)


def _reshape(x, shape):
    if _DIFF_HOOK is not None and _DIFF_HOOK[0](x):
        return _DIFF_HOOK[3](x, shape)
    return np.reshape(x, shape, order="F")


def optimize_code(ixs, iy, size_dict, optimizer="greedy", nslices=0):
    """``optimize_code(EinCode(ixs, iy), size_dict, optimizer, simplifier)`` (``src/peps.jl:87,91``)."""
    if isinstance(optimizer, (list, tuple)):
        nested = optimize_sequence(ixs, iy, list(optimizer))
    else:
        nested = optimize_greedy(ixs, iy, size_dict)
    if nslices > 0:
        return slice_code(nested, ixs, iy, size_dict, nslices)
    return nested


def tree_cost(nested, size_dict):
    """(flops, bytes, largest intermediate elements) with the accounting of SURVEY.md 8(d):
    ``flops = sum 8*prod(size(labels(a) U labels(b)))``, ``bytes = sum 16*(|a|+|b|+|out|)``."""
    tot = [0.0, 0.0, 0.0]
    size = lambda ls: float(np.prod([size_dict[l] for l in ls])) if ls else 1.0

    def walk(nd):
        if nd.isleaf():
            return
        for a in nd.args:
            walk(a)
        ia, ib = nd.eins.ixs
        tot[0] += 8.0 * size(list(dict.fromkeys(ia + ib)))
        tot[1] += 16.0 * (size(ia) + size(ib) + size(nd.eins.iy))
        tot[2] = max(tot[2], size(nd.eins.iy))
    walk(nested)
    return tuple(tot)


def einsum_grad(ixs, xs, iy, size_dict, cdy, i, conj=np.conj, project=None):
    """OMEinsum ``einsum_grad``: ``conj(einsum(ixs[i->iy] -> ixs[i]; xs[i->cdy]))`` then ProjectTo."""
    nixs = [list(ix) for ix in ixs]
    nxs = list(xs)
    nixs[i] = list(iy)
    nxs[i] = cdy
    y = einsum(nixs, nxs, list(ixs[i]), size_dict)
    y = conj(y)
    return project(xs[i], y) if project is not None else y
