"""Host mirror of the OMEinsum objects the reference holds (``EinCode``, ``NestedEinsum``,
``SlicedEinsum``, ``optimize_code``; used at src/peps.jl:1,84-93,109,321) with the device
executor behind them.

* ``EinCode(ixs, iy)(xs...)``         -> one ``tne_einsum`` call (unary or binary);
* ``NestedEinsum`` / ``SlicedEinsum`` -> one ``tne_contract_tree`` call: the whole tree runs as a
  chain of contraction kernels on the device, intermediates never leave HBM;
* on ``DiffTensor`` operands the calls are recorded on the TensorAD tape
  (src/TensorAD/rules.jl:1-8): per binary einsum (second-order capable) or, when
  ``FUSED_TREES`` is on, as one instruction whose pullback is the fused device reverse pass.

``optimize_code`` offers a deterministic greedy order (``GreedyMethod`` semantics:
minimise ``size(out) - size(a) - size(b)``) and an explicit left-deep order (boundary sweep).
"""
import ctypes as C
import itertools

import numpy as np

from . import _lib
from .array import B200Array, einsum_device, _i32, _i64

FUSED_TREES = True          # one tape instruction per tree (first-order only)
_DIFF_HOOK = None           # installed by .tensorad


class EinCode:
    def __init__(self, ixs, iy):
        self.ixs = [list(ix) for ix in ixs]
        self.iy = list(iy)

    def __call__(self, *xs):
        return einsum(self.ixs, list(xs), self.iy)


def einsum(ixs, xs, iy, conj=None, size_dict=None):
    """The single dispatch point (``OMEinsum.einsum(code, xs, size_dict)``)."""
    if _DIFF_HOOK is not None and all(_DIFF_HOOK["is"](x) for x in xs):
        return _DIFF_HOOK["einsum"](ixs, xs, iy, size_dict)
    xs = [x if isinstance(x, B200Array) else B200Array.from_numpy(np.asarray(x)) for x in xs]
    return einsum_device(ixs, xs, iy, conj, size_dict)


class NestedEinsum:
    """Binary contraction tree; a leaf has ``tensorindex >= 0`` (0-based)."""

    def __init__(self, args=None, eins=None, tensorindex=-1):
        self.args = args or []
        self.eins = eins
        self.tensorindex = tensorindex
        self._ser = None
        # optional checkpointed segments (root only): [(first post-order node index, [sliced labels]), ...]
        self.segments = None
        self.seg_shard = False   # deal the segment-local slices to the ranks of the context's communicator

    def isleaf(self):
        return self.tensorindex >= 0

    def leaf_ixs(self):
        """labels of every leaf, by tensor index, recovered from the tree."""
        out = {}

        def walk(nd):
            if nd.isleaf():
                return
            for a, ix in zip(nd.args, nd.eins.ixs):
                if a.isleaf():
                    out[a.tensorindex] = list(ix)
                else:
                    walk(a)
        walk(self)
        return [out[i] for i in range(len(out))]

    def __call__(self, *xs):
        return run_tree(self, [], list(xs))


class SlicedEinsum:
    def __init__(self, slicing, eins):
        self.slicing = list(slicing)
        self.eins = eins

    def __call__(self, *xs):
        return run_tree(self.eins, self.slicing, list(xs))


def _unwrap(code):
    if isinstance(code, SlicedEinsum):
        return code.eins, code.slicing
    return code, []


# -- serialisation to the C ABI's tne_tree ----------------------------------------------
class SerializedTree:
    def __init__(self, nested, slicing, shard=(0, 1)):
        leaf_ixs = nested.leaf_ixs()
        n = len(leaf_ixs)
        children, out_rank, out_labels = [], [], []
        counter = [n]

        def walk(nd):
            if nd.isleaf():
                return nd.tensorindex
            ids = [walk(a) for a in nd.args]
            if len(ids) != 2:
                raise ValueError("only binary trees are supported")
            children.extend(ids)
            out_rank.append(len(nd.eins.iy))
            out_labels.extend(nd.eins.iy)
            me = counter[0]
            counter[0] += 1
            return me
        walk(nested)
        self.n_leaves = n
        self.leaf_ixs = leaf_ixs
        self._keep = dict(
            leaf_rank=_i32([len(ix) for ix in leaf_ixs]),
            leaf_labels=_i32([l for ix in leaf_ixs for l in ix]),
            node_children=_i32(children),
            node_out_rank=_i32(out_rank),
            node_out_labels=_i32(out_labels),
            sliced=_i32(slicing),
        )
        t = _lib.TneTree()
        t.n_leaves = n
        t.n_nodes = len(out_rank)
        t.leaf_rank = self._keep["leaf_rank"]
        t.leaf_labels = self._keep["leaf_labels"]
        t.node_children = self._keep["node_children"]
        t.node_out_rank = self._keep["node_out_rank"]
        t.node_out_labels = self._keep["node_out_labels"]
        t.n_sliced = len(slicing)
        t.sliced_labels = self._keep["sliced"]
        t.shard_rank, t.shard_world = shard
        t.seg_shard = 1 if nested.seg_shard else 0
        segs = nested.segments or []
        if len(segs) > 1:
            off = [0]
            for _, ls in segs:
                off.append(off[-1] + len(ls))
            self._keep["seg_first"] = _i32([f for f, _ in segs])
            self._keep["seg_off"] = _i32(off)
            self._keep["seg_labels"] = _i32([l for _, ls in segs for l in ls])
            t.n_segments = len(segs)
            t.seg_first_node = self._keep["seg_first"]
            t.seg_sliced_off = self._keep["seg_off"]
            t.seg_sliced_labels = self._keep["seg_labels"]
        else:
            t.n_segments = 0
        self.struct = t
        self.root_iy = list(nested.eins.iy)


def serialize(code, shard=(0, 1)):
    nested, slicing = _unwrap(code)
    key = (tuple(slicing), shard, bool(nested.seg_shard))
    if nested._ser is None or nested._ser[0] != key:
        nested._ser = (key, SerializedTree(nested, slicing, shard))
    return nested._ser[1]


def run_tree(nested, slicing, xs, conj=None):
    if _DIFF_HOOK is not None and all(_DIFF_HOOK["is"](x) for x in xs):
        return _DIFF_HOOK["tree"](nested, slicing, xs)
    xs = [x if isinstance(x, B200Array) else B200Array.from_numpy(np.asarray(x)) for x in xs]
    return contract_tree_device(SlicedEinsum(slicing, nested) if slicing else nested, xs, conj)


def contract_tree_device(code, xs, conj=None, shard=(0, 1)):
    """``tne_contract_tree``: the whole NestedEinsum/SlicedEinsum on the device."""
    ser = serialize(code, shard)
    ctx = xs[0].ctx
    h = C.c_int64()
    cj = _i32(conj if conj is not None else [0] * len(xs))
    ctx.check(ctx.L.tne_contract_tree(ctx.h, C.byref(ser.struct), _i64([x.handle for x in xs]), cj, C.byref(h)))
    sizes = {}
    for ix, x in zip(ser.leaf_ixs, xs):
        for l, s in zip(ix, x.shape):
            sizes[l] = s
    return B200Array(h.value, tuple(sizes[l] for l in ser.root_iy), all(x.is_real for x in xs), ctx)


def expect_terms_device(code, xs, conj, terms, grad_leaves=(), seed=1.0 + 0j, shard=(0, 1)):
    """``tne_expect_terms``: value (0-dim) and pullbacks w.r.t. ``grad_leaves``.

    ``terms``: list of ``(coef, [(leaf_index, B200Array replacement), ...])``."""
    ser = serialize(code, shard)
    ctx = xs[0].ctx
    arr = (_lib.TneTerm * max(1, len(terms)))()
    keep = []
    for k, (coef, repl) in enumerate(terms):
        coef = complex(coef)
        arr[k].n_repl = len(repl)
        for r, (leaf, t) in enumerate(repl):
            arr[k].leaf[r] = int(leaf)
            arr[k].repl[r] = t.handle
            keep.append(t)
        arr[k].coef_re, arr[k].coef_im = coef.real, coef.imag
    hv = C.c_int64()
    hg = (C.c_int64 * max(1, len(grad_leaves)))()
    cj = _i32(conj if conj is not None else [0] * len(xs))
    seed = complex(seed)
    ctx.check(ctx.L.tne_expect_terms(ctx.h, C.byref(ser.struct), _i64([x.handle for x in xs]), cj, len(terms), arr,
                                     len(grad_leaves), _i32(list(grad_leaves)), seed.real, seed.imag,
                                     C.byref(hv), hg))
    sizes = {}
    for ix, x in zip(ser.leaf_ixs, xs):
        for l, s in zip(ix, x.shape):
            sizes[l] = s
    val = B200Array(hv.value, tuple(sizes[l] for l in ser.root_iy), False, ctx)
    grads = [B200Array(hg[i], xs[g].shape, False, ctx) for i, g in enumerate(grad_leaves)]
    return val, grads


# -- contraction order ------------------------------------------------------------------
def optimize_greedy(ixs, iy, size_dict):
    ixs = [list(ix) for ix in ixs]
    iy = list(iy)
    n = len(ixs)
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
        keys = sorted(nodes)
        lsets = {k: set(labels[k]) for k in keys}
        pairs = [(a, b) for ai, a in enumerate(keys) for b in keys[ai + 1:] if lsets[a] & lsets[b]]
        if not pairs:
            pairs = [(a, b) for ai, a in enumerate(keys) for b in keys[ai + 1:]]
        best = None
        for a, b in pairs:
            la, lb = labels[a], labels[b]
            out = [l for l in dict.fromkeys(la + lb) if count[l] - (l in lsets[a]) - (l in lsets[b]) > 0]
            cost = size(out) - size(la) - size(lb)
            if best is None or cost < best[0]:
                best = (cost, a, b, out)
        _, a, b, out = best
        if len(nodes) == 2:
            out = iy
        for l in lsets[a]:
            count[l] -= 1
        for l in lsets[b]:
            count[l] -= 1
        for l in set(out):
            count[l] += 1
        nodes[nxt] = NestedEinsum([nodes[a], nodes[b]], EinCode([labels[a], labels[b]], out))
        labels[nxt] = list(out)
        del nodes[a], nodes[b], labels[a], labels[b]
        nxt += 1
    return nodes[max(nodes)]


def optimize_sequence(ixs, iy, order):
    """Left-deep tree absorbing the items of ``order`` one after the other (boundary sweep).  An item is a
    leaf index or a tuple ``(i, j)``: the two leaves are contracted with each other first and the pair is
    absorbed in one step (worth it where it costs no extra flops, e.g. a PEPS site without down legs whose
    left legs are sliced: the boundary is then read and written once for the site instead of twice).
    The root gets ``first_node_of_item[k]``: the post-order index of the first node created for item k."""
    ixs = [list(ix) for ix in ixs]
    count = {}
    for ix in ixs:
        for l in set(ix):
            count[l] = count.get(l, 0) + 1
    for l in iy:
        count[l] = count.get(l, 0) + 1

    def item_node(t):
        if isinstance(t, (tuple, list)):
            i, j = t
            la, lb = ixs[i], ixs[j]
            for l in set(la):
                count[l] -= 1
            for l in set(lb):
                count[l] -= 1
            out = [l for l in dict.fromkeys(la + lb) if count[l] > 0]
            for l in set(out):
                count[l] += 1
            return NestedEinsum([NestedEinsum(tensorindex=i), NestedEinsum(tensorindex=j)], EinCode([la, lb], out)), out, 1
        return NestedEinsum(tensorindex=t), list(ixs[t]), 0
    cur, lab, n_nodes = item_node(order[0])
    first_node = [0]
    for step, t in enumerate(order[1:]):
        first_node.append(n_nodes)
        nd, lb, extra = item_node(t)
        n_nodes += extra
        for l in set(lab):
            count[l] -= 1
        for l in set(lb):
            count[l] -= 1
        out = [l for l in dict.fromkeys(lab + lb) if count[l] > 0]
        if step == len(order) - 2:
            out = list(iy)
        for l in set(out):
            count[l] += 1
        cur = NestedEinsum([cur, nd], EinCode([lab, lb], out))
        n_nodes += 1
        lab = out
    cur.first_node_of_item = first_node
    return cur


def sweep_order(ixs, iy, size_dict, groups=None):
    """Leaf order of a boundary sweep for ANY network.  ``groups``: lists of leaves that are absorbed one right after the other
    (the bra and the ket tensor of a PEPS site; default: every leaf on its own).  Starting from the group with the fewest
    neighbours, the sweep keeps absorbing the group that leaves the smallest boundary (product of the open label sizes), ties
    broken by the number of already absorbed neighbours and then by index.  On the double-layer network of an ``L x L`` grid
    PEPS this reproduces the hand-built column sweep ``bra_1, ket_1, bra_2, ...`` exactly (2.38e12 flop at 6x6 D=4, where the
    greedy pairwise optimiser of ``optimize_greedy`` -- the reference's default -- gives 2.4e15); it is the order
    ``optimize_code(..., optimizer="sweep")`` hands to :func:`optimize_sequence`."""
    n = len(ixs)
    groups = [list(g) for g in groups] if groups is not None else [[i] for i in range(n)]
    ng = len(groups)
    gsets = [set(l for i in g for l in ixs[i]) for g in groups]
    count = {}
    for ix in ixs:
        for l in set(ix):
            count[l] = count.get(l, 0) + 1
    for l in iy:
        count[l] = count.get(l, 0) + 1
    gcount = [dict() for _ in range(ng)]          # how many leaves of the group carry the label
    for k, g in enumerate(groups):
        for i in g:
            for l in set(ixs[i]):
                gcount[k][l] = gcount[k].get(l, 0) + 1
    where = {}
    for k, st in enumerate(gsets):
        for l in st:
            where.setdefault(l, []).append(k)
    nbrs = [set(j for l in gsets[k] for j in where[l] if j != k) for k in range(ng)]

    def size(labels):
        out = 1
        for l in labels:
            out *= size_dict.get(l, 2)
        return out
    start = min(range(ng), key=lambda k: (len(nbrs[k]), size(gsets[k]), k))
    order, done = [], set()
    seen = {}

    def absorb(k):
        order.append(k)
        done.add(k)
        for l, c in gcount[k].items():
            seen[l] = seen.get(l, 0) + c
    absorb(start)
    while len(order) < ng:
        cand = {j for k in done for j in nbrs[k] if j not in done} or {j for j in range(ng) if j not in done}
        best, best_key = None, None
        for j in sorted(cand):
            after = [l for l in set(seen) | gsets[j] if seen.get(l, 0) + gcount[j].get(l, 0) < count[l]]
            key = (size(after), -len(nbrs[j] & done), j)
            if best_key is None or key < best_key:
                best, best_key = j, key
        absorb(best)
    return [i for k in order for i in groups[k]]


def sweep_segments(ixs, iy, size_dict, order, groups, slice_entering=True, slice_leaving=False):
    """Checkpointed segments with segment-local slices for a sweep ``order`` (see :func:`sweep_order`), derived from the
    network alone -- the general form of ``peps.sweep_plan``.  A segment ends after every group at which the boundary
    (product of the open label sizes) has a local minimum: on a lattice, the end of every column.  A segment loops over
    * ``slice_entering``: the labels that are open when it starts and are closed by its LAST group -- they are open during
      the whole segment, so looping over them repeats nothing and divides the segment's tape by their size;
    * ``slice_leaving``: the labels its FIRST group creates that are still open when it ends (only that first, small
      absorption is repeated; the next boundary is then written slice by slice).
    Returns ``[(position in order of the segment's first leaf, [labels]), ...]`` as :func:`optimize_code` takes them."""
    count = {}
    for ix in ixs:
        for l in set(ix):
            count[l] = count.get(l, 0) + 1
    for l in iy:
        count[l] = count.get(l, 0) + 1
    pos_of_leaf = {leaf: k for k, leaf in enumerate(order)}
    gorder = sorted(range(len(groups)), key=lambda k: min(pos_of_leaf[i] for i in groups[k]))
    seen, bsize, open_after = {}, [], []
    for k in gorder:
        for i in groups[k]:
            for l in set(ixs[i]):
                seen[l] = seen.get(l, 0) + 1
        op = {l for l, c in seen.items() if c < count[l]}
        open_after.append(op)
        sz = 1
        for l in op:
            sz *= size_dict.get(l, 2)
        bsize.append(sz)
    ends = [q for q in range(1, len(gorder) - 1)
            if (bsize[q] <= bsize[q - 1] and bsize[q] < bsize[q + 1]) or (bsize[q] < bsize[q - 1] and bsize[q] <= bsize[q + 1])]
    starts = [0] + [q + 1 for q in ends]
    segs = []
    for si, q0 in enumerate(starts):
        q1 = (starts[si + 1] - 1) if si + 1 < len(starts) else len(gorder) - 1
        before = open_after[q0 - 1] if q0 > 0 else set()
        labels = []
        if slice_entering and q1 > q0:
            prev = open_after[q1 - 1]
            labels += sorted(l for l in before if l in prev and l not in open_after[q1])
        if slice_leaving and si + 1 < len(starts):
            first = set(l for i in groups[gorder[q0]] for l in ixs[i])
            labels += sorted(l for l in first if l not in before and l in open_after[q1])
        segs.append((min(pos_of_leaf[i] for i in groups[gorder[q0]]), labels))
    return segs


def _pair_where_free(ixs, iy, size_dict, order, groups, segs):
    """Items for :func:`optimize_sequence`: the two leaves of the LAST group of a sliced segment are contracted with each other
    first and absorbed in one step where that costs no more flops than absorbing them one after the other (a PEPS site whose
    entering bond is sliced and that has no legs further down: the boundary is read and written once instead of twice).
    Returns (items, segments with positions in item units)."""
    count = {}
    for ix in ixs:
        for l in set(ix):
            count[l] = count.get(l, 0) + 1
    for l in iy:
        count[l] = count.get(l, 0) + 1
    group_of = {i: k for k, g in enumerate(groups) for i in g}
    seg_start = [pos for pos, _ in segs]
    sliced_at = {}
    for si, (pos, labels) in enumerate(segs):
        end = (seg_start[si + 1] if si + 1 < len(segs) else len(order)) - 1       # last leaf of the segment
        sliced_at[group_of[order[end]]] = set(labels)

    def vol(labels):
        v = 1.0
        for l in labels:
            v *= size_dict.get(l, 2)
        return v
    items, item_of_pos, seen, k = [], {}, {}, 0
    while k < len(order):
        i = order[k]
        g = groups[group_of[i]]
        item_of_pos[k] = len(items)
        if len(g) == 2 and k + 1 < len(order) and order[k + 1] == g[1] and group_of[i] in sliced_at and sliced_at[group_of[i]]:
            sl = sliced_at[group_of[i]]
            a, b = set(ixs[g[0]]) - sl, set(ixs[g[1]]) - sl
            bnd = {l for l, c in seen.items() if c < count[l]} - sl
            # one after the other: (boundary . a) then (. b);  paired: (a . b) then (boundary . pair); cost = volume of all labels involved
            t = {l for l in bnd | a if seen.get(l, 0) + (1 if l in a else 0) < count[l]}
            seq = vol(bnd | a) + vol(t | b)
            pr = {l for l in a | b if (1 if l in a else 0) + (1 if l in b else 0) < count[l] - seen.get(l, 0) or l in bnd}
            paired = vol(a | b) + vol(bnd | pr)
            if paired <= 1.01 * seq:      # forming the pair itself is a few hundred flops
                items.append((g[0], g[1]))
                item_of_pos[k + 1] = len(items) - 1
                for j in g:
                    for l in set(ixs[j]):
                        seen[l] = seen.get(l, 0) + 1
                k += 2
                continue
        items.append(i)
        for l in set(ixs[i]):
            seen[l] = seen.get(l, 0) + 1
        k += 1
    return items, [(item_of_pos[pos], ls) for pos, ls in segs]


def optimize_treesa(ixs, iy, size_dict, sc_target=None, betas=None, ntrials=1, niters=20, sc_weight=3.0, seed=0):
    """Simulated annealing over binary contraction trees (the role of OMEinsumContractionOrders' ``TreeSA``, which the
    reference's tests use: ``TreeSA(ntrials=1, nslices=...)``, test/peps.jl:9).  Starts from the greedy tree and applies
    the local rewrites of Kalachev et al. -- ``((a b) c) -> ((a c) b)`` or ``(a (b c))`` and their mirror images -- accepted
    with the Metropolis rule on ``log2(flops) + sc_weight * max(0, log2(largest intermediate) - sc_target)``.  Deterministic
    for a given ``seed`` (the reference's optimiser is randomised).  Host-side only; the result is an ordinary NestedEinsum."""
    rng = np.random.default_rng(seed)
    ixs = [list(ix) for ix in ixs]
    iy = list(iy)
    n = len(ixs)
    lsz = {l: float(np.log2(size_dict[l])) for ix in ixs for l in ix}
    for l in iy:
        lsz.setdefault(l, float(np.log2(size_dict[l])))
    ext = {}                      # label -> number of tensors (plus the output) it sits on
    for ix in ixs:
        for l in set(ix):
            ext[l] = ext.get(l, 0) + 1
    for l in iy:
        ext[l] = ext.get(l, 0) + 1

    # a tree is a nested tuple of leaf indices; labels of a subtree = labels that still have an owner outside it
    def from_nested(nd):
        return nd.tensorindex if nd.isleaf() else (from_nested(nd.args[0]), from_nested(nd.args[1]))

    def counts(t, acc):
        if isinstance(t, tuple):
            counts(t[0], acc); counts(t[1], acc)
        else:
            for l in set(ixs[t]):
                acc[l] = acc.get(l, 0) + 1
        return acc

    def out_labels(t):
        inside = counts(t, {})
        return [l for l, c in inside.items() if ext[l] - c > 0]

    def cost(t):
        """(log2 sum of flops, log2 of the largest intermediate) of the whole tree."""
        tcs, sc = [], [0.0]

        def walk(u):
            if not isinstance(u, tuple):
                return set(ixs[u]), counts(u, {})
            la, ca = walk(u[0])
            lb, cb = walk(u[1])
            c = dict(ca)
            for l, k in cb.items():
                c[l] = c.get(l, 0) + k
            out = {l for l in (la | lb) if ext[l] - c[l] > 0}
            tcs.append(sum(lsz[l] for l in (la | lb)))
            sc[0] = max(sc[0], sum(lsz[l] for l in out))
            return out, c
        walk(t)
        m = max(tcs)
        return m + float(np.log2(sum(2.0 ** (x - m) for x in tcs))), sc[0]

    def score(t):
        tc, sc = cost(t)
        return tc + (sc_weight * max(0.0, sc - sc_target) if sc_target is not None else 0.0)

    def subtrees(t, path=()):
        if isinstance(t, tuple):
            yield path
            yield from subtrees(t[0], path + (0,))
            yield from subtrees(t[1], path + (1,))

    def get(t, path):
        for k in path:
            t = t[k]
        return t

    def put(t, path, new):
        if not path:
            return new
        k = path[0]
        return (put(t[0], path[1:], new), t[1]) if k == 0 else (t[0], put(t[1], path[1:], new))

    def rewrites(u):
        a, b = u
        out = []
        if isinstance(a, tuple):
            out += [((a[0], b), a[1]), (a[0], (a[1], b))]
        if isinstance(b, tuple):
            out += [((a, b[0]), b[1]), (b[0], (a, b[1]))]
        return out

    start = from_nested(optimize_greedy(ixs, iy, size_dict))
    if betas is None:
        betas = np.arange(0.5, 12.0, 0.5)
    best = (score(start), start)
    nit = niters if niters else 1
    for _ in range(max(1, ntrials)):
        cur, cur_s = start, score(start)
        for beta in betas:
            for _ in range(max(nit, 2 * n)):
                paths = list(subtrees(cur))
                path = paths[int(rng.integers(len(paths)))]
                opts = rewrites(get(cur, path))
                if not opts:
                    continue
                cand = put(cur, path, opts[int(rng.integers(len(opts)))])
                cs = score(cand)
                if cs <= cur_s or rng.random() < np.exp(-beta * (cs - cur_s)):
                    cur, cur_s = cand, cs
                    if cur_s < best[0]:
                        best = (cur_s, cur)     # the best tree visited, not the last one

    def to_nested(t, root=False):
        if not isinstance(t, tuple):
            return NestedEinsum(tensorindex=t), list(ixs[t])
        na, la = to_nested(t[0])
        nb, lb = to_nested(t[1])
        out = list(iy) if root else [l for l in dict.fromkeys(la + lb) if l in set(out_labels(t))]
        return NestedEinsum([na, nb], EinCode([la, lb], out)), out
    return to_nested(best[1], root=True)[0]


def pick_slices(nested, iy, size_dict, nslices):
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
    return chosen


def optimize_code(ixs, iy, size_dict, optimizer="greedy", nslices=0, slicing=None, segments=None, groups=None):
    """``optimize_code(EinCode(ixs, iy), size_dict, optimizer, simplifier)`` (src/peps.jl:87,91).

    ``optimizer``: ``"greedy"``, ``"treesa"`` (simulated annealing from the greedy tree), ``"sweep"`` (boundary sweep of any
    network, :func:`sweep_order`: reproduces the hand-built column sweep on lattices), ``"auto"`` (the cheaper of greedy and sweep)
    or an explicit leaf order (list).  ``nslices`` / ``slicing``
    produce a ``SlicedEinsum`` (the role of ``TreeSA(nslices=k)``, test/peps.jl:9).
    ``segments`` (explicit leaf order only): ``[(position in the order, [labels]), ...]`` -- a
    checkpointed segment starts at the node that absorbs ``order[position]`` and loops over the
    given labels (see ``tne_tree`` in include/tne.h); the first entry must have position 0 or 1."""
    if len(ixs) == 1:
        raise ValueError("a single-tensor network needs no order")
    if isinstance(optimizer, (list, tuple)):
        nested = optimize_sequence(ixs, iy, list(optimizer))
        if segments:
            # a segment starts at the first node created for order[pos] (post-order numbering)
            nested.segments = [(nested.first_node_of_item[pos], list(ls)) for pos, ls in segments]
    elif optimizer == "sweep":
        order = sweep_order(ixs, iy, size_dict, groups)
        if segments in ("auto", "auto2"):      # memory plan from the network alone ("auto2": also the bond leaving a segment's first group)
            gs = groups if groups is not None else [[i] for i in range(len(ixs))]
            segs = sweep_segments(ixs, iy, size_dict, order, gs, slice_leaving=segments == "auto2")
            items, seg_items = _pair_where_free(ixs, iy, size_dict, order, gs, segs)
            nested = optimize_sequence(ixs, iy, items)
            if len(seg_items) > 1:
                nested.segments = [(nested.first_node_of_item[pos], list(ls)) for pos, ls in seg_items]
        else:
            nested = optimize_sequence(ixs, iy, order)
    elif optimizer == "auto":     # the cheaper of the pairwise greedy tree and the boundary sweep (lattices: the sweep, by orders of magnitude)
        cands = [optimize_greedy(ixs, iy, size_dict), optimize_sequence(ixs, iy, sweep_order(ixs, iy, size_dict, groups))]
        nested = min(cands, key=lambda c: tree_cost(c, size_dict)[0])
    elif optimizer == "treesa":
        nested = optimize_treesa(ixs, iy, size_dict)
    else:
        nested = optimize_greedy(ixs, iy, size_dict)
    if slicing:
        return SlicedEinsum(list(slicing), nested)
    if nslices > 0:
        return SlicedEinsum(pick_slices(nested, iy, size_dict, nslices), nested)
    return nested


def tree_cost(code, size_dict):
    """(flops, bytes, largest intermediate) with SURVEY.md 8(d)'s accounting."""
    nested, _ = _unwrap(code)
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
