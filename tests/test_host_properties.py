"""Property-based CPU tests (hypothesis) of the host logic around the C ABI: contraction-order optimisers produce valid
binary trees that evaluate to the plain Einstein sum; the TEBD wavefront levelling is site-disjoint and order preserving
(the condition under which batching equals the reference's sequential loop, src/tebd.jl:16-21); the state file layout
round-trips ragged tensors bit-exactly."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import einsum as OE


def _network(draw):
    """A random closed-or-open tensor network: every label sits on at most two tensors (PEPS-like), dims 1..3."""
    n = draw(st.integers(2, 5))
    nl = draw(st.integers(n - 1, 8))
    ixs = [[] for _ in range(n)]
    sizes = {}
    for l in range(1, nl + 1):
        sizes[l] = draw(st.integers(1, 3))
        owners = draw(st.lists(st.integers(0, n - 1), min_size=1, max_size=2, unique=True))
        for o in owners:
            ixs[o].append(l)
    # make it connected enough: chain labels between neighbours
    for i in range(n - 1):
        l = nl + 1 + i
        sizes[l] = draw(st.integers(1, 3))
        ixs[i].append(l); ixs[i + 1].append(l)
    count = {}
    for ix in ixs:
        for l in ix:
            count[l] = count.get(l, 0) + 1
    open_labels = [l for l, c in count.items() if c == 1]
    iy = draw(st.permutations(open_labels)) if open_labels else []
    return ixs, list(iy), sizes


network = st.composite(_network)()


def _evaluate(code, xs):
    if code.isleaf():
        return xs[code.tensorindex]
    a, b = [_evaluate(c, xs) for c in code.args]
    return OE.einsum_raw(code.eins.ixs, [a, b], code.eins.iy)


def _reference(ixs, iy, xs):
    labels = sorted({l for ix in ixs for l in ix})
    sym = {l: chr(ord("a") + k) for k, l in enumerate(labels)}
    expr = ",".join("".join(sym[l] for l in ix) for ix in ixs) + "->" + "".join(sym[l] for l in iy)
    return np.einsum(expr, *xs)


@settings(max_examples=60, deadline=None)
@given(network, st.integers(0, 2 ** 31 - 1))
def test_optimisers_produce_valid_trees(net, seed):
    from tne_b200 import einsum as es
    ixs, iy, sizes = net
    rng = np.random.default_rng(seed)
    xs = [rng.standard_normal([sizes[l] for l in ix]) + 1j * rng.standard_normal([sizes[l] for l in ix]) for ix in ixs]
    want = _reference(ixs, iy, xs)
    costs = {}
    for opt in ("greedy", "treesa", list(rng.permutation(len(ixs)))):
        code = es.optimize_code(ixs, iy, sizes, opt)
        costs[str(opt)] = es.tree_cost(code, sizes)[0]
        leaves = []

        def walk(nd):
            if nd.isleaf():
                leaves.append(nd.tensorindex)
            else:
                assert len(nd.args) == 2
                for a in nd.args:
                    walk(a)
        walk(code)
        assert sorted(leaves) == list(range(len(ixs)))          # every tensor exactly once, binary nodes only
        got = _evaluate(code, xs)
        assert got.shape == want.shape and np.allclose(got, want, rtol=1e-10, atol=1e-10)
    assert costs["treesa"] <= costs["greedy"] * (1 + 1e-12)     # annealing starts from the greedy tree and keeps the best one


@settings(max_examples=100, deadline=None)
@given(st.lists(st.tuples(st.integers(1, 12), st.integers(1, 12)).filter(lambda e: e[0] != e[1]), min_size=0, max_size=40))
def test_wavefronts_are_disjoint_and_order_preserving(edges):
    from tne_b200 import tebd
    lv = tebd.wavefronts(edges)
    flat = [e for b in lv for e in b]
    assert sorted(flat) == sorted(edges)
    level = {}
    for k, b in enumerate(lv):
        sites = [s for e in b for s in e]
        assert len(set(sites)) == len(sites)                     # gates of one launch act on disjoint sites
    # two gates sharing a site keep their sequential order: the later one sits in a strictly later wavefront
    pos = []
    remaining = {k: list(b) for k, b in enumerate(lv)}
    for e in edges:
        k = next(k for k, b in remaining.items() if e in b)
        remaining[k].remove(e)
        pos.append(k)
    for a in range(len(edges)):
        for b in range(a + 1, len(edges)):
            if set(edges[a]) & set(edges[b]):
                assert pos[a] < pos[b]


@settings(max_examples=60, deadline=None)
@given(st.lists(st.lists(st.integers(1, 3), min_size=1, max_size=4), min_size=1, max_size=6), st.integers(0, 2 ** 31 - 1), st.booleans())
def test_state_file_round_trip(shapes, seed, vidal):
    import tne_b200
    sio = tne_b200.stateio
    rng = np.random.default_rng(seed)
    # labels: vertex k has physical label k+1 first, then fresh virtual labels (the format does not need a consistent graph)
    nvx = len(shapes)
    nxt = nvx + 1
    vlabels, virt = [], []
    for k, shp in enumerate(shapes):
        labs = [k + 1]
        for _ in shp[1:]:
            labs.append(nxt); virt.append(nxt); nxt += 1
        vlabels.append(labs)
    tensors = [rng.standard_normal(shp) + 1j * rng.standard_normal(shp) for shp in shapes]
    if vidal:
        tensors += [rng.standard_normal(2) + 0j for _ in virt]
    d = sio.pack_state("vidal" if vidal else "simple", vlabels, virt, tensors, 3, 1e-12)
    meta, back = sio.unpack_state(d)
    assert meta["vertex_labels"] == vlabels and meta["virtual_labels"] == virt
    assert len(back) == len(tensors) and all(a.shape == b.shape and np.array_equal(a, b) for a, b in zip(back, tensors))
    assert np.array_equal(d["variables"], np.concatenate([t.reshape(-1, order="F") for t in tensors]))
