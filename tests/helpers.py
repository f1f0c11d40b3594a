"""Shared builders: identical synthetic inputs for the oracle and the device path."""
import numpy as np

from oracle import graphs as OG, peps as OP, yao as OY, tebd as OT


def seed_for(L, D):
    return 20260000 + 100 * L + D


def oracle_peps(g, D, seed, kind="simple", **kw):
    rng = np.random.default_rng(seed)
    f = OP.rand_vidalpeps if kind == "vidal" else OP.rand_simplepeps
    return f(g, D, rng, **kw)


def device_peps_like(tne, op, g, D, **kw):
    """A device PEPS with the same structure and the same tensor values as oracle PEPS ``op``."""
    gg = tne.graph_from_edges(g.nv(), g.edges())
    f = tne.zero_vidalpeps if op.kind == "vidal" else tne.zero_simplepeps
    dp = f(gg, D, **kw)
    return tne.peps.from_numpy_tensors(dp, op.alltensors())


def sweep_order(n):
    """bra_1, ket_1, bra_2, ket_2, ... : the boundary-sweep leaf order for code_inner_product."""
    out = []
    for i in range(n):
        out += [i, n + i]
    return out


def rel(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(1e-300, np.max(np.abs(b))))


def to_device_hamiltonian(tne, h):
    """Re-create an oracle Hamiltonian (oracle.yao blocks) with the product's block classes."""
    Y = tne.yao

    def conv(b):
        if isinstance(b, OY.Add):
            return Y.Add(*[conv(x) for x in b.blocks])
        if isinstance(b, OY.Scale):
            return Y.Scale(conv(b.content), b.factor)
        if isinstance(b, OY.Put):
            return Y.Put(b.n, b.locs, b.content)
        if isinstance(b, OY.Kron):
            return Y.Kron(b.n, list(zip(b.locs, b.subblocks)))
        if isinstance(b, OY.Control):
            return Y.Control(b.n, b.ctrl, b.loc, b.content)
        raise TypeError(type(b))
    return conv(h)


def dense_apply(block, psi):
    """``mat(block) * psi`` without forming the 2^n x 2^n matrix (works up to ~20 sites)."""
    if isinstance(block, OY.Add):
        return sum(dense_apply(b, psi) for b in block.blocks)
    if isinstance(block, OY.Scale):
        return block.factor * dense_apply(block.content, psi)
    reg = OY.ArrayReg(psi)
    if isinstance(block, OY.Put):
        reg._apply_locs(block.locs, block.content)
    elif isinstance(block, OY.Kron):
        for l, m in zip(block.locs, block.subblocks):
            reg._apply_locs((l,), m)
    elif isinstance(block, OY.Control):
        reg._apply_locs((block.ctrl, block.loc), block.two_site_matrix())
    else:
        raise TypeError(type(block))
    return reg.state
