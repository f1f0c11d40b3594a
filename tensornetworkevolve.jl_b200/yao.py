"""Host-side mirror of the slice of Yao 0.8 the hot path touches (src/peps.jl:432-477,
src/tebd.jl:3-20): gate constants, ``put`` / ``kron`` / ``control`` / ``Scale`` / ``Add`` blocks,
``time_evolve`` and a dense ``ArrayReg`` used by tests.  Gate matrices are at most 4x4 and stay
on the host; they are passed to the device by value.

Conventions (pinned by test/peps.jl:16-19): little-endian, site 1 = least-significant bit; a
block on ``locs = (l1, l2)`` has matrix ``M[o1 + 2 o2, i1 + 2 i2]``; ``time_evolve(h, dt)`` is
``exp(-i dt H)``.  Sites are 1-based like the reference.
"""
import numpy as np
import scipy.linalg

X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
I2 = np.eye(2, dtype=np.complex128)
P0 = np.array([[1, 0], [0, 0]], dtype=np.complex128)
P1 = np.array([[0, 0], [0, 1]], dtype=np.complex128)
CNOT = np.array([[1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 1, 0, 0]], dtype=np.complex128)  # ctrl = bit 1


def kron_mat(*ms):
    """Yao ``kron(A, B)``: A acts on the first (least-significant) site."""
    out = np.eye(1, dtype=np.complex128)
    for m in ms:
        out = np.kron(m, out)
    return out


class Block:
    n = 0

    def __add__(self, o):
        return Add(*(_terms(self) + _terms(o)))

    def __rmul__(self, c):
        return Scale(self, c)

    def __mul__(self, c):
        return Scale(self, c)


def _terms(b):
    return list(b.blocks) if isinstance(b, Add) else [b]


class Put(Block):
    """``put(n, locs => content)`` with a 1- or 2-site matrix content."""

    def __init__(self, n, locs, content):
        self.n = n
        self.locs = (locs,) if isinstance(locs, (int, np.integer)) else tuple(locs)
        self.content = np.asarray(content, dtype=np.complex128)
        assert self.content.shape == (2 ** len(self.locs),) * 2


class Kron(Block):
    """``kron(n, l1 => A, l2 => B, ...)`` of single-site matrices."""

    def __init__(self, n, pairs):
        self.n = n
        self.locs = tuple(l for l, _ in pairs)
        self.subblocks = [np.asarray(m, dtype=np.complex128) for _, m in pairs]


class Control(Block):
    """``control(n, ctrl, loc => content)`` single control, single target."""

    def __init__(self, n, ctrl, loc, content):
        self.n = n
        self.ctrl = ctrl
        self.loc = loc
        self.content = np.asarray(content, dtype=np.complex128)

    def two_site_matrix(self):
        # control(2, 1, 2 => content): first loc (ctrl) is bit 0 of the 4x4 index
        return kron_mat(P0, I2) + kron_mat(P1, self.content)


class Scale(Block):
    def __init__(self, content, factor):
        self.content = content
        self.factor = factor
        self.n = content.n


class Add(Block):
    def __init__(self, *blocks):
        self.blocks = list(blocks)
        self.n = blocks[0].n if blocks else 0


def put(n, locs, content):
    return Put(n, locs, content)


def kron(n, *pairs):
    return Kron(n, list(pairs))


def control(n, ctrl, loc, content):
    return Control(n, ctrl, loc, content)


def time_evolve(h, dt):
    return scipy.linalg.expm(-1j * dt * np.asarray(h, dtype=np.complex128))


def _embed(n, locs, m):
    """Dense 2^n matrix of a block acting on ``locs`` (1-based, little-endian)."""
    k = len(locs)
    m = np.asarray(m, dtype=np.complex128).reshape((2,) * (2 * k), order="F")  # o1..ok, i1..ik
    full = np.zeros((2 ** n, 2 ** n), dtype=np.complex128)
    rest = [q for q in range(1, n + 1) if q not in locs]
    for col in range(2 ** n):
        bits = [(col >> (q - 1)) & 1 for q in range(1, n + 1)]
        iin = tuple(bits[l - 1] for l in locs)
        for oo in range(2 ** k):
            ob = tuple((oo >> t) & 1 for t in range(k))
            v = m[ob + iin]
            if v != 0:
                nb = list(bits)
                for t, l in enumerate(locs):
                    nb[l - 1] = ob[t]
                row = sum(b << q for q, b in enumerate(nb))
                full[row, col] += v
    return full


def mat(block):
    """``mat(block)``: the dense matrix (small n only; test oracle bridge)."""
    n = block.n
    if isinstance(block, Put):
        return _embed(n, block.locs, block.content)
    if isinstance(block, Kron):
        out = np.eye(2 ** n, dtype=np.complex128)
        for l, m in zip(block.locs, block.subblocks):
            out = _embed(n, (l,), m) @ out
        return out
    if isinstance(block, Control):
        return _embed(n, (block.ctrl, block.loc), block.two_site_matrix())
    if isinstance(block, Scale):
        return block.factor * mat(block.content)
    if isinstance(block, Add):
        return sum(mat(b) for b in block.blocks)
    raise TypeError(type(block))


class ArrayReg:
    """Dense register; ``apply_(block)`` mutates the statevector."""

    def __init__(self, state):
        self.state = np.array(state, dtype=np.complex128).reshape(-1)
        self.n = int(round(np.log2(self.state.size)))

    def apply_(self, block):
        if isinstance(block, Put):
            self._apply_locs(block.locs, block.content)
        else:
            self.state = mat(block) @ self.state
        return self

    def _apply_locs(self, locs, m):
        n, k = self.n, len(locs)
        psi = self.state.reshape((2,) * n, order="F")            # axis q-1 <-> site q
        m = np.asarray(m).reshape((2,) * (2 * k), order="F")      # o1..ok, i1..ik
        axes = [l - 1 for l in locs]
        out = np.tensordot(m, psi, axes=(list(range(k, 2 * k)), axes))  # o1..ok, rest...
        out = np.moveaxis(out, list(range(k)), axes)
        self.state = out.reshape(-1, order="F")

    def statevec(self):
        return self.state


def rand_unitary(n, rng):
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    q, r = np.linalg.qr(a)
    return q * (np.diag(r) / np.abs(np.diag(r)))
