"""TensorAD on device arrays: the reference's reverse-mode layer (src/TensorAD/TensorAD.jl,
src/TensorAD/rules.jl) with ``DiffTensor.data`` a :class:`B200Array`.

The tape, ``requires_grad`` bookkeeping, ``back!`` and every rule keep the reference's
structure and gradient convention (``g = dL/dre + i dL/dim``); each rule's arithmetic is a
device call (``tne_einsum`` / ``tne_map`` / ``tne_map2`` ...).  Pullbacks are written with
``DiffTensor`` operations, so a backward pass is itself taped (needed by ``apply_smatrix``,
src/timeevolve.jl:15-52).  Whole-tree contractions are one instruction when
``einsum.FUSED_TREES`` is on: the pullback is the fused device reverse pass
(``tne_expect_terms`` with ``n_grad > 0``), which is first-order only.
"""
import numpy as np

from . import array as _arr
from . import einsum as _es
from .array import B200Array


class PartialBack:  # TensorAD.jl:8-16
    def __init__(self, input_ids, output_ids, back, keep, active=True):
        self.input_ids = input_ids
        self.output_ids = output_ids
        self.back = back
        self.active = [active]
        self._keep = keep


class Instruction:  # TensorAD.jl:18-21
    def __init__(self, info, backs):
        self.info = info
        self.backs = tuple(backs)


class Tape:  # TensorAD.jl:22-24
    def __init__(self):
        self.instructs = []

    def __repr__(self):  # TensorAD.jl:114-124
        return "\n".join(i.info + "  " + " ".join("●" if b.active[0] else "○" for b in i.backs) for i in self.instructs)


GLOBAL_TAPE = Tape()
GLOBAL_REQUIRES_GRAD = {}


class DiffTensor:  # TensorAD.jl:30-40
    __array_priority__ = 1000
    __array_ufunc__ = None

    def __init__(self, data, requires_grad=True):
        if isinstance(data, DiffTensor):
            raise ValueError("DiffTensor in DiffTensor is forbidden to prevent errors.")
        if not isinstance(data, B200Array):
            data = B200Array.from_numpy(np.asarray(data))
        self.data = data
        requires_grad_(self, requires_grad)

    @property
    def shape(self):
        return self.data.shape

    @property
    def ndim(self):
        return self.data.ndim

    @property
    def size(self):
        return self.data.size

    def to_numpy(self):
        return self.data.to_numpy()

    def __add__(self, o):
        return add(self, o)

    def __sub__(self, o):
        return sub(self, o)

    def __neg__(self):
        return neg(self)

    def __mul__(self, o):
        return mul(self, o)

    def __rmul__(self, o):
        return mul(o, self)

    def __getitem__(self, idx):
        return getindex(self, idx)

    def __repr__(self):
        return "DiffTensor[%s]%s" % ("x".join(map(str, self.shape)), " ✓" if requires_grad(self) else "")


def getid(x):
    if isinstance(x, tuple):
        return tuple(getid(t) for t in x)
    return id(x)


def requires_grad(t):
    if isinstance(t, tuple):
        return any(requires_grad(x) for x in t)
    return GLOBAL_REQUIRES_GRAD[getid(t)]


def requires_grad_(t, val=True):
    GLOBAL_REQUIRES_GRAD[getid(t)] = bool(val)


def difftensor(data, info, *backs):  # TensorAD.jl:41-47
    mask = [requires_grad(x) for x, _ in backs]
    t = DiffTensor(data, requires_grad=any(mask))
    kept = [(x, b) for (x, b), m in zip(backs, mask) if m]
    GLOBAL_TAPE.instructs.append(Instruction(info, [PartialBack(getid(x), getid(t), b, keep=(x, t)) for x, b in kept]))
    return t


def getdata(x):
    return x.data if isinstance(x, DiffTensor) else x


def zero(x):
    return DiffTensor(_arr.zeros(x.data.shape, x.data.ctx), requires_grad=False)


def getgrad(d, x):
    if isinstance(x, tuple):
        return tuple(getgrad(d, t) for t in x)
    return d[getid(x)] if getid(x) in d else zero(x)


def accumulate_gradient_(storage, key, g):  # TensorAD.jl:89-102
    if isinstance(key, tuple):
        for k, gi in zip(key, g):
            accumulate_gradient_(storage, k, gi)
        return
    storage[key] = g if key not in storage else storage[key] + g


def gradient(f, *args, **kwargs):  # TensorAD.jl:131-140
    GLOBAL_TAPE.instructs.clear()
    y = f(*args, **kwargs)
    storage = init_storage_(y)
    back_(GLOBAL_TAPE, storage)
    return getgrad(storage, tuple(args))


def jacobian(f, x):  # TensorAD.jl:170-185
    """``transpose(cat(slices...; dims=2))``: row ``i`` is the gradient of ``y[i]`` (one backward pass per output element
    over the same tape, seeded with a device unit vector).  Returns a device matrix (a ``DiffTensor``, as the reference)."""
    GLOBAL_TAPE.instructs.clear()
    slices = []
    y = f(x)
    n = y.data.size
    for i in range(n):
        storage = {}
        e = np.zeros(n, dtype=np.float64 if y.data.is_real else np.complex128)
        e[i] = 1
        gy = DiffTensor(B200Array.from_numpy(e.reshape(y.data.shape, order="F"), y.data.ctx), requires_grad=False)
        accumulate_gradient_(storage, getid(y), gy)
        storage["_keep"] = [y, gy]
        back_(GLOBAL_TAPE, storage)
        g = getgrad(storage, x)
        slices.append(reshape(g, (g.data.size,)))
    return transpose(hcat(*slices))


def hessian(f, x):  # TensorAD.jl:166-168
    return jacobian(lambda z: gradient(f, z)[0], x)


def init_storage_(y, requires_grad=False):  # TensorAD.jl:142-146
    storage = {}
    ones = B200Array.from_numpy(np.ones(y.data.shape, dtype=np.float64 if y.data.is_real else np.complex128), y.data.ctx)
    accumulate_gradient_(storage, getid(y), DiffTensor(ones, requires_grad=requires_grad))
    storage["_keep"] = [y]
    return storage


def back_(tape, storage):  # TensorAD.jl:148-164
    for i in range(len(tape.instructs) - 1, -1, -1):
        for pb in tape.instructs[i].backs:
            if not pb.active[0]:
                continue
            dy = storage.get(pb.output_ids)
            if dy is not None:
                accumulate_gradient_(storage, pb.input_ids, pb.back(dy))


def propagate_requires_grad_(tape=GLOBAL_TAPE):  # TensorAD.jl:188-202
    for instruct in tape.instructs:
        for pb in instruct.backs:
            ids = pb.input_ids if isinstance(pb.input_ids, tuple) else (pb.input_ids,)
            rg = any(GLOBAL_REQUIRES_GRAD[i] for i in ids)
            pb.active[0] = rg
            outs = pb.output_ids if isinstance(pb.output_ids, tuple) else (pb.output_ids,)
            for o in outs:
                GLOBAL_REQUIRES_GRAD[o] = rg


# ---------------------------------------------------------------------------------------
# rules.jl
# ---------------------------------------------------------------------------------------
def _is_real(x):
    return x.data.is_real


def projectto(x, y=None):  # rules.jl:124-136
    if y is None:
        return lambda z: projectto(x, z)
    if _is_real(x) and not _is_real(y):
        return real(y)
    if not _is_real(x) and _is_real(y):
        return bcomplex(y)
    return y


def _project_eltype(x, y):  # rules.jl:165-166
    return real(y) if _is_real(x) else bcomplex(y)


def _einsum_rule(ixs, xs, iy, size_dict=None):  # rules.jl:1-8
    xs = list(xs)
    y = _arr.einsum_device(ixs, [x.data for x in xs], iy, None, size_dict)
    sd = dict(size_dict or {})   # get_size_dict: the pullbacks may need the size of a label the cotangent lacks
    for ix, x in zip(ixs, xs):
        for l, s in zip(ix, x.shape):
            sd[l] = s

    def mk(i):
        def back(dy):  # OMEinsum.einsum_grad(ixs, xs, iy, size_dict, conj(dy), i)
            nixs = [list(ix) for ix in ixs]
            nxs = list(xs)
            nixs[i] = list(iy)
            nxs[i] = conj(dy)
            return _project_eltype(xs[i], conj(_es.einsum(nixs, nxs, list(ixs[i]), size_dict=sd)))
        return back
    return difftensor(y, "einsum", *[(xs[i], mk(i)) for i in range(len(xs))])


def _tree_rule(nested, slicing, xs):
    """(::NestedEinsum)(xs::DiffTensor...).  Unfused: the recursion of binary einsum rules, as in the
    reference.  Fused: one instruction; all input pullbacks share one device reverse pass."""
    xs = list(xs)
    if not _es.FUSED_TREES:
        if slicing:
            raise NotImplementedError("second-order mode does not support sliced codes")
        return _run_nested_unfused(nested, xs)
    code = _es.SlicedEinsum(slicing, nested) if slicing else nested
    datas = [x.data for x in xs]
    leaves = [j for j in range(len(xs)) if requires_grad(xs[j])]
    if not leaves or len(nested.eins.iy) != 0:   # nothing to differentiate / non-scalar output (forward only)
        y = _es.contract_tree_device(code, datas)

        def no_back(dy):
            raise NotImplementedError("fused trees differentiate scalar outputs only (set einsum.FUSED_TREES = False)")
        return difftensor(y, "nested_einsum", *[(xs[i], no_back) for i in range(len(xs))])
    # value and the pullbacks for a unit cotangent in ONE device program (forward + reverse sweep, no second
    # forward pass); the program is linear in the cotangent, so back(dy) only rescales the stored gradients
    y, grads = _es.expect_terms_device(code, datas, None, [], leaves, 1.0 + 0j)
    return wrap_tree_result(xs, leaves, y, grads)


def wrap_tree_result(xs, leaves, y, grads):
    """The taped result of a fused scalar tree: value ``y`` and, per differentiable leaf, the pullback for a unit
    cotangent.  Also used to re-attach a memoised (value, pullbacks) pair to new leaf objects holding the same numbers."""
    gmap = dict(zip(leaves, grads))

    def mk(i):
        def back(dy):
            out = gmap[i].bmul(dy.data.to_complex())   # holomorphic: G(dy) = dy * G(1)
            out = out.real() if xs[i].data.is_real else out
            return DiffTensor(out, requires_grad=False)
        return back
    return difftensor(y, "nested_einsum", *[(xs[i], mk(i)) for i in range(len(xs))])


def _run_nested_unfused(node, xs):
    if node.isleaf():
        return xs[node.tensorindex]
    ys = [_run_nested_unfused(a, xs) for a in node.args]
    return _es.einsum(node.eins.ixs, ys, node.eins.iy)


def add(x, y):
    return difftensor(x.data + y.data, "+", (x, projectto(x)), (y, projectto(y)))


def sub(x, y):
    return difftensor(x.data - y.data, "-", (x, projectto(x)), (y, lambda dz: projectto(y, neg(dz))))


def neg(x):
    return difftensor(-x.data, "-", (x, lambda dz: neg(dz)))


def bdiv(x, y):
    return difftensor(x.data.bdiv(y.data), "./",
                      (x, lambda dz: projectto(x, bdiv(dz, conj(y)))),
                      (y, lambda dz: projectto(y, bmul(neg(dz), conj(bdiv(x, bpow(y, 2)))))))


def binv(x):
    return difftensor(x.data.inv(), "inv.", (x, lambda dz: projectto(x, bmul(neg(dz), conj(bpow(binv(x), 2))))))


def bpow(x, y):
    return difftensor(x.data.pow(y), ".^", (x, lambda dz: bmul(dz, conj(mul(y, bpow(x, y - 1))))))


def bsqrt(x):
    return difftensor(x.data.sqrt(), "sqrt.", (x, lambda dz: bdiv(mul(0.5, dz), conj(bsqrt(x)))))


def copy(x):
    return difftensor(x.data.copy(), "copy", (x, lambda dz: dz))


def getindex(x, idx):  # rules.jl:43-49 -- contiguous ranges of vectors (variables[a:b])
    if isinstance(idx, tuple) and len(idx) == 1:
        idx = idx[0]
    if not isinstance(idx, slice) or x.ndim != 1:
        raise IndexError("get element is forbidden!" if not isinstance(idx, slice) else "only ranges of vectors are supported on device")
    lo, hi, step = idx.indices(x.shape[0])
    assert step == 1
    total = x.shape[0]
    return difftensor(x.data.slice(lo, hi - lo), "getindex", (x, lambda dy: accum_zero(total, lo, dy)))


def accum_zero(total, lo, y):  # accum(zero(x), indices, dy), rules.jl:51-59
    n = y.shape[0]
    return difftensor(y.data.embed(lo, total), "accum", (y, lambda dz: projectto(y, getindex(dz, slice(lo, lo + n)))))


def transpose(x):
    return difftensor(_arr.einsum_device([[1, 2]], [x.data], [2, 1]), "transpose", (x, lambda dz: transpose(dz)))


def reshape(x, sz):  # rules.jl:65-69
    size0 = x.data.shape
    return difftensor(x.data.reshape(tuple(sz)), "reshape", (x, lambda dy: reshape(dy, size0)))


def mul(a, b):  # rules.jl:71-88
    if not isinstance(a, DiffTensor) and not isinstance(b, DiffTensor):
        return a * b
    if not isinstance(a, DiffTensor):
        return difftensor(b.data.scale(a), "*", (b, lambda dy: projectto(b, mul(dy, np.conj(a)))))
    if not isinstance(b, DiffTensor):
        return mul(b, a)
    if a.ndim == 0 and b.ndim == 0:
        return difftensor(a.data.bmul(b.data), "*",
                          (a, lambda dz: projectto(a, mul(conj(b), dz))),
                          (b, lambda dz: projectto(b, mul(conj(a), dz))))
    if a.ndim == 0:
        return difftensor(b.data.bmul(a.data), "*",
                          (a, lambda dy: projectto(a, sum_(bmul(dy, conj(b))))),
                          (b, lambda dy: projectto(b, mul(dy, conj(a)))))
    if b.ndim == 0:
        return mul(b, a)
    if a.ndim == 2 and b.ndim == 2:
        return _es.einsum([[1, 2], [2, 3]], [a, b], [1, 3])
    raise TypeError("* is not defined for these DiffTensor ranks")


def sum_(a):
    return difftensor(a.data.sum(), "sum", (a, lambda dz: fill(dz, a.data.shape)))


def fill(a, size):
    return difftensor(a.data.fill(tuple(size)), "fill", (a, lambda dz: sum_(dz)))


def bmul(a, b):
    return difftensor(a.data.bmul(b.data), ".*", (a, lambda dy: bmul(dy, conj(b))), (b, lambda dy: bmul(dy, conj(a))))


def bsin(a):
    return difftensor(a.data.sin(), "sin.", (a, lambda dy: bmul(dy, conj(bcos(a)))))


def bcos(a):
    return difftensor(a.data.cos(), "cos.", (a, lambda dy: bmul(neg(dy), conj(bsin(a)))))


def cat(tensors, dims):  # rules.jl:110-123 (``dims`` 1-based like Julia)
    """``cat(a, B...; dims)``.  Column-major storage makes concatenation along the LAST axis a plain concatenation of the
    flattened operands; any other axis is first rotated to the end with a (taped) permutation einsum.  The pullback of
    operand ``i`` is the matching slice of the cotangent, ``getindex(dy, ..., lo:hi, ...)`` in the reference."""
    tensors = list(tensors)
    nd = max(tensors[0].ndim, dims)
    ax = dims - 1
    shapes = [tuple(t.shape) + (1,) * (nd - t.ndim) for t in tensors]
    for sh in shapes:
        if any(sh[k] != shapes[0][k] for k in range(nd) if k != ax):
            raise ValueError("cat: mismatch in dimension sizes")
    if ax != nd - 1:   # rotate ``ax`` to the end, concatenate there, rotate back
        labs = list(range(1, nd + 1))
        moved = [l for l in labs if l != dims] + [dims]
        rot = [_es.einsum([labs], [reshape(t, sh)], moved) for t, sh in zip(tensors, shapes)]
        return _es.einsum([moved], [cat(rot, nd)], labs)
    lead = int(np.prod(shapes[0][:-1], dtype=np.int64))
    ends = np.cumsum([sh[-1] for sh in shapes])
    out_shape = shapes[0][:-1] + (int(ends[-1]),)

    def mk(i):
        lo = 0 if i == 0 else int(ends[i - 1])
        hi = int(ends[i])

        def back(dy):
            flat = reshape(dy, (lead * int(ends[-1]),))
            return reshape(getindex(flat, slice(lead * lo, lead * hi)), tuple(tensors[i].shape))
        return back
    data = _arr.concat([t.data for t in tensors]).reshape(out_shape)
    return difftensor(data, "cat", *[(tensors[i], mk(i)) for i in range(len(tensors))])


def vcat(*tensors):
    return cat(tensors, 1)


def hcat(*tensors):
    return cat(tensors, 2)


def real(x):
    return difftensor(x.data.real(), "real", (x, projectto(x)))


def imag(x):
    return difftensor(x.data.imag(), "imag", (x, lambda dz: mul(1j, dz)))


def conj(x):
    if not isinstance(x, DiffTensor):
        return x.conj() if isinstance(x, B200Array) else np.conj(x)
    return difftensor(x.data.conj(), "conj", (x, lambda dz: conj(dz)))


def babs(x):
    return difftensor(x.data.abs(), "abs.", (x, lambda dy: bmul(bsign(x), dy)))


def bsign(x):
    return bdiv(x, babs(x))


def bcomplex(x):
    if not _is_real(x):
        return x
    return difftensor(x.data.to_complex(), "Complex.", (x, lambda dy: real(dy)))


_es._DIFF_HOOK = {"is": lambda x: isinstance(x, DiffTensor), "einsum": _einsum_rule, "tree": _tree_rule}
