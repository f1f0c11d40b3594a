"""Literal numpy restatement of ``src/TensorAD/TensorAD.jl`` and ``src/TensorAD/rules.jl``.

TEST INFRASTRUCTURE ONLY (see oracle/__init__).  A global tape of instructions, each a
tuple of partial pullbacks; pullbacks are written with ``DiffTensor`` operations, so a
backward pass is itself taped (second order works, ``src/timeevolve.jl:15-52``).
Gradient convention: ``g = dL/d(re) + i dL/d(im)`` (e.g. ``Number*T`` pulls back
``dy*conj(a)``, ``rules.jl:71-75``).
"""
import numpy as np

from . import einsum as _es

ADTYPES = (np.float32, np.float64, np.complex64, np.complex128)


class PartialBack:  # TensorAD.jl:8-16
    def __init__(self, input_ids, output_ids, back, keep, active=True):
        self.input_ids = input_ids
        self.output_ids = output_ids
        self.back = back
        self.active = [active]
        self._keep = keep  # strong refs so python ids stay unique while the tape lives


class Instruction:  # TensorAD.jl:18-21
    def __init__(self, info, backs):
        self.info = info
        self.backs = tuple(backs)


class Tape:  # TensorAD.jl:22-24
    def __init__(self):
        self.instructs = []


GLOBAL_TAPE = Tape()            # TensorAD.jl:25
GLOBAL_REQUIRES_GRAD = {}       # TensorAD.jl:27 (never emptied in the reference)


class DiffTensor:  # TensorAD.jl:30-40
    __array_priority__ = 1000
    __array_ufunc__ = None

    def __init__(self, data, requires_grad=True):
        if isinstance(data, DiffTensor):
            raise ValueError("DiffTensor in DiffTensor is forbidden to prevent errors.")
        data = np.asarray(data)
        if data.dtype.type not in ADTYPES:
            data = data.astype(np.complex128 if np.iscomplexobj(data) else np.float64)
        self.data = data
        requires_grad_(self, requires_grad)

    # ---- array interface ----
    @property
    def shape(self):
        return self.data.shape

    @property
    def ndim(self):
        return self.data.ndim

    @property
    def size(self):
        return self.data.size

    @property
    def dtype(self):
        return self.data.dtype

    def __len__(self):
        return self.data.shape[0]

    # ---- operator overloads -> rules ----
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
        s = "DiffTensor[%s]" % "x".join(map(str, self.shape))
        return s + (" ✓" if requires_grad(self) else "")


def getid(x):
    if isinstance(x, tuple):
        return tuple(getid(t) for t in x)
    return id(x)


def requires_grad(t):  # TensorAD.jl:49-50
    if isinstance(t, tuple):
        return any(requires_grad(x) for x in t)
    return GLOBAL_REQUIRES_GRAD[getid(t)]


def requires_grad_(t, val=True):  # TensorAD.jl:51-53
    GLOBAL_REQUIRES_GRAD[getid(t)] = bool(val)
    t._rg = bool(val)


def difftensor(data, info, *backs):  # TensorAD.jl:41-47
    mask = [requires_grad(x) for x, _ in backs]
    t = DiffTensor(data, requires_grad=any(mask))
    kept = [(x, b) for (x, b), m in zip(backs, mask) if m]
    GLOBAL_TAPE.instructs.append(Instruction(info, [
        PartialBack(getid(x), getid(t), b, keep=(x, t)) for x, b in kept]))
    return t


def getdata(x):
    return x.data if isinstance(x, DiffTensor) else x


def zero(x):  # TensorAD.jl:76
    return DiffTensor(np.zeros_like(x.data), requires_grad=False)


def getgrad(d, x):  # TensorAD.jl:80-88
    if isinstance(x, tuple):
        return tuple(getgrad(d, t) for t in x)
    return d[getid(x)] if getid(x) in d else zero(x)


def accumulate_gradient_(storage, key, g):  # TensorAD.jl:89-102
    if isinstance(key, tuple):
        for k, gi in zip(key, g):
            accumulate_gradient_(storage, k, gi)
        return
    if key not in storage:
        storage[key] = g
    else:
        storage[key] = storage[key] + g


def gradient(f, *args, **kwargs):  # TensorAD.jl:131-140
    GLOBAL_TAPE.instructs.clear()
    y = f(*args, **kwargs)
    storage = init_storage_(y)
    back_(GLOBAL_TAPE, storage)
    return getgrad(storage, tuple(args))


def init_storage_(y, requires_grad=False):  # TensorAD.jl:142-146
    storage = {}
    accumulate_gradient_(storage, getid(y), DiffTensor(np.ones(y.data.shape, dtype=y.data.dtype),
                                                       requires_grad=requires_grad))
    storage["_keep"] = [y]
    return storage


def back_(tape, storage):  # TensorAD.jl:148-164
    for i in range(len(tape.instructs) - 1, -1, -1):
        for pb in tape.instructs[i].backs:
            if not pb.active[0]:
                continue
            dy = storage.get(pb.output_ids)
            if dy is not None:
                grads = pb.back(dy)
                accumulate_gradient_(storage, pb.input_ids, grads)


def jacobian(f, x):  # TensorAD.jl:170-185
    GLOBAL_TAPE.instructs.clear()
    slices = []
    y = f(x)
    for i in range(y.data.size):
        storage = {}
        gy = zero(y)
        flat = np.zeros(y.data.size, dtype=y.data.dtype)
        flat[i] = 1
        gy.data = flat.reshape(y.data.shape, order="F")
        accumulate_gradient_(storage, getid(y), gy)
        back_(GLOBAL_TAPE, storage)
        slices.append(getgrad(storage, x).data.reshape(-1, order="F"))
    return np.stack(slices, axis=0)


def hessian(f, x):  # TensorAD.jl:166-168
    return jacobian(lambda z: gradient(f, z)[0], x)


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
    return not np.iscomplexobj(x.data)


def projectto(x, y=None):  # rules.jl:124-136
    if y is None:
        return lambda z: projectto(x, z)
    if _is_real(x) and not _is_real(y):
        return real(y)
    if not _is_real(x) and _is_real(y):
        return bcomplex(y)
    return y


def _einsum_rule(ixs, xs, iy, size_dict):  # rules.jl:1-8
    sd = _es._size_dict(ixs, [x.data for x in xs], size_dict)
    y = _es.einsum_raw(ixs, [x.data for x in xs], iy, sd)
    xs = list(xs)

    def mk(i):
        return lambda dy: _es.einsum_grad(ixs, xs, iy, sd, conj(dy), i, conj=conj, project=_project_eltype)
    return difftensor(y, "einsum", *[(xs[i], mk(i)) for i in range(len(xs))])


def _project_eltype(x, y):  # rules.jl:165-166 (map(ProjectTo{T}, y))
    return real(y) if _is_real(x) else bcomplex(y)


def add(x, y):  # rules.jl:10-12
    return difftensor(x.data + y.data, "+", (x, projectto(x)), (y, projectto(y)))


def sub(x, y):  # rules.jl:13-15
    return difftensor(x.data - y.data, "-", (x, projectto(x)), (y, lambda dz: projectto(y, neg(dz))))


def neg(x):  # rules.jl:16-18
    return difftensor(-x.data, "-", (x, lambda dz: neg(dz)))


def bdiv(x, y):  # rules.jl:19-22
    return difftensor(x.data / y.data, "./",
                      (x, lambda dz: projectto(x, bdiv(dz, conj(y)))),
                      (y, lambda dz: projectto(y, bmul(neg(dz), conj(bdiv(x, bpow(y, 2)))))))


def binv(x):  # rules.jl:23-26
    return difftensor(1.0 / x.data, "inv.",
                      (x, lambda dz: projectto(x, bmul(neg(dz), conj(bpow(binv(x), 2))))))


def bpow(x, y):  # rules.jl:27-35
    return difftensor(np.asarray(x.data ** y), ".^",
                      (x, lambda dz: bmul(dz, conj(mul(y, bpow(x, y - 1))))))


def bsqrt(x):  # rules.jl:36-39
    return difftensor(np.sqrt(x.data), "sqrt.",
                      (x, lambda dz: bdiv(mul(0.5, dz), conj(bsqrt(x)))))


def copy(x):  # rules.jl:40-42
    return difftensor(x.data.copy(), "copy", (x, lambda dz: dz))


def getindex(x, idx):  # rules.jl:43-49
    if not isinstance(idx, tuple):
        idx = (idx,)
    if all(isinstance(i, (int, np.integer)) for i in idx):
        raise IndexError("get element is forbidden!")
    return difftensor(np.array(x.data[idx]), "getindex", (x, lambda dy: accum(zero(x), idx, dy)))


def accum(x, idx, y):  # rules.jl:51-59
    if not isinstance(x, DiffTensor):
        z = np.array(x, copy=True)
        z[idx] += y
        return z
    z = np.array(x.data, copy=True)
    z[idx] = z[idx] + y.data
    return difftensor(z, "accum", (x, projectto(x)), (y, lambda dz: projectto(y, getindex(dz, idx))))


def transpose(x):  # rules.jl:60-63
    return difftensor(np.array(x.data.T), "transpose", (x, lambda dz: transpose(dz)))


def reshape(x, sz):  # rules.jl:65-69 (column-major)
    size0 = x.data.shape
    return difftensor(np.reshape(x.data, sz, order="F"), "reshape", (x, lambda dy: reshape(dy, size0)))


def mul(a, b):  # rules.jl:71-88
    if not isinstance(a, DiffTensor) and not isinstance(b, DiffTensor):
        return a * b
    if not isinstance(a, DiffTensor):     # Number * T
        return difftensor(np.asarray(a * b.data), "*", (b, lambda dy: projectto(b, mul(dy, np.conj(a)))))
    if not isinstance(b, DiffTensor):
        return mul(b, a)
    if a.ndim == 0 and b.ndim == 0:
        return difftensor(np.asarray(a.data[()] * b.data[()]), "*",
                          (a, lambda dz: projectto(a, mul(conj(b), dz))),
                          (b, lambda dz: projectto(b, mul(conj(a), dz))))
    if a.ndim == 0:
        return difftensor(np.asarray(a.data[()] * b.data), "*",
                          (a, lambda dy: projectto(a, sum_(bmul(dy, conj(b))))),
                          (b, lambda dy: projectto(b, mul(dy, conj(a)))))
    if b.ndim == 0:
        return mul(b, a)
    if a.ndim == 2 and b.ndim == 2:
        return _es.einsum([[1, 2], [2, 3]], [a, b], [1, 3])
    raise TypeError("* is not defined for these DiffTensor ranks")


def sum_(a):  # rules.jl:91-93
    return difftensor(np.asarray(a.data.sum()), "sum", (a, lambda dz: fill(dz, a.data.shape)))


def fill(a, size):  # rules.jl:94-96
    return difftensor(np.full(size, a.data[()]), "fill", (a, lambda dz: sum_(dz)))


def bmul(a, b):  # rules.jl:98-101
    return difftensor(a.data * b.data, ".*", (a, lambda dy: bmul(dy, conj(b))), (b, lambda dy: bmul(dy, conj(a))))


def bsin(a):  # rules.jl:102-105
    return difftensor(np.sin(a.data), "sin.", (a, lambda dy: bmul(dy, conj(bcos(a)))))


def bcos(a):  # rules.jl:106-109
    return difftensor(np.cos(a.data), "cos.", (a, lambda dy: bmul(neg(dy), conj(bsin(a)))))


def cat(tensors, dims):  # rules.jl:110-123  (dims is 1-based like Julia)
    tensors = list(tensors)
    ax = dims - 1
    nd = max(tensors[0].ndim, dims)
    datas = [t.data.reshape(t.data.shape + (1,) * (nd - t.ndim)) for t in tensors]
    ends = np.cumsum([d.shape[ax] for d in datas])

    def mk(i):
        lo = 0 if i == 0 else int(ends[i - 1])
        hi = int(ends[i])

        def back(dy):
            idx = tuple(slice(lo, hi) if k == ax else slice(None) for k in range(dy.ndim))
            g = getindex(dy, idx)
            return g if g.data.shape == tensors[i].data.shape else reshape(g, tensors[i].data.shape)
        return back
    return difftensor(np.concatenate(datas, axis=ax), "cat", *[(tensors[i], mk(i)) for i in range(len(tensors))])


def vcat(*tensors):
    return cat(tensors, 1)


def hcat(*tensors):
    return cat(tensors, 2)


def real(x):  # rules.jl:142-144
    return difftensor(np.real(x.data).copy(), "real", (x, projectto(x)))


def imag(x):  # rules.jl:145-147
    return difftensor(np.imag(x.data).copy(), "imag", (x, lambda dz: mul(1j, dz)))


def conj(x):  # rules.jl:148-150
    if not isinstance(x, DiffTensor):
        return np.conj(x)
    return difftensor(np.conj(x.data), "conj", (x, lambda dz: conj(dz)))


def babs(x):  # rules.jl:151-154
    return difftensor(np.abs(x.data), "abs.", (x, lambda dy: bmul(bsign(x), dy)))


def bsign(x):  # rules.jl:155-157
    return bdiv(x, babs(x))


def bcomplex(x):  # rules.jl:158-162
    if not _is_real(x):
        return x
    return difftensor(x.data.astype(np.complex128), "Complex.", (x, lambda dy: real(dy)))


_es._DIFF_HOOK = (lambda x: isinstance(x, DiffTensor), _einsum_rule, getindex, reshape)
