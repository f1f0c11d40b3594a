"""``B200Array``: the device array type that slots in behind the reference's array-generic
code (``vertex_tensors::Vector{<:AbstractArray{T}}``, src/peps.jl:41; ``DiffTensor`` wraps any
``AbstractArray``, src/TensorAD/TensorAD.jl:30-40).  Storage is ComplexF64, column-major, owned
by the library; ``is_real`` mirrors Julia's real element types (norms, singular values).
"""
import ctypes as C
import os
import weakref

import numpy as np

from . import _lib
from ._lib import TneError

# tne_map_op / tne_map2_op
COPY, CONJ, REAL, IMAG, ABS, SQRT, INV, NEG, SCALE, POW, SIN, COS, SAFE_INV = range(13)
ADD, SUB, MUL, DIV = range(4)


class Context:
    """One context per process / GPU (``tne_ctx_create``)."""

    def __init__(self, device=None):
        L = _lib.lib()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        rc = L.tne_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise TneError("tne_ctx_create failed: %s" % (L.tne_last_error(None) or b"").decode())
        self.h = h
        self.L = L
        self.device = int(device)
        self._fin = weakref.finalize(self, L.tne_ctx_destroy, h)

    def check(self, rc):
        if rc != 0:
            raise TneError((self.L.tne_last_error(self.h) or b"?").decode())

    def sync(self):
        self.check(self.L.tne_sync(self.h))

    def set_option(self, key, value):
        self.check(self.L.tne_set_option(self.h, key.encode(), int(value)))

    def set_mem_limit(self, nbytes):
        self.check(self.L.tne_set_mem_limit(self.h, int(nbytes)))

    def launch_count(self):
        return int(self.L.tne_launch_count(self.h))

    def stream(self):
        return self.L.tne_stream(self.h)

    def program_stats(self):
        s = (C.c_double * 6)()
        self.check(self.L.tne_last_program_stats(self.h, s))
        return dict(flops=s[0], bytes=s[1], contractions=int(s[2]), workspace_bytes=int(s[3]),
                    recompute_flops=s[4], values=int(s[5]))

    def plan_cache_info(self):
        """(programs held, calls served from the cache) of the tree entry points (``tne_plan_cache_info``)."""
        a, b = C.c_int64(), C.c_int64()
        self.check(self.L.tne_plan_cache_info(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def profile_collect(self, fine=False):
        """Per kernel: launches, device ms, algorithmic flops and bytes since the last call.  ``fine``: keyed by
        (kernel, bucket) where the bucket of an absorb kernel is log2(padded K / 2) (5, 6: fused launches)."""
        names = ["gett_generic", "gett_absorb", "gett_reduce", "gett_absorb_tma", "gett_reduce_tma", "gett_site2", "gett_site2_bwd", "gett_gemm"]
        ncls = 64
        out = (C.c_double * (4 * ncls))()
        self.check(self.L.tne_profile_collect(self.h, out))
        res = {}
        for c in range(ncls):
            if out[4 * c] == 0:
                continue
            key = (names[c // 8], c % 8) if fine else names[c // 8]
            d = res.setdefault(key, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            d["launches"] += int(out[4 * c]); d["ms"] += out[4 * c + 1]; d["flops"] += out[4 * c + 2]; d["bytes"] += out[4 * c + 3]
        if not fine:
            for n in names:
                res.setdefault(n, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
        return res

    def measure_fp64_peak(self):
        """FP64 DMMA / DFMA issue peaks of this device right now, in TFLOP/s (``tne_measure_fp64_peak``)."""
        a, b = C.c_double(), C.c_double()
        self.check(self.L.tne_measure_fp64_peak(self.h, C.byref(a), C.byref(b)))
        return dict(dmma_tflops=a.value, dfma_tflops=b.value)

    def mem_info(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.check(self.L.tne_mem_info(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(in_use=a.value, peak=b.value, workspace=c.value)


_CTX = None


def context():
    """The process-wide context (created on first use; needs a B200)."""
    global _CTX
    if _CTX is None:
        _CTX = Context()
    return _CTX


def _i64(seq):
    return (C.c_int64 * max(1, len(seq)))(*[int(x) for x in seq])


def _i32(seq):
    return (C.c_int32 * max(1, len(seq)))(*[int(x) for x in seq])


class B200Array:
    """Device tensor handle.  Immutable from the host's point of view (ops return new arrays)."""

    __array_priority__ = 900
    __array_ufunc__ = None

    def __init__(self, handle, shape, is_real=False, ctx=None):
        self.ctx = ctx or context()
        self.handle = int(handle)
        self.shape = tuple(int(s) for s in shape)
        self.is_real = bool(is_real)
        self._fin = weakref.finalize(self, _free, self.ctx, self.handle)

    # -- construction / transfer -------------------------------------------------------
    @staticmethod
    def from_numpy(x, ctx=None):
        ctx = ctx or context()
        x = np.asarray(x)
        is_real = not np.iscomplexobj(x)
        if x.dtype == np.complex128 and x.ndim > 0 and x.flags.f_contiguous:
            host = x  # borrowed as is (e.g. a pinned staging buffer): no extra host copy
        else:
            host = np.array(x, dtype=np.complex128, order="F")  # keeps 0-dim arrays 0-dim
        h = C.c_int64()
        ctx.check(ctx.L.tne_alloc(ctx.h, host.ndim, _i64(host.shape), C.byref(h)))
        if host.size:
            ctx.check(ctx.L.tne_upload(ctx.h, h.value, host.ctypes.data_as(C.c_void_p)))
        return B200Array(h.value, host.shape, is_real, ctx)

    def to_numpy(self):
        out = np.empty(self.shape, dtype=np.complex128, order="F")
        if out.size:
            self.ctx.check(self.ctx.L.tne_download(self.ctx.h, self.handle, out.ctypes.data_as(C.c_void_p)))
        return out.real.copy() if self.is_real else out

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    @property
    def dtype(self):
        return np.dtype(np.float64 if self.is_real else np.complex128)

    def item(self):
        v = self.to_numpy().reshape(-1)[0]
        return float(v) if self.is_real else complex(v)

    def __repr__(self):
        return "B200Array(%s, %s)" % ("x".join(map(str, self.shape)) or "0-dim", "Float64" if self.is_real else "ComplexF64")

    # -- elementwise ---------------------------------------------------------------------
    def _map(self, op, s=0.0, is_real=None):
        h = C.c_int64()
        s = complex(s)
        self.ctx.check(self.ctx.L.tne_map(self.ctx.h, op, self.handle, s.real, s.imag, C.byref(h)))
        return B200Array(h.value, self.shape, self.is_real if is_real is None else is_real, self.ctx)

    def _map2(self, op, other):
        h = C.c_int64()
        self.ctx.check(self.ctx.L.tne_map2(self.ctx.h, op, self.handle, other.handle, C.byref(h)))
        shape = self.shape if self.size >= other.size and not (self.size == 1 and other.ndim > self.ndim) else other.shape
        return B200Array(h.value, shape, self.is_real and other.is_real, self.ctx)

    def copy(self):
        return self._map(COPY)

    def conj(self):
        return self._map(CONJ)

    def real(self):
        return self._map(REAL, is_real=True)

    def imag(self):
        return self._map(IMAG, is_real=True)

    def abs(self):
        return self._map(ABS, is_real=True)

    def sqrt(self):
        return self._map(SQRT)

    def inv(self):
        return self._map(INV)

    def safe_inv(self):
        return self._map(SAFE_INV)

    def sin(self):
        return self._map(SIN)

    def cos(self):
        return self._map(COS)

    def pow(self, p):
        return self._map(POW, float(p))

    def scale(self, s):
        s = complex(s)
        return self._map(SCALE, s, is_real=self.is_real and s.imag == 0.0)

    def to_complex(self):
        return B200Array.view_of(self, self.shape, is_real=False) if self.is_real else self

    def __neg__(self):
        return self._map(NEG)

    def __add__(self, o):
        return self._map2(ADD, o)

    def __sub__(self, o):
        return self._map2(SUB, o)

    def bmul(self, o):
        return self._map2(MUL, o)

    def bdiv(self, o):
        return self._map2(DIV, o)

    def __mul__(self, o):
        if isinstance(o, B200Array):
            if o.ndim == 0 or self.ndim == 0:
                return self._map2(MUL, o)
            raise TypeError("* between tensors is only defined with a 0-dim operand")
        return self.scale(o)

    __rmul__ = __mul__

    def sum(self):
        h = C.c_int64()
        self.ctx.check(self.ctx.L.tne_reduce_sum(self.ctx.h, self.handle, C.byref(h)))
        return B200Array(h.value, (), self.is_real, self.ctx)

    # -- structure -----------------------------------------------------------------------
    @staticmethod
    def view_of(x, shape, is_real=None):
        h = C.c_int64()
        x.ctx.check(x.ctx.L.tne_reshape(x.ctx.h, x.handle, len(shape), _i64(shape), C.byref(h)))
        return B200Array(h.value, shape, x.is_real if is_real is None else is_real, x.ctx)

    def reshape(self, shape):
        """Column-major reshape; shares storage (Julia ``reshape``)."""
        return B200Array.view_of(self, tuple(shape))

    def slice(self, lo, count):
        """``x[lo+1 : lo+count]`` of the flattened tensor (0-based ``lo``)."""
        h = C.c_int64()
        self.ctx.check(self.ctx.L.tne_slice(self.ctx.h, self.handle, int(lo), int(count), C.byref(h)))
        return B200Array(h.value, (int(count),), self.is_real, self.ctx)

    def embed(self, lo, total):
        """``accum(zeros(total), lo+1:lo+len, self)`` (src/TensorAD/rules.jl:46-55)."""
        h = C.c_int64()
        self.ctx.check(self.ctx.L.tne_embed(self.ctx.h, self.handle, int(lo), int(total), C.byref(h)))
        return B200Array(h.value, (int(total),), self.is_real, self.ctx)

    def fill(self, shape):
        h = C.c_int64()
        self.ctx.check(self.ctx.L.tne_fill(self.ctx.h, self.handle, len(shape), _i64(shape), C.byref(h)))
        return B200Array(h.value, shape, self.is_real, self.ctx)


def _free(ctx, handle):
    try:
        ctx.L.tne_free(ctx.h, handle)
    except Exception:
        pass


def concat(arrays):
    """``vcat(vec.(xs)...)`` (src/peps.jl:202)."""
    ctx = arrays[0].ctx
    h = C.c_int64()
    ctx.check(ctx.L.tne_concat(ctx.h, len(arrays), _i64([a.handle for a in arrays]), C.byref(h)))
    return B200Array(h.value, (sum(a.size for a in arrays),), all(a.is_real for a in arrays), ctx)


def zeros(shape, ctx=None):
    return B200Array.from_numpy(np.zeros(shape, dtype=np.complex128), ctx)


def einsum_device(ixs, xs, iy, conj=None, size_dict=None):
    """``OMEinsum.einsum(EinCode(ixs, iy), xs, size_dict)`` for one or two device tensors.  ``size_dict`` is only
    consulted for labels that no input carries (OMEinsum's repeat rule, e.g. the pullback ``ein"->ii"`` of a trace)."""
    ctx = xs[0].ctx
    flat = [l for ix in ixs for l in ix]
    h = C.c_int64()
    cj = _i32(conj if conj is not None else [0] * len(xs))
    sizes = {}
    for ix, x in zip(ixs, xs):
        for l, s in zip(ix, x.shape):
            sizes[l] = s
    extra = [l for l in dict.fromkeys(iy) if l not in sizes]
    if extra:
        if size_dict is None or any(l not in size_dict for l in extra):
            raise ValueError("einsum: output labels %r are in no input and have no size" % (extra,))
        for l in extra:
            sizes[l] = int(size_dict[l])
        ctx.check(ctx.L.tne_einsum_sized(ctx.h, len(xs), _i32(flat), _i32([len(ix) for ix in ixs]),
                                         _i64([x.handle for x in xs]), cj, _i32(iy), len(iy),
                                         len(extra), _i32(extra), _i64([sizes[l] for l in extra]), C.byref(h)))
    else:
        ctx.check(ctx.L.tne_einsum(ctx.h, len(xs), _i32(flat), _i32([len(ix) for ix in ixs]),
                                   _i64([x.handle for x in xs]), cj, _i32(iy), len(iy), C.byref(h)))
    return B200Array(h.value, tuple(sizes[l] for l in iy), all(x.is_real for x in xs), ctx)


def init_comm(ctx, dist, rank, world):
    """Create the library's NCCL communicator (``tne_comm_init``); the unique id travels through the
    host program's ``torch.distributed`` group."""
    import torch
    idbuf = (C.c_ubyte * 128)()
    if rank == 0:
        ctx.check(ctx.L.tne_nccl_unique_id(idbuf))
    t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    ctx.check(ctx.L.tne_comm_init(ctx.h, C.create_string_buffer(raw, 128), rank, world))


def allreduce_sum(ctx, arrays):
    """In-place sum over the ranks (``tne_allreduce_sum``: NCCL over NVLink)."""
    ctx.check(ctx.L.tne_allreduce_sum(ctx.h, len(arrays), _i64([a.handle for a in arrays])))


def svd_trunc(a, Dmax, eps=0.0):
    """``U, S, V = svd(A)`` truncated like src/peps.jl:381-390 (``A ~= U * Diagonal(S) * V'``); returns
    ``(U, S, V, kept, truncation_error)``.  Sizes of a bond update (even m, even n <= 16)."""
    ctx = a.ctx
    hu, hs, hv = C.c_int64(), C.c_int64(), C.c_int64()
    kept, err = C.c_int32(), C.c_double()
    ctx.check(ctx.L.tne_svd_trunc(ctx.h, a.handle, int(Dmax), float(eps), C.byref(hu), C.byref(hs), C.byref(hv),
                                  C.byref(kept), C.byref(err)))
    k = kept.value
    return (B200Array(hu.value, (a.shape[0], k), False, ctx), B200Array(hs.value, (k,), True, ctx),
            B200Array(hv.value, (a.shape[1], k), False, ctx), k, err.value)
