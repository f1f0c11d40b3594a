"""ctypes binding of ``libtne_b200.so`` (the C ABI declared in ``include/tne.h``).

This is the same call-for-call surface the Julia shim ``julia/TensorNetworkEvolveB200.jl``
binds with ``ccall``.  There is no CPU fallback: if the shared library is missing or no
CUDA device is present, importing :func:`lib` / creating a context raises.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtne_b200.so")
_LIB = None

MAX_RANK = 32


class TneTree(C.Structure):
    _fields_ = [
        ("n_leaves", C.c_int32),
        ("n_nodes", C.c_int32),
        ("leaf_rank", C.POINTER(C.c_int32)),
        ("leaf_labels", C.POINTER(C.c_int32)),
        ("node_children", C.POINTER(C.c_int32)),
        ("node_out_rank", C.POINTER(C.c_int32)),
        ("node_out_labels", C.POINTER(C.c_int32)),
        ("n_sliced", C.c_int32),
        ("sliced_labels", C.POINTER(C.c_int32)),
        ("shard_rank", C.c_int32),
        ("shard_world", C.c_int32),
        ("seg_shard", C.c_int32),
        ("n_segments", C.c_int32),
        ("seg_first_node", C.POINTER(C.c_int32)),
        ("seg_sliced_off", C.POINTER(C.c_int32)),
        ("seg_sliced_labels", C.POINTER(C.c_int32)),
    ]


class TneTerm(C.Structure):
    _fields_ = [
        ("n_repl", C.c_int32),
        ("leaf", C.c_int32 * 4),
        ("repl", C.c_int64 * 4),
        ("coef_re", C.c_double),
        ("coef_im", C.c_double),
    ]


class TneBondJob(C.Structure):
    _fields_ = [
        ("ti", C.c_int64), ("tj", C.c_int64),
        ("axis_i", C.c_int32), ("axis_j", C.c_int32),
        ("gate", C.c_double * 32),
        ("Dmax", C.c_int32),
        ("eps", C.c_double),
        ("vidal", C.c_int32),
        ("lambda_i", C.c_int64 * MAX_RANK),
        ("lambda_j", C.c_int64 * MAX_RANK),
        ("ti_out", C.c_int64), ("tj_out", C.c_int64), ("s_out", C.c_int64),
        ("kept", C.c_int32),
        ("trunc_err", C.c_double),
    ]


class TneSweepGate(C.Structure):
    _fields_ = [
        ("site_i", C.c_int32), ("site_j", C.c_int32),
        ("axis_i", C.c_int32), ("axis_j", C.c_int32),
        ("bond", C.c_int32),
        ("gate", C.c_double * 32),
        ("lambda_i", C.c_int32 * MAX_RANK),
        ("lambda_j", C.c_int32 * MAX_RANK),
        ("kept", C.c_int32),
        ("trunc_err", C.c_double),
    ]


# every symbol include/tne.h declares: name -> (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "tne_version": (C.c_int, []),
    "tne_ctx_create": (C.c_int, [C.c_int, _P(C.c_void_p)]),
    "tne_ctx_destroy": (C.c_int, [C.c_void_p]),
    "tne_last_error": (C.c_char_p, [C.c_void_p]),
    "tne_sync": (C.c_int, [C.c_void_p]),
    "tne_set_mem_limit": (C.c_int, [C.c_void_p, C.c_int64]),
    "tne_mem_info": (C.c_int, [C.c_void_p, _P(C.c_int64), _P(C.c_int64), _P(C.c_int64)]),
    "tne_stream": (C.c_void_p, [C.c_void_p]),
    "tne_launch_count": (C.c_int64, [C.c_void_p]),
    "tne_last_program_stats": (C.c_int, [C.c_void_p, _P(C.c_double)]),
    "tne_plan_cache_info": (C.c_int, [C.c_void_p, _P(C.c_int64), _P(C.c_int64)]),
    "tne_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "tne_profile_collect": (C.c_int, [C.c_void_p, _P(C.c_double)]),
    "tne_measure_fp64_peak": (C.c_int, [C.c_void_p, _P(C.c_double), _P(C.c_double)]),
    "tne_alloc": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_int64), _P(C.c_int64)]),
    "tne_free": (C.c_int, [C.c_void_p, C.c_int64]),
    "tne_shape": (C.c_int, [C.c_void_p, C.c_int64, _P(C.c_int), _P(C.c_int64)]),
    "tne_upload": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "tne_download": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    "tne_device_ptr": (C.c_void_p, [C.c_void_p, C.c_int64]),
    "tne_einsum": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_int32), _P(C.c_int32), _P(C.c_int64), _P(C.c_int32),
                             _P(C.c_int32), C.c_int, _P(C.c_int64)]),
    "tne_einsum_sized": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_int32), _P(C.c_int32), _P(C.c_int64), _P(C.c_int32),
                                   _P(C.c_int32), C.c_int, C.c_int, _P(C.c_int32), _P(C.c_int64), _P(C.c_int64)]),
    "tne_map": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_double, _P(C.c_int64)]),
    "tne_map2": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int64, _P(C.c_int64)]),
    "tne_reduce_sum": (C.c_int, [C.c_void_p, C.c_int64, _P(C.c_int64)]),
    "tne_fill": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, _P(C.c_int64), _P(C.c_int64)]),
    "tne_slice": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, _P(C.c_int64)]),
    "tne_embed": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, _P(C.c_int64)]),
    "tne_reshape": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, _P(C.c_int64), _P(C.c_int64)]),
    "tne_concat": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_int64), _P(C.c_int64)]),
    "tne_contract_tree": (C.c_int, [C.c_void_p, _P(TneTree), _P(C.c_int64), _P(C.c_int32), _P(C.c_int64)]),
    "tne_expect_terms": (C.c_int, [C.c_void_p, _P(TneTree), _P(C.c_int64), _P(C.c_int32), C.c_int, _P(TneTerm),
                                   C.c_int, _P(C.c_int32), C.c_double, C.c_double, _P(C.c_int64), _P(C.c_int64)]),
    "tne_svd_trunc": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_double, _P(C.c_int64), _P(C.c_int64),
                                _P(C.c_int64), _P(C.c_int32), _P(C.c_double)]),
    "tne_apply_onbond_batched": (C.c_int, [C.c_void_p, C.c_int, _P(TneBondJob)]),
    "tne_tebd_sweep": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_int64), C.c_int, _P(C.c_int64), C.c_int, _P(TneSweepGate),
                                 C.c_int, C.c_double, C.c_int]),
    "tne_tebd_sweeps": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_int64), C.c_int, _P(C.c_int64), C.c_int, _P(TneSweepGate),
                                  C.c_int, C.c_double, C.c_int, C.c_int, _P(C.c_int32), _P(C.c_double)]),
    "tne_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "tne_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "tne_allreduce_sum": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_int64)]),
}


def build(verbose=False):
    """Compile ``csrc/*.cu`` for sm_100a into ``libtne_b200.so`` (in-tree, via the Makefile)."""
    res = subprocess.run(["make", "-j8", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libtne_b200.so failed")
    return LIB_PATH


def lib():
    """The loaded shared library with prototypes set.  Fails loudly if it is not built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libtne_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


class TneError(RuntimeError):
    """Raised for a non-zero ``tne_status`` (the Julia shim raises ErrorException)."""
