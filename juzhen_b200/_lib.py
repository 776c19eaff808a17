"""ctypes binding of the C-ABI product library (include/jz_b200.h).

The product path has NO CPU fallback: if ``libjz_b200.so`` is missing this module raises, and
every compute entry point fails with JZ_ERR_CUDA when no CUDA device is present.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_size_t, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JZ_B200_LIB") or os.path.join(_HERE, "libjz_b200.so")   # override: an instrumented build (scripts/build_prof_lib.sh)

MG_HANDLE_BYTES = 128
JZ_OK, JZ_ERR_SHAPE, JZ_ERR_CUDA, JZ_ERR_OOM, JZ_ERR_ARG, JZ_ERR_UNSUPPORTED = range(6)

UNARY = {"exp": 0, "log": 1, "tanh": 2, "dtanh": 3, "square": 4, "sqrt": 5, "relu": 6, "drelu": 7}
STEP_AFFINE, STEP_ELEMINV = 100, 101
GEMM_MODES = {"3xtf32": 0, "tf32": 1, "fp32": 2}


class jz_step(Structure):
    _fields_ = [("kind", c_int), ("s1", c_float), ("a", c_float)]


class JzError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"jz error {code}: {msg}")
        self.code = code


class JzShapeError(ValueError):
    """JZ_ERR_SHAPE -- what the reference throws as std::invalid_argument."""


_F = c_void_p  # float* (device or host), passed as integer addresses
_S = c_void_p  # stream

_SIGS = {
    "jz_abi_version": (c_int, []),
    "jz_init": (c_int, [c_int]),
    "jz_shutdown": (c_int, []),
    "jz_last_error": (c_char_p, []),
    "jz_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_size_t)]),
    "jz_sync": (c_int, [_S]),
    "jz_launch_count": (c_uint64, []),
    "jz_set_gemm_mode": (c_int, [c_int]),
    "jz_get_gemm_mode": (c_int, []),
    "jz_gemm_last_path": (c_int, []),
    "jz_malloc": (c_int, [POINTER(c_void_p), c_size_t, _S]),
    "jz_free": (c_int, [_F, _S]),
    "jz_pool_trim": (c_int, []),
    "jz_pool_stats": (c_int, [POINTER(c_size_t)] * 4),
    "jz_memcpy_h2d": (c_int, [_F, _F, c_size_t, _S]),
    "jz_upload": (c_int, [_F, _F, c_size_t, _S]),
    "jz_memcpy_d2h": (c_int, [_F, _F, c_size_t, _S]),
    "jz_memcpy_d2d": (c_int, [_F, _F, c_size_t, _S]),
    "jz_fill": (c_int, [_F, c_size_t, c_float, _S]),
    "jz_copy": (c_int, [_F, _F, c_size_t, _S]),
    "jz_affine": (c_int, [_F, _F, c_size_t, c_float, c_float, _S]),
    "jz_eleminv": (c_int, [_F, _F, c_size_t, c_float, _S]),
    "jz_unary": (c_int, [c_int, _F, _F, c_size_t, _S]),
    "jz_axpby": (c_int, [_F, _F, _F, c_size_t, c_float, c_float, _S]),
    "jz_hadamard": (c_int, [_F, _F, _F, c_size_t, _S]),
    "jz_div": (c_int, [_F, _F, _F, c_size_t, _S]),
    "jz_chain": (c_int, [_F, _F, c_size_t, POINTER(jz_step), c_int, _S]),
    "jz_axpby2d": (c_int, [_F, c_size_t, c_size_t, c_size_t, _F, c_size_t, c_int, _F, c_size_t, c_int,
                           c_float, c_float, _S]),
    "jz_hadamard2d": (c_int, [_F, c_size_t, c_size_t, c_size_t, _F, c_size_t, c_int, _F, c_size_t, c_int, _S]),
    "jz_copy2d": (c_int, [_F, c_size_t, _F, c_size_t, c_size_t, c_size_t, c_int, _S]),
    "jz_sum": (c_int, [_F, _F, c_size_t, c_size_t, c_size_t, c_int, _S]),
    "jz_max": (c_int, [_F, _F, c_size_t, c_size_t, c_size_t, c_int, _S]),
    "jz_softmax_cols": (c_int, [_F, _F, c_size_t, c_size_t, c_size_t, _S]),
    "jz_softmax_ce_grad": (c_int, [_F, _F, _F, c_size_t, c_size_t, c_float, _S]),
    "jz_softmax_ce_grad_scaled": (c_int, [_F, _F, _F, c_size_t, c_size_t, c_float, _S]),
    "jz_nrm2": (c_int, [_F, c_size_t, POINTER(c_float), _S]),
    "jz_add_bcast": (c_int, [_F, _F, c_size_t, c_size_t, _F, c_int, c_float, c_float, _S]),
    "jz_outer": (c_int, [_F, c_size_t, _F, c_size_t, _F, c_size_t, _S]),
    "jz_gemm": (c_int, [c_int, c_int, c_size_t, c_size_t, c_size_t, c_float, _F, c_size_t, _F, c_size_t,
                        c_float, _F, c_size_t, c_int, _S]),
    "jz_gemm_chain": (c_int, [c_int, c_int, c_size_t, c_size_t, c_size_t, c_float, _F, c_size_t, _F, c_size_t,
                              _F, c_size_t, POINTER(jz_step), c_int, c_int, _S]),
    "jz_gemm_bias_chain": (c_int, [c_int, c_int, c_size_t, c_size_t, c_size_t, c_float, _F, c_size_t, _F, c_size_t,
                                   _F, c_size_t, _F, c_int, c_float, c_float, POINTER(jz_step), c_int, c_int, _S]),
    "jz_gemm_chain_bcast": (c_int, [c_int, c_int, c_size_t, c_size_t, c_size_t, c_float, _F, c_size_t, _F, c_size_t,
                                    _F, c_size_t, POINTER(c_void_p), c_int, POINTER(jz_step), c_int, c_int, _S]),
    "jz_gemm_chain_mcast": (c_int, [c_int, c_int, c_size_t, c_size_t, c_size_t, c_float, _F, c_size_t, _F, c_size_t,
                                    _F, c_size_t, _F, POINTER(jz_step), c_int, c_int, _S]),
    "jz_gemm_last_splits": (c_int, []),
    "jz_gemm_last_cluster_split": (c_int, []),
    "jz_gemm_last_walk": (c_int, []),
    "jz_gemm_strided_batched": (c_int, [c_int, c_int, c_size_t, c_size_t, c_size_t, c_float, _F, c_size_t, c_size_t, _F, c_size_t,
                                        c_size_t, c_float, _F, c_size_t, c_size_t, c_size_t, c_int, _S]),
    "jz_mg_block_range": (c_int, [c_size_t, c_int, c_int, POINTER(c_size_t), POINTER(c_size_t)]),
    "jz_mg_export": (c_int, [_F, c_void_p]),
    "jz_mg_import": (c_int, [c_void_p, POINTER(c_void_p)]),
    "jz_mg_release": (c_int, [_F]),
    "jz_mg_barrier": (c_int, [POINTER(c_void_p), c_int, c_int, c_uint32, _S]),
    "jz_mg_gemm_allgather": (c_int, [c_int, c_size_t, c_size_t, c_size_t, c_float, _F, c_size_t, _F, c_size_t,
                                     POINTER(c_void_p), c_int, c_int, POINTER(jz_step), c_int, c_int, _S]),
    "jz_mg_allreduce_sum": (c_int, [_F, POINTER(c_void_p), c_size_t, c_int, c_int, _S]),
    "jz_softmax_rows_batched": (c_int, [_F, _F, c_size_t, c_size_t, c_int, c_float, _S]),
    "jz_causal_mask": (c_int, [_F, c_size_t, c_size_t, c_float, _S]),
    "jz_softmax_rows_backward": (c_int, [_F, _F, _F, c_size_t, c_size_t, c_float, _S]),
    "jz_layernorm_forward": (c_int, [_F, _F, _F, _F, _F, _F, c_size_t, c_size_t, _S]),
    "jz_layernorm_backward": (c_int, [_F, _F, _F, _F, _F, c_size_t, c_size_t, _S]),
    "jz_rand_uniform": (c_int, [_F, c_size_t, c_uint64, c_uint64, _S]),
    "jz_rand_normal": (c_int, [_F, c_size_t, c_uint64, c_uint64, _S]),
    "jz_adam_update": (c_int, [_F, _F, _F, c_size_t] + [c_float] * 6 + [_S]),
    "jz_unary_ulp_sweep": (c_int, [c_int, c_uint32, c_uint32, POINTER(c_uint32), POINTER(c_uint32), _S]),
}

_lib = None


def lib():
    """Load libjz_b200.so (built by ``python -m juzhen_b200.build`` / ``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m juzhen_b200.build` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError = header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def check(rc):
    if rc == JZ_OK:
        return
    msg = lib().jz_last_error().decode(errors="replace")
    if rc == JZ_ERR_SHAPE:
        raise JzShapeError("Matrix dimensions are not compatible" + (f" ({msg})" if msg else ""))
    raise JzError(rc, msg)


def make_steps(steps):
    """steps: list of ('exp'|'log'|..., ) | ('affine', s1, a) | ('eleminv', l)"""
    arr = (jz_step * max(len(steps), 1))()
    for i, st in enumerate(steps):
        kind = st[0]
        if kind == "affine":
            arr[i] = jz_step(STEP_AFFINE, float(st[1]), float(st[2]))
        elif kind == "eleminv":
            arr[i] = jz_step(STEP_ELEMINV, float(st[1]), 0.0)
        else:
            arr[i] = jz_step(UNARY[kind], 0.0, 0.0)
    return arr, len(steps)
