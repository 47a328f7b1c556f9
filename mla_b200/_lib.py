"""ctypes binding of libmla_b200.so (the C ABI declared in include/mla_b200.h).

There is no fallback: if the library is missing or the device is not sm_100 every op raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libmla_b200.so"
_lib = None


class MlaError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("b", C.c_void_p), ("c", C.c_void_p),
        ("m", C.c_int64), ("n", C.c_int64), ("k", C.c_int64),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldc", C.c_int64),
        ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
        ("c_dtype", C.c_int32), ("accumulate", C.c_int32), ("activation", C.c_int32),
        ("alpha", C.c_float),
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("pre_act", C.c_void_p), ("ldp", C.c_int64), ("sched_ws", C.c_void_p),
        ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p), ("rope_seq", C.c_int32), ("rope_cols", C.c_int32),
        ("swiglu_out", C.c_void_p), ("ld_swiglu", C.c_int64),
        ("rope_pos", C.c_void_p),
        ("swiglu_bwd_gu", C.c_void_p), ("ld_swiglu_bwd_gu", C.c_int64),
        ("swiglu_bwd_dgu", C.c_void_p), ("ld_swiglu_bwd_dgu", C.c_int64),
        ("swiglu_bwd_act", C.c_void_p), ("ld_swiglu_bwd_act", C.c_int64),
        ("sumsq", C.c_void_p),
    ]


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises MlaError if it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise MlaError(
                f"{_LIB_PATH} not found: build it with `python -m mla_b200.build` "
                "(there is no CPU or PyTorch fallback for the MLA hot path)")
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.mla_version.restype = C.c_char_p
        _lib.mla_last_error.restype = C.c_char_p
        _lib.mla_launch_count.restype = C.c_int64
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise MlaError(f"libmla_b200 error {rc}: {lib().mla_last_error().decode()}")


def launch_count() -> int:
    return int(lib().mla_launch_count())


def tmap_cache_stats() -> tuple:
    """(hits, misses) of the TMA-descriptor cache inside the library."""
    h, m = C.c_int64(0), C.c_int64(0)
    lib().mla_tmap_cache_stats(C.byref(h), C.byref(m))
    return int(h.value), int(m.value)


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("ld_qkv", C.c_int64),
        ("o", C.c_void_p), ("ld_o", C.c_int64), ("lse", C.c_void_p), ("mask", C.c_void_p),
        ("batch", C.c_int32), ("seq", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("scale", C.c_float),
        ("d_o", C.c_void_p), ("delta", C.c_void_p),
        ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p), ("ld_dqkv", C.c_int64),
    ]


class MhaArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
        ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64),
        ("o", C.c_void_p), ("ldo", C.c_int64), ("lse", C.c_void_p),
        ("keep_mask", C.c_void_p), ("keep_scale", C.c_float),
        ("batch", C.c_int32), ("heads", C.c_int32), ("len_q", C.c_int32), ("len_k", C.c_int32), ("head_dim", C.c_int32),
        ("scale", C.c_float),
        ("d_o", C.c_void_p), ("ld_do", C.c_int64), ("delta", C.c_void_p),
        ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
        ("ld_dq", C.c_int64), ("ld_dk", C.c_int64), ("ld_dv", C.c_int64),
    ]


class GenImageArgs(C.Structure):
    _fields_ = [
        ("cur", C.c_void_p), ("nxt", C.c_void_p),
        ("cur_stride_b", C.c_int64), ("cur_stride_c", C.c_int64), ("nxt_stride_b", C.c_int64), ("nxt_stride_c", C.c_int64),
        ("n_images", C.c_int32), ("width", C.c_int32), ("patch", C.c_int32), ("grid", C.c_int32), ("n_patches", C.c_int32),
        ("delta_raw", C.c_void_p), ("ld_delta", C.c_int64), ("ao_raw", C.c_void_p), ("ld_ao", C.c_int64),
        ("roi", C.c_void_p),
        ("delta_clip", C.c_float), ("max_shift", C.c_float), ("gen_weight", C.c_float),
        ("blended", C.c_void_p), ("delta_all", C.c_void_p), ("alpha_all", C.c_void_p), ("offset_all", C.c_void_p),
        ("sums", C.c_void_p), ("losses", C.c_void_p), ("coef", C.c_void_p),
        ("grad_scale", C.c_void_p), ("d_delta_raw", C.c_void_p), ("d_ao_raw", C.c_void_p),
    ]


class GemvArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("w", C.c_void_p), ("out", C.c_void_p), ("residual", C.c_void_p), ("ln_weight", C.c_void_p),
        ("m", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
        ("ldx", C.c_int64), ("ldw", C.c_int64), ("ldo", C.c_int64), ("ldr", C.c_int64),
        ("prologue", C.c_int32), ("eps", C.c_float),
    ]


class DecodeStackArgs(C.Structure):
    """mla_decode_stack_args (include/mla_b200.h)."""
    _fields_ = [
        ("w_qkv", C.c_void_p), ("w_o", C.c_void_p), ("w_gate_up", C.c_void_p), ("w_down", C.c_void_p),
        ("ln1", C.c_void_p), ("ln2", C.c_void_p), ("kv_cache", C.c_void_p),
        ("x", C.c_void_p), ("qkv", C.c_void_p), ("ctx", C.c_void_p), ("x_mid", C.c_void_p), ("gate_up", C.c_void_p),
        ("cos_t", C.c_void_p), ("sin_t", C.c_void_p), ("workspace", C.c_void_p),
        ("layers", C.c_int32), ("batch", C.c_int32), ("n", C.c_int32), ("prefix", C.c_int32), ("heads", C.c_int32),
        ("head_dim", C.c_int32), ("ffn", C.c_int32),
        ("eps", C.c_float), ("scale", C.c_float),
        ("trace", C.c_void_p),
    ]
