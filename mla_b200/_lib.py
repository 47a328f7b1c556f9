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


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("ld_qkv", C.c_int64),
        ("o", C.c_void_p), ("ld_o", C.c_int64), ("lse", C.c_void_p), ("mask", C.c_void_p),
        ("batch", C.c_int32), ("seq", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("scale", C.c_float),
        ("d_o", C.c_void_p), ("delta", C.c_void_p),
        ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p), ("ld_dqkv", C.c_int64),
    ]
