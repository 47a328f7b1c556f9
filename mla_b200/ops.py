"""Tensor-level wrappers over the C ABI: torch tensors in, kernels launched on torch's current stream.

torch is used only for device memory and streams; every computation is a libmla_b200 kernel.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GemmArgs, check

ACT_NONE, ACT_RELU, ACT_GELU_ERF, ACT_GELU_TANH, ACT_SILU = 0, 1, 2, 3, 4


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise _lib.MlaError(f"{name}: expected a CUDA tensor (libmla_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.MlaError(f"{name}: expected {dtype}, got {t.dtype}")


def _rowmajor_2d(t: torch.Tensor, name: str) -> int:
    if t.dim() != 2 or t.stride(1) != 1:
        raise _lib.MlaError(f"{name}: expected a 2-D tensor with unit inner stride, got {tuple(t.shape)} / {t.stride()}")
    return t.stride(0)


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
         out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.bfloat16,
         bias: Optional[torch.Tensor] = None, act: int = ACT_NONE, residual: Optional[torch.Tensor] = None,
         pre_act: Optional[torch.Tensor] = None, alpha: float = 1.0, accumulate: bool = False) -> torch.Tensor:
    """C = epilogue(alpha * A_op @ B_op).

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True);  b: [N,K] (b_mn=False, nn.Linear weight) or [K,N] (b_mn=True).
    """
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    lda = _rowmajor_2d(a, "a")
    ldb = _rowmajor_2d(b, "b")
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb:
        raise _lib.MlaError(f"gemm: contraction mismatch {K} vs {Kb}")
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    else:
        if tuple(out.shape) != (M, N):
            raise _lib.MlaError(f"gemm: out shape {tuple(out.shape)} != {(M, N)}")
        out_dtype = out.dtype
    ldc = _rowmajor_2d(out, "out")
    g = GemmArgs()
    g.a, g.b, g.c = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.m, g.n, g.k = M, N, K
    g.lda, g.ldb, g.ldc = lda, ldb, ldc
    g.a_mn_major, g.b_mn_major = int(a_mn), int(b_mn)
    g.c_dtype = {torch.bfloat16: 0, torch.float32: 1}[out_dtype]
    g.accumulate = int(accumulate)
    g.activation = act
    g.alpha = alpha
    if bias is not None:
        _req(bias, torch.bfloat16, "bias")
    g.bias = _ptr(bias)
    if residual is not None:
        _req(residual, torch.bfloat16, "residual")
        g.ldr = _rowmajor_2d(residual, "residual")
    g.residual = _ptr(residual)
    if pre_act is not None:
        _req(pre_act, torch.bfloat16, "pre_act")
        g.ldp = _rowmajor_2d(pre_act, "pre_act")
    g.pre_act = _ptr(pre_act)
    check(_lib.lib().mla_gemm_bf16(C.byref(g), _stream()))
    return out
