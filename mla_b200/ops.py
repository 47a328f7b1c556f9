"""Tensor-level wrappers over the C ABI: torch tensors in, kernels launched on torch's current stream.

torch is used only for device memory and streams; every computation is a libmla_b200 kernel.
"""
from __future__ import annotations

import ctypes as C
import os
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
         pre_act: Optional[torch.Tensor] = None, alpha: float = 1.0, accumulate: bool = False,
         rope: Optional[tuple] = None, swiglu_out: Optional[torch.Tensor] = None,
         store_c: bool = True, swiglu_bwd: Optional[tuple] = None, sumsq: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C = epilogue(alpha * A_op @ B_op).

    a: [M,K] (a_mn=False) or [K,M] (a_mn=True);  b: [N,K] (b_mn=False, nn.Linear weight) or [K,N] (b_mn=True).
    rope = (cos [S,64] bf16, sin [S,64] bf16, S, n_cols[, pos int32 [M]]): rotate the leading n_cols columns (head_dim 128)
    in the epilogue; position = row % S, or pos[row] when the table is given (shared-prefix layout).
    swiglu_out (see include/mla_b200.h): bf16 [M, N/2] receiving SwiGLU of the [gate | up] projection in
    the epilogue; with store_c=False the projection itself is not written (returns None).
    swiglu_bwd = (gu [M, 2N] bf16, dgu [M, 2N] bf16, act [M, N] bf16 or None): the product is d_act of a SwiGLU whose
    gate|up is gu; the epilogue writes d(gate|up) into dgu and the re-materialised act, and nothing else (returns None).
    """
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    lda = _rowmajor_2d(a, "a")
    ldb = _rowmajor_2d(b, "b")
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb:
        raise _lib.MlaError(f"gemm: contraction mismatch {K} vs {Kb}")
    if swiglu_bwd is not None or (swiglu_out is not None and not store_c):
        out, ldc = None, N
    else:
        if out is None:
            out = torch.empty((M, N), dtype=out_dtype, device=a.device)
        else:
            if tuple(out.shape) != (M, N):
                raise _lib.MlaError(f"gemm: out shape {tuple(out.shape)} != {(M, N)}")
            out_dtype = out.dtype
        ldc = _rowmajor_2d(out, "out")
    g = GemmArgs()
    g.a, g.b, g.c = a.data_ptr(), b.data_ptr(), (out.data_ptr() if out is not None else None)
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
    if rope is not None:
        cos_t, sin_t, seq, ncols = rope[:4]
        if len(rope) > 4 and rope[4] is not None:
            _req(rope[4], torch.int32, "rope pos")
            if rope[4].numel() != M or not rope[4].is_contiguous():
                raise _lib.MlaError("gemm: rope position table must be a contiguous int32 [M]")
            g.rope_pos = rope[4].data_ptr()
        _req(cos_t, torch.bfloat16, "rope cos")
        _req(sin_t, torch.bfloat16, "rope sin")
        if tuple(cos_t.shape) != (seq, 64) or tuple(sin_t.shape) != (seq, 64) or not (cos_t.is_contiguous() and sin_t.is_contiguous()):
            raise _lib.MlaError("gemm: fused RoPE needs contiguous [seq, 64] tables (head_dim 128)")
        g.rope_cos, g.rope_sin, g.rope_seq, g.rope_cols = cos_t.data_ptr(), sin_t.data_ptr(), seq, ncols
    if swiglu_out is not None:
        _req(swiglu_out, torch.bfloat16, "swiglu_out")
        if tuple(swiglu_out.shape) != (M, N // 2):
            raise _lib.MlaError(f"gemm: swiglu_out shape {tuple(swiglu_out.shape)} != {(M, N // 2)}")
        g.swiglu_out, g.ld_swiglu = swiglu_out.data_ptr(), _rowmajor_2d(swiglu_out, "swiglu_out")
    if swiglu_bwd is not None:
        gu_t, dgu_t, act_t = swiglu_bwd
        _req(gu_t, torch.bfloat16, "swiglu_bwd gu")
        _req(dgu_t, torch.bfloat16, "swiglu_bwd dgu")
        if tuple(gu_t.shape) != (M, 2 * N) or tuple(dgu_t.shape) != (M, 2 * N):
            raise _lib.MlaError(f"gemm: swiglu_bwd gu / dgu must be [{M}, {2 * N}]")
        g.swiglu_bwd_gu, g.ld_swiglu_bwd_gu = gu_t.data_ptr(), _rowmajor_2d(gu_t, "swiglu_bwd gu")
        g.swiglu_bwd_dgu, g.ld_swiglu_bwd_dgu = dgu_t.data_ptr(), _rowmajor_2d(dgu_t, "swiglu_bwd dgu")
        if act_t is not None:
            _req(act_t, torch.bfloat16, "swiglu_bwd act")
            if tuple(act_t.shape) != (M, N):
                raise _lib.MlaError(f"gemm: swiglu_bwd act must be [{M}, {N}]")
            g.swiglu_bwd_act, g.ld_swiglu_bwd_act = act_t.data_ptr(), _rowmajor_2d(act_t, "swiglu_bwd act")
    if sumsq is not None:       # fp32 outputs: *sumsq += sum of squares of what this launch writes
        _req(sumsq, torch.float32, "sumsq")
        g.sumsq = sumsq.data_ptr()
    if DYNAMIC_TILES["on"]:
        g.sched_ws = _sched_ws(a.device).data_ptr()
    check(_lib.lib().mla_gemm_bf16(C.byref(g), _stream()))
    return out


# Dynamic tile scheduling of the persistent GEMM (see include/mla_b200.h: sched_ws).  One 8-byte counter per
# (device, stream).  On by default: measured on one B200 (tools/ab_inproc.py, interleaved) 678.9 -> 660.1 ms per step,
# and under DDP the NCCL all-reduces share the SMs with the GEMMs, where a static order stalls whole waves.
DYNAMIC_TILES = {"on": __import__("os").environ.get("MLA_DYNAMIC_TILES", "1") == "1"}
_SCHED_WS: dict = {}


def _sched_ws(device) -> torch.Tensor:
    key = (str(device), torch.cuda.current_stream().cuda_stream)
    t = _SCHED_WS.get(key)
    if t is None:
        t = torch.zeros(2, dtype=torch.int32, device=device)
        _SCHED_WS[key] = t
    return t


# ---------------------------------------------------------------------------------------------- row kernels
def _p(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(None if t is None else t.data_ptr())


def rmsnorm_fwd(x: torch.Tensor, w: torch.Tensor, eps: float, *, mode: int = 0,
                rstd: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: [..., h] bf16 (last dim contiguous, rows uniformly strided) -> same shape."""
    _req(x, torch.bfloat16, "x")
    _req(w, torch.bfloat16, "w")
    h = x.shape[-1]
    x2 = x.reshape(-1, h)
    if out is None:
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    o2 = out.view(-1, h)
    check(_lib.lib().mla_rmsnorm_fwd(_p(x2), _p(w), _p(o2), _p(rstd), C.c_int64(x2.shape[0]), C.c_int32(h),
                                     C.c_int64(x2.stride(0)), C.c_int64(o2.stride(0)), C.c_float(eps),
                                     C.c_int32(mode), _stream()))
    return out


def rmsnorm_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, eps: float, *,
                dres: Optional[torch.Tensor] = None, dw: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Returns dx (+ dres); accumulates the weight gradient into dw (f32 [h]) when given."""
    for t, n in ((dy, "dy"), (x, "x"), (w, "w")):
        _req(t, torch.bfloat16, n)
    h = x.shape[-1]
    dy2, x2 = dy.reshape(-1, h), x.reshape(-1, h)
    if not (dy2.is_contiguous() and x2.is_contiguous()):
        raise _lib.MlaError("rmsnorm_bwd: dy and x must be contiguous")
    if dres is not None:
        _req(dres, torch.bfloat16, "dres")
        if not dres.is_contiguous():
            raise _lib.MlaError("rmsnorm_bwd: dres must be contiguous")
    if dw is not None:
        _req(dw, torch.float32, "dw")
    dx = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().mla_rmsnorm_bwd(_p(dy2), _p(x2), _p(w), _p(dres), _p(dx), _p(dw), C.c_int64(x2.shape[0]),
                                     C.c_int32(h), C.c_float(eps), _stream()))
    return dx


def rope_(buf: torch.Tensor, col0: int, heads: int, head_dim: int, seq: int, cos_t: torch.Tensor,
          sin_t: torch.Tensor, transpose: bool = False) -> None:
    """Rotate, in place, `heads` heads starting at column col0 of the 2-D buffer buf [tokens, ld]."""
    _req(buf, torch.bfloat16, "buf")
    ld = _rowmajor_2d(buf, "buf")
    base = buf.data_ptr() + 2 * col0
    check(_lib.lib().mla_rope_inplace(C.c_void_p(base), _p(cos_t), _p(sin_t), C.c_int64(buf.shape[0]), C.c_int32(seq),
                                      C.c_int32(heads), C.c_int32(head_dim), C.c_int64(ld), C.c_int32(int(transpose)),
                                      _stream()))


def swiglu_fwd(gu: torch.Tensor) -> torch.Tensor:
    _req(gu, torch.bfloat16, "gu")
    rows, f2 = gu.shape
    if not gu.is_contiguous():
        raise _lib.MlaError("swiglu_fwd: gu must be contiguous")
    out = torch.empty((rows, f2 // 2), dtype=torch.bfloat16, device=gu.device)
    check(_lib.lib().mla_swiglu_fwd(_p(gu), _p(out), C.c_int64(rows), C.c_int32(f2 // 2), _stream()))
    return out


def swiglu_bwd(dact: torch.Tensor, gu: torch.Tensor) -> torch.Tensor:
    _req(dact, torch.bfloat16, "dact")
    _req(gu, torch.bfloat16, "gu")
    rows, f2 = gu.shape
    if not (gu.is_contiguous() and dact.is_contiguous()):
        raise _lib.MlaError("swiglu_bwd: inputs must be contiguous")
    dgu = torch.empty_like(gu)
    check(_lib.lib().mla_swiglu_bwd(_p(dact), _p(gu), _p(dgu), C.c_int64(rows), C.c_int32(f2 // 2), _stream()))
    return dgu


def swiglu_bwd_act(dact: torch.Tensor, gu: torch.Tensor):
    """(d_gu, act): swiglu_bwd fused with the recompute of act = swiglu_fwd(gu)."""
    _req(dact, torch.bfloat16, "dact")
    _req(gu, torch.bfloat16, "gu")
    rows, f2 = gu.shape
    if not (gu.is_contiguous() and dact.is_contiguous()):
        raise _lib.MlaError("swiglu_bwd_act: inputs must be contiguous")
    dgu = torch.empty_like(gu)
    act = torch.empty((rows, f2 // 2), dtype=torch.bfloat16, device=gu.device)
    check(_lib.lib().mla_swiglu_bwd_act(_p(dact), _p(gu), _p(dgu), _p(act), C.c_int64(rows), C.c_int32(f2 // 2),
                                        _stream()))
    return dgu, act


# ---------------------------------------------------------------------------------------------- attention
# Which kernel serves head_dim 128: "sm100" = tcgen05/TMEM/TMA (attention_sm100.cu), "mma" = mma.sync (attention.cu).
import os as _os

_attn_default = _os.environ.get("MLA_ATTN_IMPL", "sm100")      # A/B switch for benchmarking: "sm100" | "mma"
# backward generations of the tcgen05 path: "sm100" = attention_bwd_sm100.cu, "sm100v2" = the pipelined kernel
# (attention_bwd2_sm100.cu), which can also apply the transposed RoPE in its epilogue (attn_bwd(..., rope=(cos, sin)))
ATTN_IMPL = {"fwd": _attn_default,
             "bwd": _os.environ.get("MLA_ATTN_BWD", "sm100v2") if _attn_default == "sm100" else _attn_default}


def _attn_args(qkv: torch.Tensor, B: int, S: int, H: int, D: int, mask: Optional[torch.Tensor]) -> "_lib.AttnArgs":
    _req(qkv, torch.bfloat16, "qkv")
    ld = _rowmajor_2d(qkv, "qkv")
    if qkv.shape[0] != B * S or qkv.shape[1] != 3 * H * D:
        raise _lib.MlaError(f"attention: qkv shape {tuple(qkv.shape)} != [{B * S}, {3 * H * D}]")
    a = _lib.AttnArgs()
    base = qkv.data_ptr()
    a.q, a.k, a.v = base, base + 2 * H * D, base + 4 * H * D
    a.ld_qkv = ld
    a.batch, a.seq, a.heads, a.head_dim = B, S, H, D
    a.scale = D ** -0.5
    if mask is not None:
        if mask.dtype not in (torch.uint8, torch.bool) or not mask.is_contiguous() or tuple(mask.shape) != (B, S):
            raise _lib.MlaError("attention: mask must be a contiguous uint8/bool [B,S] tensor")
        a.mask = mask.data_ptr()
    return a


def _grouped_args(grouped, B: int, name: str):
    """(prefix_len int32 [B], group) of a shared-prefix layout -> (pointer, group) after validation."""
    if grouped is None:
        return None, 0
    prefix_len, group = grouped
    _req(prefix_len, torch.int32, f"{name}: prefix_len")
    if prefix_len.numel() != B or not prefix_len.is_contiguous() or group <= 0:
        raise _lib.MlaError(f"{name}: shared-prefix layout needs a contiguous int32 prefix_len [B] and group > 0")
    return prefix_len, int(group)


def attn_fwd(qkv: torch.Tensor, B: int, S: int, H: int, D: int, mask: Optional[torch.Tensor] = None,
             grouped: Optional[tuple] = None):
    """qkv: fused projection [B*S, 3*H*D] (q | k | v, RoPE already applied). Returns (ctx [B*S, H*D], lse [B,H,S]).
    grouped = (prefix_len int32 [B], group): shared-prefix layout — rows from prefix_len[b] on are groups of `group` rows
    that see the whole prefix and, causally, only their own group (tcgen05 kernel, head_dim 128)."""
    a = _attn_args(qkv, B, S, H, D, mask)
    ctx = torch.empty((B * S, H * D), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((B, H, S), dtype=torch.float32, device=qkv.device)
    a.o, a.ld_o, a.lse = ctx.data_ptr(), ctx.stride(0), lse.data_ptr()
    pl, group = _grouped_args(grouped, B, "attention fwd")
    if D == 128 and ATTN_IMPL["fwd"] == "sm100":
        # tcgen05 / TMEM / TMA kernel (head_dim 128); the mma.sync kernel serves head_dim 32 / 64
        check(_lib.lib().mla_attn_fwd_sm100_grouped(C.c_void_p(qkv.data_ptr()), C.c_int64(a.ld_qkv), C.c_void_p(a.o),
                                                    C.c_int64(a.ld_o), C.c_void_p(a.lse), C.c_void_p(a.mask), _p(pl),
                                                    C.c_int32(group), C.c_int32(B), C.c_int32(S), C.c_int32(H),
                                                    C.c_float(a.scale), _stream()))
    else:
        if group:
            raise _lib.MlaError("attention fwd: the shared-prefix layout needs the tcgen05 kernel (head_dim 128)")
        check(_lib.lib().mla_attn_fwd(C.byref(a), _stream()))
    return ctx, lse


def attn_bwd_fuses_rope(D: int) -> bool:
    """True when attn_bwd can apply the transposed RoPE itself (the pipelined tcgen05 kernel, head_dim 128)."""
    return D == 128 and ATTN_IMPL["bwd"] == "sm100v2"


def attn_bwd(dctx: torch.Tensor, qkv: torch.Tensor, ctx: torch.Tensor, lse: torch.Tensor, B: int, S: int, H: int,
             D: int, mask: Optional[torch.Tensor] = None, rope: Optional[tuple] = None,
             grouped: Optional[tuple] = None, rope_pos: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Returns d_qkv [B*S, 3*H*D] (dq | dk | dv) — gradients w.r.t. the post-RoPE q,k and v; with rope = (cos, sin)
    (bf16 [S, 64]; only where attn_bwd_fuses_rope) w.r.t. the PRE-RoPE q and k: the transposed rotation is applied in
    the kernel's epilogue."""
    a = _attn_args(qkv, B, S, H, D, mask)
    _req(dctx, torch.bfloat16, "dctx")
    if not (dctx.is_contiguous() and ctx.is_contiguous()):
        raise _lib.MlaError("attention bwd: ctx / dctx must be contiguous")
    if rope is not None and not attn_bwd_fuses_rope(D):
        raise _lib.MlaError("attention bwd: fused RoPE transpose needs the pipelined tcgen05 kernel (head_dim 128)")
    pl, group = _grouped_args(grouped, B, "attention bwd")
    if group and not attn_bwd_fuses_rope(D):
        raise _lib.MlaError("attention bwd: the shared-prefix layout needs the pipelined tcgen05 kernel (head_dim 128)")
    if rope_pos is not None:
        _req(rope_pos, torch.int32, "rope_pos")
        if rope_pos.numel() != B * S or not rope_pos.is_contiguous():
            raise _lib.MlaError("attention bwd: rope_pos must be a contiguous int32 [B*S]")
    dqkv = torch.empty_like(qkv, memory_format=torch.contiguous_format)
    if D == 128 and ATTN_IMPL["bwd"] in ("sm100", "sm100v2"):
        lib = _lib.lib()
        lib.mla_attn_bwd_sm100_workspace.restype = C.c_size_t
        ws = torch.empty(lib.mla_attn_bwd_sm100_workspace(C.c_int32(B), C.c_int32(S), C.c_int32(H)) // 4,
                         dtype=torch.float32, device=qkv.device)
        if ATTN_IMPL["bwd"] == "sm100v2":
            cos_t = sin_t = None
            if rope is not None:
                cos_t, sin_t = rope
                _req(cos_t, torch.bfloat16, "rope cos")
                _req(sin_t, torch.bfloat16, "rope sin")
                if tuple(cos_t.shape) != (S, 64) or tuple(sin_t.shape) != (S, 64) or not (
                        cos_t.is_contiguous() and sin_t.is_contiguous()):
                    raise _lib.MlaError("attention bwd: fused RoPE needs contiguous [seq, 64] tables")
            check(lib.mla_attn_bwd2_sm100_grouped(C.c_void_p(qkv.data_ptr()), C.c_int64(a.ld_qkv), _p(ctx), _p(dctx),
                                                  C.c_int64(ctx.stride(0)), _p(lse), C.c_void_p(a.mask), _p(dqkv),
                                                  C.c_int64(dqkv.stride(0)), _p(ws), _p(cos_t), _p(sin_t), _p(rope_pos),
                                                  _p(pl), C.c_int32(group), C.c_int32(B), C.c_int32(S), C.c_int32(H),
                                                  C.c_float(a.scale), _stream()))
            return dqkv
        check(lib.mla_attn_bwd_sm100(C.c_void_p(qkv.data_ptr()), C.c_int64(a.ld_qkv), _p(ctx), _p(dctx),
                                     C.c_int64(ctx.stride(0)), _p(lse), C.c_void_p(a.mask), _p(dqkv),
                                     C.c_int64(dqkv.stride(0)), _p(ws), C.c_int32(B), C.c_int32(S), C.c_int32(H),
                                     C.c_float(a.scale), _stream()))
        return dqkv
    delta = torch.empty_like(lse)
    a.o, a.ld_o, a.lse = ctx.data_ptr(), ctx.stride(0), lse.data_ptr()
    a.d_o, a.delta = dctx.data_ptr(), delta.data_ptr()
    base = dqkv.data_ptr()
    a.dq, a.dk, a.dv = base, base + 2 * H * D, base + 4 * H * D
    a.ld_dqkv = dqkv.stride(0)
    check(_lib.lib().mla_attn_bwd(C.byref(a), _stream()))
    return dqkv


# ---------------------------------------------------------------------------------------------- casts / caches
def cast_bf16(src: torch.Tensor, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 -> bf16 copy (contiguous)."""
    _req(src, torch.float32, "src")
    s = src.detach()
    if not s.is_contiguous():
        raise _lib.MlaError("cast_bf16: src must be contiguous")
    if dst is None:
        dst = torch.empty(s.shape, dtype=torch.bfloat16, device=s.device)
    elif not dst.is_contiguous() or dst.numel() != s.numel():
        raise _lib.MlaError("cast_bf16: dst must be contiguous with the same number of elements")
    check(_lib.lib().mla_cast_f32_bf16(_p(s), _p(dst), C.c_int64(s.numel()), _stream()))
    return dst


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


_BF16_CACHE: dict = {}


def bf16_of(param: torch.Tensor, pad2d: bool = False) -> torch.Tensor:
    """bf16 compute copy of an fp32 parameter, cached on (storage, version).  With pad2d a [N,K] matrix is copied
    into a zero-padded [pad8(N), pad8(K)] buffer so that it satisfies the TMA 16-byte pitch rule."""
    if param.dtype == torch.bfloat16 and not pad2d:
        return param.detach()
    # keyed on the storage (kept alive by the entry, so its address cannot be recycled under us) + view geometry
    st = param.untyped_storage()
    key = (st._cdata, param.storage_offset(), tuple(param.shape), tuple(param.stride()), pad2d)
    tag = (param._version, str(param.device), param.dtype)
    hit = _BF16_CACHE.get(key)
    if hit is not None and hit[0] == tag:
        return hit[1]
    p = param.detach()
    if p.dtype != torch.float32:
        p = p.float()
    if pad2d and p.dim() == 2 and (p.shape[0] % 8 or p.shape[1] % 8):
        n8, k8 = _pad8(p.shape[0]), _pad8(p.shape[1])
        out = hit[1] if hit is not None and tuple(hit[1].shape) == (n8, k8) else torch.empty(
            (n8, k8), dtype=torch.bfloat16, device=p.device)
        p = p.contiguous()
        check(_lib.lib().mla_cast_pad_f32_bf16(_p(p), _p(out), C.c_int64(n8), C.c_int64(k8), C.c_int64(p.shape[0]),
                                               C.c_int64(p.shape[1]), C.c_int64(p.stride(0)), _stream()))
    else:
        out = hit[1] if hit is not None and hit[1].shape == p.shape else None
        out = cast_bf16(p.contiguous(), out)
    _BF16_CACHE[key] = (tag, out, st)
    return out


def clear_bf16_cache() -> None:
    _BF16_CACHE.clear()


def pad_cols_bf16(x: torch.Tensor, k8: int) -> torch.Tensor:
    """[M,K] f32/bf16 activations -> bf16 [M,k8] zero padded (inputs with K not a multiple of 8: 7-dof actions)."""
    M, K = x.shape
    xf = x.detach().float().contiguous()
    out = torch.empty((M, k8), dtype=torch.bfloat16, device=x.device)
    check(_lib.lib().mla_cast_pad_f32_bf16(_p(xf), _p(out), C.c_int64(M), C.c_int64(k8), C.c_int64(M), C.c_int64(K),
                                           C.c_int64(K), _stream()))
    return out


def act_bwd(dy: torch.Tensor, pre: torch.Tensor, act: int) -> torch.Tensor:
    dy = dy.contiguous()
    dx = torch.empty_like(dy)
    check(_lib.lib().mla_act_bwd(_p(dy), _p(pre), _p(dx), C.c_int64(dy.numel()), C.c_int32(act), _stream()))
    return dx


def colsum(x: torch.Tensor, n_cols: int) -> torch.Tensor:
    out = torch.zeros(n_cols, dtype=torch.float32, device=x.device)
    check(_lib.lib().mla_colsum_bf16(_p(x), _p(out), C.c_int64(x.shape[0]), C.c_int32(n_cols), C.c_int64(x.stride(0)),
                                     _stream()))
    return out


def add_bf16(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    a, b = a.contiguous(), b.contiguous()
    out = torch.empty_like(a)
    check(_lib.lib().mla_add_bf16(_p(a), _p(b), _p(out), C.c_int64(a.numel()), _stream()))
    return out


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) on the tcgen05 GEMM; x bf16 [M,K], W/b fp32 parameters (bf16 compute copies cached).
    Gradients: dx bf16, dW/db fp32 returned to autograd (these are the small non-decoder modules)."""

    @staticmethod
    def forward(ctx, x, weight, bias, act):
        M, K = x.shape
        N = weight.shape[0]
        padded = bool(K % 8 or N % 8)
        w = bf16_of(weight, pad2d=True)
        N8, K8 = w.shape
        xin = x
        if K8 != K:
            xin = torch.zeros((M, K8), dtype=torch.bfloat16, device=x.device)
            xin[:, :K] = x
        b = None
        if bias is not None:
            b = bf16_of(bias)
            if N8 != N:
                bp = torch.zeros(N8, dtype=torch.bfloat16, device=x.device)
                bp[:N] = b
                b = bp
        pre = torch.empty((M, N8), dtype=torch.bfloat16, device=x.device) if act != ACT_NONE else None
        y = gemm(xin, w, bias=b, act=act, pre_act=pre)
        ctx.save_for_backward(xin, pre)
        ctx.weight, ctx.act, ctx.has_bias, ctx.dims = weight, act, bias is not None, (M, N, K, N8, K8)
        ctx.x_needs = x.requires_grad
        return y[:, :N] if padded and N8 != N else y

    @staticmethod
    def backward(ctx, dy):
        xin, pre = ctx.saved_tensors
        M, N, K, N8, K8 = ctx.dims
        if N8 != N:
            d = torch.zeros((M, N8), dtype=torch.bfloat16, device=dy.device)
            d[:, :N] = dy
            dy = d
        else:
            dy = dy.contiguous()
        if ctx.act != ACT_NONE:
            dy = act_bwd(dy, pre, ctx.act)
        w = bf16_of(ctx.weight, pad2d=True)
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dy, w, b_mn=True)
            if K8 != K:
                dx = dx[:, :K]
        if ctx.needs_input_grad[1]:
            dW = gemm(dy, xin, a_mn=True, b_mn=True, out_dtype=torch.float32)
            if N8 != N or K8 != K:
                dW = dW[:N, :K].contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy, N8)[:N]
        return dx, dW, db, None


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE):
    """nn.Linear (+activation) over the last dim; x any leading shape, bf16."""
    lead = x.shape[:-1]
    y = LinearFn.apply(x.reshape(-1, x.shape[-1]), weight, bias, act)
    return y.reshape(*lead, weight.shape[0])


class RMSNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, eps, mode=0):
        w = bf16_of(weight)
        ctx.save_for_backward(x)
        ctx.weight, ctx.eps, ctx.mode = weight, eps, mode
        return rmsnorm_fwd(x, w, eps, mode=mode)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        if ctx.mode != 0:
            raise _lib.MlaError("RMSNorm backward is implemented for the mean-square mode only")
        dw = torch.zeros(x.shape[-1], dtype=torch.float32, device=x.device)
        dx = rmsnorm_bwd(dy.contiguous(), x.contiguous(), bf16_of(ctx.weight), ctx.eps, dw=dw)
        return dx, dw, None, None


# ---------------------------------------------------------------------------------------------- gather / scatter
def gather_rows(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """src bf16 [R,h]; idx int32/int64 [n] (negative -> zero row) -> [n,h]."""
    _req(src, torch.bfloat16, "src")
    lds = _rowmajor_2d(src, "src")
    n, h = idx.numel(), src.shape[1]
    dst = torch.empty((n, h), dtype=torch.bfloat16, device=src.device)
    check(_lib.lib().mla_gather_rows(_p(src), _p(idx), _p(dst), C.c_int64(n), C.c_int32(h), C.c_int64(lds),
                                     C.c_int32(int(idx.dtype == torch.int64)), _stream()))
    return dst


class GatherRowsFn(torch.autograd.Function):
    """dst = src[idx] with an injective int32 idx (sequence splice); backward is the inverse permutation."""

    @staticmethod
    def forward(ctx, src, idx):
        ctx.save_for_backward(idx)
        ctx.rows = src.shape[0]
        return gather_rows(src, idx)

    @staticmethod
    def backward(ctx, d):
        (idx,) = ctx.saved_tensors
        d = d.contiguous()
        dsrc = torch.zeros((ctx.rows, d.shape[1]), dtype=torch.bfloat16, device=d.device)
        check(_lib.lib().mla_scatter_rows(_p(dsrc), _p(idx), _p(d), C.c_int64(d.shape[0]), C.c_int32(d.shape[1]),
                                          C.c_int64(dsrc.stride(0)), _stream()))
        return dsrc, None


class EmbeddingFn(torch.autograd.Function):
    """nn.Embedding lookup from the bf16 compute copy; backward scatter-adds into an fp32 gradient."""

    @staticmethod
    def forward(ctx, ids, weight, padding_idx):
        ctx.save_for_backward(ids)
        ctx.weight, ctx.padding_idx = weight, padding_idx
        return gather_rows(bf16_of(weight), ids.reshape(-1).contiguous())

    @staticmethod
    def backward(ctx, d):
        (ids,) = ctx.saved_tensors
        w = ctx.weight
        d = d.contiguous()
        g = torch.zeros(w.shape, dtype=torch.float32, device=d.device)
        pad = -1 if ctx.padding_idx is None else int(ctx.padding_idx)
        check(_lib.lib().mla_embedding_bwd(_p(g), _p(ids.reshape(-1).contiguous()), _p(d), C.c_int64(d.shape[0]),
                                           C.c_int32(d.shape[1]), C.c_int64(pad), _stream()))
        return None, g, None


class MSEFn(torch.autograd.Function):
    """mean((pred - target)^2) with bf16 pred and fp32 target -> fp32 scalar (models/mla/model_mla.py:215)."""

    @staticmethod
    def forward(ctx, pred, target):
        pred = pred.contiguous()
        target = target.contiguous()
        ctx.save_for_backward(pred, target)
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
        check(_lib.lib().mla_mse_fwd(_p(pred), _p(target), _p(loss), C.c_int64(pred.numel()), _stream()))
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        gs = g.reshape(1).float().contiguous()
        dp = torch.empty_like(pred)
        check(_lib.lib().mla_mse_bwd(_p(pred), _p(target), _p(gs), _p(dp), C.c_int64(pred.numel()), _stream()))
        return dp, None


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    """nn.LayerNorm over the last dim; x bf16 [rows,h] contiguous, w/b fp32. Forward only (frozen tokenizer)."""
    _req(x, torch.bfloat16, "x")
    y = torch.empty_like(x)
    h = x.shape[-1]
    check(_lib.lib().mla_layernorm_fwd(_p(x), _p(w.detach()), _p(b.detach()), _p(y), C.c_int64(x.numel() // h),
                                       C.c_int32(h), C.c_float(eps), _stream()))
    return y


def q_sample(a: torch.Tensor, noise: torch.Tensor, t: torch.Tensor, sqrt_ac: torch.Tensor, sqrt_1mac: torch.Tensor):
    a, noise = a.contiguous(), noise.contiguous()
    out = torch.empty_like(a)
    check(_lib.lib().mla_q_sample(_p(a), _p(noise), _p(t), _p(sqrt_ac), _p(sqrt_1mac), _p(out), C.c_int64(a.numel()),
                                  C.c_int32(a.numel() // a.shape[0]), _stream()))
    return out


class CrossEntropyFn(torch.autograd.Function):
    """Shifted CE of bf16 logits [B*S, V] against labels [B, S] (modeling_llama.py:1258-1269): fp32 math, mean over
    the rows whose next-token label is not -100.  Backward reuses the logits buffer for d(logits)."""

    @staticmethod
    def forward(ctx, logits, labels):
        B, S = labels.shape
        V = logits.shape[1]
        labels = labels.contiguous()
        lse = torch.empty(B * S, dtype=torch.float32, device=logits.device)
        acc = torch.empty(2, dtype=torch.float32, device=logits.device)
        loss = torch.empty(1, dtype=torch.float32, device=logits.device)
        check(_lib.lib().mla_ce_fwd(_p(logits), C.c_int64(logits.stride(0)), _p(labels), C.c_int64(B * S), C.c_int32(S),
                                    C.c_int32(V), _p(lse), _p(acc), _p(loss), _stream()))
        ctx.save_for_backward(logits, labels, lse, acc)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        logits, labels, lse, acc = ctx.saved_tensors
        B, S = labels.shape
        gs = g.reshape(1).float().contiguous()
        d = logits.detach().clone()
        check(_lib.lib().mla_ce_bwd(_p(d), C.c_int64(d.stride(0)), _p(labels), C.c_int64(B * S), C.c_int32(S),
                                    C.c_int32(logits.shape[1]), _p(lse), _p(acc), _p(gs), _stream()))
        return d, None


# ---------------------------------------------------------------------------------------------- inference (decode)
# Skinny linears of the denoise step on the tensor cores (csrc/skinny_sm100.cu, swap-AB tcgen05) instead of the CUDA-core
# weight-streaming kernels (csrc/decode.cu).  Correct and tested, but measured slower end to end (4.70 vs 3.60 ms per
# DDIM step at 2 rows, 11.2 vs 7.8 at 17: every launch pays its prologue and a split-K finish with the memory pipe idle;
# profiles/r02_decode_stack_findings.md), so it is opt-in: MLA_DECODE_SKINNY=1.
SKINNY = {"on": os.environ.get("MLA_DECODE_SKINNY", "0") == "1"}
_skinny_ws = {}


def _skinny_gemm(x, w, residual, out, norm, swiglu):
    m = x.shape[0]
    n, k = w.shape
    if out is None:
        out = torch.empty((m, n), dtype=torch.bfloat16, device=x.device)
    lib = _lib.lib()
    key = (x.device, n, m <= 16)
    ws = _skinny_ws.get(key)
    if ws is None:
        lib.mla_skinny_gemm_workspace.restype = C.c_size_t
        ws = torch.zeros(lib.mla_skinny_gemm_workspace(C.c_int32(n), C.c_int32(m)), dtype=torch.uint8, device=x.device)
        _skinny_ws[key] = ws
    a = _lib.GemvArgs()
    a.x, a.w, a.out = x.data_ptr(), w.data_ptr(), out.data_ptr()
    a.m, a.n, a.k = m, n, k
    a.ldx, a.ldw, a.ldo = _rowmajor_2d(x, "x"), _rowmajor_2d(w, "w"), _rowmajor_2d(out, "out")
    if residual is not None:
        _req(residual, torch.bfloat16, "residual")
        a.residual, a.ldr = residual.data_ptr(), _rowmajor_2d(residual, "residual")
    if norm is not None:
        _req(norm[0], torch.bfloat16, "ln_weight")
        a.prologue, a.ln_weight, a.eps = 1, norm[0].data_ptr(), float(norm[1])
    elif swiglu:
        a.prologue = 2
    check(lib.mla_skinny_gemm(C.byref(a), _p(ws), _stream()))
    return out


def gemv(x: torch.Tensor, w: torch.Tensor, residual: Optional[torch.Tensor] = None,
         out: Optional[torch.Tensor] = None, norm: Optional[tuple] = None, swiglu: bool = False) -> torch.Tensor:
    """Skinny nn.Linear for a handful of rows (HBM-bound weight streaming, csrc/decode.cu): x bf16 [m,k], w bf16 [n,k]
    -> bf16 [m,n] (+ residual).  Optional prologues on the activations, fused for m <= 4: norm = (ln_weight bf16 [k],
    eps) applies LlamaRMSNorm first; swiglu=True takes x = [gate | up] bf16 [m, 2k].  Otherwise the prologue
    runs as its own kernel; above 16 rows this is the tcgen05 GEMM."""
    _req(x, torch.bfloat16, "x")
    _req(w, torch.bfloat16, "w")
    m = x.shape[0]
    n, k = w.shape
    if x.shape[1] != (2 * k if swiglu else k):
        raise _lib.MlaError(f"gemv: contraction mismatch {x.shape[1]} vs {k}")
    if SKINNY["on"] and m <= 32 and k % 8 == 0 and w.is_contiguous():
        return _skinny_gemm(x, w, residual, out, norm, swiglu)
    fused = (m <= 4 and k <= 4096) or (m <= 2 and k <= 12288)      # activations fit the kernel's registers
    if not fused:
        if norm is not None:
            x, norm = rmsnorm_fwd(x, norm[0], norm[1]), None
        if swiglu:
            x, swiglu = swiglu_fwd(x.contiguous()), False
    if m > 16:          # SIMT dot products stop paying: the tensor-core GEMM streams the same weights
        return gemm(x.contiguous(), w, residual=residual, out=out)
    if out is None:
        out = torch.empty((m, n), dtype=torch.bfloat16, device=x.device)
    a = _lib.GemvArgs()
    a.x, a.w, a.out = x.data_ptr(), w.data_ptr(), out.data_ptr()
    a.m, a.n, a.k = m, n, k
    a.ldx, a.ldw, a.ldo = _rowmajor_2d(x, "x"), _rowmajor_2d(w, "w"), _rowmajor_2d(out, "out")
    if residual is not None:
        _req(residual, torch.bfloat16, "residual")
        a.residual, a.ldr = residual.data_ptr(), _rowmajor_2d(residual, "residual")
    if norm is not None:
        _req(norm[0], torch.bfloat16, "ln_weight")
        a.prologue, a.ln_weight, a.eps = 1, norm[0].data_ptr(), float(norm[1])
    elif swiglu:
        a.prologue = 2
    check(_lib.lib().mla_gemv_fused(C.byref(a), _stream()))
    return out


def rope_cache(qkv: torch.Tensor, cache: torch.Tensor, cos_t: torch.Tensor, sin_t: torch.Tensor, B: int, n: int, P: int,
               H: int, D: int) -> None:
    """RoPE on the q and k of the n new rows per sample of qkv [B*n, 3*H*D] (q in place, k into the cache) + v copy;
    cache bf16 [B*(P+n), 2*H*D], cos/sin = table rows P..P+n-1."""
    _req(qkv, torch.bfloat16, "qkv")
    _req(cache, torch.bfloat16, "cache")
    if tuple(qkv.shape) != (B * n, 3 * H * D) or not qkv.is_contiguous():
        raise _lib.MlaError(f"rope_cache: qkv must be contiguous [{B * n}, {3 * H * D}]")
    if tuple(cache.shape) != (B * (P + n), 2 * H * D) or not cache.is_contiguous():
        raise _lib.MlaError(f"rope_cache: cache must be contiguous [{B * (P + n)}, {2 * H * D}]")
    if tuple(cos_t.shape) != (n, D // 2) or not (cos_t.is_contiguous() and sin_t.is_contiguous()):
        raise _lib.MlaError("rope_cache: cos/sin must be contiguous [n, D/2]")
    check(_lib.lib().mla_rope_cache(_p(qkv), _p(cache), _p(cos_t), _p(sin_t), C.c_int32(B), C.c_int32(n), C.c_int32(P),
                                    C.c_int32(H), C.c_int32(D), _stream()))


def decode_attn(q: torch.Tensor, kv: torch.Tensor, B: int, H: int, Lq: int, Lk: int, D: int) -> torch.Tensor:
    """q: bf16 [B*Lq, >= H*D] (queries in the leading H*D columns, any row pitch); kv: bf16 [B*Lk, 2*H*D] (k | v), the
    cache of a length-Lk sequence whose last Lq positions are the queries.  Returns ctx bf16 [B*Lq, H*D]."""
    _req(q, torch.bfloat16, "q")
    _req(kv, torch.bfloat16, "kv")
    if tuple(kv.shape) != (B * Lk, 2 * H * D) or not kv.is_contiguous():
        raise _lib.MlaError(f"decode_attn: kv cache must be contiguous [{B * Lk}, {2 * H * D}], got {tuple(kv.shape)}")
    if q.shape[0] != B * Lq or q.shape[1] < H * D:
        raise _lib.MlaError(f"decode_attn: q shape {tuple(q.shape)} does not hold [{B * Lq}, {H * D}] queries")
    ctx = torch.empty((B * Lq, H * D), dtype=torch.bfloat16, device=q.device)
    check(_lib.lib().mla_decode_attn(_p(q), C.c_int64(_rowmajor_2d(q, "q")), _p(kv),
                                     C.c_void_p(kv.data_ptr() + 2 * H * D), C.c_int64(kv.stride(0)), _p(ctx),
                                     C.c_int64(ctx.stride(0)), C.c_int32(B), C.c_int32(H), C.c_int32(Lq), C.c_int32(Lk),
                                     C.c_int32(D), C.c_float(D ** -0.5), _stream()))
    return ctx


def decode_attn_rope(qkv: torch.Tensor, kv: torch.Tensor, cos_t: torch.Tensor, sin_t: torch.Tensor, B: int, H: int,
                     Lq: int, P: int, D: int, split_k: bool = False) -> torch.Tensor:
    """Attention of the Lq new rows per sample against a head-major prefix cache + themselves: qkv bf16 [B*Lq, 3*H*D]
    un-rotated (RoPE of q and of the new keys happens in the kernel; cos/sin = table rows P..P+Lq-1); kv bf16
    [B, 2, H, P, D] = the rotated prefix keys | values.  Returns ctx bf16 [B*Lq, H*D].  split_k spreads the keys of
    each query over ceil(keys/128) CTAs with a last-arrival merge; measured slower at 546 keys (20.6 + 3.4 us for the
    counter memset against 17.6 us, profiles/r01_launches_decode_layer_v8_splitk.csv), so it is off by default."""
    _req(qkv, torch.bfloat16, "qkv")
    _req(kv, torch.bfloat16, "kv")
    if tuple(qkv.shape) != (B * Lq, 3 * H * D) or not qkv.is_contiguous():
        raise _lib.MlaError(f"decode_attn_rope: qkv must be contiguous [{B * Lq}, {3 * H * D}]")
    if tuple(kv.shape) != (B, 2, H, P, D) or not kv.is_contiguous():
        raise _lib.MlaError(f"decode_attn_rope: kv cache must be contiguous [{B}, 2, {H}, {P}, {D}], got {tuple(kv.shape)}")
    if tuple(cos_t.shape) != (Lq, D // 2) or not (cos_t.is_contiguous() and sin_t.is_contiguous()):
        raise _lib.MlaError("decode_attn_rope: cos/sin must be contiguous [Lq, D/2]")
    ctx = torch.empty((B * Lq, H * D), dtype=torch.bfloat16, device=qkv.device)
    lib = _lib.lib()
    ws = cnt = None
    if split_k and P + Lq > 128:
        lib.mla_decode_attn_workspace.restype = C.c_size_t
        nbytes = lib.mla_decode_attn_workspace(C.c_int32(B), C.c_int32(H), C.c_int32(Lq), C.c_int32(P + Lq), C.c_int32(D))
        ws = torch.empty(nbytes // 4, dtype=torch.float32, device=qkv.device)
        cnt = torch.zeros(B * H * Lq, dtype=torch.int32, device=qkv.device)
    check(lib.mla_decode_attn_rope(_p(qkv), C.c_int64(qkv.stride(0)), _p(kv),
                                   C.c_void_p(kv.data_ptr() + 2 * H * P * D), C.c_int64(2 * H * P * D),
                                   C.c_int64(P * D), C.c_int64(D), _p(cos_t), _p(sin_t), _p(ctx),
                                   C.c_int64(ctx.stride(0)), C.c_int32(B), C.c_int32(H), C.c_int32(Lq),
                                   C.c_int32(P + Lq), C.c_int32(D), C.c_float(D ** -0.5), _p(ws), _p(cnt), _stream()))
    return ctx


def decode_stack_supported(rows: int, h: int, f: int, D: int) -> bool:
    """Shapes the one-launch decoder stack (csrc/decode_stack.cu) takes: activations of <= 2 rows live in registers."""
    return rows <= 2 and D in (32, 64, 128) and h % 8 == 0 and f % 8 == 0 and h <= 12288 and f <= 12288


def decode_stack_workspace(B: int, n: int, P: int, H: int, D: int, device) -> torch.Tensor:
    """Zeroed once; the kernel's arrival counters and grid barrier re-arm themselves."""
    lib = _lib.lib()
    lib.mla_decode_stack_workspace.restype = C.c_size_t
    nbytes = lib.mla_decode_stack_workspace(C.c_int32(B), C.c_int32(n), C.c_int32(P), C.c_int32(H), C.c_int32(D))
    return torch.zeros(nbytes, dtype=torch.uint8, device=device)


def decode_stack(x: torch.Tensor, table: torch.Tensor, cos_t: torch.Tensor, sin_t: torch.Tensor, workspace: torch.Tensor,
                 B: int, n: int, P: int, H: int, D: int, f: int, eps: float,
                 trace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All decoder layers over the n suffix rows per sample in ONE persistent launch (see mla_decode_stack).
    x bf16 [B*n, h] (not modified); table int64 [7, L] on the device = per-layer pointers of w_qkv, w_o, w_gate_up,
    w_down, ln1, ln2, kv_cache; returns the last layer's output bf16 [B*n, h] (final norm not applied)."""
    _req(x, torch.bfloat16, "x")
    h, M, L = H * D, B * n, table.shape[1]
    if tuple(x.shape) != (M, h):
        raise _lib.MlaError(f"decode_stack: x must be [{M}, {h}]")
    if table.dtype != torch.int64 or table.shape[0] != 7 or not table.is_contiguous():
        raise _lib.MlaError("decode_stack: the pointer table must be contiguous int64 [7, layers]")
    dev = x.device
    xb = x.contiguous().clone()
    qkv = torch.empty((M, 3 * h), dtype=torch.bfloat16, device=dev)
    ctx = torch.empty((M, h), dtype=torch.bfloat16, device=dev)
    xmid = torch.empty((M, h), dtype=torch.bfloat16, device=dev)
    gu = torch.empty((M, 2 * f), dtype=torch.bfloat16, device=dev)
    a = _lib.DecodeStackArgs()
    base = table.data_ptr()
    a.w_qkv, a.w_o, a.w_gate_up, a.w_down, a.ln1, a.ln2, a.kv_cache = (base + 8 * L * i for i in range(7))
    a.x, a.qkv, a.ctx, a.x_mid, a.gate_up = xb.data_ptr(), qkv.data_ptr(), ctx.data_ptr(), xmid.data_ptr(), gu.data_ptr()
    a.cos_t, a.sin_t, a.workspace = cos_t.data_ptr(), sin_t.data_ptr(), workspace.data_ptr()
    a.layers, a.batch, a.n, a.prefix, a.heads, a.head_dim, a.ffn = L, B, n, P, H, D, f
    a.eps, a.scale = float(eps), float(D ** -0.5)
    a.trace = trace.data_ptr() if trace is not None else None      # int64 [SMs, L, 5, 3] phase timestamps (profiling)
    check(_lib.lib().mla_decode_stack(C.byref(a), _stream()))
    return xb


def ddim_step(x: torch.Tensor, eps: torch.Tensor, coef: torch.Tensor) -> torch.Tensor:
    """x_{t-1} = ddim_sample(x_t, eps) with eta = 0 (see mla_ddim_step); x f32, eps bf16/f32, coef f32 [4] on device."""
    _req(x, torch.float32, "x")
    x, eps = x.contiguous(), eps.contiguous()
    if eps.dtype not in (torch.float32, torch.bfloat16) or eps.numel() != x.numel():
        raise _lib.MlaError("ddim_step: eps must be bf16 or f32 with as many elements as x")
    out = torch.empty_like(x)
    check(_lib.lib().mla_ddim_step(_p(x), _p(eps), C.c_int32(int(eps.dtype == torch.float32)), _p(coef), _p(out),
                                   C.c_int64(x.numel()), _stream()))
    return out
