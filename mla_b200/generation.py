"""Post-training generation heads — drop-in for models/mla/generation/models.py.

`MultimodalGenerationManager` owns the same sub-modules, constructor arguments and state-dict keys as the reference
(the parameter containers ARE torch's nn.TransformerDecoder / nn.MultiheadAttention / nn.LayerNorm / nn.Linear /
nn.Conv1d / nn.BatchNorm1d, so keys, shapes and default initialisation are identical), but none of their forwards is
ever called: every linear layer runs on the tcgen05 GEMM (`ops.linear`), the attention cores on `mla_mha_*`, the fp32
LayerNorms, BatchNorm, token plumbing, image warp/blend + losses and Chamfer distance on the kernels of
csrc/generation.cu.  The image and point-cloud losses of PrismaticVLM.compute_generation_losses
(models/vlm/prismatic.py:771-838) are fused into the last kernels of their heads.

dtype policy = the reference under FSDP bf16 parameters + CUDA autocast: linears in bf16, layer_norm in fp32 (so the
TransformerDecoder residual stream is fp32 after the first norm), bf16 + bf16 adds rounded to bf16.

Dropout / DropPath (p = 0.1 in the reference's train mode): the masks are drawn on the host side with torch's generator
and applied by our kernels; they cannot reproduce the reference's fused-dropout Philox stream, so parity tests run the
heads with p = 0 (distributional equivalence otherwise).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import GenImageArgs, MhaArgs, check

_p, _stream = ops._p, ops._stream


# ================================================================================================ autograd nodes
class ParamRowsFn(torch.autograd.Function):
    """Learned [P, h] table (fp32 master) -> f32 [B*P, h] holding its bf16 values, one copy per sample."""

    @staticmethod
    def forward(ctx, table, B):
        t = table.detach().contiguous()
        out = torch.empty((B * t.shape[-2], t.shape[-1]), dtype=torch.float32, device=t.device)
        check(_lib.lib().mla_tile_rows_fwd(_p(t), _p(out), C.c_int64(t.numel()), C.c_int32(B), _stream()))
        ctx.B, ctx.shape = B, table.shape
        return out

    @staticmethod
    def backward(ctx, d):
        d = d.contiguous()
        g = torch.empty(ctx.shape, dtype=torch.float32, device=d.device)
        check(_lib.lib().mla_tile_rows_bwd(_p(d), _p(g), C.c_int64(g.numel()), C.c_int32(ctx.B), _stream()))
        return g, None


class CastF32Fn(torch.autograd.Function):
    """bf16 -> f32 (exact); backward rounds the gradient to bf16, as autograd does for a .float() of a bf16 tensor."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
        check(_lib.lib().mla_cast_bf16_f32(_p(x), _p(y), C.c_int64(x.numel()), _stream()))
        return y

    @staticmethod
    def backward(ctx, d):
        return ops.cast_bf16(d.contiguous())


def _cast_f32(x16: torch.Tensor) -> torch.Tensor:
    y = torch.empty(x16.shape, dtype=torch.float32, device=x16.device)
    check(_lib.lib().mla_cast_bf16_f32(_p(x16.contiguous()), _p(y), C.c_int64(x16.numel()), _stream()))
    return y


class AddF32Bf16Fn(torch.autograd.Function):
    """out(f32) = a(f32) + b(bf16); round=True reproduces a bf16 + bf16 add (a then holds bf16 values)."""

    @staticmethod
    def forward(ctx, a, b, round_bf16):
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        check(_lib.lib().mla_add_f32_bf16(_p(a), _p(b), _p(out), C.c_int64(a.numel()), C.c_int32(int(round_bf16)),
                                          _stream()))
        return out

    @staticmethod
    def backward(ctx, d):
        d = d.contiguous()
        return d, ops.cast_bf16(d), None


class AddBf16Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return ops.add_bf16(a, b)

    @staticmethod
    def backward(ctx, d):
        return d, d


class LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm on the fp32 stream: returns (y f32, y bf16 copy for the next GEMM)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        ctx.set_materialize_grads(False)
        x = x.contiguous()
        h = x.shape[-1]
        rows = x.numel() // h
        y = torch.empty_like(x)
        y16 = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        # the reference's parameters are bf16 (FSDP MixedPrecision): use the bf16 values of the fp32 masters
        w = _cast_f32(ops.bf16_of(weight))
        b = _cast_f32(ops.bf16_of(bias))
        check(_lib.lib().mla_ln_f32_fwd(_p(x), _p(w), _p(b), _p(y), _p(y16), _p(mean), _p(rstd), C.c_int64(rows),
                                        C.c_int32(h), C.c_float(eps), _stream()))
        ctx.save_for_backward(x, w, mean, rstd)
        return y, y16

    @staticmethod
    def backward(ctx, dy, dy16):
        x, w, mean, rstd = ctx.saved_tensors
        if dy is None and dy16 is None:
            return None, None, None, None
        if dy is None:
            g = _cast_f32(dy16)
        elif dy16 is None:
            g = dy.contiguous()
        else:
            g = torch.empty_like(x)
            check(_lib.lib().mla_add_f32_bf16(_p(dy.contiguous()), _p(dy16.contiguous()), _p(g), C.c_int64(g.numel()),
                                              C.c_int32(0), _stream()))
        h = x.shape[-1]
        dx = torch.empty_like(x)
        dw = torch.zeros(h, dtype=torch.float32, device=x.device)
        db = torch.zeros(h, dtype=torch.float32, device=x.device)
        check(_lib.lib().mla_ln_f32_bwd(_p(g), _p(x), _p(w), _p(mean), _p(rstd), _p(dx), _p(dw), _p(db),
                                        C.c_int64(x.numel() // h), C.c_int32(h), _stream()))
        return dx, dw, db, None


class MaskScaleFn(torch.autograd.Function):
    """y = keep[i // per_mask] ? x / (1 - p) : 0 — dropout (per_mask = 1) or DropPath (per_mask = elements/sample)."""

    @staticmethod
    def forward(ctx, x, keep, per_mask, scale):
        x = x.contiguous()
        y = torch.empty_like(x)
        check(_lib.lib().mla_mask_scale_bf16(_p(x), _p(keep), _p(y), C.c_int64(x.numel()), C.c_int64(per_mask),
                                             C.c_float(scale), _stream()))
        ctx.save_for_backward(keep)
        ctx.per_mask, ctx.scale = per_mask, scale
        return y

    @staticmethod
    def backward(ctx, d):
        (keep,) = ctx.saved_tensors
        d = d.contiguous()
        g = torch.empty_like(d)
        check(_lib.lib().mla_mask_scale_bf16(_p(d), _p(keep), _p(g), C.c_int64(d.numel()), C.c_int64(ctx.per_mask),
                                             C.c_float(ctx.scale), _stream()))
        return g, None, None, None


def dropout(x: torch.Tensor, p: float, training: bool, per_sample_of: int = 0) -> torch.Tensor:
    """nn.Dropout(p) (per_sample_of = 0) or timm DropPath(p) over `per_sample_of` samples; identity when inactive."""
    if not training or p <= 0.0:
        return x
    n_masks = per_sample_of if per_sample_of else x.numel()
    keep = (torch.rand(n_masks, device=x.device) >= p).to(torch.uint8)     # host-side generator draw (RNG plumbing)
    return MaskScaleFn.apply(x, keep, x.numel() // n_masks, 1.0 / (1.0 - p))


class MhaFn(torch.autograd.Function):
    """Attention core.  Self-attention: a = packed [B*L, 3d] (q | k | v), b = None.  Cross-attention: a = q [B*Lq, d],
    b = packed [B*Lk, 2d] (k | v).  Returns ctx [B*Lq, d]."""

    @staticmethod
    def forward(ctx, a, b, B, H, Lq, Lk, keep, keep_scale):
        a = a.contiguous()
        d = a.shape[1] // 3 if b is None else a.shape[1]
        D = d // H
        if b is not None:
            b = b.contiguous()
        o = torch.empty((B * Lq, d), dtype=torch.bfloat16, device=a.device)
        lse = torch.empty((B, H, Lq), dtype=torch.float32, device=a.device)
        args = MhaFn._args(a, b, d, o, lse, B, H, Lq, Lk, D, keep, keep_scale)
        check(_lib.lib().mla_mha_fwd(C.byref(args), _stream()))
        ctx.save_for_backward(a, b, o, lse, keep)
        ctx.dims = (B, H, Lq, Lk, D, d, keep_scale)
        return o

    @staticmethod
    def _args(a, b, d, o, lse, B, H, Lq, Lk, D, keep, keep_scale):
        g = MhaArgs()
        if b is None:
            g.q, g.k, g.v = a.data_ptr(), a.data_ptr() + 2 * d, a.data_ptr() + 4 * d
            g.ldq = g.ldk = g.ldv = a.stride(0)
        else:
            g.q, g.k, g.v = a.data_ptr(), b.data_ptr(), b.data_ptr() + 2 * d
            g.ldq, g.ldk, g.ldv = a.stride(0), b.stride(0), b.stride(0)
        g.o, g.ldo, g.lse = o.data_ptr(), o.stride(0), lse.data_ptr()
        g.keep_mask = None if keep is None else keep.data_ptr()
        g.keep_scale = keep_scale
        g.batch, g.heads, g.len_q, g.len_k, g.head_dim = B, H, Lq, Lk, D
        g.scale = D ** -0.5
        return g

    @staticmethod
    def backward(ctx, do):
        a, b, o, lse, keep = ctx.saved_tensors
        B, H, Lq, Lk, D, d, keep_scale = ctx.dims
        do = do.contiguous()
        g = MhaFn._args(a, b, d, o, lse, B, H, Lq, Lk, D, keep, keep_scale)
        delta = torch.empty_like(lse)
        da = torch.empty_like(a)
        g.d_o, g.ld_do, g.delta = do.data_ptr(), do.stride(0), delta.data_ptr()
        if b is None:
            db = None
            g.dq, g.dk, g.dv = da.data_ptr(), da.data_ptr() + 2 * d, da.data_ptr() + 4 * d
            g.ld_dq = g.ld_dk = g.ld_dv = da.stride(0)
        else:
            db = torch.empty_like(b)
            g.dq, g.dk, g.dv = da.data_ptr(), db.data_ptr(), db.data_ptr() + 2 * d
            g.ld_dq, g.ld_dk, g.ld_dv = da.stride(0), db.stride(0), db.stride(0)
        check(_lib.lib().mla_mha_bwd(C.byref(g), _stream()))
        return da, db, None, None, None, None, None, None


def _attn_keep(p: float, training: bool, B: int, H: int, Lq: int, Lk: int, device):
    if not training or p <= 0.0:
        return None, 1.0
    return (torch.rand((B, H, Lq, Lk), device=device) >= p).to(torch.uint8), 1.0 / (1.0 - p)


class SeqMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, B, S):
        x = x.contiguous()
        Cc = x.shape[-1]
        y = torch.empty((B, Cc), dtype=torch.bfloat16, device=x.device)
        check(_lib.lib().mla_seq_mean_fwd(_p(x), _p(y), C.c_int32(B), C.c_int32(S), C.c_int32(Cc), _stream()))
        ctx.dims = (B, S, Cc)
        return y

    @staticmethod
    def backward(ctx, d):
        B, S, Cc = ctx.dims
        d = d.contiguous()
        dx = torch.empty((B * S, Cc), dtype=torch.bfloat16, device=d.device)
        check(_lib.lib().mla_seq_mean_bwd(_p(d), _p(dx), C.c_int32(B), C.c_int32(S), C.c_int32(Cc), _stream()))
        return dx, None, None


class BnRowsFn(torch.autograd.Function):
    """nn.BatchNorm1d (train-mode statistics, running stats updated) + ReLU over rows [R, C] bf16."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, training):
        if not training:
            raise _lib.MlaError("BatchNorm1d of the point-cloud generation head: eval mode is not part of the training path")
        x = x.contiguous()
        R, Cc = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(Cc, dtype=torch.float32, device=x.device)
        rstd = torch.empty(Cc, dtype=torch.float32, device=x.device)
        w = _cast_f32(ops.bf16_of(weight))
        b = _cast_f32(ops.bf16_of(bias))
        check(_lib.lib().mla_bn_rows_fwd(_p(x), _p(w), _p(b), _p(y), _p(mean), _p(rstd), _p(running_mean),
                                         _p(running_var), C.c_int32(R), C.c_int32(Cc), C.c_float(eps),
                                         C.c_float(momentum), C.c_int32(1), _stream()))
        ctx.save_for_backward(x, y, w, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, w, mean, rstd = ctx.saved_tensors
        R, Cc = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dw = torch.empty(Cc, dtype=torch.float32, device=x.device)
        db = torch.empty(Cc, dtype=torch.float32, device=x.device)
        check(_lib.lib().mla_bn_rows_bwd(_p(dy), _p(x), _p(y), _p(w), _p(mean), _p(rstd), _p(dx), _p(dw), _p(db),
                                         C.c_int32(R), C.c_int32(Cc), C.c_int32(1), _stream()))
        return dx, dw, db, None, None, None, None, None


class MaskTokensFn(torch.autograd.Function):
    """MAE decoder input: (roi ? mask_token : image feature) + pos  (generation/models.py:183-187)."""

    @staticmethod
    def forward(ctx, feat, roi, mask_token, pos, B, P):
        feat = feat.contiguous()
        h = feat.shape[-1]
        out = torch.empty((B * P, h), dtype=torch.float32, device=feat.device)
        check(_lib.lib().mla_mask_tokens_fwd(_p(feat), _p(roi), _p(mask_token.detach().contiguous()),
                                             _p(pos.detach().contiguous()), _p(out), C.c_int32(B), C.c_int32(P),
                                             C.c_int32(h), _stream()))
        ctx.save_for_backward(roi)
        ctx.dims = (B, P, h, mask_token.shape, pos.shape)
        return out

    @staticmethod
    def backward(ctx, d):
        (roi,) = ctx.saved_tensors
        B, P, h, mshape, pshape = ctx.dims
        d = d.contiguous()
        dfeat = torch.empty((B * P, h), dtype=torch.bfloat16, device=d.device)
        dm = torch.zeros(mshape, dtype=torch.float32, device=d.device)
        dpos = torch.empty(pshape, dtype=torch.float32, device=d.device)
        check(_lib.lib().mla_mask_tokens_bwd(_p(d), _p(roi), _p(dfeat), _p(dm), _p(dpos), C.c_int32(B), C.c_int32(P),
                                             C.c_int32(h), _stream()))
        return dfeat, None, dm, dpos, None, None


class GenImageFn(torch.autograd.Function):
    """Image head tail + image losses.  Returns (image_gen_loss, losses[4], blended, delta_all, alpha_all, offset_all);
    only the first output is differentiable."""

    @staticmethod
    def forward(ctx, delta_raw, ao_raw, roi, cur, nxt, B, G, ps, delta_clip, max_shift, gen_weight):
        N, E = B * G * G, 3 * ps * ps
        dev = delta_raw.device
        a = GenImageArgs()
        a.cur, a.nxt = cur.data_ptr(), nxt.data_ptr()
        a.cur_stride_b, a.cur_stride_c = cur.stride(0), cur.stride(1)
        a.nxt_stride_b, a.nxt_stride_c = nxt.stride(0), nxt.stride(1)
        if cur.stride(3) != 1 or nxt.stride(3) != 1 or cur.stride(2) != cur.shape[3] or nxt.stride(2) != nxt.shape[3]:
            raise _lib.MlaError("generation: images must be row-contiguous [B, C, H, W] tensors")
        a.n_images, a.width, a.patch, a.grid, a.n_patches = cur.shape[0], cur.shape[3], ps, G, N
        a.delta_raw, a.ld_delta = delta_raw.data_ptr(), delta_raw.stride(0)
        a.ao_raw, a.ld_ao = ao_raw.data_ptr(), ao_raw.stride(0)
        a.roi = roi.data_ptr()
        a.delta_clip, a.max_shift, a.gen_weight = delta_clip, max_shift, gen_weight
        blended = torch.empty((N, E), dtype=torch.float32, device=dev)
        delta_all = torch.empty((N, E), dtype=torch.bfloat16, device=dev)
        alpha_all = torch.empty(N, dtype=torch.bfloat16, device=dev)
        offset_all = torch.empty((N, 2), dtype=torch.bfloat16, device=dev)
        sums = torch.empty(5, dtype=torch.float32, device=dev)
        losses = torch.empty(4, dtype=torch.float32, device=dev)
        coef = torch.empty(4, dtype=torch.float32, device=dev)
        a.blended, a.delta_all, a.alpha_all, a.offset_all = (blended.data_ptr(), delta_all.data_ptr(),
                                                             alpha_all.data_ptr(), offset_all.data_ptr())
        a.sums, a.losses, a.coef = sums.data_ptr(), losses.data_ptr(), coef.data_ptr()
        check(_lib.lib().mla_gen_image_fwd(C.byref(a), _stream()))
        ctx.save_for_backward(delta_raw, ao_raw, roi, cur, nxt, coef)
        ctx.consts = (B, G, ps, delta_clip, max_shift, gen_weight)
        for t in (losses, blended, delta_all, alpha_all, offset_all):
            ctx.mark_non_differentiable(t)
        return losses[0].clone(), losses, blended, delta_all, alpha_all, offset_all

    @staticmethod
    def backward(ctx, g, *unused):
        delta_raw, ao_raw, roi, cur, nxt, coef = ctx.saved_tensors
        B, G, ps, delta_clip, max_shift, gen_weight = ctx.consts
        N = B * G * G
        a = GenImageArgs()
        a.cur, a.nxt = cur.data_ptr(), nxt.data_ptr()
        a.cur_stride_b, a.cur_stride_c = cur.stride(0), cur.stride(1)
        a.nxt_stride_b, a.nxt_stride_c = nxt.stride(0), nxt.stride(1)
        a.n_images, a.width, a.patch, a.grid, a.n_patches = cur.shape[0], cur.shape[3], ps, G, N
        a.delta_raw, a.ld_delta = delta_raw.data_ptr(), delta_raw.stride(0)
        a.ao_raw, a.ld_ao = ao_raw.data_ptr(), ao_raw.stride(0)
        a.roi = roi.data_ptr()
        a.delta_clip, a.max_shift, a.gen_weight = delta_clip, max_shift, gen_weight
        gs = g.reshape(1).float().contiguous()
        # gradients laid out like the raw tensors (same pitch), returned as views of the right shape
        dd = torch.zeros((N, delta_raw.stride(0)), dtype=torch.bfloat16, device=g.device)
        dao = torch.zeros((N, ao_raw.stride(0)), dtype=torch.bfloat16, device=g.device)
        a.coef, a.grad_scale = coef.data_ptr(), gs.data_ptr()
        a.d_delta_raw, a.d_ao_raw = dd.data_ptr(), dao.data_ptr()
        check(_lib.lib().mla_gen_image_bwd(C.byref(a), _stream()))
        return (dd[:, :delta_raw.shape[1]], dao[:, :ao_raw.shape[1]]) + (None,) * 9


class ChamferFn(torch.autograd.Function):
    """chamfer_distance_l2(pred bf16 [B, N1, 3], gt f32 [n_gt, N2, 3]) — generation/gen_loss.py:12-18."""

    @staticmethod
    def forward(ctx, pred, gt):
        pred, gt = pred.contiguous(), gt.contiguous()
        B, N1, _ = pred.shape
        n_gt, N2, _ = gt.shape
        dev = pred.device
        idx1 = torch.empty((B, N1), dtype=torch.int32, device=dev)
        idx2 = torch.empty((B, N2), dtype=torch.int32, device=dev)
        d1 = torch.empty((B, N1), dtype=torch.float32, device=dev)
        d2 = torch.empty((B, N2), dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        check(_lib.lib().mla_chamfer_fwd(_p(pred), _p(gt), C.c_int32(B), C.c_int32(N1), C.c_int32(N2), C.c_int32(n_gt),
                                         _p(idx1), _p(d1), _p(idx2), _p(d2), _p(loss), _stream()))
        ctx.save_for_backward(pred, gt, idx1, d1, idx2, d2)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        pred, gt, idx1, d1, idx2, d2 = ctx.saved_tensors
        B, N1, _ = pred.shape
        n_gt, N2, _ = gt.shape
        gs = g.reshape(1).float().contiguous()
        dp = torch.zeros(pred.shape, dtype=torch.float32, device=pred.device)
        check(_lib.lib().mla_chamfer_bwd(_p(pred), _p(gt), C.c_int32(B), C.c_int32(N1), C.c_int32(N2), C.c_int32(n_gt),
                                         _p(idx1), _p(d1), _p(idx2), _p(d2), _p(gs), _p(dp), _stream()))
        return ops.cast_bf16(dp), None


def roi_mask(patch_indices: torch.Tensor, grid: int, ksize: int) -> torch.Tensor:
    """create_roi_mask_from_indices + dilate_mask -> u8 [B, grid*grid]."""
    pi = patch_indices.contiguous()
    B, n_pts, _ = pi.shape
    out = torch.empty((B, grid * grid), dtype=torch.uint8, device=pi.device)
    check(_lib.lib().mla_roi_mask(_p(pi), _p(out), C.c_int32(B), C.c_int32(n_pts), C.c_int32(grid), C.c_int32(ksize),
                                  _stream()))
    return out


# ================================================================================================ building blocks
def _mha_self(x16, mha: nn.MultiheadAttention, B, L, training):
    d, H = mha.embed_dim, mha.num_heads
    qkv = ops.linear(x16, mha.in_proj_weight, mha.in_proj_bias)
    keep, ks = _attn_keep(mha.dropout, training, B, H, L, L, x16.device)
    ctx = MhaFn.apply(qkv, None, B, H, L, L, keep, ks)
    return ops.linear(ctx, mha.out_proj.weight, mha.out_proj.bias)


def _mha_cross(x16, mem16, mha: nn.MultiheadAttention, B, Lq, Lk, training):
    d, H = mha.embed_dim, mha.num_heads
    q = ops.linear(x16, mha.in_proj_weight[:d], mha.in_proj_bias[:d])
    kv = ops.linear(mem16, mha.in_proj_weight[d:], mha.in_proj_bias[d:])
    keep, ks = _attn_keep(mha.dropout, training, B, H, Lq, Lk, x16.device)
    ctx = MhaFn.apply(q, kv, B, H, Lq, Lk, keep, ks)
    return ops.linear(ctx, mha.out_proj.weight, mha.out_proj.bias)


def decoder_layer(layer: nn.TransformerDecoderLayer, x32, x16, bf16_valued, mem16, B, Lq, Lk, training):
    """nn.TransformerDecoderLayer(norm_first=False, activation='gelu', batch_first=True).forward on our kernels.
    x32 [B*Lq, d] f32 stream (+ its bf16 copy x16); returns (y32, y16)."""
    sa = dropout(_mha_self(x16, layer.self_attn, B, Lq, training), layer.dropout1.p, training)
    x32, x16 = LayerNormFn.apply(AddF32Bf16Fn.apply(x32, sa, bf16_valued), layer.norm1.weight, layer.norm1.bias,
                                 layer.norm1.eps)
    ca = dropout(_mha_cross(x16, mem16, layer.multihead_attn, B, Lq, Lk, training), layer.dropout2.p, training)
    x32, x16 = LayerNormFn.apply(AddF32Bf16Fn.apply(x32, ca, False), layer.norm2.weight, layer.norm2.bias,
                                 layer.norm2.eps)
    hdn = dropout(ops.linear(x16, layer.linear1.weight, layer.linear1.bias, ops.ACT_GELU_ERF), layer.dropout.p, training)
    ff = dropout(ops.linear(hdn, layer.linear2.weight, layer.linear2.bias), layer.dropout3.p, training)
    return LayerNormFn.apply(AddF32Bf16Fn.apply(x32, ff, False), layer.norm3.weight, layer.norm3.bias, layer.norm3.eps)


def run_decoder(dec: nn.TransformerDecoder, x32, mem16, B, Lq, Lk, training):
    """x32: f32 [B*Lq, d] holding bf16 values (the reference's tgt is bf16)."""
    x16 = ops.cast_bf16(x32.detach()) if not x32.requires_grad else _ToBf16Fn.apply(x32)
    bf16_valued = True
    for layer in dec.layers:
        x32, x16 = decoder_layer(layer, x32, x16, bf16_valued, mem16, B, Lq, Lk, training)
        bf16_valued = False
    if dec.norm is not None:
        x32, x16 = LayerNormFn.apply(x32, dec.norm.weight, dec.norm.bias, dec.norm.eps)
    return x32, x16


class _ToBf16Fn(torch.autograd.Function):
    """f32 -> bf16 cast with a gradient (the autocast cast in front of a linear layer)."""

    @staticmethod
    def forward(ctx, x):
        return ops.cast_bf16(x.contiguous())

    @staticmethod
    def backward(ctx, d):
        return _cast_f32(d)


class DropPath(nn.Module):
    """timm DropPath (stochastic depth per sample); parameter-free, so it does not appear in the state dict."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = drop_prob


# ================================================================================================ modules
class TransformerBlock(nn.Module):
    """generation/models.py:39-65 (pre-norm block of the point-cloud head)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, drop=0.0, attn_drop=0.0, drop_path=0.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = nn.MultiheadAttention(dim, num_heads, dropout=attn_drop, batch_first=True)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = nn.LayerNorm(dim)
        hid = int(dim * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(dim, hid), nn.GELU(), nn.Dropout(drop), nn.Linear(hid, dim), nn.Dropout(drop))

    def run(self, x16, pos32, B, G):
        tr = self.training
        dp = self.drop_path.drop_prob if isinstance(self.drop_path, DropPath) else 0.0
        _, n16 = LayerNormFn.apply(AddF32Bf16Fn.apply(pos32, x16, True), self.norm1.weight, self.norm1.bias,
                                   self.norm1.eps)
        a = _mha_self(n16, self.attn, B, G, tr)
        x16 = AddBf16Fn.apply(x16, dropout(a, dp, tr, per_sample_of=B))
        _, n16 = LayerNormFn.apply(CastF32Fn.apply(x16), self.norm2.weight, self.norm2.bias, self.norm2.eps)
        hdn = dropout(ops.linear(n16, self.mlp[0].weight, self.mlp[0].bias, ops.ACT_GELU_ERF), self.mlp[2].p, tr)
        m = dropout(ops.linear(hdn, self.mlp[3].weight, self.mlp[3].bias), self.mlp[4].p, tr)
        return AddBf16Fn.apply(x16, dropout(m, dp, tr, per_sample_of=B))


class ImageGenerationModule(nn.Module):
    """generation/models.py:68-286."""

    def __init__(self, token_size=4096, num_gen_queries=64, decoder_layers=3, decoder_heads=8, image_patch_size=42,
                 use_roi=True, roi_dilation_kernel_size=3, gen_delta_clip=5.0, max_patch_shift_pixels=8,
                 use_patch_offset=True):
        super().__init__()
        self.token_size, self.num_gen_queries, self.image_patch_size = token_size, num_gen_queries, image_patch_size
        self.use_roi, self.roi_dilation_kernel_size = use_roi, roi_dilation_kernel_size
        self.gen_delta_clip, self.max_patch_shift_pixels = gen_delta_clip, max_patch_shift_pixels
        self.use_patch_offset = use_patch_offset
        self.image_num_patches = 256
        self.image_gen_queries = nn.Parameter(torch.zeros(1, num_gen_queries, token_size))
        self.mae_mask_token = nn.Parameter(torch.zeros(1, 1, token_size))
        self.mae_pos_embed = nn.Parameter(torch.zeros(1, self.image_num_patches, token_size))
        self.intent_decoder = nn.TransformerDecoder(
            nn.TransformerDecoderLayer(d_model=token_size, nhead=decoder_heads, dim_feedforward=token_size * 2,
                                       dropout=0.1, activation="gelu", batch_first=True), num_layers=2)
        self.mae_decoder = nn.TransformerDecoder(
            nn.TransformerDecoderLayer(d_model=token_size, nhead=decoder_heads, dim_feedforward=token_size * 4,
                                       dropout=0.1, activation="gelu", batch_first=True), num_layers=decoder_layers)
        patch_dim = image_patch_size ** 2 * 3
        self.mae_patch_norm = nn.LayerNorm(token_size)
        self.mae_delta_head = nn.Linear(token_size, patch_dim)
        self.mae_alpha_head = nn.Linear(token_size, 1)
        self.mae_offset_head = nn.Linear(token_size, 2)
        nn.init.normal_(self.image_gen_queries, std=0.02)
        nn.init.normal_(self.mae_mask_token, std=0.02)
        nn.init.normal_(self.mae_pos_embed, std=0.02)
        nn.init.normal_(self.mae_delta_head.weight, std=0.02)
        nn.init.constant_(self.mae_delta_head.bias, 0.0)
        nn.init.normal_(self.mae_alpha_head.weight, std=0.02)
        nn.init.constant_(self.mae_alpha_head.bias, -3.0)
        nn.init.normal_(self.mae_offset_head.weight, std=0.001)
        nn.init.constant_(self.mae_offset_head.bias, 0.0)

    def run(self, hidden16, B, S, img_feat16, cur_images, next_images, roi_u8) -> Dict[str, torch.Tensor]:
        """hidden16 [B*S, h] last LLM hidden state; img_feat16 [B*256, h] front-image tokens; cur/next images f32
        [n_img, >=3, 672, 672] (sample b uses image b % n_img); roi_u8 [B, 256] (already dilated) or None."""
        if not self.use_patch_offset:
            raise NotImplementedError("use_patch_offset=False is not built (the reference never sets it)")
        tr, h, P, Q = self.training, self.token_size, self.image_num_patches, self.num_gen_queries
        q32 = ParamRowsFn.apply(self.image_gen_queries, B)
        _, intent16 = run_decoder(self.intent_decoder, q32, hidden16, B, Q, S, tr)
        if roi_u8 is None:
            roi_u8 = torch.ones((B, P), dtype=torch.uint8, device=hidden16.device)
        tok32 = MaskTokensFn.apply(img_feat16, roi_u8, self.mae_mask_token, self.mae_pos_embed, B, P)
        gen32, _ = run_decoder(self.mae_decoder, tok32, intent16, B, P, Q, tr)
        _, fn16 = LayerNormFn.apply(gen32, self.mae_patch_norm.weight, self.mae_patch_norm.bias, self.mae_patch_norm.eps)
        delta_raw = ops.linear(fn16, self.mae_delta_head.weight, self.mae_delta_head.bias)
        ao_raw = torch.cat([ops.linear(fn16, self.mae_alpha_head.weight, self.mae_alpha_head.bias),
                            ops.linear(fn16, self.mae_offset_head.weight, self.mae_offset_head.bias)], dim=1)   # [N, 3]
        G = int(P ** 0.5)
        loss, losses, blended, delta_all, alpha_all, offset_all = GenImageFn.apply(
            delta_raw, ao_raw, roi_u8, cur_images, next_images, B, G, self.image_patch_size, float(self.gen_delta_clip),
            float(self.max_patch_shift_pixels), 0.95)
        E = 3 * self.image_patch_size ** 2
        return {"image_generation": blended.view(B, P, E), "generation_roi_mask": roi_u8.bool(),
                "delta_all": delta_all.view(B, P, E), "alpha_all": alpha_all.view(B, P),
                "offset_all": offset_all.view(B, P, 2), "_image_gen_loss": loss, "_image_loss_terms": losses}


class PointCloudGenerationModule(nn.Module):
    """generation/models.py:289-386 (current_point_cloud is always None on the training path, prismatic.py:1098, so the
    geometric-prior FPS branch never runs)."""

    def __init__(self, prismatic_hidden_dim=4096, trans_dim=1024, decoder_depth=4, decoder_num_heads=8, group_size=32,
                 num_groups=128, loss="cdl2", use_geometric_prior=True):
        super().__init__()
        self.prismatic_hidden_dim, self.trans_dim = prismatic_hidden_dim, trans_dim
        self.decoder_depth, self.decoder_num_heads = decoder_depth, decoder_num_heads
        self.group_size, self.num_groups, self.loss, self.use_geometric_prior = group_size, num_groups, loss, use_geometric_prior
        self.feature_projector = nn.Linear(prismatic_hidden_dim, trans_dim)
        self.seq_to_patch = nn.Linear(trans_dim, num_groups * trans_dim)
        self.pos_embed = nn.Parameter(torch.zeros(1, num_groups, trans_dim))
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        self.decoder_blocks = nn.ModuleList([
            TransformerBlock(dim=trans_dim, num_heads=decoder_num_heads, mlp_ratio=4.0, qkv_bias=True, drop=0.1,
                             attn_drop=0.1, drop_path=0.1) for _ in range(decoder_depth)])
        self.future_predictor = nn.Sequential(nn.Conv1d(trans_dim, trans_dim, 1), nn.BatchNorm1d(trans_dim),
                                              nn.ReLU(inplace=True), nn.Conv1d(trans_dim, 3 * group_size, 1))
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def run(self, hidden16, B, S) -> Dict[str, torch.Tensor]:
        G, td = self.num_groups, self.trans_dim
        proj = ops.linear(hidden16, self.feature_projector.weight, self.feature_projector.bias)     # [B*S, td]
        agg = SeqMeanFn.apply(proj, B, S)                                                             # [B, td]
        x16 = ops.linear(agg, self.seq_to_patch.weight, self.seq_to_patch.bias).reshape(B * G, td)
        pos32 = ParamRowsFn.apply(self.pos_embed, B)
        for blk in self.decoder_blocks:
            x16 = blk.run(x16, pos32, B, G)
        c0, bn, c3 = self.future_predictor[0], self.future_predictor[1], self.future_predictor[3]
        y = ops.linear(x16, c0.weight[:, :, 0], c0.bias)
        y = BnRowsFn.apply(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum, self.training)
        if self.training:
            bn.num_batches_tracked += 1
        y = ops.linear(y, c3.weight[:, :, 0], c3.bias)                                               # [B*G, 3*M]
        return {"pointcloud_coord_generation": y.reshape(B, G * self.group_size, 3)}


class TactileGenerationModule(nn.Module):
    """generation/models.py:389-430."""

    def __init__(self, token_size=4096, tactile_dim=128, decoder_layers=2, decoder_heads=4):
        super().__init__()
        self.token_size, self.tactile_dim = token_size, tactile_dim
        self.feature_projector = nn.Linear(token_size, token_size)
        self.tactile_query = nn.Parameter(torch.zeros(1, 1, token_size))
        nn.init.normal_(self.tactile_query, std=0.02)
        self.decoder = nn.TransformerDecoder(
            nn.TransformerDecoderLayer(d_model=token_size, nhead=decoder_heads, dim_feedforward=token_size * 2,
                                       dropout=0.1, activation="gelu", batch_first=True), num_layers=decoder_layers)
        self.output_head = nn.Linear(token_size, tactile_dim)

    def run(self, hidden16, B, S) -> Dict[str, torch.Tensor]:
        mem = ops.linear(hidden16, self.feature_projector.weight, self.feature_projector.bias)
        q32 = ParamRowsFn.apply(self.tactile_query, B)
        _, dec16 = run_decoder(self.decoder, q32, mem, B, 1, S, self.training)
        return {"tactile_generation": ops.linear(dec16, self.output_head.weight, self.output_head.bias)}


class MultimodalGenerationManager(nn.Module):
    """generation/models.py:433-539 — same constructor arguments and sub-module names."""

    def __init__(self, token_size=4096, use_image_generation=False, num_image_gen_queries=64, image_decoder_layers=3,
                 image_decoder_heads=8, image_patch_size=42, use_roi=True, roi_dilation_kernel_size=3,
                 use_pointcloud_generation=False, pointcloud_trans_dim=1024, pointcloud_decoder_layers=4,
                 pointcloud_decoder_heads=8, pointcloud_group_size=16, pointcloud_num_groups=64,
                 use_tactile_generation=False, tactile_dim=128, tactile_decoder_layers=2, tactile_decoder_heads=4):
        super().__init__()
        self.use_image_generation = use_image_generation
        self.use_pointcloud_generation = use_pointcloud_generation
        self.use_tactile_generation = use_tactile_generation
        if use_image_generation:
            self.image_gen_module = ImageGenerationModule(
                token_size=token_size, num_gen_queries=num_image_gen_queries, decoder_layers=image_decoder_layers,
                decoder_heads=image_decoder_heads, image_patch_size=image_patch_size, use_roi=use_roi,
                roi_dilation_kernel_size=roi_dilation_kernel_size)
        if use_pointcloud_generation:
            self.pointcloud_gen_module = PointCloudGenerationModule(
                prismatic_hidden_dim=token_size, trans_dim=pointcloud_trans_dim, decoder_depth=pointcloud_decoder_layers,
                decoder_num_heads=pointcloud_decoder_heads, group_size=pointcloud_group_size,
                num_groups=pointcloud_num_groups, loss="cdl2", use_geometric_prior=True)
        if use_tactile_generation:
            self.tactile_gen_module = TactileGenerationModule(
                token_size=token_size, tactile_dim=tactile_dim, decoder_layers=tactile_decoder_layers,
                decoder_heads=tactile_decoder_heads)

    def get_module_keys(self) -> list:
        keys = []
        if self.use_image_generation:
            keys.append("image_gen_module")
        if self.use_pointcloud_generation:
            keys.append("pointcloud_gen_module")
        if self.use_tactile_generation:
            keys.append("tactile_gen_module")
        return keys

    def run(self, hidden16, B, S, img_feat16=None, cur_images=None, next_images=None, roi_u8=None):
        out: Dict[str, torch.Tensor] = {}
        if self.use_image_generation:
            out.update(self.image_gen_module.run(hidden16, B, S, img_feat16, cur_images, next_images, roi_u8))
        if self.use_pointcloud_generation:
            out.update(self.pointcloud_gen_module.run(hidden16, B, S))
        if self.use_tactile_generation:
            out.update(self.tactile_gen_module.run(hidden16, B, S))
        return out


def compute_generation_losses(outputs: Dict[str, torch.Tensor], gen_image: bool, gen_pointcloud: bool,
                              gen_tactile: bool, next_point_cloud: Optional[torch.Tensor] = None,
                              next_tactile: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """PrismaticVLM.compute_generation_losses (prismatic.py:771-838); the image terms come fused out of GenImageFn."""
    losses: Dict[str, torch.Tensor] = {}
    total = None
    if gen_image and "_image_gen_loss" in outputs:
        terms = outputs.pop("_image_loss_terms")
        losses["image_roi_generation_loss"] = terms[1]
        losses["bg_consistency_loss"] = terms[2]
        losses["delta_magnitude_reward"] = terms[3]
        losses["image_gen_loss"] = outputs.pop("_image_gen_loss")
        total = losses["image_gen_loss"]
    if gen_pointcloud and next_point_cloud is not None and "pointcloud_coord_generation" in outputs:
        assert next_point_cloud.shape[2] == 3, "Point cloud must have 3 dimensions (XYZ)"
        pc = ChamferFn.apply(outputs["pointcloud_coord_generation"], next_point_cloud.float())
        losses["point_cloud_gen_loss"] = pc
        total = pc if total is None else total + pc
    if gen_tactile and next_tactile is not None and "tactile_generation" in outputs:
        pred = outputs["tactile_generation"]
        tgt = next_tactile.float().reshape(next_tactile.shape[0], -1)
        if tgt.shape[0] != pred.shape[0]:        # targets arrive un-repeated: sample b uses row b % B
            tgt = tgt.repeat(pred.shape[0] // tgt.shape[0], 1)
        tl = ops.MSEFn.apply(pred, tgt)
        losses["tactile_gen_loss"] = tl
        total = tl if total is None else total + tl
    losses["total_generation_loss"] = total if total is not None else 0.0
    return losses
