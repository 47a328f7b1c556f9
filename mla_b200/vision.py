"""Encoder-free image tokenizer (reference: models/mla/image/vision_tokenizer.py:92-160) on our kernels.

patchify (im2col + tcgen05 GEMM) -> 3x3 window pooling -> LocalAttention (LN + q / kv GEMMs + 9-way attention +
proj GEMM with fused residual) -> projector (applied by the caller's `modules` argument, as in the reference).

Same parameters and names as the reference module (patch_embedding, class_embedding, split_embedding,
local_attention.{q,kv,proj}, global_attention.*).  GlobalAttention is computed and discarded by the reference
(:141-142,:149), so it has no effect on any output and is not evaluated here; its parameters are kept for
checkpoint compatibility and, as in the reference, never receive gradients.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import check


class _AttnParams(nn.Module):
    def __init__(self, input_size: int):
        super().__init__()
        self.q = nn.Sequential(nn.LayerNorm(input_size), nn.Linear(input_size, input_size, bias=False))
        self.kv = nn.Sequential(nn.LayerNorm(input_size), nn.Linear(input_size, input_size * 2, bias=False))
        self.proj = nn.Linear(input_size, input_size)


class LocalAttention(_AttnParams):
    def __init__(self, input_size: int, conv_stride: int, num_heads: int = 8):
        super().__init__(input_size)
        self.conv_stride, self.num_heads, self.scale = conv_stride, num_heads, input_size ** -0.5


class GlobalAttention(_AttnParams):
    def __init__(self, input_size: int, num_heads: int = 8):
        super().__init__(input_size)
        self.num_heads, self.scale = num_heads, input_size ** -0.5


def layernorm_bwd(dy: torch.Tensor, x: torch.Tensor, w: torch.Tensor, eps: float, dres=None, dgrp=None, win: int = 1):
    """Backward of ops.layernorm on bf16 rows: returns (dx bf16, dw f32, db f32); see mla_layernorm_bwd."""
    rows, h = x.shape
    dx = torch.empty_like(x)
    dw = torch.zeros(h, dtype=torch.float32, device=x.device)
    db = torch.zeros(h, dtype=torch.float32, device=x.device)
    check(_lib.lib().mla_layernorm_bwd(ops._p(dy.contiguous()), ops._p(x), ops._p(w.detach()), ops._p(dres), ops._p(dgrp),
                                       C.c_int32(win), ops._p(dx), ops._p(dw), ops._p(db), C.c_int64(rows), C.c_int32(h),
                                       C.c_float(eps), ops._stream()))
    return dx, dw, db


class _VisionTowerFn(torch.autograd.Function):
    """VisionTokenizer.forward up to the pooled patch features as ONE autograd node (stage 'pretrain').

    backward = the reference's autograd through LocalAttention.forward (vision_tokenizer.py:27-47) and the patchify
    conv (:124): every linear's dgrad / wgrad on the tcgen05 GEMM, the 9-way attention core, the two LayerNorms and
    the avg-pool / residual fan-in on the kernels of csrc/tower_bwd.cu.  Pixels get no gradient."""

    @staticmethod
    def forward(ctx, tower, px, *params):
        tape: dict = {}
        pooled, _, _ = tower._pooled_impl(px, tape)
        ctx.tower, ctx.tape = tower, tape
        return pooled

    @staticmethod
    def backward(ctx, d):
        tw, t = ctx.tower, ctx.tape
        ctx.tape = None
        la = tw.local_attention
        need = ctx.needs_input_grad[2:]
        Cdim, G, win = tw.hidden_size, t["G"], t["win"]
        d = d.contiguous()
        f32 = torch.float32
        wg = lambda dy, x: ops.gemm(dy, x, a_mn=True, b_mn=True, out_dtype=f32)
        # proj (+ residual): pooled = agg W^T + b + red
        d_b = ops.colsum(d, Cdim) if need[8] else None
        d_wproj = wg(d, t["agg"]) if need[7] else None
        d_agg = ops.gemm(d, ops.bf16_of(la.proj.weight), b_mn=True)
        # 9-way attention core
        dq = torch.empty_like(t["q"])
        dkv = torch.empty_like(t["kv"])
        check(_lib.lib().mla_local_attn_bwd(ops._p(t["q"]), ops._p(t["kv"]), ops._p(d_agg), ops._p(dq), ops._p(dkv),
                                            C.c_int64(G), C.c_int32(Cdim), C.c_int32(la.num_heads), C.c_int32(win),
                                            C.c_float(la.scale), ops._stream()))
        # query path: q = LN(red) Wq^T ; red also feeds the residual
        d_wq = wg(dq, t["qn"]) if need[3] else None
        d_qn = ops.gemm(dq, ops.bf16_of(la.q[1].weight), b_mn=True)
        d_red, d_lnq_w, d_lnq_b = layernorm_bwd(d_qn, t["red"], la.q[0].weight, la.q[0].eps, dres=d)
        # key/value path: kv = LN(feat) Wkv^T ; feat also feeds the 3x3 average (red)
        d_wkv = wg(dkv, t["kvn"]) if need[6] else None
        d_kvn = ops.gemm(dkv, ops.bf16_of(la.kv[1].weight), b_mn=True)
        d_feat, d_lnkv_w, d_lnkv_b = layernorm_bwd(d_kvn, t["feat"], la.kv[0].weight, la.kv[0].eps, dgrp=d_red, win=win)
        # patchify conv as a GEMM over im2col rows
        d_wpatch = None
        if need[0]:
            pw = tw.patch_embedding.weight
            d_wpatch = wg(d_feat, t["cols"])[:Cdim, :pw[0].numel()].reshape(pw.shape)
        grads = [d_wpatch, d_lnq_w, d_lnq_b, d_wq, d_lnkv_w, d_lnkv_b, d_wkv, d_wproj, d_b]
        return (None, None, *[g if n else None for g, n in zip(grads, need)])


class VisionTokenizer(nn.Module):
    def __init__(self, input_size: int):
        super().__init__()
        self.half_precision_dtype = torch.float16
        self.is_loaded = True
        self.hidden_size = input_size
        self._image_processor = None
        self.patch_stride = 14
        self.conv_stride = 3
        self.image_size = 672          # CLIPImageProcessor(size=672, crop_size=672), vision_tokenizer.py:98-105
        self.patch_embedding = nn.Conv2d(3, input_size, kernel_size=14, stride=14, bias=False)
        self.class_embedding = nn.Parameter(torch.randn(input_size))
        self.split_embedding = nn.Parameter(torch.randn(input_size))
        self.local_attention = LocalAttention(input_size, self.conv_stride)
        self.global_attention = GlobalAttention(input_size)

    @property
    def image_processor(self):
        """CLIPImageProcessor(672) as the data pipeline expects (vision_tokenizer.py:98-105, scripts/train.py:346)."""
        if self._image_processor is None:
            from transformers import CLIPImageProcessor
            self._image_processor = CLIPImageProcessor(do_resize=True, size=672, do_center_crop=True, crop_size=672,
                                                       do_normalize=True, do_rescale=True)
        return self._image_processor

    @property
    def dtype(self):
        return self.patch_embedding.weight.dtype

    @property
    def device(self):
        return self.patch_embedding.weight.device

    def _tower_params(self):
        la = self.local_attention
        return [self.patch_embedding.weight, la.q[0].weight, la.q[0].bias, la.q[1].weight, la.kv[0].weight,
                la.kv[0].bias, la.kv[1].weight, la.proj.weight, la.proj.bias]

    def _pooled_impl(self, pixel_values: torch.Tensor, tape: Optional[dict]) -> Tuple[torch.Tensor, int, int]:
        """The kernel sequence of the forward; with `tape` the intermediates the backward needs are kept in it."""
        P, cs, Cdim = self.patch_stride, self.conv_stride, self.hidden_size
        lib = _lib.lib()
        s = ops._stream()
        wpatch = ops.bf16_of(self.patch_embedding.weight.view(Cdim, -1), pad2d=True)           # [C, 592]
        k_pad = wpatch.shape[1]
        if pixel_values.dtype == torch.uint8:
            # raw camera frames [B, H, W, 3]: CLIP preprocessing fused into the im2col (csrc/preprocess.cu)
            from .preprocess import patchify_frames
            B = pixel_values.shape[0]
            h = w = self.image_size // (P * cs)
            px = pixel_values
            cols = patchify_frames(pixel_values, self.image_size, P, cs, k_pad)
        else:
            B, Ct, H, W = pixel_values.shape
            h, w = H // (P * cs), W // (P * cs)
            px = pixel_values.float().contiguous()
            cols = torch.empty((B * h * w * cs * cs, k_pad), dtype=torch.bfloat16, device=px.device)
            check(lib.mla_patchify(ops._p(px), ops._p(cols), C.c_int32(B), C.c_int32(Ct), C.c_int32(H), C.c_int32(W),
                                   C.c_int32(P), C.c_int32(cs), C.c_int32(k_pad), s))
        feat = ops.gemm(cols, wpatch)                                                          # [B*G*9, C]
        G = B * h * w
        red = torch.empty((G, Cdim), dtype=torch.bfloat16, device=px.device)
        check(lib.mla_window_mean(ops._p(feat), ops._p(red), C.c_int64(G), C.c_int32(Cdim), C.c_int32(cs * cs), s))
        la = self.local_attention
        qn = ops.layernorm(red, la.q[0].weight, la.q[0].bias, la.q[0].eps)
        q = ops.gemm(qn, ops.bf16_of(la.q[1].weight))
        kvn = ops.layernorm(feat, la.kv[0].weight, la.kv[0].bias, la.kv[0].eps)
        kv = ops.gemm(kvn, ops.bf16_of(la.kv[1].weight))                                       # [B*G*9, 2C]
        agg = torch.empty((G, Cdim), dtype=torch.bfloat16, device=px.device)
        check(lib.mla_local_attn(ops._p(q), ops._p(kv), ops._p(agg), C.c_int64(G), C.c_int32(Cdim),
                                 C.c_int32(la.num_heads), C.c_int32(cs * cs), C.c_float(la.scale), s))
        pooled = ops.gemm(agg, ops.bf16_of(la.proj.weight), bias=ops.bf16_of(la.proj.bias), residual=red)
        if tape is not None:
            tape.update(cols=cols, feat=feat, red=red, qn=qn, q=q, kvn=kvn, kv=kv, agg=agg, G=G, win=cs * cs)
        return pooled, h, w

    def pooled_features(self, pixel_values: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
        """pixel_values f32 [B, 4, H, W] (RGB + all-ones mask channel) -> pooled bf16 [B*h*w, C], (h, w).
        Also accepts the raw uint8 camera frames [B, Hc, Wc, 3]: the reference's CPU-side CLIP preprocessing is then
        applied inside the im2col kernel, bit-exactly, and the 7.2 MB/sample f32 image is never materialised.

        Frozen in the finetune / post-training stages (prismatic.py:460,:493): runs without autograd.  Stage
        'pretrain' trains the tokenizer (:427-428): the same kernels run inside one autograd node whose backward is
        `_VisionTowerFn.backward`.  The crop to the mask's bounding box (:129-137) is the identity for the all-ones
        masks the data pipeline emits (vla/datasets/datasets.py:68-69); other masks are rejected because downstream
        asserts 256 tokens anyway."""
        P, cs = self.patch_stride, self.conv_stride
        if pixel_values.dtype == torch.uint8:
            h = w = self.image_size // (P * cs)
        else:
            h, w = pixel_values.shape[2] // (P * cs), pixel_values.shape[3] // (P * cs)
        params = self._tower_params()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _VisionTowerFn.apply(self, pixel_values, *params), h, w
        with torch.no_grad():
            return self._pooled_impl(pixel_values, None)

    def forward(self, pixel_values: torch.Tensor, modules) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        """Reference signature: returns (list of per-sample [h*w, token] tensors, list of [h, w] LongTensors)."""
        pooled, h, w = self.pooled_features(pixel_values)
        tokens = modules(pooled)                                                              # [B*h*w, token]
        B = pixel_values.shape[0]
        tokens = tokens.view(B, h * w, -1)
        hw = torch.tensor([h, w], dtype=torch.long, device=tokens.device)
        return [tokens[i] for i in range(B)], [hw for _ in range(B)]
