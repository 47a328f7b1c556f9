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
from typing import List, Tuple

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


class VisionTokenizer(nn.Module):
    def __init__(self, input_size: int):
        super().__init__()
        self.half_precision_dtype = torch.float16
        self.is_loaded = True
        self.hidden_size = input_size
        self._image_processor = None
        self.patch_stride = 14
        self.conv_stride = 3
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

    def pooled_features(self, pixel_values: torch.Tensor) -> Tuple[torch.Tensor, int, int]:
        """pixel_values f32 [B, 4, H, W] (RGB + all-ones mask channel) -> pooled bf16 [B*h*w, C], (h, w).

        Frozen in the finetune / post-training stages (prismatic.py:460,:493): runs without autograd.  The crop
        to the mask's bounding box (:129-137) is the identity for the all-ones masks the data pipeline emits
        (vla/datasets/datasets.py:68-69); other masks are rejected because downstream asserts 256 tokens anyway."""
        if any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError(
                "VisionTokenizer backward (stage 'pretrain') is not built yet: freeze vision_tower_2d "
                "(stages 'finetune' / 'post-training', as scripts/sft_rlbench.sh does)")
        B, Ct, H, W = pixel_values.shape
        P, cs, Cdim = self.patch_stride, self.conv_stride, self.hidden_size
        h, w = H // (P * cs), W // (P * cs)
        px = pixel_values.float().contiguous()
        lib = _lib.lib()
        s = ops._stream()
        with torch.no_grad():
            wpatch = ops.bf16_of(self.patch_embedding.weight.view(Cdim, -1), pad2d=True)       # [C, 592]
            k_pad = wpatch.shape[1]
            cols = torch.empty((B * h * w * cs * cs, k_pad), dtype=torch.bfloat16, device=px.device)
            check(lib.mla_patchify(ops._p(px), ops._p(cols), C.c_int32(B), C.c_int32(Ct), C.c_int32(H), C.c_int32(W),
                                   C.c_int32(P), C.c_int32(cs), C.c_int32(k_pad), s))
            feat = ops.gemm(cols, wpatch)                                                      # [B*G*9, C]
            del cols
            G = B * h * w
            red = torch.empty((G, Cdim), dtype=torch.bfloat16, device=px.device)
            check(lib.mla_window_mean(ops._p(feat), ops._p(red), C.c_int64(G), C.c_int32(Cdim), C.c_int32(cs * cs), s))
            la = self.local_attention
            qn = ops.layernorm(red, la.q[0].weight, la.q[0].bias, la.q[0].eps)
            q = ops.gemm(qn, ops.bf16_of(la.q[1].weight))
            kvn = ops.layernorm(feat, la.kv[0].weight, la.kv[0].bias, la.kv[0].eps)
            kv = ops.gemm(kvn, ops.bf16_of(la.kv[1].weight))                                   # [B*G*9, 2C]
            del kvn, feat
            agg = torch.empty((G, Cdim), dtype=torch.bfloat16, device=px.device)
            check(lib.mla_local_attn(ops._p(q), ops._p(kv), ops._p(agg), C.c_int64(G), C.c_int32(Cdim),
                                     C.c_int32(la.num_heads), C.c_int32(cs * cs), C.c_float(la.scale), s))
            pooled = ops.gemm(agg, ops.bf16_of(la.proj.weight), bias=ops.bf16_of(la.proj.bias), residual=red)
        return pooled, h, w

    def forward(self, pixel_values: torch.Tensor, modules) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        """Reference signature: returns (list of per-sample [h*w, token] tensors, list of [h, w] LongTensors)."""
        pooled, h, w = self.pooled_features(pixel_values)
        tokens = modules(pooled)                                                              # [B*h*w, token]
        B = pixel_values.shape[0]
        tokens = tokens.view(B, h * w, -1)
        hw = torch.tensor([h, w], dtype=torch.long, device=tokens.device)
        return [tokens[i] for i in range(B)], [hw for _ in range(B)]
