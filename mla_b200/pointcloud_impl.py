"""CUDA forward of the point-cloud tokenizer (PointTokenizer.forward -> Point_PN_scan -> EncP.forward,
models/mla/pointcloud/backbone/pointvit.py:59-82, Point_PN.py:284-298).

Frozen in the finetune / post-training stages (prismatic.py:464,:497) but kept in train mode by the training loop
(base_strategy_mla.py:291): BatchNorm uses batch statistics and updates its running buffers; no gradients flow.
Rows are laid out (batch, group, neighbour) x channels so every 1x1 conv is one tcgen05 GEMM.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch

from . import _lib, ops
from ._lib import check

# test hooks: fixed FPS start indices / neighbour sets (torch.topk ties are implementation-defined in the reference)
_OVERRIDE = {"fps_starts": None, "knn_idx": None}


def set_test_overrides(fps_starts: Optional[List[torch.Tensor]] = None, knn_idx: Optional[List[torch.Tensor]] = None):
    _OVERRIDE["fps_starts"], _OVERRIDE["knn_idx"] = fps_starts, knn_idx


def _batchnorm(y: torch.Tensor, bn: torch.nn.BatchNorm2d, training: bool):
    """Returns coef (mean | invstd) of bf16 y [rows, c]; updates running stats like nn.BatchNorm in train mode."""
    rows, c = y.shape
    lib, s = _lib.lib(), ops._stream()
    coef = torch.empty(2 * c, dtype=torch.float32, device=y.device)
    if training or bn.running_mean is None:
        sums = torch.zeros(2 * c, dtype=torch.float32, device=y.device)
        check(lib.mla_bn_stats(ops._p(y), ops._p(sums), C.c_int64(rows), C.c_int32(c), s))
        upd = training and bn.track_running_stats and bn.running_mean is not None
        mom = bn.momentum if bn.momentum is not None else 0.1
        check(lib.mla_bn_finalize(ops._p(sums), ops._p(coef), ops._p(bn.running_mean if upd else None),
                                  ops._p(bn.running_var if upd else None), C.c_int64(rows), C.c_int32(c),
                                  C.c_float(bn.eps), C.c_float(mom), s))
        if upd and bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
    else:
        coef[:c] = bn.running_mean
        coef[c:] = torch.rsqrt(bn.running_var + bn.eps)
    return coef


def point_tokenizer_forward(tok, p: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """p f32 [B, N, 3] -> (patch tokens bf16 [B, G, 768], centres f32 [B, G, 3])."""
    if any(q.requires_grad for q in tok.parameters()):
        raise NotImplementedError("PointTokenizer backward (stage 'pretrain') is not built yet: freeze vision_tower_3d")
    enc = tok.patch_embed.EncP
    lib, s = _lib.lib(), ops._stream()
    dev = tok.proj.weight.device
    xyz = p.to(dev).float().contiguous()
    B, N, _ = xyz.shape
    with torch.no_grad():
        # raw point embedding: conv1d(3 -> embed, k=1, no bias) + BN + ReLU        (Point_PN.py:173-185,:286)
        conv, bn = enc.raw_point_embed.net[0], enc.raw_point_embed.net[1]
        w = ops.bf16_of(conv.weight.view(conv.out_channels, -1), pad2d=True)                   # [96, 8]
        xin = ops.pad_cols_bf16(xyz.view(B * N, 3), w.shape[1])
        y = ops.gemm(xin, w)
        coef = _batchnorm(y, bn, tok.training)
        check(lib.mla_bn_relu(ops._p(y), ops._p(coef), ops._p(bn.weight), ops._p(bn.bias), ops._p(y),
                              C.c_int64(B * N), C.c_int32(y.shape[1]), s))
        feat, feat_bf16 = y, True                                                             # [B*N, C] bf16
        Ccur, Ncur = y.shape[1], N
        K = enc.k_neighbors
        for i in range(enc.num_stages):
            G = Ncur // 2
            if _OVERRIDE["fps_starts"] is not None:
                start = _OVERRIDE["fps_starts"][i].to(dev).to(torch.int64).contiguous()
            else:
                start = torch.randint(0, Ncur, (B,), dtype=torch.long, device=dev)            # Point_PN.py:10
            fps_idx = torch.empty((B, G), dtype=torch.int32, device=dev)
            centers = torch.empty((B, G, 3), dtype=torch.float32, device=dev)
            check(lib.mla_fps(ops._p(xyz), ops._p(start), ops._p(fps_idx), ops._p(centers), C.c_int32(B),
                              C.c_int32(Ncur), C.c_int32(G), s))
            if _OVERRIDE["knn_idx"] is not None:
                knn_idx = _OVERRIDE["knn_idx"][i].to(dev).to(torch.int32).contiguous()
            else:
                knn_idx = torch.empty((B, G, K), dtype=torch.int32, device=dev)
                check(lib.mla_knn(ops._p(xyz), ops._p(centers), ops._p(knn_idx), C.c_int32(B), C.c_int32(Ncur),
                                  C.c_int32(G), C.c_int32(K), C.c_int32(1), s))
            lga = enc.LGA_list[i]
            out_dim = 2 * Ccur
            fd = out_dim // 6
            dim_embed = torch.pow(torch.tensor(float(lga.alpha)), torch.arange(fd, dtype=torch.float32) / fd).to(dev)
            rows = B * G * K
            xf = torch.empty((rows, out_dim), dtype=torch.float32, device=dev)
            xb = torch.empty((rows, out_dim), dtype=torch.bfloat16, device=dev)
            check(lib.mla_group_pose(ops._p(xyz), ops._p(feat), C.c_int32(int(feat_bf16)), ops._p(fps_idx),
                                     ops._p(knn_idx), ops._p(dim_embed), ops._p(xf), ops._p(xb), C.c_int32(B),
                                     C.c_int32(Ncur), C.c_int32(G), C.c_int32(K), C.c_int32(Ccur),
                                     C.c_float(float(lga.beta)), s))
            n_blocks = len(lga.linear2)
            pooled = None
            for j, blk in enumerate(lga.linear2):
                c1, b1 = blk.net1[0], blk.net1[1]
                c2, b2 = blk.net2[0], blk.net2[1]
                y1 = ops.gemm(xb, ops.bf16_of(c1.weight.view(c1.out_channels, -1)), bias=ops.bf16_of(c1.bias))
                coef1 = _batchnorm(y1, b1, tok.training)
                check(lib.mla_bn_relu(ops._p(y1), ops._p(coef1), ops._p(b1.weight), ops._p(b1.bias), ops._p(y1),
                                      C.c_int64(rows), C.c_int32(y1.shape[1]), s))
                y2 = ops.gemm(y1, ops.bf16_of(c2.weight.view(c2.out_channels, -1)), bias=ops.bf16_of(c2.bias))
                coef2 = _batchnorm(y2, b2, tok.training)
                last = j == n_blocks - 1
                if last:
                    pooled = torch.empty((B * G, out_dim), dtype=torch.float32, device=dev)
                check(lib.mla_bn_res_relu(ops._p(y2), ops._p(coef2), ops._p(b2.weight), ops._p(b2.bias), ops._p(xf),
                                          ops._p(None if last else xf), ops._p(None if last else xb),
                                          ops._p(pooled), C.c_int64(B * G), C.c_int32(K), C.c_int32(out_dim), s))
            feat, feat_bf16 = pooled, False                                                   # [B*G, out_dim] f32
            xyz, Ncur, Ccur = centers, G, out_dim
        tokens_in = ops.cast_bf16(feat)
        tokens = ops.gemm(tokens_in, ops.bf16_of(tok.proj.weight), bias=ops.bf16_of(tok.proj.bias))
    return tokens.view(B, Ncur, -1), xyz
