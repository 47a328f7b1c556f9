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


def _tower_params(tok) -> List[torch.nn.Parameter]:
    """Every parameter the forward reads, in a fixed order (raw embed, per stage / block conv+BN pairs, proj)."""
    enc = tok.patch_embed.EncP
    ps = [enc.raw_point_embed.net[0].weight, enc.raw_point_embed.net[1].weight, enc.raw_point_embed.net[1].bias]
    for lga in enc.LGA_list:
        for blk in lga.linear2:
            ps += [blk.net1[0].weight, blk.net1[0].bias, blk.net1[1].weight, blk.net1[1].bias,
                   blk.net2[0].weight, blk.net2[0].bias, blk.net2[1].weight, blk.net2[1].bias]
    return ps + [tok.proj.weight, tok.proj.bias]


def point_tokenizer_forward(tok, p: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """p f32 [B, N, 3] -> (patch tokens bf16 [B, G, 768], centres f32 [B, G, 3]).

    Frozen stages run the kernels without autograd; stage 'pretrain' (prismatic.py:431-432) wraps the same kernel
    sequence in one autograd node (`_PointTowerFn`) that keeps the pre-BatchNorm activations for its backward."""
    params = _tower_params(tok)
    if torch.is_grad_enabled() and any(q.requires_grad for q in params):
        if not tok.training:
            raise NotImplementedError("PointTokenizer backward with eval-mode BatchNorm is not part of the training path")
        return _PointTowerFn.apply(tok, p, *params)
    with torch.no_grad():
        return _forward_impl(tok, p, None)


def _forward_impl(tok, p: torch.Tensor, tape: Optional[dict]) -> Tuple[torch.Tensor, torch.Tensor]:
    """The kernel sequence.  Without a tape BatchNorm+ReLU run in place and the residual stream is overwritten block by
    block; with one, every buffer the backward reads (pre-BN conv outputs, block inputs/outputs, indices) is kept."""
    keep = tape is not None
    enc = tok.patch_embed.EncP
    lib, s = _lib.lib(), ops._stream()
    dev = tok.proj.weight.device
    xyz = p.to(dev).float().contiguous()
    B, N, _ = xyz.shape
    # raw point embedding: conv1d(3 -> embed, k=1, no bias) + BN + ReLU        (Point_PN.py:173-185,:286)
    conv, bn = enc.raw_point_embed.net[0], enc.raw_point_embed.net[1]
    w = ops.bf16_of(conv.weight.view(conv.out_channels, -1), pad2d=True)                   # [96, 8]
    xin = ops.pad_cols_bf16(xyz.view(B * N, 3), w.shape[1])
    y = ops.gemm(xin, w)
    coef = _batchnorm(y, bn, tok.training)
    feat = torch.empty_like(y) if keep else y
    check(lib.mla_bn_relu(ops._p(y), ops._p(coef), ops._p(bn.weight), ops._p(bn.bias), ops._p(feat),
                          C.c_int64(B * N), C.c_int32(y.shape[1]), s))
    if keep:
        tape["raw"] = dict(xin=xin, y=y, coef=coef)
        tape["stages"] = []
    feat_bf16 = True                                                                      # [B*N, C] bf16
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
        st = dict(fps_idx=fps_idx, knn_idx=knn_idx, N=Ncur, G=G, K=K, C=Ccur, feat_bf16=feat_bf16, blocks=[])
        n_blocks = len(lga.linear2)
        pooled = None
        for j, blk in enumerate(lga.linear2):
            c1, b1 = blk.net1[0], blk.net1[1]
            c2, b2 = blk.net2[0], blk.net2[1]
            y1 = ops.gemm(xb, ops.bf16_of(c1.weight.view(c1.out_channels, -1)), bias=ops.bf16_of(c1.bias))
            coef1 = _batchnorm(y1, b1, tok.training)
            a1 = torch.empty_like(y1) if keep else y1
            check(lib.mla_bn_relu(ops._p(y1), ops._p(coef1), ops._p(b1.weight), ops._p(b1.bias), ops._p(a1),
                                  C.c_int64(rows), C.c_int32(y1.shape[1]), s))
            y2 = ops.gemm(a1, ops.bf16_of(c2.weight.view(c2.out_channels, -1)), bias=ops.bf16_of(c2.bias))
            coef2 = _batchnorm(y2, b2, tok.training)
            last = j == n_blocks - 1
            if last:
                pooled = torch.empty((B * G, out_dim), dtype=torch.float32, device=dev)
            xf_out = torch.empty_like(xf) if keep else (None if last else xf)
            xb_out = None if last else (torch.empty_like(xb) if keep else xb)
            check(lib.mla_bn_res_relu(ops._p(y2), ops._p(coef2), ops._p(b2.weight), ops._p(b2.bias), ops._p(xf),
                                      ops._p(xf_out), ops._p(xb_out), ops._p(pooled), C.c_int64(B * G), C.c_int32(K),
                                      C.c_int32(out_dim), s))
            if keep:
                st["blocks"].append(dict(xb_in=xb, y1=y1, coef1=coef1, y2=y2, coef2=coef2, xnew=xf_out))
                xf, xb = xf_out, xb_out
        if keep:
            tape["stages"].append(st)
        feat, feat_bf16 = pooled, False                                                   # [B*G, out_dim] f32
        xyz, Ncur, Ccur = centers, G, out_dim
    tokens_in = ops.cast_bf16(feat)
    tokens = ops.gemm(tokens_in, ops.bf16_of(tok.proj.weight), bias=ops.bf16_of(tok.proj.bias))
    if keep:
        tape["tokens_in"] = tokens_in
    return tokens.view(B, Ncur, -1), xyz


# ---------------------------------------------------------------------------------------------------- backward
def _wgrad(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """dW f32 [n_out, n_in] = dy^T x for bf16 [rows, n_out], [rows, n_in].  The 1x1 convs of the tokenizer have
    ~1e6 neighbour rows against <= 384 channels, i.e. a single output tile with a very long reduction: eight
    interleaved row classes are laid side by side (a free view), one GEMM fills an 8x8 grid of tiles and the
    diagonal blocks are summed (mla_diag_block_sum) — 8x the tensor work, which is idle anyway, for 8-64x the SMs."""
    rows, m = dy.shape
    n = x.shape[1]
    parts = 8
    if rows < 65536 or rows % parts or not (dy.is_contiguous() and x.is_contiguous()):
        return ops.gemm(dy, x, a_mn=True, b_mn=True, out_dtype=torch.float32)
    big = ops.gemm(dy.view(rows // parts, parts * m), x.view(rows // parts, parts * n), a_mn=True, b_mn=True,
                   out_dtype=torch.float32)
    out = torch.empty((m, n), dtype=torch.float32, device=dy.device)
    check(_lib.lib().mla_diag_block_sum(ops._p(big), ops._p(out), C.c_int32(m), C.c_int32(n), C.c_int32(parts),
                                        ops._stream()))
    return out


def _bn_bwd(up: torch.Tensor, y: torch.Tensor, coef: torch.Tensor, bn, mode: int, xnew: Optional[torch.Tensor] = None,
            want_res: bool = False):
    """Train-mode BatchNorm backward over rows (mla_bn_bwd).  Returns (dy bf16, d_weight, d_bias, d_residual)."""
    rows, c = y.shape
    sums = torch.empty(2 * c, dtype=torch.float32, device=y.device)
    dy = torch.empty_like(y)
    dres = torch.empty((rows, c), dtype=torch.float32, device=y.device) if want_res else None
    check(_lib.lib().mla_bn_bwd(ops._p(up), C.c_int32(int(up.dtype == torch.float32)), ops._p(y), ops._p(coef),
                                ops._p(bn.weight), ops._p(bn.bias), ops._p(xnew), C.c_int32(mode), ops._p(sums),
                                ops._p(dy), ops._p(dres), C.c_int64(rows), C.c_int32(c), ops._stream()))
    return dy, sums[c:], sums[:c], dres


class _PointTowerFn(torch.autograd.Function):
    """PointTokenizer.forward as ONE autograd node (stage 'pretrain').  backward follows the reference's autograd
    through EncP.forward (Point_PN.py:284-298): proj Linear, per stage max-pool -> Linear2Layer blocks (conv1x1 / BN /
    ReLU / residual, :188-219) -> neighbour gathers (:116-122), then the raw point embedding (:173-185).  FPS / kNN
    indices and the positional term carry no parameters; the point coordinates get no gradient."""

    @staticmethod
    def forward(ctx, tok, p, *params):
        tape: dict = {}
        tokens, centers = _forward_impl(tok, p, tape)
        ctx.tok, ctx.tape, ctx.n_params = tok, tape, len(params)
        ctx.mark_non_differentiable(centers)
        return tokens, centers

    @staticmethod
    def backward(ctx, d_tokens, _d_centers):
        tok, t = ctx.tok, ctx.tape
        ctx.tape = None
        enc = tok.patch_embed.EncP
        lib, s = _lib.lib(), ops._stream()
        f32 = torch.float32
        grads: dict = {}
        dt = d_tokens.reshape(-1, d_tokens.shape[-1]).contiguous()
        B = d_tokens.shape[0]
        # proj: Linear(384 -> 768) on the pooled stage output (pointvit.py:80)
        grads[id(tok.proj.bias)] = ops.colsum(dt, dt.shape[1])
        grads[id(tok.proj.weight)] = ops.gemm(dt, t["tokens_in"], a_mn=True, b_mn=True, out_dtype=f32)
        d16 = ops.gemm(dt, ops.bf16_of(tok.proj.weight), b_mn=True)
        d_pooled = torch.empty(d16.shape, dtype=f32, device=d16.device)
        check(lib.mla_cast_bf16_f32(ops._p(d16), ops._p(d_pooled), C.c_int64(d16.numel()), s))
        for i in reversed(range(enc.num_stages)):
            st = t["stages"][i]
            G, K, Ccur, Ncur = st["G"], st["K"], st["C"], st["N"]
            D = 2 * Ccur
            blocks = st["blocks"]
            d_x = torch.empty_like(blocks[-1]["xnew"])
            check(lib.mla_maxpool_bwd(ops._p(blocks[-1]["xnew"]), ops._p(d_pooled), ops._p(d_x), C.c_int64(B * G),
                                      C.c_int32(K), C.c_int32(D), s))
            rows = B * G * K
            for j in reversed(range(len(blocks))):
                bk, blk = blocks[j], enc.LGA_list[i].linear2[j]
                c1, b1, c2, b2 = blk.net1[0], blk.net1[1], blk.net2[0], blk.net2[1]
                # x_new = relu(bn2(conv2(a1)) + x)
                dy2, gw, gb, d_res = _bn_bwd(d_x, bk["y2"], bk["coef2"], b2, 2, xnew=bk["xnew"], want_res=True)
                grads[id(b2.weight)], grads[id(b2.bias)] = gw, gb
                a1 = torch.empty_like(bk["y1"])                                            # recomputed relu(bn1(y1))
                check(lib.mla_bn_relu(ops._p(bk["y1"]), ops._p(bk["coef1"]), ops._p(b1.weight), ops._p(b1.bias),
                                      ops._p(a1), C.c_int64(rows), C.c_int32(a1.shape[1]), s))
                grads[id(c2.bias)] = ops.colsum(dy2, D)
                grads[id(c2.weight)] = _wgrad(dy2, a1).reshape(c2.weight.shape)
                d_a1 = ops.gemm(dy2, ops.bf16_of(c2.weight.view(c2.out_channels, -1)), b_mn=True)
                del a1, dy2
                # a1 = relu(bn1(conv1(x_bf16)))
                dy1, gw, gb, _ = _bn_bwd(d_a1, bk["y1"], bk["coef1"], b1, 1)
                grads[id(b1.weight)], grads[id(b1.bias)] = gw, gb
                grads[id(c1.bias)] = ops.colsum(dy1, dy1.shape[1])
                grads[id(c1.weight)] = _wgrad(dy1, bk["xb_in"]).reshape(c1.weight.shape)
                d_xb = ops.gemm(dy1, ops.bf16_of(c1.weight.view(c1.out_channels, -1)), b_mn=True)
                # d(block input) = residual path (f32) + conv path (bf16)
                check(lib.mla_add_f32_bf16(ops._p(d_res), ops._p(d_xb), ops._p(d_x), C.c_int64(d_res.numel()),
                                           C.c_int32(0), s))                               # d_x's old content is dead
                del d_xb, dy1, d_a1, d_res
            d_feat = torch.zeros((B * Ncur, Ccur), dtype=f32, device=d_x.device)
            check(lib.mla_group_pose_bwd(ops._p(d_x), ops._p(st["fps_idx"]), ops._p(st["knn_idx"]), ops._p(d_feat),
                                         C.c_int32(B), C.c_int32(Ncur), C.c_int32(G), C.c_int32(K), C.c_int32(Ccur), s))
            d_pooled = d_feat
            del d_x
        # raw point embedding: relu(bn(conv1d(xyz)))
        conv, bn = enc.raw_point_embed.net[0], enc.raw_point_embed.net[1]
        raw = t["raw"]
        dy0, gw, gb, _ = _bn_bwd(d_pooled, raw["y"], raw["coef"], bn, 1)
        grads[id(bn.weight)], grads[id(bn.bias)] = gw, gb
        gw0 = ops.gemm(dy0, raw["xin"], a_mn=True, b_mn=True, out_dtype=f32)
        grads[id(conv.weight)] = gw0[:conv.out_channels, :conv.weight[0].numel()].reshape(conv.weight.shape)
        need = ctx.needs_input_grad[2:]
        return (None, None, *[grads.get(id(q)) if n else None for q, n in zip(_tower_params(tok), need)])
