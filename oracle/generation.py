"""Oracle (test infrastructure): the post-training generation heads restated over a reference-keyed state dict.

Follows models/mla/generation/models.py (ImageGenerationModule :68-286, PointCloudGenerationModule :289-386,
TactileGenerationModule :389-430), generation/utils.py, generation/gen_loss.py and
PrismaticVLM.compute_generation_losses (models/vlm/prismatic.py:771-838) op by op, with every dropout / DropPath at
p = 0 (their masks cannot be shared with the reference's fused kernels; the goldens are recorded the same way).

dtype policy as oracle.mla.Ctx: compute_dtype bf16 + flavor "cuda" = the reference's training arithmetic
(autocast: linear / bmm in bf16, layer_norm / softmax / cdist / losses in fp32); flavor "cpu" = torch.autocast("cpu")
(layer_norm and softmax stay bf16) — only to replay the CPU-recorded goldens; compute_dtype fp32 = truth.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _ln(c, pre: str, x: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm under autocast: CUDA -> fp32 in, fp32 out; CPU -> runs in the input dtype with the bf16 params."""
    w, b = c.p(pre + ".weight"), c.p(pre + ".bias")
    if c.dt == torch.float32:
        return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), eps)
    if c.cpu_bf16:
        dt = torch.promote_types(x.dtype, torch.bfloat16)
        return F.layer_norm(x.to(dt), (x.shape[-1],), w.to(dt), b.to(dt), eps)
    return F.layer_norm(x.float(), (x.shape[-1],), w.to(c.dt).float(), b.to(c.dt).float(), eps)


def _lin(c, x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    return F.linear(x.to(c.dt), w.to(c.dt), None if b is None else b.to(c.dt))


def _attn_core(c, q: Tensor, k: Tensor, v: Tensor, H: int, explicit: bool) -> Tensor:
    """q [B,Lq,d], k/v [B,Lk,d] (already projected, compute dtype) -> [B,Lq,d].
    explicit=False: F.scaled_dot_product_attention (need_weights=False path of F.multi_head_attention_forward);
    explicit=True: the need_weights=True path (q * sqrt(1/D), bmm, softmax, bmm)."""
    B, Lq, d = q.shape
    Lk, D = k.shape[1], d // H
    qh = q.view(B, Lq, H, D).transpose(1, 2)
    kh = k.view(B, Lk, H, D).transpose(1, 2)
    vh = v.view(B, Lk, H, D).transpose(1, 2)
    if c.dt == torch.float32:
        p = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(D), dim=-1)
        o = p @ vh
    elif not explicit:
        if c.cpu_bf16:
            o = F.scaled_dot_product_attention(qh, kh, vh)
        else:   # fused kernels: fp32 scores / softmax from bf16 operands, P rounded to bf16 for the PV product
            s = (qh.float() @ kh.float().transpose(-1, -2)) / math.sqrt(D)
            p = torch.softmax(s, dim=-1).to(c.dt)
            o = (p.float() @ vh.float()).to(c.dt)
    else:
        qs = qh * math.sqrt(1.0 / float(D))
        s = (qs.float() @ kh.float().transpose(-1, -2)).to(c.dt)          # bmm -> bf16
        p = torch.softmax(s, dim=-1) if c.cpu_bf16 else torch.softmax(s.float(), dim=-1)
        o = (p.to(c.dt).float() @ vh.float()).to(c.dt)
    return o.transpose(1, 2).reshape(B, Lq, d)


def mha(c, pre: str, xq: Tensor, xkv: Tensor, H: int, explicit: bool = False) -> Tensor:
    """nn.MultiheadAttention(batch_first=True) with packed in_proj (generation/models.py:44 and the decoder layers)."""
    W, bias = c.p(pre + ".in_proj_weight"), c.p(pre + ".in_proj_bias")
    d = W.shape[1]
    q = _lin(c, xq, W[:d], bias[:d])
    k = _lin(c, xkv, W[d:2 * d], bias[d:2 * d])
    v = _lin(c, xkv, W[2 * d:], bias[2 * d:])
    o = _attn_core(c, q, k, v, H, explicit)
    return _lin(c, o, c.p(pre + ".out_proj.weight"), c.p(pre + ".out_proj.bias"))


def decoder_layer(c, pre: str, x: Tensor, mem: Tensor, H: int) -> Tensor:
    """nn.TransformerDecoderLayer(norm_first=False, activation='gelu', batch_first=True), dropout off."""
    x = _ln(c, pre + ".norm1", x + mha(c, pre + ".self_attn", x, x, H))
    x = _ln(c, pre + ".norm2", x + mha(c, pre + ".multihead_attn", x, mem, H))
    ff = _lin(c, F.gelu(_lin(c, x, c.p(pre + ".linear1.weight"), c.p(pre + ".linear1.bias"))),
              c.p(pre + ".linear2.weight"), c.p(pre + ".linear2.bias"))
    return _ln(c, pre + ".norm3", x + ff)


def decoder(c, pre: str, x: Tensor, mem: Tensor, H: int) -> Tensor:
    n = 1 + max(int(k.split(".")[len(pre.split(".")) + 1]) for k in c.sd if k.startswith(pre + ".layers."))
    for i in range(n):
        x = decoder_layer(c, f"{pre}.layers.{i}", x, mem, H)
    return x


def images_to_patches(images: Tensor, ps: int = 42) -> Tensor:
    """generation/utils.py:7-20."""
    B, C, H, W = images.shape
    p = images.unfold(2, ps, ps).unfold(3, ps, ps).contiguous().view(B, C, -1, ps, ps)
    return p.permute(0, 2, 1, 3, 4).contiguous().view(B, -1, C * ps * ps)


def roi_mask(patch_indices: Tensor, ksize: int = 3, grid: int = 16) -> Tensor:
    """create_roi_mask_from_indices + dilate_mask (generation/utils.py:41-70) -> bool [B, grid*grid]."""
    B = patch_indices.shape[0]
    m = torch.zeros(B, grid, grid, dtype=torch.bool)
    bi = torch.arange(B).view(B, 1)
    m[bi, patch_indices[..., 0], patch_indices[..., 1]] = True
    dil = F.max_pool2d(m.float().unsqueeze(1), kernel_size=ksize, stride=1, padding=(ksize - 1) // 2)
    return (dil > 0).squeeze(1).view(B, -1)


def image_head(c, pre: str, hidden: Tensor, img_feat: Tensor, cur_patches: Tensor, roi: Tensor, heads: int = 8,
               delta_clip: float = 5.0, max_shift: float = 8.0, ps: int = 42) -> Dict[str, Tensor]:
    """ImageGenerationModule.forward (:160-231) + _generate_generated_patches (:233-286).  roi bool [B, 256]."""
    B = hidden.shape[0]
    dt = c.dt
    q = c.p(pre + ".image_gen_queries").to(dt).expand(B, -1, -1)
    intent = decoder(c, pre + ".intent_decoder", q, hidden, heads)
    tok = img_feat.clone()
    tok[roi] = c.p(pre + ".mae_mask_token").to(dt).view(-1).to(tok.dtype)
    tok = tok + c.p(pre + ".mae_pos_embed").to(dt)
    gen = decoder(c, pre + ".mae_decoder", tok, intent, heads)
    fn = _ln(c, pre + ".mae_patch_norm", gen.reshape(-1, gen.shape[-1]))
    delta = _lin(c, fn, c.p(pre + ".mae_delta_head.weight"), c.p(pre + ".mae_delta_head.bias"))
    alpha = torch.sigmoid(_lin(c, fn, c.p(pre + ".mae_alpha_head.weight"), c.p(pre + ".mae_alpha_head.bias")).squeeze(-1))
    off = _lin(c, fn, c.p(pre + ".mae_offset_head.weight"), c.p(pre + ".mae_offset_head.bias"))
    P = gen.shape[1]
    E = 3 * ps * ps
    delta = (torch.tanh(delta) * delta_clip).view(B, P, E)
    alpha = alpha.view(B, P)
    off = (torch.tanh(off) * float(max_shift)).view(B, P, 2)
    # _generate_generated_patches
    cur = cur_patches.view(B * P, 3, ps, ps)
    o = off.view(B * P, 2)
    txn = 2.0 * o[:, 0] / float(ps - 1)
    tyn = 2.0 * o[:, 1] / float(ps - 1)
    aff = torch.zeros(B * P, 2, 3, dtype=o.dtype)
    aff[:, 0, 0] = 1.0
    aff[:, 1, 1] = 1.0
    aff[:, 0, 2] = txn
    aff[:, 1, 2] = tyn
    grid = F.affine_grid(aff.float(), size=(B * P, 3, ps, ps), align_corners=True)
    warped = F.grid_sample(cur.float(), grid, mode="bilinear", padding_mode="border", align_corners=True).to(cur.dtype)
    d_img = delta.view(B * P, 3, ps, ps)
    w = 0.95
    roi_pred = (1 - w) * (cur + d_img) + w * d_img
    non_roi = warped + d_img
    rf = roi.view(B * P, 1, 1, 1)
    pred = torch.where(rf, roi_pred, non_roi)
    a = torch.where(roi, torch.ones_like(alpha), alpha).view(B * P, 1, 1, 1)
    blended = a * pred + (1.0 - a) * cur
    return {"image_generation": blended.view(B, P, -1), "generation_roi_mask": roi, "delta_all": delta,
            "alpha_all": alpha, "offset_all": off}


def image_losses(out: Dict[str, Tensor], next_patches: Tensor) -> Dict[str, Tensor]:
    """compute_generation_losses, image part (prismatic.py:779-816)."""
    gen, roi = out["image_generation"], out["generation_roi_mask"]
    losses: Dict[str, Tensor] = {}
    total = 0.0
    pr, gr = gen[roi], next_patches[roi]
    if pr.numel() > 0:
        l = F.mse_loss(pr.float(), gr.float()) + 0.5 * F.l1_loss(pr.float(), gr.float())
        losses["image_roi_generation_loss"] = l
        total = total + l
    pb, gb = gen[~roi], next_patches[~roi]
    if pb.numel() > 0:
        losses["bg_consistency_loss"] = 0.01 * F.l1_loss(pb.float(), gb.float())
        total = total + losses["bg_consistency_loss"]
    dl = -0.1 * out["delta_all"].abs().mean()
    losses["delta_magnitude_reward"] = dl
    losses["image_gen_loss"] = total + dl
    return losses


def pointcloud_head(c, pre: str, hidden: Tensor, heads: int = 8) -> Tensor:
    """PointCloudGenerationModule.forward (:346-386) with current_pointcloud=None (prismatic.py:1098)."""
    B = hidden.shape[0]
    dt = c.dt
    proj = _lin(c, hidden, c.p(pre + ".feature_projector.weight"), c.p(pre + ".feature_projector.bias"))
    agg = proj.mean(dim=1)
    pos = c.p(pre + ".pos_embed").to(dt)
    G, td = pos.shape[1], pos.shape[2]
    x = _lin(c, agg, c.p(pre + ".seq_to_patch.weight"), c.p(pre + ".seq_to_patch.bias")).reshape(B, G, td)
    pos = pos.expand(B, -1, -1)
    n_blk = 1 + max(int(k.split(".")[len(pre.split(".")) + 1]) for k in c.sd if k.startswith(pre + ".decoder_blocks."))
    for i in range(n_blk):
        b = f"{pre}.decoder_blocks.{i}"
        xn = _ln(c, b + ".norm1", x + pos)
        x = x + mha(c, b + ".attn", xn, xn, heads, explicit=True)
        m = _lin(c, F.gelu(_lin(c, _ln(c, b + ".norm2", x), c.p(b + ".mlp.0.weight"), c.p(b + ".mlp.0.bias"))),
                 c.p(b + ".mlp.3.weight"), c.p(b + ".mlp.3.bias"))
        x = x + m
    fp = pre + ".future_predictor"
    y = _lin(c, x.reshape(B * G, td), c.p(fp + ".0.weight")[:, :, 0], c.p(fp + ".0.bias"))
    yf = y.float()                                                     # BatchNorm1d, train-mode batch statistics
    mu, var = yf.mean(0), yf.var(0, unbiased=False)
    y = ((yf - mu) * torch.rsqrt(var + 1e-5) * c.p(fp + ".1.weight").to(dt).float() + c.p(fp + ".1.bias").to(dt).float()).to(y.dtype)
    y = torch.relu(y)
    y = _lin(c, y, c.p(fp + ".3.weight")[:, :, 0], c.p(fp + ".3.bias"))
    M = y.shape[1] // 3
    return y.reshape(B, G * M, 3)


def chamfer_l2(pred: Tensor, gt: Tensor) -> Tensor:
    """generation/gen_loss.py:12-18 (cdist is on autocast's fp32 list)."""
    d = torch.cdist(pred.float(), gt.float())
    return (d.min(dim=2)[0].mean(dim=1) + d.min(dim=1)[0].mean(dim=1)).mean()


def tactile_head(c, pre: str, hidden: Tensor, heads: int = 4) -> Tensor:
    """TactileGenerationModule.forward (:417-430)."""
    B = hidden.shape[0]
    q = c.p(pre + ".tactile_query").to(c.dt).expand(B, -1, -1)
    mem = _lin(c, hidden, c.p(pre + ".feature_projector.weight"), c.p(pre + ".feature_projector.bias"))
    dec = decoder(c, pre + ".decoder", q, mem, heads)
    return _lin(c, dec.squeeze(1), c.p(pre + ".output_head.weight"), c.p(pre + ".output_head.bias"))
