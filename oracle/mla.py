"""Oracle (test infrastructure): the MLA training forward restated over a reference-keyed state_dict.

`forward(sd, batch, cfg, draws)` follows MLA.forward -> PrismaticVLM.forward -> LlamaForCausalLM.forward of the
reference op by op (citations inline) but takes the parameters as a plain dict with the REFERENCE's state_dict keys
(`vlm.llm_backbone.llm.model.layers.0.self_attn.q_proj.weight`, ...), so the same weights can be fed to the
reference (tests/golden/make_golden.py), to this oracle and to the CUDA implementation.

dtype policy: `compute_dtype=torch.bfloat16` emulates the reference's autocast(bf16) arithmetic — every nn.Linear /
conv / matmul input is cast to bf16 and its output rounded to bf16, while the ops on autocast's fp32 list
(layer_norm, softmax, norm, cross_entropy, mse) run in fp32; `torch.float32` gives the exact-arithmetic truth.
Random draws (diffusion noise, timesteps, FPS start indices) are inputs, never drawn here.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import llama as L

Tensor = torch.Tensor


class Ctx:
    """Parameter access + autocast emulation."""

    def __init__(self, sd: Dict[str, Tensor], compute_dtype=torch.bfloat16, flavor: str = "cuda"):
        """flavor: which autocast op lists apply.  "cuda" = torch.autocast("cuda") — layer_norm / softmax / norm /
        cross_entropy / mse run in fp32 (what the reference does in training); "cpu" = torch.autocast("cpu"), whose
        fp32 list is much shorter (softmax, normalize stay in bf16) — used only to replay the CPU-recorded goldens."""
        self.sd = sd
        self.dt = compute_dtype
        self.flavor = flavor
        self.cpu_bf16 = flavor == "cpu" and compute_dtype == torch.bfloat16

    def p(self, key: str) -> Tensor:
        return self.sd[key]

    def has(self, key: str) -> bool:
        return key in self.sd

    def lin(self, x: Tensor, prefix: str, bias: bool = True) -> Tensor:
        """nn.Linear under autocast: inputs/weights cast to the compute dtype, output in the compute dtype."""
        w = self.p(prefix + ".weight").to(self.dt)
        b = self.p(prefix + ".bias").to(self.dt) if bias and self.has(prefix + ".bias") else None
        return F.linear(x.to(self.dt), w, b)


# ------------------------------------------------------------------------------------------------ image tokenizer
def local_attention(c: Ctx, pre: str, features: Tensor, conv_stride: int = 3, num_heads: int = 8) -> Tensor:
    """LocalAttention.forward — models/mla/image/vision_tokenizer.py:27-47.  features [B,C,H,W]."""
    B, C, H, W = features.shape
    scale = C ** -0.5
    red = F.avg_pool2d(features, kernel_size=conv_stride, stride=conv_stride)
    h, w = red.shape[-2:]
    N = conv_stride ** 2
    red = red.flatten(2).transpose(-2, -1)                                                    # [B, hw, C]

    def ln(x, p_):
        return F.layer_norm(x.float(), (C,), c.p(p_ + ".weight").float(), c.p(p_ + ".bias").float(), 1e-5)
    q = c.lin(ln(red, pre + ".q.0"), pre + ".q.1", bias=False)
    q = q.reshape(B, h * w, num_heads, -1).permute(0, 2, 1, 3).unsqueeze(-2)
    f = features.unfold(2, conv_stride, conv_stride).unfold(3, conv_stride, conv_stride)
    f = f.contiguous().view(B, C, h * w, conv_stride, conv_stride)
    kv = c.lin(ln(f.flatten(3).permute(0, 2, 3, 1), pre + ".kv.0"), pre + ".kv.1", bias=False)
    kv = kv.reshape(B, h * w, N, 2, num_heads, -1).permute(3, 0, 4, 1, 2, 5)
    attn = (q * scale * kv[0]).sum(-1)
    attn = attn.softmax(dim=-1) if c.cpu_bf16 else attn.float().softmax(dim=-1)               # cuda autocast: fp32
    agg = (attn.unsqueeze(-1) * kv[1]).sum(-2)
    agg = agg.transpose(1, 2).reshape(B, h * w, -1)
    return red + c.lin(agg, pre + ".proj")


def image_tokens(c: Ctx, pre_tower: str, pre_proj: str, pixel_values: Tensor) -> Tensor:
    """VisionTokenizer.forward + MLP_GELU — vision_tokenizer.py:119-152,:79-89, for all-ones pixel masks (the crop
    at :129-137 is then the identity).  pixel_values [B,4,H,W] -> [B, hw, token]."""
    px, mask = pixel_values[:, :-1], pixel_values[:, -1:]
    assert bool((mask == 1).all()), "oracle restates the all-ones-mask path only"
    w = c.p(pre_tower + ".patch_embedding.weight").to(c.dt)
    emb = F.conv2d(px.to(c.dt), w, stride=14)                                                 # [B,C,Hp,Wp]
    out = []
    for i in range(emb.shape[0]):
        pe = local_attention(c, pre_tower + ".local_attention", emb[i:i + 1])                 # [1, hw, C]
        # GlobalAttention (:141-142) is evaluated and dropped by the reference: no effect on the output
        t = c.lin(pe[0], pre_proj + ".mlp.0")
        t = F.gelu(t)
        out.append(c.lin(t, pre_proj + ".mlp.2"))
    return torch.stack(out, 0)


# ------------------------------------------------------------------------------------------------ diffusion bits
def timm_mlp(c: Ctx, pre: str, x: Tensor) -> Tensor:
    """timm Mlp (fc1 -> GELU(tanh) -> fc2), models/diffusion/models.py:115-123."""
    return c.lin(F.gelu(c.lin(x, pre + ".fc1"), approximate="tanh"), pre + ".fc2")


def timestep_embed(c: Ctx, pre: str, t: Tensor) -> Tensor:
    """TimestepEmbedder — models/diffusion/models.py:41-65 (t arrives as bf16, prismatic.py:877-878)."""
    half = 128
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    return c.lin(F.silu(c.lin(emb, pre + ".mlp.0")), pre + ".mlp.2")


def timm_rmsnorm(x: Tensor, w: Tensor, eps: float = 1e-6, variance_mode: bool = False) -> Tensor:
    """timm RmsNorm used by FinalLayer (models/diffusion/models.py:179).  NOT in the reference tree (timm 0.9.10):
    mean-square form by default, torch.var form selectable — parity unpinned."""
    xf = x.float()
    v = xf.var(dim=-1, keepdim=True) if variance_mode else xf.pow(2).mean(-1, keepdim=True)
    return (xf * torch.rsqrt(v + eps)).to(x.dtype) * w.to(x.dtype)


def sqrt_alpha_tables(steps: int = 100) -> Tuple[np.ndarray, np.ndarray]:
    """squaredcos_cap_v2 — models/diffusion/gaussian_diffusion.py:112-140,:152-186."""
    ab = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    betas = np.array([min(1 - ab((i + 1) / steps) / ab(i / steps), 0.999) for i in range(steps)], dtype=np.float64)
    ac = np.cumprod(1.0 - betas, axis=0)
    return np.sqrt(ac), np.sqrt(1.0 - ac)


def q_sample(a: Tensor, t: Tensor, noise: Tensor) -> Tensor:
    """GaussianDiffusion.q_sample — gaussian_diffusion.py:214-229 with _extract_into_tensor (:866-881)."""
    sa, sb = sqrt_alpha_tables()
    ca = torch.from_numpy(sa)[t.cpu()].float().to(a.device).view(-1, *([1] * (a.dim() - 1)))
    cb = torch.from_numpy(sb)[t.cpu()].float().to(a.device).view(-1, *([1] * (a.dim() - 1)))
    return ca * a + cb * noise


# ------------------------------------------------------------------------------------------------ contrastive
CAMERAS = {
    "rlbench_front": dict(
        K=[[-307.7174807, 0.0, 112.0], [0.0, -307.7174807, 112.0], [0.0, 0.0, 1.0]],
        R=[[1.19209290e-07, -4.22617942e-01, -9.06307936e-01], [-1.00000000e+00, -5.96046448e-07, 1.49011612e-07],
           [-5.66244125e-07, 9.06307936e-01, -4.22617912e-01]],
        t=[1.34999919e+00, 3.71546562e-08, 1.57999933e+00], orig=(224, 224)),
}


def project_3d_to_2d(xyz: Tensor, camera: str) -> Tuple[Tensor, Tensor]:
    """project_3d_to_2d_672_rlbench — models/mla/fuser/contrastive.py:5-45 with camera.py:13-26."""
    cam = CAMERAS[camera]
    K = torch.tensor(cam["K"], dtype=torch.float32, device=xyz.device)
    R = torch.tensor(cam["R"], dtype=torch.float32, device=xyz.device)
    t = torch.tensor(cam["t"], dtype=torch.float32, device=xyz.device)
    sx, sy = 672 / cam["orig"][1], 672 / cam["orig"][0]
    Ks = K.clone()
    Ks[0, 0] *= sx; Ks[1, 1] *= sy; Ks[0, 2] *= sx; Ks[1, 2] *= sy
    Rw = R.T
    tw = -Rw @ t
    cam_xyz = xyz @ Rw.T + tw
    uvw = cam_xyz @ Ks.T
    z = uvw[..., 2:]
    xy = uvw[..., :2] / (z + 1e-6)
    row = (xy[..., 1] / 42).floor().long()
    col = (xy[..., 0] / 42).floor().long()
    valid = (z.squeeze(-1) > 0) & (xy[..., 0] >= 0) & (xy[..., 0] < 672) & (xy[..., 1] >= 0) & (xy[..., 1] < 672)
    return torch.stack([row.clamp(0, 15), col.clamp(0, 15)], -1), valid


def proj_head(c: Ctx, pre: str, x: Tensor) -> Tensor:
    return c.lin(F.relu(c.lin(x, pre + ".0")), pre + ".2")


def normalize(x: Tensor, cpu_bf16: bool = False) -> Tensor:
    """F.normalize under cuda autocast: `norm` is on the fp32 list, so the quotient is fp32 (cpu autocast: bf16)."""
    if cpu_bf16:
        return F.normalize(x, p=2, dim=-1)
    n = x.float().norm(2, -1, keepdim=True).clamp_min(1e-12)
    return x.float() / n


def coordinate_contrastive(c: Ctx, pre: str, img_f: Tensor, pc_f: Tensor, patch_idx: Tensor, valid: Tensor,
                           temperature: float = 0.07) -> Tensor:
    """CoordinateAwareContrastiveLoss.forward — contrastive.py:185-215."""
    img = normalize(proj_head(c, pre + ".image_projection_head", img_f), c.cpu_bf16)
    pc = normalize(proj_head(c, pre + ".pointcloud_projection_head", pc_f), c.cpu_bf16)
    B, n_p, _ = img_f.shape
    lin = patch_idx[:, :, 0] * int(n_p ** 0.5) + patch_idx[:, :, 1]
    tgt = torch.gather(img, 1, lin.unsqueeze(-1).expand(-1, -1, img.shape[-1]))
    a, b = pc[valid], tgt[valid]
    if a.shape[0] == 0:
        return torch.tensor(0.0, device=img_f.device)
    logits = (torch.matmul(a.to(c.dt), b.to(c.dt).t()) / temperature)
    lab = torch.arange(a.shape[0], device=logits.device)
    return (F.cross_entropy(logits.float(), lab) + F.cross_entropy(logits.float().t(), lab)) / 2


def tactile_contrastive(c: Ctx, pre: str, tac_f: Tensor, pc_f: Tensor, img_f: Tensor, pos_pc: Tensor,
                        pos_img: Tensor, temperature: float = 0.07) -> Tensor:
    """TactileContrastiveLoss.forward — contrastive.py:241-258."""
    tac = normalize(proj_head(c, pre + ".tactile_projection_head", tac_f), c.cpu_bf16)
    pc = normalize(proj_head(c, pre + ".pointcloud_projection_head", pc_f), c.cpu_bf16)
    img = normalize(proj_head(c, pre + ".image_projection_head", img_f), c.cpu_bf16)
    l_pc = torch.bmm(tac.to(c.dt), pc.to(c.dt).transpose(1, 2)) / temperature
    l_img = torch.bmm(tac.to(c.dt), img.to(c.dt).transpose(1, 2)) / temperature
    return (F.cross_entropy(l_pc.float().view(-1, pc.shape[1]), pos_pc.view(-1)) +
            F.cross_entropy(l_img.float().view(-1, img.shape[1]), pos_img.view(-1))) / 2


# ------------------------------------------------------------------------------------------------ point tokenizer
def fps(xyz: Tensor, npoint: int, start: Tensor) -> Tensor:
    """furthest_point_sample — models/mla/pointcloud/backbone/Point_PN.py:6-21 with the random start index given."""
    B, N, _ = xyz.shape
    idx = torch.zeros(B, npoint, dtype=torch.long, device=xyz.device)
    far = start.clone()
    dist = torch.ones(B, N, device=xyz.device) * 1e10
    ar = torch.arange(B, device=xyz.device)
    for i in range(npoint):
        idx[:, i] = far
        cen = xyz[ar, far, :].view(B, 1, 3)
        d = torch.sum((xyz - cen) ** 2, -1)
        m = d < dist
        dist[m] = d[m]
        far = torch.max(dist, -1)[1]
    return idx


def index_points(points: Tensor, idx: Tensor) -> Tensor:
    """Point_PN.py:41-58."""
    B = points.shape[0]
    view = [B] + [1] * (idx.dim() - 1)
    bi = torch.arange(B, device=points.device).view(view).expand_as(idx)
    return points[bi, idx, :]


def knn(c: Ctx, k: int, xyz: Tensor, new_xyz: Tensor) -> Tensor:
    """knn_point / square_distance — Point_PN.py:23-39,:62-73.  Under autocast the matmul runs in the compute dtype
    and the in-place `+=` keep that dtype, so the ranking is done on (possibly bf16) distances; ties between equal
    distances are implementation-defined in torch.topk (unpinned) — resolved here by lowest index."""
    d = -2 * torch.matmul(new_xyz.to(c.dt), xyz.to(c.dt).permute(0, 2, 1))
    d += torch.sum(new_xyz ** 2, -1).unsqueeze(-1)
    d += torch.sum(xyz ** 2, -1).unsqueeze(1)
    order = torch.sort(d.float(), dim=-1, stable=True)[1]
    return order[..., :k]


def batchnorm_train(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5, native: bool = False) -> Tensor:
    """nn.BatchNorm{1,2}d in train mode (the tokenizer is frozen but `self.vlm.train()` keeps batch statistics,
    training/strategies/base_strategy_mla.py:291): statistics over every dim but channel (dim 1), fp32."""
    if native:   # replaying CPU goldens: ATen's own bf16 CPU kernel (its internal rounding differs by <= 1 ulp)
        return F.batch_norm(x, None, None, w.to(x.dtype), b.to(x.dtype), True, 0.1, eps)
    dims = [0] + list(range(2, x.dim()))
    xf = x.float()
    mean = xf.mean(dims, keepdim=True)
    var = xf.var(dims, unbiased=False, keepdim=True)
    shape = [1, -1] + [1] * (x.dim() - 2)
    return ((xf - mean) * torch.rsqrt(var + eps) * w.float().view(shape) + b.float().view(shape)).to(x.dtype)


def batchnorm_eval(x: Tensor, w: Tensor, b: Tensor, mean: Tensor, var: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm{1,2}d in eval mode (inference, MLA.predict_action_diff calls self.vlm.eval()): running statistics."""
    shape = [1, -1] + [1] * (x.dim() - 2)
    xf = x.float()
    return ((xf - mean.float().view(shape)) * torch.rsqrt(var.float().view(shape) + eps) * w.float().view(shape)
            + b.float().view(shape)).to(x.dtype)


def point_tokens(c: Ctx, pre: str, p: Tensor, starts: List[Tensor], k_neighbors: int = 81,
                 alpha: float = 1000.0, beta: float = 100.0, knn_override: Optional[List[Tensor]] = None,
                 record: Optional[dict] = None, bn_eval: bool = False) -> Tuple[Tensor, Tensor]:
    """PointTokenizer.forward -> Point_PN_scan -> EncP.forward — pointvit.py:59-82, Point_PN.py:284-298 with
    FPS_kNN :84-94, LGA :112-158 ('scan' normalisation), PosE_Geo :228-249, Linear2Layer :188-219, max pooling.
    p [B,N,3] f32; starts = FPS start indices per stage.  Returns (tokens [B,G,768], centres [B,G,3])."""
    e = pre + ".patch_embed.EncP"

    def bn(x_, q_):
        if bn_eval:
            return batchnorm_eval(x_, c.p(q_ + ".weight"), c.p(q_ + ".bias"), c.p(q_ + ".running_mean"), c.p(q_ + ".running_var"))
        return batchnorm_train(x_, c.p(q_ + ".weight"), c.p(q_ + ".bias"), native=c.cpu_bf16)
    xyz = p.float()
    x = p.float().transpose(1, 2).contiguous()                                                # [B,3,N]
    w = c.p(e + ".raw_point_embed.net.0.weight").to(c.dt)
    x = F.conv1d(x.to(c.dt), w)
    x = F.relu(bn(x, e + ".raw_point_embed.net.1"))
    out_dim, blocks = x.shape[1], [2, 1]
    for i in range(2):
        out_dim *= 2
        G = xyz.shape[1] // 2
        xt = x.permute(0, 2, 1).contiguous()
        fidx = fps(xyz.contiguous(), G, starts[i])
        lc_xyz, lc_x = index_points(xyz, fidx), index_points(xt, fidx)
        kidx = knn_override[i] if knn_override is not None else knn(c, k_neighbors, xyz, lc_xyz)
        if record is not None:
            record.setdefault("knn_idx", []).append(kidx)
            record.setdefault("fps_idx", []).append(fidx)
        knn_xyz, knn_x = index_points(xyz, kidx), index_points(xt, kidx)
        xyz = lc_xyz
        # LGA 'scan' normalisation (:125-134)
        kx = knn_xyz.permute(0, 3, 1, 2) - lc_xyz.permute(0, 2, 1).unsqueeze(-1)
        mx = torch.abs(kx).max(dim=-1, keepdim=True)[0].clamp(min=1e-6)
        kx = kx / mx                                                                          # [B,3,G,K]
        B, Gn, K, C = knn_x.shape
        feat = torch.cat([knn_x, lc_x.reshape(B, Gn, 1, -1).repeat(1, 1, K, 1)], dim=-1).permute(0, 3, 1, 2)
        # PosE_Geo (:228-249)
        fd = out_dim // 6
        rng = torch.arange(fd, dtype=torch.float32, device=p.device)
        div = (beta * kx.unsqueeze(-1)) / torch.pow(alpha, rng / fd)
        pe = torch.cat([torch.sin(div), torch.cos(div)], -1).permute(0, 1, 4, 2, 3).contiguous().view(B, out_dim, Gn, K)
        xw = feat + pe
        for j in range(blocks[i]):
            q = f"{e}.LGA_list.{i}.linear2.{j}"
            y = F.conv2d(xw.to(c.dt), c.p(q + ".net1.0.weight").to(c.dt), c.p(q + ".net1.0.bias").to(c.dt))
            y = F.relu(bn(y, q + ".net1.1"))
            y = F.conv2d(y.to(c.dt), c.p(q + ".net2.0.weight").to(c.dt), c.p(q + ".net2.0.bias").to(c.dt))
            y = bn(y, q + ".net2.1")
            xw = F.relu(y + xw)
        x = xw.max(-1)[0]
    tok = c.lin(x.transpose(1, 2), pre + ".proj")
    return tok, xyz


# ------------------------------------------------------------------------------------------------ full forward
def forward(sd: Dict[str, Tensor], batch: Dict, cfg: Dict, draws: Dict, compute_dtype=torch.bfloat16,
            flavor: str = "cuda") -> Dict:
    """MLA.forward (model_mla.py:118-234) with use_diff=True.

    cfg: n_heads, rms_eps, future_action_window_size, repeated_diffusion_steps, use_pointcloud, use_tactile,
         use_contrastive, camera_name, rmsnorm_variance_mode; post-training: use_generation, gen_image, use_roi,
         gen_pointcloud, gen_tactile (+ the heads' head counts).
    draws: noise [B_eff,T+1,A], timestep [B_eff] (and fps_starts: list of [B_eff] per stage when use_pointcloud).
    Returns losses, noise_pred and the boundary tensors the parity tests compare."""
    c = Ctx(sd, compute_dtype, flavor)
    R = cfg["repeated_diffusion_steps"]
    T = cfg["future_action_window_size"]
    rep = lambda v: v.repeat(R, *([1] * (v.dim() - 1)))                                        # model_mla.py:147-176
    evalm = bool(cfg.get("eval"))          # inference: PrismaticVLM.forward in eval mode, x_t and t given by the sampler
    proprio = rep(batch["proprio"])
    ids = rep(batch["input_ids"])
    am = rep(batch["attention_mask"]) if batch.get("attention_mask") is not None else torch.ones_like(ids, dtype=torch.bool)
    images = {k: rep(v) for k, v in batch["images"].items()}
    t = draws["timestep"]
    if "x" in draws:
        x, noise = draws["x"], draws.get("noise")
    else:
        a_future = rep(batch["actions"])[:, -(T + 1):, :]
        noise = draws["noise"]
        x = q_sample(a_future, t, noise)                                                      # :180
    tag = 29871 if evalm else 2                                                               # prismatic.py:882-887

    V = "vlm."
    # ---- get_fused_tokens (prismatic.py:598-769)
    front = image_tokens(c, V + "vision_tower_2d", V + "projector_2d", images["front_image"])
    B, n_img, h = front.shape
    centers = None
    if cfg.get("use_pointcloud"):
        pc_emb, centers = point_tokens(c, V + "vision_tower_3d", rep(batch["point_cloud"]), draws["fps_starts"],
                                       k_neighbors=cfg.get("k_neighbors", 81), knn_override=draws.get("knn_idx"),
                                       bn_eval=evalm)
        pc_tok = c.lin(F.gelu(c.lin(pc_emb, V + "projector_3d.projector.0")), V + "projector_3d.projector.2")
        patch_idx, valid = project_3d_to_2d(centers, cfg["camera_name"])
    else:
        pc_tok = torch.zeros(B, n_img, h, dtype=front.dtype)
        patch_idx = torch.zeros(B, n_img, 2, dtype=torch.long)
        valid = torch.zeros(B, n_img, dtype=torch.bool)
    parts = [pc_tok, front]
    for k_ in images:
        if k_ != "front_image":
            parts.append(image_tokens(c, V + "vision_tower_2d", V + "projector_2d", images[k_]))
    pos_pc = lin_img = None
    if cfg.get("use_tactile"):
        tac, grip = rep(batch["tactile"]), rep(batch["gripper_xyz"])
        n_arms = grip.shape[-1] // 3
        tac_emb = torch.cat([timm_mlp(c, V + "tactile_embedder.mlp", ts).unsqueeze(1)
                             for ts in torch.chunk(tac.view(B, -1), n_arms, dim=-1)], 1)
        parts.append(tac_emb)
        d = torch.cdist(grip.view(B, n_arms, 3).float(), centers)
        pos_pc = torch.topk(d, k=1, dim=2, largest=False)[1]
        sel = torch.gather(patch_idx.unsqueeze(1).expand(-1, n_arms, -1, -1), 2, pos_pc.unsqueeze(-1).expand(-1, -1, -1, 2))
        lin_img = sel[..., 0] * int(n_img ** 0.5) + sel[..., 1]
    else:
        parts.append(torch.zeros(B, 1, h, dtype=front.dtype))
    fused = torch.cat(parts, 1)
    F_ = fused.shape[1]

    # ---- splice (prismatic.py:946-1042)
    emb = F.embedding(ids, c.p(V + "llm_backbone.llm.model.embed_tokens.weight")).to(front.dtype)
    z = torch.cat([emb[:, :1], fused, emb[:, 1:]], 1)
    pr = timm_mlp(c, V + "proprio_embedder.mlp", proprio.to(torch.bfloat16))
    xe = timm_mlp(c, V + "x_embedder.mlp", x.to(torch.bfloat16))
    te = timestep_embed(c, V + "t_embedder", t.to(torch.bfloat16)).unsqueeze(1)
    seqs, masks, ltis = [], [], []
    for i in range(B):
        lti = int(torch.where(ids[i] == tag)[0][-1]) + F_
        ltis.append(lti)
        seqs.append(torch.cat([z[i, :lti], pr[i], te[i], xe[i], z[i, lti:]], 0).unsqueeze(0))
        masks.append(torch.cat([am[i, :1], torch.ones(F_, dtype=torch.bool), am[i, 1:lti - F_],
                                torch.ones(2 + xe.shape[1], dtype=torch.bool), am[i, lti - F_:]], 0).unsqueeze(0))
    embeds, mask = torch.cat(seqs, 0), torch.cat(masks, 0)

    # ---- decoder (modeling_llama.py)
    P = V + "llm_backbone.llm.model."
    n_layers = 1 + max(int(k_.split(".")[5]) for k_ in sd if k_.startswith(P + "layers."))
    cast = lambda w: w.to(compute_dtype)
    layers = []
    for li in range(n_layers):
        q = f"{P}layers.{li}."
        layers.append(dict(q_proj=cast(sd[q + "self_attn.q_proj.weight"]), k_proj=cast(sd[q + "self_attn.k_proj.weight"]),
                           v_proj=cast(sd[q + "self_attn.v_proj.weight"]), o_proj=cast(sd[q + "self_attn.o_proj.weight"]),
                           gate_proj=cast(sd[q + "mlp.gate_proj.weight"]), up_proj=cast(sd[q + "mlp.up_proj.weight"]),
                           down_proj=cast(sd[q + "mlp.down_proj.weight"]), ln1=cast(sd[q + "input_layernorm.weight"]),
                           ln2=cast(sd[q + "post_attention_layernorm.weight"])))
    hs = L.decoder(embeds.to(compute_dtype), layers, cast(sd[P + "norm.weight"]), cfg["n_heads"], cfg["rms_eps"], mask)

    out = dict(fused=fused, embeds=embeds, mask=mask, hidden_states=hs, last_true_indices=ltis, x=x,
               patch_indices=patch_idx, valid_mask=valid, centers=centers)
    # ---- contrastive (modeling_llama.py:1271-1303)
    total_extra = 0.0
    if cfg.get("use_contrastive"):
        h8 = hs[8]
        lm = V + "llm_backbone.llm."
        out["img_pc_contrastive_loss"] = coordinate_contrastive(
            c, lm + "coordinate_aware_contrastive_loss_module", h8[:, 1 + n_img:1 + 2 * n_img], h8[:, 1:1 + n_img],
            patch_idx, valid)
        total_extra = total_extra + out["img_pc_contrastive_loss"]
        if cfg.get("use_tactile"):
            out["tactile_contrastive_loss"] = tactile_contrastive(
                c, lm + "tactile_contrastive_loss_module", h8[:, 1 + 2 * n_img:2 + 2 * n_img], h8[:, 1:1 + n_img],
                h8[:, 1 + n_img:1 + 2 * n_img], pos_pc, lin_img)
            total_extra = total_extra + out["tactile_contrastive_loss"]
    # ---- post-training generation heads (prismatic.py:1075-1113, :771-838; model_mla.py:218-226)
    if cfg.get("use_generation") and (cfg.get("gen_image") or cfg.get("gen_pointcloud") or cfg.get("gen_tactile")):
        from . import generation as G
        GM = V + "generation_manager."
        gen_total = 0.0
        if cfg.get("gen_image"):
            cur_p = G.images_to_patches(images["front_image"][:, :3], 42)
            nxt_p = G.images_to_patches(rep(batch["next_images"]), 42)
            roi = (G.roi_mask(patch_idx, cfg.get("roi_dilation_kernel_size", 3)) if cfg.get("use_roi")
                   else torch.ones(B, n_img, dtype=torch.bool))
            io = G.image_head(c, GM + "image_gen_module", hs[-1], fused[:, n_img:2 * n_img], cur_p, roi,
                              heads=cfg.get("image_decoder_heads", 8))
            il = G.image_losses(io, nxt_p)
            out.update(image_generation=io["image_generation"], generation_roi_mask=roi, delta_all=io["delta_all"],
                       alpha_all=io["alpha_all"], offset_all=io["offset_all"], image_gen_loss=il["image_gen_loss"],
                       image_loss_terms=il)
            gen_total = gen_total + il["image_gen_loss"]
        if cfg.get("gen_pointcloud"):
            pc_pred = G.pointcloud_head(c, GM + "pointcloud_gen_module", hs[-1], heads=cfg.get("pointcloud_decoder_heads", 8))
            out["pointcloud_coord_generation"] = pc_pred
            out["point_cloud_gen_loss"] = G.chamfer_l2(pc_pred, rep(batch["next_point_cloud"]))
            gen_total = gen_total + out["point_cloud_gen_loss"]
        if cfg.get("gen_tactile"):
            tp = G.tactile_head(c, GM + "tactile_gen_module", hs[-1], heads=cfg.get("tactile_decoder_heads", 4))
            out["tactile_generation"] = tp
            out["tactile_gen_loss"] = F.mse_loss(tp.float(), rep(batch["next_tactile"]).float())
            gen_total = gen_total + out["tactile_gen_loss"]
        total_extra = total_extra + gen_total
    # ---- final layer + loss (prismatic.py:1115-1126, model_mla.py:205-232)
    last = hs[-1]
    y = timm_rmsnorm(last, sd[V + "final_layer.norm_final.weight"], 1e-6, cfg.get("rmsnorm_variance_mode", False))
    y = timm_mlp(c, V + "final_layer.mlp", y)
    noise_pred = torch.cat([y[i, l + 2:l + T + 3].unsqueeze(0) for i, l in enumerate(ltis)], 0)
    out["noise_pred"] = noise_pred
    if noise is not None:
        diff = ((noise_pred - noise) ** 2).mean()
        out.update(diff_loss=diff, total_loss=diff + total_extra)
    return out
