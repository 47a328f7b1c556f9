"""Cross-modal alignment: 3D->2D patch correspondence and the two InfoNCE losses.

Reference: models/mla/fuser/contrastive.py (project_3d_to_2d_672_* :5-131, CoordinateAwareContrastiveLoss :170-215,
TactileContrastiveLoss :219-258) and models/mla/fuser/camera.py (CAMERA_CONFIGS, get_camera_params,
get_projection_func).  Parameter names match; the projection heads run on the tcgen05 GEMM, the normalisation and
softmax statistics on bandwidth-bound warp kernels (csrc/contrastive.cu).  No host synchronisation: the reference's
boolean-mask compaction (:203-206) becomes an on-device validity mask.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import check


@dataclass
class CameraParams:
    K: torch.Tensor
    R: torch.Tensor
    t: torch.Tensor


# models/mla/fuser/camera.py:12-51 (calibration constants of the three supported cameras)
CAMERA_CONFIGS = {
    "rlbench_front": CameraParams(
        K=torch.tensor([[-307.7174807, 0.0, 112.0], [0.0, -307.7174807, 112.0], [0.0, 0.0, 1.0]], dtype=torch.float32),
        R=torch.tensor([[1.19209290e-07, -4.22617942e-01, -9.06307936e-01],
                        [-1.00000000e+00, -5.96046448e-07, 1.49011612e-07],
                        [-5.66244125e-07, 9.06307936e-01, -4.22617912e-01]], dtype=torch.float32),
        t=torch.tensor([1.34999919e+00, 3.71546562e-08, 1.57999933e+00], dtype=torch.float32)),
    "franka_right": CameraParams(
        K=torch.tensor([[387.414794921875, 0.0, 319.47052001953125], [0.0, 386.8714904785156, 241.13287353515625],
                        [0.0, 0.0, 1.0]], dtype=torch.float32),
        R=torch.tensor([[0.91300858, 0.26157042, -0.31304353], [0.39730357, -0.7442472, 0.53688545],
                        [-0.09254842, -0.61455433, -0.78342694]], dtype=torch.float32),
        t=torch.tensor([0.8591219242556176, -0.5851783639922448, 0.7535876808722389], dtype=torch.float32)),
    "franka_front": CameraParams(
        K=torch.tensor([[388.2638244628906, 0.0, 328.3757019042969], [0.0, 387.84130859375, 240.24295043945312],
                        [0.0, 0.0, 1.0]], dtype=torch.float32),
        R=torch.tensor([[-0.01750229, 0.95018522, -0.31119403], [0.99984609, 0.01625676, -0.00659609],
                        [-0.0012085, -0.31126158, -0.95032351]], dtype=torch.float32),
        t=torch.tensor([0.8545415959817313, 0.5748472977587156, 1.0411478820663598], dtype=torch.float32)),
}
# original image sizes (H, W) each project_3d_to_2d_672_* assumes (contrastive.py:8,:50,:92)
_ORIG_SIZE = {"rlbench_front": (224, 224), "franka_right": (480, 640), "franka_front": (720, 1280)}
_CAM_DEV = {}


def get_camera_params(config_name: str = "default", device=None) -> CameraParams:
    if config_name not in CAMERA_CONFIGS:
        raise ValueError(f"Unknown camera config: {config_name}. Available configs: {list(CAMERA_CONFIGS.keys())}")
    p = CAMERA_CONFIGS[config_name]
    if device is not None:
        p.K, p.R, p.t = p.K.to(device), p.R.to(device), p.t.to(device)
    return p


def project_points(xyz: torch.Tensor, camera_name: str, image_size_resize=(672, 672), patch_stride: int = 14,
                   conv_stride: int = 3) -> Tuple[torch.Tensor, torch.Tensor]:
    """project_3d_to_2d_672_{rlbench,franka_right,franka_front}: xyz f32 [B,N,3] -> (patch_idx int64 [B,N,2] =
    (row, col), valid bool [B,N])."""
    if camera_name not in CAMERA_CONFIGS:
        raise ValueError(f"Unknown projection func for camera {camera_name}. Available: {list(CAMERA_CONFIGS.keys())}")
    key = (camera_name, str(xyz.device))
    if key not in _CAM_DEV:
        p = CAMERA_CONFIGS[camera_name]
        _CAM_DEV[key] = torch.cat([p.R.reshape(-1).cpu(), p.t.reshape(-1).cpu(), p.K.reshape(-1).cpu()]).float().to(xyz.device)
    cam = _CAM_DEV[key]
    oh, ow = _ORIG_SIZE[camera_name]
    B, N, _ = xyz.shape
    pts = xyz.detach().float().contiguous()
    idx = torch.empty((B, N, 2), dtype=torch.int64, device=xyz.device)
    valid = torch.empty((B, N), dtype=torch.uint8, device=xyz.device)
    total = patch_stride * conv_stride
    check(_lib.lib().mla_project_points(
        ops._p(pts), ops._p(cam), C.c_int64(B * N), C.c_float(image_size_resize[1] / ow),
        C.c_float(image_size_resize[0] / oh), C.c_float(total), C.c_int32(image_size_resize[0] // total),
        C.c_int32(image_size_resize[1] // total), C.c_float(image_size_resize[1]), C.c_float(image_size_resize[0]),
        ops._p(idx), ops._p(valid), ops._stream()))
    return idx, valid.bool()


class _L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        rows, d = x.shape
        y = torch.empty_like(x)
        norms = torch.empty(rows, dtype=torch.float32, device=x.device)
        check(_lib.lib().mla_l2norm_fwd(ops._p(x), ops._p(y), ops._p(norms), C.c_int64(rows), C.c_int32(d),
                                        C.c_int64(x.stride(0)), C.c_float(1e-12), ops._stream()))
        ctx.save_for_backward(x, norms)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, norms = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        check(_lib.lib().mla_l2norm_bwd(ops._p(x), ops._p(norms), ops._p(dy), ops._p(dx), C.c_int64(x.shape[0]),
                                        C.c_int32(x.shape[1]), C.c_int64(x.stride(0)), ops._stream()))
        return dx


class _InfoNCEFn(torch.autograd.Function):
    """Symmetric InfoNCE of rows a_i vs b_i over the valid subset; a, b bf16 [N,D] L2-normalised."""

    @staticmethod
    def forward(ctx, a, b, valid, temperature):
        N = a.shape[0]
        sim = ops.gemm(a, b)                                            # bf16 [N,N] = bf16(a_i . b_j)
        ws = torch.empty(_lib.lib().mla_infonce_workspace(C.c_int32(N)) // 4, dtype=torch.float32, device=a.device)
        out = torch.empty(2, dtype=torch.float32, device=a.device)
        check(_lib.lib().mla_infonce_fwd(ops._p(sim), ops._p(valid), ops._p(ws), ops._p(out), C.c_int32(N),
                                         C.c_float(temperature), ops._stream()))
        ctx.save_for_backward(a, b, valid, sim, ws, out)
        ctx.temperature = temperature
        return out[0]

    @staticmethod
    def backward(ctx, g):
        a, b, valid, sim, ws, out = ctx.saved_tensors
        gs = g.reshape(1).float().contiguous()
        check(_lib.lib().mla_infonce_bwd(ops._p(sim), ops._p(valid), ops._p(ws), ops._p(out), ops._p(gs),
                                         C.c_int32(a.shape[0]), C.c_float(ctx.temperature), ops._stream()))
        da = ops.gemm(sim, b, b_mn=True)                                # dsim . b
        db = ops.gemm(sim, a, a_mn=True, b_mn=True)                     # dsim^T . a
        return da, db, None, None


class _TacNCEFn(torch.autograd.Function):
    """mean CE of bf16(q.k/T) against a positive key index; q bf16 [B,A,D], keys bf16 [B,K,D], pos int64 [B,A]."""

    @staticmethod
    def forward(ctx, q, keys, pos, temperature):
        B, A, D = q.shape
        K = keys.shape[1]
        q, keys, pos = q.contiguous(), keys.contiguous(), pos.reshape(-1).contiguous()
        probs = torch.empty((B * A, K), dtype=torch.float32, device=q.device)
        rows = torch.empty(B * A, dtype=torch.float32, device=q.device)
        check(_lib.lib().mla_tac_nce_fwd(ops._p(q), ops._p(keys), ops._p(pos), ops._p(probs), ops._p(rows), C.c_int32(B),
                                         C.c_int32(A), C.c_int32(K), C.c_int32(D), C.c_float(temperature), ops._stream()))
        ctx.save_for_backward(q, keys, pos, probs)
        ctx.temperature = temperature
        return rows.mean()

    @staticmethod
    def backward(ctx, g):
        q, keys, pos, probs = ctx.saved_tensors
        B, A, D = q.shape
        K = keys.shape[1]
        gs = g.reshape(1).float().contiguous()
        dq = torch.empty((B, A, D), dtype=torch.float32, device=q.device)
        dk = torch.zeros((B, K, D), dtype=torch.float32, device=q.device)
        check(_lib.lib().mla_tac_nce_bwd(ops._p(q), ops._p(keys), ops._p(pos), ops._p(probs), ops._p(gs), ops._p(dq),
                                         ops._p(dk), C.c_int32(B), C.c_int32(A), C.c_int32(K), C.c_int32(D),
                                         C.c_float(ctx.temperature), ops._stream()))
        return dq.to(torch.bfloat16), dk.to(torch.bfloat16), None, None


def _head(feature_dim: int, projection_dim: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(feature_dim, feature_dim), nn.ReLU(inplace=True), nn.Linear(feature_dim, projection_dim))


def _run_head(head: nn.Sequential, x2d: torch.Tensor) -> torch.Tensor:
    x = ops.linear(x2d, head[0].weight, head[0].bias, ops.ACT_RELU)
    return ops.linear(x, head[2].weight, head[2].bias)


class CoordinateAwareContrastiveLoss(nn.Module):
    def __init__(self, feature_dim: int, projection_dim: int = 256, temperature: float = 0.07):
        super().__init__()
        self.temperature = temperature
        self.image_projection_head = _head(feature_dim, projection_dim)
        self.pointcloud_projection_head = _head(feature_dim, projection_dim)

    def forward(self, image_features, pointcloud_features, patch_indices, valid_mask):
        B, n_patches, D = image_features.shape
        n_points = pointcloud_features.shape[1]
        img = _L2NormFn.apply(_run_head(self.image_projection_head, image_features.reshape(B * n_patches, D)))
        pc = _L2NormFn.apply(_run_head(self.pointcloud_projection_head, pointcloud_features.reshape(B * n_points, D)))
        patch_w = int(n_patches ** 0.5)
        lin = patch_indices[:, :, 0] * patch_w + patch_indices[:, :, 1]                       # [B, n_points]
        rows = (lin + torch.arange(B, device=lin.device).view(B, 1) * n_patches).reshape(-1).to(torch.int32)
        # rows may repeat (several points project into one patch): gather with autograd through an index-add
        target = _GatherDupFn.apply(img, rows)
        valid = valid_mask.reshape(-1).to(torch.uint8).contiguous()
        return _InfoNCEFn.apply(pc, target, valid, self.temperature)


class _GatherDupFn(torch.autograd.Function):
    """dst = src[idx] where idx may contain duplicates; backward accumulates in fp32."""

    @staticmethod
    def forward(ctx, src, idx):
        ctx.save_for_backward(idx)
        ctx.shape = src.shape
        return ops.gather_rows(src, idx.contiguous())

    @staticmethod
    def backward(ctx, d):
        (idx,) = ctx.saved_tensors
        g = torch.zeros(ctx.shape, dtype=torch.float32, device=d.device)
        d = d.contiguous()
        check(_lib.lib().mla_embedding_bwd(ops._p(g), ops._p(idx.to(torch.int64)), ops._p(d), C.c_int64(d.shape[0]),
                                           C.c_int32(d.shape[1]), C.c_int64(-1), ops._stream()))
        return g.to(torch.bfloat16), None


class TactileContrastiveLoss(nn.Module):
    def __init__(self, feature_dim: int, projection_dim: int = 256, temperature: float = 0.07):
        super().__init__()
        self.temperature = temperature
        self.tactile_projection_head = _head(feature_dim, projection_dim)
        self.pointcloud_projection_head = _head(feature_dim, projection_dim)
        self.image_projection_head = _head(feature_dim, projection_dim)

    def forward(self, tac_features, pc_features, img_features, positive_pc_indices, linear_positive_img_indices):
        if tac_features.shape[0] == 0:
            return torch.tensor(0.0, device=tac_features.device, requires_grad=True)
        B, A, D = tac_features.shape
        K = pc_features.shape[1]
        tac = _L2NormFn.apply(_run_head(self.tactile_projection_head, tac_features.reshape(B * A, D))).view(B, A, -1)
        pc = _L2NormFn.apply(_run_head(self.pointcloud_projection_head, pc_features.reshape(B * K, D))).view(B, K, -1)
        img = _L2NormFn.apply(_run_head(self.image_projection_head, img_features.reshape(B * img_features.shape[1], D)))
        img = img.view(B, img_features.shape[1], -1)
        loss_pc = _TacNCEFn.apply(tac, pc, positive_pc_indices.reshape(B, A).to(torch.int64), self.temperature)
        loss_img = _TacNCEFn.apply(tac, img, linear_positive_img_indices.reshape(B, A).to(torch.int64), self.temperature)
        return (loss_pc + loss_img) / 2
