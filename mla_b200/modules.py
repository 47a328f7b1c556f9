"""Small trainable modules of the MLA hot path, parameter-compatible with the reference, computed by our kernels.

  MLP_GELU          models/mla/image/vision_tokenizer.py:79-89   (projector_2d)
  MLPProjector      util/nn_utils.py:22-36                        (projector_3d)
  Mlp / RmsNorm     timm==0.9.10 layers used by models/diffusion/models.py:18 (restated; timm is not vendored)
  ActionEmbedder    models/diffusion/models.py:112-123
  TimestepEmbedder  models/diffusion/models.py:28-65
  LabelEmbedder     models/diffusion/models.py:67-97
  FinalLayer        models/diffusion/models.py:173-189
  GaussianDiffusion models/diffusion/gaussian_diffusion.py:152-229 (schedule + q_sample only; samplers out of scope)

All forwards take/return bf16 CUDA tensors and run through `ops.linear` (tcgen05 GEMM with fused bias+activation
epilogue) so the parameters (fp32 masters, reference names) receive ordinary autograd gradients.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import ops


class MLP_GELU(nn.Module):
    def __init__(self, input_size: int, hidden_size: int, depth: int):
        super().__init__()
        layers = [nn.Linear(input_size, hidden_size)]
        for _ in range(1, depth):
            layers.append(nn.GELU())
            layers.append(nn.Linear(hidden_size, hidden_size))
        self.mlp = nn.Sequential(*layers)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        lin = [m for m in self.mlp if isinstance(m, nn.Linear)]
        for i, m in enumerate(lin):
            x = ops.linear(x, m.weight, m.bias, ops.ACT_GELU_ERF if i + 1 < len(lin) else ops.ACT_NONE)
        return x


class MLPProjector(nn.Module):
    def __init__(self, vision_dim: int, llm_dim: int, mlp_type: str = "gelu-mlp"):
        super().__init__()
        if mlp_type != "gelu-mlp":
            raise ValueError(f"Projector with `{mlp_type = }` is not supported!")
        self.projector = nn.Sequential(nn.Linear(vision_dim, llm_dim, bias=True), nn.GELU(),
                                       nn.Linear(llm_dim, llm_dim, bias=True))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = ops.linear(x, self.projector[0].weight, self.projector[0].bias, ops.ACT_GELU_ERF)
        return ops.linear(x, self.projector[2].weight, self.projector[2].bias)


class Mlp(nn.Module):
    """timm.layers.Mlp with drop=0 / norm=Identity: fc1 -> act -> fc2."""

    def __init__(self, in_features: int, hidden_features: Optional[int] = None, out_features: Optional[int] = None,
                 act: int = ops.ACT_GELU_TANH):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self._act = act

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = ops.linear(x, self.fc1.weight, self.fc1.bias, self._act)
        return ops.linear(x, self.fc2.weight, self.fc2.bias)


class RmsNorm(nn.Module):
    """timm RmsNorm(channels, eps=1e-6, affine=True).  `variance_mode` selects the arithmetic: False = mean of
    squares (timm >= 1.0 and the survey's shim), True = torch.var as timm 0.9.x is believed to do;
    parity for this module is unpinned because timm is not in the reference tree."""

    def __init__(self, channels: int, eps: float = 1e-6, variance_mode: bool = False):
        super().__init__()
        self.eps = eps
        self.variance_mode = variance_mode
        self.weight = nn.Parameter(torch.ones(channels))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.RMSNormFn.apply(x, self.weight, self.eps, int(self.variance_mode))


class ActionEmbedder(nn.Module):
    def __init__(self, action_size: int, hidden_size: int):
        super().__init__()
        self.mlp = Mlp(action_size, hidden_size, hidden_size, act=ops.ACT_GELU_TANH)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [..., action_size] (any float dtype) -> bf16 [..., hidden]."""
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        if x2.dtype != torch.bfloat16:
            x2 = x2.to(torch.bfloat16)
        return self.mlp(x2.contiguous()).reshape(*lead, -1)


class TimestepEmbedder(nn.Module):
    def __init__(self, hidden_size: int, frequency_embedding_size: int = 256, max_t: int = 1000):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size, bias=True))
        self.frequency_embedding_size = frequency_embedding_size
        self._max_t = max_t
        self._table = None

    @staticmethod
    def timestep_embedding(t: torch.Tensor, dim: int, max_period: int = 10000) -> torch.Tensor:
        """models/diffusion/models.py:41-60 verbatim semantics (fp32)."""
        half = dim // 2
        freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(t.device)
        args = t[:, None].float() * freqs[None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        if dim % 2:
            emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
        return emb

    def _freq_table(self, device) -> torch.Tensor:
        # The sinusoid depends only on the integer timestep: one constant bf16 table, built once with the reference
        # formula (timesteps reach the embedder as bf16, prismatic.py:877-878; integers < 256 are exact in bf16).
        if self._table is None or self._table.device != device:
            t = torch.arange(self._max_t, dtype=torch.float32).to(torch.bfloat16)
            self._table = self.timestep_embedding(t, self.frequency_embedding_size).to(torch.bfloat16).to(device)
        return self._table

    def forward(self, t: torch.Tensor) -> torch.Tensor:
        """t: integer-valued timesteps [B] -> bf16 [B, hidden]."""
        idx = t.to(torch.int64).contiguous()
        t_freq = ops.gather_rows(self._freq_table(t.device), idx)
        x = ops.linear(t_freq, self.mlp[0].weight, self.mlp[0].bias, ops.ACT_SILU)
        return ops.linear(x, self.mlp[2].weight, self.mlp[2].bias)


class LabelEmbedder(nn.Module):
    """Identity unless dropout_prob > 0, where whole-sequence conditions are zeroed per sample (token_drop :82-92)."""

    def __init__(self, in_size: int, hidden_size: int, dropout_prob: float = -1, conditions_shape=(1, 1, 4096)):
        super().__init__()
        self.dropout_prob = dropout_prob

    def forward(self, conditions: torch.Tensor, train: bool, force_drop_ids=None) -> torch.Tensor:
        use_dropout = self.dropout_prob > 0
        if (train and use_dropout) or (force_drop_ids is not None):
            if force_drop_ids is None:
                drop = torch.rand(conditions.shape[0], device=conditions.device) < self.dropout_prob
            else:
                drop = force_drop_ids == 1
            conditions = torch.where(drop.view(-1, 1, 1), torch.zeros_like(conditions), conditions)
        return conditions


class FinalLayer(nn.Module):
    def __init__(self, hidden_size: int, out_channels: int, variance_mode: bool = False):
        super().__init__()
        self.norm_final = RmsNorm(hidden_size, eps=1e-6, variance_mode=variance_mode)
        self.mlp = Mlp(hidden_size, hidden_size, out_channels, act=ops.ACT_GELU_TANH)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        lead = x.shape[:-1]
        y = self.mlp(self.norm_final(x.reshape(-1, x.shape[-1])))
        return y.reshape(*lead, -1)


# ------------------------------------------------------------------------------------------------ diffusion
def betas_for_alpha_bar(num_diffusion_timesteps: int, alpha_bar, max_beta: float = 0.999) -> np.ndarray:
    """models/diffusion/gaussian_diffusion.py:124-140."""
    betas = []
    for i in range(num_diffusion_timesteps):
        t1 = i / num_diffusion_timesteps
        t2 = (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return np.array(betas)


class GaussianDiffusion:
    """Schedule + q_sample of the reference's SpacedDiffusion for the training step (`create_diffusion("",
    "squaredcos_cap_v2", diffusion_steps=100, ...)`, models/mla/model_mla.py:97): with empty respacing every
    timestep is kept, so the spaced betas equal the base betas (respace.py:77-96)."""

    def __init__(self, betas: np.ndarray):
        betas = np.array(betas, dtype=np.float64)
        assert len(betas.shape) == 1 and (betas > 0).all() and (betas <= 1).all()
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self._dev_tables = {}

    def _tables(self, device):
        key = str(device)
        if key not in self._dev_tables:
            # _extract_into_tensor (:866-881) casts the float64 table entry to float32 before the multiply
            self._dev_tables[key] = (torch.from_numpy(self.sqrt_alphas_cumprod).to(device).float(),
                                     torch.from_numpy(self.sqrt_one_minus_alphas_cumprod).to(device).float())
        return self._dev_tables[key]

    def q_sample(self, x_start: torch.Tensor, t: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        if noise is None:
            noise = torch.randn_like(x_start)
        assert noise.shape == x_start.shape
        sa, sb = self._tables(x_start.device)
        return ops.q_sample(x_start.float(), noise.float(), t.to(torch.int64).contiguous(), sa, sb)


def create_diffusion(timestep_respacing="", noise_schedule="squaredcos_cap_v2", diffusion_steps=100, **_unused):
    """models/diffusion/__init__.py:11-47: the training process (empty respacing) or, for "ddim<N>", the respaced
    sampler MLA.create_ddim builds for inference (mla_b200/sampler.py)."""
    ddim = isinstance(timestep_respacing, str) and timestep_respacing.startswith("ddim")
    if not ddim and timestep_respacing not in (None, "", [diffusion_steps]):
        raise NotImplementedError("section-count respacing is not used by any MLA recipe")
    if noise_schedule == "squaredcos_cap_v2":
        betas = betas_for_alpha_bar(diffusion_steps, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    elif noise_schedule == "linear":
        scale = 1000 / diffusion_steps
        betas = np.linspace(scale * 0.0001, scale * 0.02, diffusion_steps, dtype=np.float64)
    else:
        raise NotImplementedError(f"unknown beta schedule: {noise_schedule}")
    if ddim:
        from .sampler import SpacedDiffusion, space_timesteps_ddim
        base = GaussianDiffusion(betas)
        return SpacedDiffusion(base.alphas_cumprod, space_timesteps_ddim(diffusion_steps, int(timestep_respacing[4:])))
    return GaussianDiffusion(betas)
