"""DDIM action sampler of the inference path — the part of models/diffusion/{respace,gaussian_diffusion}.py that
MLA.predict_action_diff drives (models/mla/model_mla.py:743-755: `create_ddim(num_ddim_steps)` then
`ddim_diffusion.ddim_sample_loop(model, noise.shape, noise, clip_denoised=False, model_kwargs=..., eta=0.0)`).

`SpacedDiffusion` keeps that call surface (timestep_map, num_timesteps, ddim_sample_loop); the schedule is built on
the host in float64 exactly as the reference does (space_timesteps "ddimN" respace.py:34-43; re-derived betas and
their cumprod :82-91; tables gaussian_diffusion.py:160-186) and the per-step update runs in `mla_ddim_step`, which
reproduces ddim_sample's fp32 op order bit for bit (eta = 0: no noise is drawn).  DDPM ancestral sampling
(p_sample_loop, use_ddim=False) and eta > 0 are not built.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from . import ops


def space_timesteps_ddim(num_timesteps: int, count: int) -> List[int]:
    """respace.py:34-43 — fixed DDIM striding; `count == 1` is the reference's special case {50}."""
    if count == 1:
        return [50]
    for i in range(1, num_timesteps):
        if len(range(0, num_timesteps, i)) == count:
            return list(range(0, num_timesteps, i))
    raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")


class SpacedDiffusion:
    """The respaced process of `create_diffusion("ddim<N>", ...)`: epsilon-predicting model, fixed-small variance."""

    def __init__(self, base_alphas_cumprod: np.ndarray, use_timesteps: List[int]):
        self.original_num_steps = len(base_alphas_cumprod)
        self.timestep_map = sorted(set(use_timesteps))
        last, betas = 1.0, []
        for i in self.timestep_map:
            betas.append(1 - base_alphas_cumprod[i] / last)
            last = base_alphas_cumprod[i]
        self.betas = np.array(betas, dtype=np.float64)
        self.num_timesteps = len(betas)
        self.alphas_cumprod = np.cumprod(1.0 - self.betas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        ac, acp = self.alphas_cumprod, self.alphas_cumprod_prev
        # what ddim_sample reads per step, cast to fp32 the way _extract_into_tensor does (:866-881), then sqrt in fp32
        acp32 = acp.astype(np.float32)
        self._coef = np.stack([np.sqrt(1.0 / ac).astype(np.float32), np.sqrt(1.0 / ac - 1.0).astype(np.float32),
                               np.sqrt(acp32), np.sqrt(np.float32(1.0) - acp32)], axis=1).astype(np.float32)
        self._dev: Dict[str, torch.Tensor] = {}

    def coef(self, device) -> torch.Tensor:
        k = str(device)
        if k not in self._dev:
            self._dev[k] = torch.from_numpy(self._coef).to(device).contiguous()
        return self._dev[k]

    def ddim_sample_loop(self, model: Callable, shape=None, noise: Optional[torch.Tensor] = None,
                         clip_denoised: bool = False, denoised_fn=None, cond_fn=None, model_kwargs: Optional[dict] = None,
                         device=None, progress: bool = False, eta: float = 0.0) -> torch.Tensor:
        """gaussian_diffusion.py:603-689 for the arguments predict_action_diff passes.  `model(x, t, **model_kwargs)`
        returns the noise prediction or a tuple whose last element is it (p_mean_variance :280-284); t holds ORIGINAL
        timesteps (the reference's _WrappedModel maps them, respace.py:120-131)."""
        if clip_denoised or denoised_fn is not None or cond_fn is not None or eta != 0.0:
            raise NotImplementedError("only the predict_action_diff configuration is built: clip_denoised=False, "
                                      "no denoised_fn / cond_fn, eta = 0")
        kw = model_kwargs or {}
        if noise is None:
            noise = torch.randn(*shape, device=device)
        x = noise.float().contiguous()
        coef = self.coef(x.device)
        for i in reversed(range(self.num_timesteps)):
            t = torch.full((x.shape[0],), self.timestep_map[i], dtype=torch.long, device=x.device)
            out = model(x, t, **kw)
            eps = out[-1] if isinstance(out, tuple) else out
            x = ops.ddim_step(x, eps, coef[i])
        return x
