"""Oracle (test infrastructure): the reference's DDIM action sampler restated in numpy / PyTorch.

Follows models/diffusion/__init__.py:11-47 (`create_diffusion`), respace.py:12-66 (`space_timesteps`), :75-92
(`SpacedDiffusion.__init__`: re-derived betas over the kept timesteps), gaussian_diffusion.py:124-140
(`betas_for_alpha_bar`, squaredcos_cap_v2), :152-186 (schedule tables in float64), :342-352 (x_start <-> eps), :522-571
(`ddim_sample`, eta = 0) and :640-689 (`ddim_sample_loop_progressive`) as called by MLA.predict_action_diff
(models/mla/model_mla.py:746-755: clip_denoised=False, eta=0.0) with create_ddim (:1166-1173: "ddim<N>", 100 steps,
squaredcos_cap_v2, epsilon prediction, fixed-small variance).  Pinned by tests/golden/ddim_*.npz (recorded from the
unmodified reference) in tests/test_sampler_cpu.py.
"""
from __future__ import annotations

import math
from typing import Callable, List, Tuple

import numpy as np
import torch


def base_alphas_cumprod(steps: int = 100) -> np.ndarray:
    """gaussian_diffusion.py:112-140,:160-163 — squaredcos_cap_v2 betas (max 0.999) -> cumprod(1 - beta), float64."""
    ab = lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    betas = np.array([min(1 - ab((i + 1) / steps) / ab(i / steps), 0.999) for i in range(steps)], dtype=np.float64)
    return np.cumprod(1.0 - betas, axis=0)


def ddim_timesteps(num_timesteps: int, count: int) -> List[int]:
    """space_timesteps(num_timesteps, f"ddim{count}") — respace.py:34-43: the first integer stride that yields exactly
    `count` steps (count == 1 is special-cased to {50})."""
    if count == 1:
        return [50]
    for i in range(1, num_timesteps):
        if len(range(0, num_timesteps, i)) == count:
            return list(range(0, num_timesteps, i))
    raise ValueError(f"cannot create exactly {count} steps with an integer stride")


def ddim_schedule(ddim_steps: int, diffusion_steps: int = 100) -> Tuple[List[int], np.ndarray]:
    """(timestep_map, alphas_cumprod of the respaced process).  SpacedDiffusion rebuilds betas as
    1 - ac[i] / ac[last kept] and GaussianDiffusion takes their cumprod again (respace.py:82-91), which reproduces
    ac[kept] up to float64 rounding — the round trip is kept so the tables are bit-identical to the reference's."""
    ac = base_alphas_cumprod(diffusion_steps)
    keep = ddim_timesteps(diffusion_steps, ddim_steps)
    last, betas = 1.0, []
    for i in keep:
        betas.append(1 - ac[i] / last)
        last = ac[i]
    return keep, np.cumprod(1.0 - np.array(betas, dtype=np.float64), axis=0)


def ddim_tables(ddim_steps: int, diffusion_steps: int = 100) -> Tuple[List[int], np.ndarray]:
    """Per respaced step i the four fp32 coefficients the update uses (the reference extracts float64 tables and casts
    `.float()`, gaussian_diffusion.py:866-881): sqrt(1/ac), sqrt(1/ac - 1), sqrt(ac_prev), sqrt(1 - ac_prev)."""
    keep, ac = ddim_schedule(ddim_steps, diffusion_steps)
    ac_prev = np.append(1.0, ac[:-1])
    tab = np.stack([np.sqrt(1.0 / ac), np.sqrt(1.0 / ac - 1.0), ac_prev, ac], axis=1)      # float64
    return keep, tab


def ddim_step(x: torch.Tensor, eps: torch.Tensor, i: int, tab: np.ndarray) -> torch.Tensor:
    """One ddim_sample with eta = 0, clip_denoised=False, epsilon-predicting model (fp32, the reference's op order)."""
    r1, r2 = np.float32(tab[i, 0]), np.float32(tab[i, 1])
    ab_prev, ab = torch.tensor(np.float32(tab[i, 2])), torch.tensor(np.float32(tab[i, 3]))
    x = x.float()
    eps = eps.float()
    pred_xstart = float(r1) * x - float(r2) * eps                       # _predict_xstart_from_eps
    eps2 = (float(r1) * x - pred_xstart) / float(r2)                    # _predict_eps_from_xstart (re-derived, :549)
    sigma = 0.0 * torch.sqrt((1 - ab_prev) / (1 - ab)) * torch.sqrt(1 - ab / ab_prev)
    return pred_xstart * torch.sqrt(ab_prev) + torch.sqrt(1 - ab_prev - sigma ** 2) * eps2


def ddim_sample_loop(model: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], noise: torch.Tensor, ddim_steps: int,
                     diffusion_steps: int = 100, trace: list = None) -> torch.Tensor:
    """ddim_sample_loop: i = N-1 .. 0; the model sees the ORIGINAL timestep timestep_map[i] (respace.py:120-131)."""
    keep, tab = ddim_tables(ddim_steps, diffusion_steps)
    x = noise.float()
    for i in reversed(range(len(keep))):
        t = torch.full((x.shape[0],), keep[i], dtype=torch.long)
        eps = model(x, t)
        if trace is not None:
            trace.append((t, x, eps))
        x = ddim_step(x, eps, i, tab)
    return x
