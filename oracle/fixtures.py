"""Deterministic weights and synthetic RLBench-shaped batches shared by the golden generator, the CPU tests, the GPU
parity tests and bench.py (test infrastructure; shapes/distributions from SURVEY.md §8d)."""
from __future__ import annotations

import zlib
from typing import Dict

import torch


def fill_state_dict(sd: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Overwrite every floating tensor of `sd` IN PLACE with values that depend only on (seed, key name, shape):
    matrices ~ N(0, 1/fan_in), biases ~ N(0, 0.02), 1-D `weight`s (norm scales) ~ 1 + N(0, 0.1); BatchNorm running
    statistics and integer buffers are left alone.  The reference, the oracle and the CUDA modules have identical
    keys, so they all get identical weights without shipping them."""
    for name in sorted(sd):
        t = sd[name]
        if not torch.is_floating_point(t) or name.endswith(("running_mean", "running_var")):
            continue
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 31))
        if t.dim() >= 2:
            fan_in = t[0].numel()
            v = torch.randn(t.shape, generator=g) * (fan_in ** -0.5)
        elif name.endswith("weight"):
            v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
        else:
            v = 0.02 * torch.randn(t.shape, generator=g)
        with torch.no_grad():
            t.copy_(v.to(t.dtype))
    return sd


def synthetic_batch(B: int, Lt: int, T: int = 0, image_hw: int = 672, n_points: int = 1024, seed: int = 1234,
                    use_pointcloud: bool = False, use_tactile: bool = False, pad_last: int = 0,
                    extra_views: int = 0, generation: bool = False) -> Dict:
    """CPU fp32 batch in the collator's contract (util/data_utils.py:100-195): images f32 [B,4,H,W] with an all-ones
    mask channel, point clouds inside the RLBench workspace box, ids with BOS first and EOS (id 2) last."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, 31000, (B, Lt), generator=g)
    ids[:, 0] = 1
    ids[:, -1] = 2
    am = torch.ones(B, Lt, dtype=torch.bool)
    if pad_last:                      # right-pad the last sample: EOS moves left, pad id 32000 follows
        ids[-1, Lt - pad_last - 1] = 2
        ids[-1, Lt - pad_last:] = 32000
        am[-1, Lt - pad_last:] = False
    labels = ids.clone()
    labels[~am] = -100
    img = lambda: torch.cat([torch.randn(B, 3, image_hw, image_hw, generator=g), torch.ones(B, 1, image_hw, image_hw)], 1)
    images = {"front_image": img()}
    for v in range(extra_views):
        images[f"wrist_{v}"] = img()
    batch = dict(input_ids=ids, attention_mask=am, labels=labels, images=images,
                 actions=torch.rand(B, T + 1, 7, generator=g) * 2 - 1, proprio=torch.rand(B, 1, 7, generator=g) * 2 - 1,
                 action_masks=torch.ones(B, T + 1, dtype=torch.bool), camera_name="rlbench_front")
    if use_pointcloud:
        box_lo = torch.tensor([-0.1, -0.5, 0.75])
        box_sz = torch.tensor([0.8, 1.0, 0.8])
        batch["point_cloud"] = torch.rand(B, n_points, 3, generator=g) * box_sz + box_lo
    if use_tactile:
        batch["tactile"] = torch.rand(B, 12, generator=g)
        batch["gripper_xyz"] = torch.rand(B, 3, generator=g) * torch.tensor([0.8, 1.0, 0.8]) + torch.tensor([-0.1, -0.5, 0.75])
    if generation:      # post-training targets: next frame, next point cloud, next tactile reading
        g2 = torch.Generator().manual_seed(seed + 77)
        batch["next_images"] = torch.randn(B, 3, image_hw, image_hw, generator=g2)
        batch["next_point_cloud"] = (torch.rand(B, n_points, 3, generator=g2) * torch.tensor([0.8, 1.0, 0.8])
                                     + torch.tensor([-0.1, -0.5, 0.75]))
        batch["next_tactile"] = torch.rand(B, 12, generator=g2)
    return batch
