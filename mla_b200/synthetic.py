"""Synthetic RLBench-shaped batches in the collator's contract (util/data_utils.py:100-195), for bench.py / smoke.
Shapes and distributions follow SURVEY.md §8d; tensors are CPU fp32 (optionally pinned), as the DataLoader emits."""
from __future__ import annotations

from typing import Dict

import torch


def make_batch(B: int, Lt: int = 32, T: int = 0, image_hw: int = 672, n_points: int = 1024, seed: int = 1234,
               use_pointcloud: bool = False, use_tactile: bool = False, extra_views: int = 0, pin: bool = False,
               generation: bool = False) -> Dict:
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, 31000, (B, Lt), generator=g)
    ids[:, 0] = 1
    ids[:, -1] = 2
    am = torch.ones(B, Lt, dtype=torch.bool)
    labels = ids.clone()

    def img():
        return torch.cat([torch.randn(B, 3, image_hw, image_hw, generator=g), torch.ones(B, 1, image_hw, image_hw)], 1)
    images = {"front_image": img()}
    for v in range(extra_views):
        images[f"wrist_{v}"] = img()
    batch = dict(input_ids=ids, attention_mask=am, labels=labels, images=images,
                 actions=torch.rand(B, T + 1, 7, generator=g) * 2 - 1, proprio=torch.rand(B, 1, 7, generator=g) * 2 - 1,
                 action_masks=torch.ones(B, T + 1, dtype=torch.bool))
    if use_pointcloud:
        batch["point_cloud"] = (torch.rand(B, n_points, 3, generator=g) * torch.tensor([0.8, 1.0, 0.8]) +
                                torch.tensor([-0.1, -0.5, 0.75]))
    if use_tactile:
        batch["tactile"] = torch.rand(B, 12, generator=g)
        batch["gripper_xyz"] = torch.rand(B, 3, generator=g) * torch.tensor([0.8, 1.0, 0.8]) + torch.tensor([-0.1, -0.5, 0.75])
    if generation:      # post-training targets (next frame / point cloud / tactile reading)
        batch["next_images"] = torch.randn(B, 3, image_hw, image_hw, generator=g)
        batch["next_point_cloud"] = (torch.rand(B, n_points, 3, generator=g) * torch.tensor([0.8, 1.0, 0.8]) +
                                     torch.tensor([-0.1, -0.5, 0.75]))
        batch["next_tactile"] = torch.rand(B, 12, generator=g)
    if pin and torch.cuda.is_available():
        batch = map_tensors(batch, lambda t: t.pin_memory())
    return batch


def map_tensors(batch: Dict, fn) -> Dict:
    out = {}
    for k, v in batch.items():
        if isinstance(v, dict):
            out[k] = {kk: fn(vv) for kk, vv in v.items()}
        elif torch.is_tensor(v):
            out[k] = fn(v)
        else:
            out[k] = v
    return out


def batch_bytes(batch: Dict) -> int:
    n = 0
    for v in batch.values():
        if isinstance(v, dict):
            n += sum(t.numel() * t.element_size() for t in v.values())
        elif torch.is_tensor(v):
            n += v.numel() * v.element_size()
    return n
