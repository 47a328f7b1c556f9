"""MLA policy wrapper — drop-in for models/mla/model_mla.py:47-276 (training forward).

`MLA.forward` keeps the reference's signature, RNG draw order (noise ~ randn_like(actions_future), then
t ~ randint(0, 100)), loss assembly — including the in-place aliasing that makes the reported `diff_loss` equal
`total_loss` (:215-232) — and returns `(loss_dict, output)`.  Differences that do not change results:
  * the batch is moved to the GPU here (FSDP's root pre-forward does it in the reference);
  * identical image tensors of the `repeated_diffusion_steps` copies are tokenised once (the frozen tokenizer is
    deterministic) and the tokens are tiled, instead of tiling 4x the 7 MB/sample pixel tensors (:159-165);
  * no `print(loss_dict)` host sync per step (:233) unless MLA.verbose is set.
Inference (`predict_action_*`), `from_pretrained` and EMA are out of the hot-path scope.
"""
from __future__ import annotations

from functools import partial
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .modules import MLP_GELU, MLPProjector, create_diffusion
from .vision import VisionTokenizer
from .vlm import PrismaticVLM

IGNORE_INDEX = -100


class MLA(nn.Module):
    verbose = False

    def __init__(self, vlm: PrismaticVLM, action_tokenizer=None, token_size: int = 4096, action_dim: int = 7,
                 future_action_window_size: int = 15, past_action_window_size: int = 0, use_ema: bool = False,
                 norm_stats=None, use_diff: bool = False, use_pointcloud: bool = False, use_tactile: bool = False,
                 use_contrastive: bool = False, use_generation: bool = False, gen_image: bool = False,
                 use_roi: bool = False, gen_pointcloud: bool = False, gen_tactile: bool = False, **kwargs) -> None:
        super().__init__()
        self.action_tokenizer = action_tokenizer
        self.use_diff, self.use_pointcloud, self.use_tactile = use_diff, use_pointcloud, use_tactile
        self.use_contrastive, self.use_generation = use_contrastive, use_generation
        self.gen_image, self.use_roi, self.gen_pointcloud, self.gen_tactile = gen_image, use_roi, gen_pointcloud, gen_tactile
        self.vlm = vlm
        self.future_action_window_size = future_action_window_size
        self.vlm.future_action_window_size = future_action_window_size
        self.past_action_window_size = past_action_window_size
        self.all_module_keys = ["vlm." + k for k in self.vlm.all_module_keys]
        self.use_ema = use_ema
        self.norm_stats = norm_stats
        self._trainable_module_keys: List[str] = []
        if self.use_diff:
            self.ddim_diffusion = None
            self.diffusion_steps = 100
            self.diffusion = create_diffusion(timestep_respacing="", noise_schedule="squaredcos_cap_v2",
                                              diffusion_steps=100, sigma_small=True, learn_sigma=False)

    @property
    def trainable_module_keys(self) -> List[str]:
        return ["vlm." + k for k in self.vlm.trainable_module_keys] + self._trainable_module_keys

    @property
    def llm_backbone(self):
        return self.vlm.llm_backbone

    def freeze_backbones(self, stage):
        self.vlm.freeze_backbones(stage)

    def get_fsdp_wrapping_policy(self) -> Callable:
        from torch.distributed.fsdp.wrap import _module_wrap_policy, _or_policy
        from .pointcloud import PointTokenizer
        return partial(_or_policy, policies=[
            partial(_module_wrap_policy, module_classes={PointTokenizer, VisionTokenizer}),
            self.vlm.llm_backbone.get_fsdp_wrapping_policy(),
            partial(_module_wrap_policy, module_classes={MLPProjector, MLP_GELU}),
        ])

    def forward(self, input_ids=None, attention_mask=None, images=None, next_images=None, camera_name=None,
                point_cloud=None, next_point_cloud=None, tactile=None, next_tactile=None, labels=None, actions=None,
                proprio=None, gripper_xyz=None, inputs_embeds=None, past_key_values=None, use_cache=None,
                output_attentions=None, output_hidden_states=None, return_dict=None,
                repeated_diffusion_steps: int = 4, action_masks=None, use_diff: Optional[bool] = None) -> Tuple:
        if use_diff is not None:
            self.use_diff = use_diff
        if not self.use_diff:
            raise NotImplementedError("autoregressive action training (use_diff=False) needs the vocabulary CE; "
                                      "the MLA recipes (scripts/*_rlbench.sh) train the diffusion head")
        dev = self.vlm.llm_backbone.llm.lm_head.weight.device
        R = repeated_diffusion_steps

        def rep(v):
            v = v.to(dev, non_blocking=True)
            return v.repeat(R, *([1] * (v.ndimension() - 1)))

        proprio = rep(proprio)
        actions = rep(actions)
        actions_future = actions[:, -(self.future_action_window_size + 1):, :]
        input_ids, attention_mask, labels = rep(input_ids), rep(attention_mask), rep(labels)
        if action_masks is not None:
            action_masks = rep(action_masks)
        if self.use_pointcloud:
            point_cloud = rep(point_cloud)       # FPS draws a fresh random start per copy: copies are NOT shared
        if self.use_tactile:
            tactile, gripper_xyz = rep(tactile), rep(gripper_xyz)

        noise = torch.randn_like(actions_future)
        timestep = torch.randint(0, self.diffusion.num_timesteps, (actions_future.size(0),), device=actions.device)
        x = self.diffusion.q_sample(actions_future, timestep, noise)

        output, noise_pred, generation_outputs, generation_losses = self.vlm(
            input_ids=input_ids, attention_mask=attention_mask, images=images, next_images=next_images,
            camera_name=camera_name, point_cloud=point_cloud, next_point_cloud=next_point_cloud, tactile=tactile,
            next_tactile=next_tactile, labels=labels, x=x, t=timestep, proprio=proprio, gripper_xyz=gripper_xyz,
            inputs_embeds=inputs_embeds, past_key_values=past_key_values, use_cache=use_cache,
            output_attentions=output_attentions, output_hidden_states=output_hidden_states, return_dict=return_dict,
            use_diff=self.use_diff, image_repeat=R)
        assert noise_pred.shape == noise.shape == actions.shape
        zero = torch.tensor(0, dtype=torch.float32)
        loss_dict = {k: zero for k in ("total_loss", "img_pc_contrastive_loss", "tactile_contrastive_loss",
                                       "diff_loss", "image_gen_loss", "point_cloud_gen_loss", "tactile_gen_loss")}
        diff_loss = ops.MSEFn.apply(noise_pred, noise)
        total = diff_loss
        # model_mla.py:218-226 — generation losses are added before the contrastive ones
        if self.use_generation and self.gen_image:
            loss_dict["image_gen_loss"] = generation_losses["image_gen_loss"]
            total = total + generation_losses["image_gen_loss"]
        if self.use_generation and self.gen_pointcloud:
            loss_dict["point_cloud_gen_loss"] = generation_losses["point_cloud_gen_loss"]
            total = total + generation_losses["point_cloud_gen_loss"]
        if self.use_generation and self.gen_tactile:
            loss_dict["tactile_gen_loss"] = generation_losses["tactile_gen_loss"]
            total = total + generation_losses["tactile_gen_loss"]
        if self.use_contrastive:
            loss_dict["img_pc_contrastive_loss"] = output.img_pc_contrastive_loss
            total = total + output.img_pc_contrastive_loss
            if self.use_tactile:
                loss_dict["tactile_contrastive_loss"] = output.tactile_contrastive_loss
                total = total + output.tactile_contrastive_loss
        # model_mla.py:215-232: `total_loss` and `diff_loss` alias one tensor that is `+=`-ed in place, so both keys
        # report the total.  Same observable values here, without the in-place op on a graph leaf.
        loss_dict["total_loss"] = total
        loss_dict["diff_loss"] = total
        output.noise = noise
        output.noise_pred = noise_pred
        output.timestep = timestep
        output.diff_loss_only = diff_loss
        if self.verbose:
            print(loss_dict)
        return loss_dict, output
