"""MLA policy wrapper — drop-in for models/mla/model_mla.py:47-276 (training forward).

`MLA.forward` keeps the reference's signature, RNG draw order (noise ~ randn_like(actions_future), then
t ~ randint(0, 100)), loss assembly — including the in-place aliasing that makes the reported `diff_loss` equal
`total_loss` (:215-232) — and returns `(loss_dict, output)`.  Differences that do not change results:
  * the batch is moved to the GPU here (FSDP's root pre-forward does it in the reference);
  * identical image tensors of the `repeated_diffusion_steps` copies are tokenised once (the frozen tokenizer is
    deterministic) and the tokens are tiled, instead of tiling 4x the 7 MB/sample pixel tensors (:159-165);
  * no `print(loss_dict)` host sync per step (:233) unless MLA.verbose is set.
Inference: `denoise_actions` is the device-side core of `predict_action_diff` (:592-775) — the DDIM loop over the
diffusion head — with the decoder prefix computed once and K/V-cached (the reference re-encodes all 548 tokens at each
of the 8 steps); `predict_action_diff` wraps it with the reference's prompt / normalisation plumbing.  The
autoregressive samplers (`predict_action_ar`, `_diff_ar`, `_batch`), `from_pretrained` and EMA are out of scope.
"""
from __future__ import annotations

import os

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
    # SURVEY 8 f2 (opt-in, MLA_SHARE_PREFIX=1 or mla.share_diffusion_prefix = True): the R diffusion copies of a sample
    # share one decoder prefix in training (image-only configurations: no per-copy randomness in front of the suffix).
    # Same loss and gradients as the repeated batch, ~R x fewer decoder rows.
    share_diffusion_prefix = os.environ.get("MLA_SHARE_PREFIX", "0") == "1"

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
        share = (self.share_diffusion_prefix and R > 1 and self.training and not (
            self.use_pointcloud or self.use_tactile or self.use_contrastive or self.use_generation))

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

        if share:
            # the R copies differ only in (t, x): run the decoder over ONE prefix per sample + R suffix
            # groups (B * (S + (R-1)*(T+3)) rows instead of B * R * S) — identical loss and gradients
            B0 = input_ids.shape[0] // R
            output, noise_pred, generation_outputs, generation_losses = self.vlm.forward_shared_prefix(
                x=x, t=timestep, proprio=proprio[:B0], input_ids=input_ids[:B0],
                attention_mask=attention_mask[:B0] if attention_mask is not None else None, images=images,
                camera_name=camera_name, repeats=R)
        else:
            output, noise_pred, generation_outputs, generation_losses = self._vlm_repeated(
                input_ids, attention_mask, images, next_images, camera_name, point_cloud, next_point_cloud, tactile,
                next_tactile, labels, x, timestep, proprio, gripper_xyz, inputs_embeds, past_key_values, use_cache,
                output_attentions, output_hidden_states, return_dict, R)
        return self._losses(output, noise_pred, noise, actions, timestep, generation_losses)

    def _vlm_repeated(self, input_ids, attention_mask, images, next_images, camera_name, point_cloud, next_point_cloud,
                      tactile, next_tactile, labels, x, timestep, proprio, gripper_xyz, inputs_embeds, past_key_values,
                      use_cache, output_attentions, output_hidden_states, return_dict, R):
        return self.vlm(
            input_ids=input_ids, attention_mask=attention_mask, images=images, next_images=next_images,
            camera_name=camera_name, point_cloud=point_cloud, next_point_cloud=next_point_cloud, tactile=tactile,
            next_tactile=next_tactile, labels=labels, x=x, t=timestep, proprio=proprio, gripper_xyz=gripper_xyz,
            inputs_embeds=inputs_embeds, past_key_values=past_key_values, use_cache=use_cache,
            output_attentions=output_attentions, output_hidden_states=output_hidden_states, return_dict=return_dict,
            use_diff=self.use_diff, image_repeat=R)

    def _losses(self, output, noise_pred, noise, actions, timestep, generation_losses):
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

    # ------------------------------------------------------------------ inference (models/mla/model_mla.py:592-775)
    def create_ddim(self, ddim_step: int = 10, noise_schedule: str = "squaredcos_cap_v2", diffusion_steps: int = 100):
        """model_mla.py:1166-1173."""
        self.ddim_diffusion = create_diffusion(timestep_respacing="ddim" + str(ddim_step), noise_schedule=noise_schedule,
                                               diffusion_steps=diffusion_steps, sigma_small=True, learn_sigma=False)
        self._ddim_steps = ddim_step
        return self.ddim_diffusion

    @torch.no_grad()
    def denoise_actions(self, input_ids, images, point_cloud=None, proprio=None, camera_name: str = "rlbench_front",
                        noise: Optional[torch.Tensor] = None, num_ddim_steps: int = 8, use_kv_cache: bool = True,
                        tactile=None, gripper_xyz=None, use_cuda_graph: bool = True) -> torch.Tensor:
        """`prepare_diffusion` + `sample_diffusion` of predict_action_diff (:709-766) without classifier-free guidance
        (cfg_scale <= 1: the only live branch — the reference's CFG branch calls a `forward_with_cfg` that
        PrismaticVLM does not define).  input_ids end with the tag token 29871 (the reference strips its last three
        ids, :714-715).  Returns the normalised action chunk f32 [B, T+1, action_dim] (on the device).

        use_kv_cache=True: prefix once + per-step suffix (see PrismaticVLM.denoise_prefill); with use_cuda_graph the
        decoder prefill and the whole DDIM loop are replayed as two CUDA graphs (PrismaticVLM.denoise_session) — the
        tokenizers and the splice stay eager.  use_kv_cache=False: the reference's schedule — the whole eval forward
        (tokenizers included) at every step."""
        self.vlm.eval()
        dev = self.vlm.llm_backbone.llm.lm_head.weight.device
        if getattr(self, "ddim_diffusion", None) is None or getattr(self, "_ddim_steps", None) != num_ddim_steps:
            self.create_ddim(ddim_step=num_ddim_steps)
        B = input_ids.shape[0]
        if noise is None:
            noise = torch.randn(B, self.future_action_window_size + 1, self.vlm.action_dim, device=dev)
        noise = noise.to(dev).float()
        if use_kv_cache and use_cuda_graph:
            q = self.vlm.denoise_prefill(input_ids, images, point_cloud=point_cloud, proprio=proprio,
                                         camera_name=camera_name, tactile=tactile, gripper_xyz=gripper_xyz,
                                         n_x=noise.shape[1], embeds_only=True)
            sess = self.vlm.denoise_session(q.B, q.P, q.n_x, self.ddim_diffusion)
            sess.prefix.copy_(q.prefix)
            sess.noise.copy_(noise)
            sess.g_prefill.replay()
            sess.g_loop.replay()
            return sess.out.clone()
        if use_kv_cache:
            st = self.vlm.denoise_prefill(input_ids, images, point_cloud=point_cloud, proprio=proprio,
                                          camera_name=camera_name, tactile=tactile, gripper_xyz=gripper_xyz,
                                          n_x=noise.shape[1])
            model = lambda x, t: self.vlm.denoise_step(st, x, t)
            return self.ddim_diffusion.ddim_sample_loop(model, noise.shape, noise, clip_denoised=False, eta=0.0)
        kw = dict(input_ids=input_ids, images=images, point_cloud=point_cloud, proprio=proprio, camera_name=camera_name,
                  tactile=tactile, gripper_xyz=gripper_xyz)
        return self.ddim_diffusion.ddim_sample_loop(self.vlm.forward, noise.shape, noise, clip_denoised=False,
                                                    model_kwargs=kw, eta=0.0)

    @torch.no_grad()
    def predict_action_diff(self, image=None, pointcloud=None, instruction: Optional[str] = None, cur_robot_state=None,
                            unnorm_key: Optional[str] = None, cfg_scale: float = 0.0, use_ddim: bool = True,
                            num_ddim_steps: int = 8, action_dim: int = 7, camera_name: str = "rlbench_front",
                            use_kv_cache: bool = True, **kwargs):
        """model_mla.py:592-775 — prompt, CLIP preprocessing, proprio normalisation and action un-normalisation are the
        reference's host-side plumbing (HF tokenizer / image processor objects owned by the backbone); the denoise
        loop is `denoise_actions`.  Returns the un-normalised action(s) as a numpy array, like the reference."""
        import numpy as np
        if cfg_scale > 1.0:
            raise NotImplementedError("classifier-free guidance: the reference's branch calls PrismaticVLM."
                                      "forward_with_cfg, which it does not define")
        if not use_ddim or num_ddim_steps is None:
            raise NotImplementedError("DDPM ancestral sampling (use_ddim=False) is not built; the reference's "
                                      "evaluation uses DDIM")
        self.vlm.eval()
        dev = self.vlm.llm_backbone.llm.lm_head.weight.device
        tokenizer = self.vlm.llm_backbone.tokenizer
        prompt_builder = self.vlm.get_prompt_builder()
        prompt_builder.add_turn(role="human", message=f"What action should the robot take to {instruction.lower()}?")
        input_ids = tokenizer(prompt_builder.get_prompt(), truncation=True, return_tensors="pt").input_ids.to(dev)
        if not torch.all(input_ids[:, -1] == 29871):                                          # :642-643
            input_ids = torch.cat((input_ids, torch.tensor([[29871, 32001, 32002, 29871]], device=dev)), dim=1)
        input_ids = input_ids[:, :-3]                                                         # :714-715
        px = self.vlm.get_vision_tower_2d().image_processor.preprocess(image, return_tensors="pt")["pixel_values"][0]
        px = torch.cat([px, torch.ones(1, px.shape[-2], px.shape[-1])], dim=0).unsqueeze(0).to(dev)
        if isinstance(pointcloud, np.ndarray):
            pointcloud = torch.from_numpy(pointcloud)
        if pointcloud is not None:
            pointcloud = pointcloud.to(dev).contiguous()
        proprio = None
        if cur_robot_state is not None:
            st = self.get_proprio_stats(unnorm_key)
            mask = st.get("mask", np.ones_like(st["q01"], dtype=bool))
            hi, lo = np.array(st["q99"]), np.array(st["q01"])
            cur = np.clip(np.where(mask, 2 * (cur_robot_state - lo) / (hi - lo + 1e-8) - 1, cur_robot_state), -1, 1)
            proprio = torch.tensor(cur, dtype=torch.float32).unsqueeze(0).unsqueeze(0).to(dev)
        samples = self.denoise_actions(input_ids, {"front_image": px}, point_cloud=pointcloud, proprio=proprio,
                                       camera_name=camera_name, num_ddim_steps=num_ddim_steps,
                                       use_kv_cache=use_kv_cache)
        normalized = np.clip(samples[0].cpu().numpy(), -1, 1)
        if normalized.ndim == 1:                                                              # :683-700 gripper bit(s)
            for g in range(6, normalized.shape[0], 7):
                normalized[g] = np.where(normalized[g] < 0.5, 0, 1)
        else:
            for g in range(6, normalized.shape[1], 7):
                normalized[:, g] = np.where(normalized[:, g] < 0.5, 0, 1)
        st = self.get_action_stats(unnorm_key)
        mask = st.get("mask", np.ones_like(st["q01"], dtype=bool))
        hi, lo = np.array(st["q99"]), np.array(st["q01"])
        return np.where(mask, 0.5 * (normalized + 1) * (hi - lo) + lo, normalized)

    @staticmethod
    def _check_unnorm_key(norm_stats, unnorm_key):
        """model_mla.py:1175-1192."""
        if unnorm_key is None:
            assert len(norm_stats) == 1, (f"Your model was trained on more than one dataset, please pass a `unnorm_key` "
                                          f"from the following options to choose the statistics used for "
                                          f"un-normalizing actions: {norm_stats.keys()}")
            unnorm_key = next(iter(norm_stats.keys()))
        assert unnorm_key in norm_stats, (f"The `unnorm_key` you chose is not in the set of available statistics; "
                                          f"choose from: {norm_stats.keys()}")
        return unnorm_key

    def get_action_stats(self, unnorm_key=None):
        """model_mla.py:1199-1204."""
        return self.norm_stats[self._check_unnorm_key(self.norm_stats, unnorm_key)]["action"]

    def get_proprio_stats(self, unnorm_key=None):
        return self.norm_stats[self._check_unnorm_key(self.norm_stats, unnorm_key)]["proprio"]
