"""PrismaticVLM: multimodal fusion + LLM + diffusion head — drop-in for models/vlm/prismatic.py:148-1144.

Same constructor flags, attribute names, module keys (`all_module_keys`, `trainable_module_keys`), `freeze_backbones`
stages and `forward` returns as the reference; the compute is libmla_b200 kernels:

    tokenizers/projectors -> splice [BOS | pc | img (| views) | tac | text | proprio,t,x | EOS] (one index kernel +
    one row gather, no per-sample python loop or .item() sync, prismatic.py:981-1038) -> decoder stack ->
    FinalLayer on the T+1 noisy-action rows only (the reference runs it on all S rows and slices, :1117-1126 —
    row-wise ops, identical values).

Post-training generation heads (models/mla/generation, config 5): `generation_manager` (mla_b200/generation.py) runs
on the last hidden state in train mode exactly where the reference calls it (prismatic.py:1075-1113).
"""
from __future__ import annotations

import ctypes as C
from functools import partial
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import check
from .backbone import CausalLMOutputWithPast, LLMBackbone
from .contrastive import get_camera_params, project_points
from .generation import MultimodalGenerationManager, compute_generation_losses, roi_mask
from .modules import (ActionEmbedder, FinalLayer, LabelEmbedder, MLP_GELU, MLPProjector, TimestepEmbedder)
from .vision import VisionTokenizer

IGNORE_INDEX = -100


class PrismaticVLM(nn.Module):
    def __init__(self, model_id: str, llm_backbone: LLMBackbone, enable_mixed_precision_training: bool = True,
                 action_dim: int = 7, token_size: int = 4096, future_action_window_size: int = 0,
                 past_action_window_size: int = 0, class_dropout_prob: float = 0.0, norm_stats=None,
                 use_diff: bool = False, use_pointcloud: bool = False, use_tactile: bool = False,
                 use_contrastive: bool = False, llm_vision_layers: int = 1, use_generation: bool = True,
                 gen_image: bool = False, num_image_gen_queries: int = 128, image_decoder_layers: int = 3,
                 image_decoder_heads: int = 8, image_patch_size: int = 42, use_roi: bool = False,
                 roi_dilation_kernel_size: int = 3, gen_pointcloud: bool = True, gen_tactile: bool = True,
                 pointcloud_trans_dim: int = 1024, pointcloud_decoder_layers: int = 4, pointcloud_decoder_heads: int = 8,
                 pointcloud_group_size: int = 8, pointcloud_num_groups: int = 128, tactile_decoder_layers: int = 2,
                 tactile_decoder_heads: int = 4, image_hidden_dim: int = 1024, **kwargs) -> None:
        # positional order = the reference's (prismatic.py:149-185); image_hidden_dim (hard-coded 1024 there, :219) is ours
        super().__init__()
        self.model_family, self.model_id = "prismatic", model_id
        self.llm_backbone = llm_backbone
        self.enable_mixed_precision_training = enable_mixed_precision_training
        self.token_size = token_size
        self.use_diff, self.use_pointcloud, self.use_tactile = use_diff, use_pointcloud, use_tactile
        self.use_contrastive, self.llm_vision_layers = use_contrastive, llm_vision_layers
        self.use_generation = use_generation
        self.gen_image = gen_image and use_generation
        self.use_roi = use_roi
        self.gen_pointcloud = gen_pointcloud and use_generation
        self.gen_tactile = gen_tactile and use_generation
        self.roi_dilation_kernel_size = roi_dilation_kernel_size

        # prismatic.py:208-212 — likelihood helper tokens
        self.string2idx = {}
        for s in ["True", "False", "Yes", "No"] + [chr(ord("A") + i) for i in range(26)]:
            ids = self.llm_backbone.tokenizer.encode(s, add_special_tokens=False)
            assert len(ids) == 1, f'String "{s}" is tokenized as more than one token!'
            self.string2idx[s] = ids[0]

        self.norm_stats = norm_stats
        self.class_dropout_prob = class_dropout_prob
        self.future_action_window_size = future_action_window_size
        self.action_dim = action_dim

        self.image_hidden_dim = image_hidden_dim
        self.vision_tower_2d = VisionTokenizer(input_size=self.image_hidden_dim)
        self.projector_2d = MLP_GELU(self.image_hidden_dim, token_size, 2)
        if self.use_pointcloud:
            from .pointcloud import PointTokenizer
            self.vision_tower_3d = PointTokenizer(in_channels=3, embed_dim=768, depth=12, num_heads=12)
            self.projector_3d = MLPProjector(self.vision_tower_3d.embed_dim, token_size)
        self.tactile_dim = 12      # the reference sets it under use_tactile only and then crashes at :267 without it
        if self.use_tactile:
            self.tactile_embedder = ActionEmbedder(action_size=self.tactile_dim, hidden_size=token_size)
        self.proprio_embedder = ActionEmbedder(action_size=action_dim, hidden_size=token_size)
        if self.use_diff:
            self.x_embedder = ActionEmbedder(action_size=action_dim, hidden_size=token_size)
            self.t_embedder = TimestepEmbedder(token_size)
            self.z_embedder = LabelEmbedder(in_size=token_size, hidden_size=token_size, dropout_prob=class_dropout_prob)
            self.final_layer = FinalLayer(token_size, action_dim)
        if self.use_generation:         # prismatic.py:247-270
            self.generation_manager = MultimodalGenerationManager(
                token_size=token_size, use_image_generation=self.gen_image, num_image_gen_queries=num_image_gen_queries,
                image_decoder_layers=image_decoder_layers, image_decoder_heads=image_decoder_heads,
                image_patch_size=image_patch_size, use_roi=use_roi, roi_dilation_kernel_size=roi_dilation_kernel_size,
                use_pointcloud_generation=self.gen_pointcloud, pointcloud_trans_dim=pointcloud_trans_dim,
                pointcloud_decoder_layers=pointcloud_decoder_layers, pointcloud_decoder_heads=pointcloud_decoder_heads,
                pointcloud_group_size=pointcloud_group_size, pointcloud_num_groups=pointcloud_num_groups,
                use_tactile_generation=self.gen_tactile, tactile_dim=self.tactile_dim,
                tactile_decoder_layers=tactile_decoder_layers, tactile_decoder_heads=tactile_decoder_heads)

        self.all_module_keys = ["vision_tower_2d", "projector_2d", "llm_backbone", "proprio_embedder"]
        if self.use_diff:
            self.all_module_keys.extend(["x_embedder", "t_embedder", "final_layer"])
        if self.use_pointcloud:
            self.all_module_keys.extend(["vision_tower_3d", "projector_3d"])
        if self.use_tactile:
            self.all_module_keys.extend(["tactile_embedder"])
        if self.use_generation:
            self.all_module_keys.append("generation_manager")
        self.trainable_module_keys: List[str] = []
        self.vision_backbone_requires_grad = False
        self.initialize_weights()
        if self.use_pointcloud:
            self.vision_tower_3d.initialize_weights()
        self._err_flag = None

    # ------------------------------------------------------------------ reference API surface
    @property
    def device(self) -> torch.device:
        """base_vlm.py:46-48: the device of the first parameter."""
        return next(self.parameters()).device

    def get_prompt_builder(self, system_prompt: Optional[str] = None):
        """prismatic.py:411-413."""
        return self.llm_backbone.prompt_builder_fn(self.model_family, system_prompt=system_prompt)

    def get_vision_tower_2d(self):
        return self.vision_tower_2d

    def encode_images(self, images):
        return self.vision_tower_2d(images, self.projector_2d)

    def initialize_weights(self):
        """prismatic.py:299-321 — xavier-uniform for every nn.Linear (the LLM included, as `self.apply` does),
        unit LayerNorm, N(0, 0.02) embedders, zero-initialised final projection."""
        from .llama import _Proj

        def _basic_init(m):
            if isinstance(m, nn.Linear):
                torch.nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, _Proj):
                # the decoder's q/k/v/o/gate/up/down are nn.Linear in the reference, so `self.apply` re-initialises them
                # too; here they are bias-free parameter holders with nn.Linear's [out, in] layout
                torch.nn.init.xavier_uniform_(m.weight)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.weight, 1.0)
                nn.init.constant_(m.bias, 0)
        self.apply(_basic_init)
        if self.use_diff:
            nn.init.normal_(self.x_embedder.mlp.fc1.weight, std=0.02)
            nn.init.normal_(self.x_embedder.mlp.fc2.weight, std=0.02)
            nn.init.normal_(self.proprio_embedder.mlp.fc1.weight, std=0.02)
            nn.init.normal_(self.proprio_embedder.mlp.fc2.weight, std=0.02)
            nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
            nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
            nn.init.constant_(self.final_layer.mlp.fc2.weight, 0)
            nn.init.constant_(self.final_layer.mlp.fc2.bias, 0)

    def freeze_backbones(self, stage: str) -> None:
        """prismatic.py:415-536."""
        if stage not in {"pretrain", "finetune", "post-training"}:
            raise ValueError(f"Stage `{stage}` is not supported! Try < pretrain | finetune | post-training >")
        train_towers = stage == "pretrain"
        self.vision_tower_2d.requires_grad_(train_towers)
        self.llm_backbone.requires_grad_(True)
        self.projector_2d.requires_grad_(True)
        if self.use_pointcloud:
            self.vision_tower_3d.requires_grad_(train_towers)
            self.projector_3d.requires_grad_(True)
        if self.use_tactile:
            self.tactile_embedder.requires_grad_(True)
        if stage == "post-training":
            self.generation_manager.requires_grad_(True)       # prismatic.py:501 (requires use_generation, as there)
        if stage == "finetune":
            keys = ["llm_backbone", "projector_2d", "proprio_embedder"]
            if self.use_diff:
                keys += ["x_embedder", "t_embedder", "final_layer"]
            if self.use_pointcloud:
                keys += ["projector_3d"]
            if self.use_tactile:
                keys += ["tactile_embedder"]
        else:
            keys = ["vision_tower_2d", "projector_2d", "llm_backbone", "proprio_embedder"]
            if self.use_diff:
                keys += ["x_embedder", "t_embedder", "final_layer"]
            if self.use_pointcloud:
                keys += ["vision_tower_3d", "projector_3d"]
            if self.use_tactile:
                keys += ["tactile_embedder"]
            if stage == "post-training":
                keys += ["generation_manager"]
        self.trainable_module_keys = keys
        self.vision_backbone_requires_grad = train_towers

    def get_fsdp_wrapping_policy(self) -> Callable:
        from torch.distributed.fsdp.wrap import _module_wrap_policy, _or_policy
        from .pointcloud import PointTokenizer
        return partial(_or_policy, policies=[
            partial(_module_wrap_policy, module_classes={PointTokenizer, VisionTokenizer}),
            self.llm_backbone.get_fsdp_wrapping_policy(),
            partial(_module_wrap_policy, module_classes={MLP_GELU, MultimodalGenerationManager}),
        ])

    # ------------------------------------------------------------------ fusion
    def _image_tokens(self, px: torch.Tensor, repeat: int) -> torch.Tensor:
        """[B,4,H,W] -> bf16 [B*repeat, n_tok, token]; the frozen tokenizer runs once per distinct image."""
        pooled, h, w = self.vision_tower_2d.pooled_features(px)
        tok = self.projector_2d(pooled).view(px.shape[0], h * w, -1)
        if repeat > 1:
            tok = tok.repeat(repeat, 1, 1)
        return tok

    def get_fused_tokens(self, images, pointcloud, tactile, gripper_xyz, camera_name, image_repeat: int = 1):
        """prismatic.py:598-769.  Returns (fused [B, F, token], patch_indices [B,N,2] i64, valid_mask [B,N] bool,
        positive_pc_indices_for_tac, linear_positive_img_indices_for_tac, pointcloud_centers)."""
        views = images if isinstance(images, dict) else {"front_image": images}
        assert "front_image" in views, "front_image must be present in multi-view images"
        dev = self.llm_backbone.llm.lm_head.weight.device
        self._front_px = views["front_image"].to(dev, non_blocking=True)     # kept for the image generation head
        front = self._image_tokens(self._front_px, image_repeat)
        if self._front_px.dtype == torch.uint8:
            # raw frames were tokenised through the fused preprocess+im2col; the generation head's pixel losses need
            # the reference's f32 tensor, everything else does not
            if self.use_generation and self.gen_image and self.training:
                from .preprocess import clip_preprocess
                self._front_px = clip_preprocess(self._front_px, self.vision_tower_2d.image_size)
            else:
                self._front_px = None
        B, n_img, _ = front.shape
        centers = None
        if self.use_pointcloud and pointcloud is not None:
            pc_emb, centers = self.vision_tower_3d(pointcloud.to(dev, non_blocking=True))
            pc_tok = self.projector_3d(pc_emb.reshape(-1, pc_emb.shape[-1])).view(B, -1, self.token_size)
            patch_indices, valid_mask = project_points(centers, camera_name)
        else:
            pc_tok = torch.zeros((B, n_img, self.token_size), dtype=front.dtype, device=dev)
            patch_indices = torch.zeros((B, n_img, 2), dtype=torch.long, device=dev)
            valid_mask = torch.zeros((B, n_img), dtype=torch.bool, device=dev)
        assert pc_tok.shape[1] == front.shape[1], f"Token count mismatch: PC={pc_tok.shape[1]}, Front Img={front.shape[1]}"
        parts = [pc_tok, front]
        for key in views:
            if key != "front_image":
                parts.append(self._image_tokens(views[key].to(dev, non_blocking=True), image_repeat))
        pos_pc = lin_img = None
        if self.use_tactile and tactile is not None:
            last_dim = gripper_xyz.shape[-1]
            if last_dim % 3 != 0:
                raise ValueError(f"gripper_xyz last dimension ({last_dim}) is not divisible by 3")
            n_arms = last_dim // 3
            t_flat = tactile.to(dev).view(tactile.shape[0], -1)
            if t_flat.shape[-1] != self.tactile_dim * n_arms:
                raise ValueError(f"Unexpected tactile shape {tuple(tactile.shape)}.")
            tac = self.tactile_embedder(t_flat.reshape(B * n_arms, self.tactile_dim)).view(B, n_arms, -1)
            parts.append(tac)
            # nearest point-cloud centre to each gripper and its image patch (:742-749)
            g = gripper_xyz.to(dev).view(B, n_arms, 3).float().contiguous()
            patch_w = int(front.shape[1] ** 0.5)
            pos_pc = torch.empty((B, n_arms, 1), dtype=torch.long, device=dev)
            lin_img = torch.empty((B, n_arms, 1), dtype=torch.long, device=dev)
            _lib.check(_lib.lib().mla_nearest_center(
                ops._p(g), ops._p(centers.contiguous()), ops._p(patch_indices.contiguous()), C.c_int32(B),
                C.c_int32(n_arms), C.c_int32(centers.shape[1]), C.c_int32(patch_w), ops._p(pos_pc), ops._p(lin_img),
                ops._stream()))
        else:
            parts.append(torch.zeros((B, 1, self.token_size), dtype=front.dtype, device=dev))
        fused = torch.cat(parts, dim=1)
        return fused, patch_indices, valid_mask, pos_pc, lin_img, centers

    def _fused_sequence(self, x, t, proprio, input_ids, attention_mask, labels, images, point_cloud, tactile,
                        gripper_xyz, camera_name, image_repeat: int = 1):
        """Tokenizers, embedders and the splice of prismatic.py:926-1042: returns the decoder input rows and the index
        tensors around them (shared by the training forward and the inference prefill)."""
        dev = self.llm_backbone.llm.lm_head.weight.device
        h = self.token_size
        eos_tag = 2 if self.training else 29871            # tag_0 (:882-887)

        fused, patch_indices, valid_mask, pos_pc, lin_img, _ = self.get_fused_tokens(
            images, point_cloud, tactile, gripper_xyz, camera_name, image_repeat)
        B, F, _ = fused.shape
        N_pc = N_img = patch_indices.shape[1]     # 256 in the reference (prismatic.py:932-933)
        input_ids = input_ids.to(dev, non_blocking=True)
        Lt = input_ids.shape[1]
        text = self.llm_backbone.embed_input_ids(input_ids)                                   # [B, Lt, h]

        n_ins = n_x = 0
        ins_parts = []
        if self.use_diff:
            pr = self.proprio_embedder(proprio.to(dev).to(torch.bfloat16))                     # [B, 1, h]
            if self.training and self.z_embedder.dropout_prob > 0:
                # LabelEmbedder.token_drop (models/diffusion/models.py:82-92) zeroes the whole condition sequence
                # z = [BOS | fused | text] of a dropped sample: apply one draw to both row groups
                drop = (torch.rand(B, device=dev) < self.z_embedder.dropout_prob).view(B, 1, 1)
                fused = torch.where(drop, torch.zeros_like(fused), fused)
                text = torch.where(drop, torch.zeros_like(text), text)
            xe = self.x_embedder(x.to(dev).to(torch.bfloat16))                                # [B, T+1, h]
            if t is not None:
                te = self.t_embedder(t.to(dev)).unsqueeze(1)
            else:
                te = torch.zeros_like(pr)
            ins_parts = [pr, te, xe]
            n_x = xe.shape[1]
            n_ins = pr.shape[1] + 1 + n_x
        # row table: text | fused | inserted
        table = torch.cat([text.reshape(B * Lt, h), fused.reshape(B * F, h)] +
                          ([torch.cat(ins_parts, dim=1).reshape(B * n_ins, h)] if n_ins else []), dim=0)
        S = F + Lt + n_ins
        src_idx = torch.empty((B, S), dtype=torch.int32, device=dev)
        mask = torch.empty((B, S), dtype=torch.uint8, device=dev)
        lti = torch.empty(B, dtype=torch.int32, device=dev)
        head_rows = torch.empty((B, max(n_x, 1)), dtype=torch.int32, device=dev)
        fused_labels = torch.empty((B, S), dtype=torch.int64, device=dev) if labels is not None else None
        if self._err_flag is None or self._err_flag.device != dev:
            self._err_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        am = None
        if attention_mask is not None:
            am = attention_mask.to(dev, non_blocking=True)
            am = (am if am.dtype in (torch.bool, torch.uint8) else am != 0).contiguous()
        lab = labels.to(dev, non_blocking=True).contiguous() if labels is not None else None
        check(_lib.lib().mla_splice_index(
            ops._p(input_ids.contiguous()), ops._p(am), ops._p(lab), C.c_int32(B), C.c_int32(Lt), C.c_int32(F),
            C.c_int32(n_ins), C.c_int32(n_x), C.c_int64(eos_tag), C.c_int32(0), C.c_int32(B * Lt),
            C.c_int32(B * Lt + B * F), ops._p(src_idx), ops._p(mask), ops._p(fused_labels), ops._p(lti),
            ops._p(head_rows), ops._p(self._err_flag), ops._stream()))
        embeds = ops.GatherRowsFn.apply(table, src_idx.view(-1))                              # [B*S, h]

        return SimpleNamespace(embeds=embeds, mask=mask, fused_labels=fused_labels, lti=lti, head_rows=head_rows, B=B, S=S,
                               F=F, h=h, n_x=n_x, n_ins=n_ins, fused=fused, patch_indices=patch_indices,
                               valid_mask=valid_mask, pos_pc=pos_pc, lin_img=lin_img, N_pc=N_pc, N_img=N_img, dev=dev)

    def _fused_sequence_shared(self, x, t, proprio, input_ids, attention_mask, images, camera_name, repeats: int):
        """Shared-prefix packing: the R = `repeats` diffusion copies of a sample (model_mla.py:147-176)
        differ only in their [t | x] rows, so sample b becomes ONE sequence [prefix | R x (t_e | x_e.. | EOS)] and the
        decoder runs B * S' rows instead of B * R * S.  x, t arrive per copy (B*R rows, copy e = r*B + b), everything
        else per sample.  Exact for image-only training (no per-copy randomness in front of the suffix)."""
        dev = self.llm_backbone.llm.lm_head.weight.device
        h = self.token_size
        fused, patch_indices, valid_mask, _, _, _ = self.get_fused_tokens(images, None, None, None, camera_name, 1)
        B, F, _ = fused.shape
        input_ids = input_ids.to(dev, non_blocking=True)
        Lt = input_ids.shape[1]
        R = repeats
        text = self.llm_backbone.embed_input_ids(input_ids)                                   # [B, Lt, h]
        pr = self.proprio_embedder(proprio.to(dev).to(torch.bfloat16))                         # [B, 1, h]
        xe = self.x_embedder(x.to(dev).to(torch.bfloat16))                                    # [B*R, n_x, h]
        te = self.t_embedder(t.to(dev)).unsqueeze(1)                                           # [B*R, 1, h]
        n_x = xe.shape[1]
        if pr.shape[1] != 1 or xe.shape[0] != B * R:
            raise ValueError("shared-prefix packing expects one proprio token per sample and B*R noisy-action rows")
        table = torch.cat([text.reshape(B * Lt, h), fused.reshape(B * F, h), pr.reshape(B, h), te.reshape(B * R, h),
                           xe.reshape(B * R * n_x, h)], dim=0)
        Sp = F + Lt + R * (n_x + 2)
        src_idx = torch.empty((B, Sp), dtype=torch.int32, device=dev)
        mask = torch.empty((B, Sp), dtype=torch.uint8, device=dev)
        rope_pos = torch.empty((B, Sp), dtype=torch.int32, device=dev)
        prefix_len = torch.empty(B, dtype=torch.int32, device=dev)
        lti = torch.empty(B * R, dtype=torch.int32, device=dev)
        head_rows = torch.empty((B * R, n_x), dtype=torch.int32, device=dev)
        if self._err_flag is None or self._err_flag.device != dev:
            self._err_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        am = None
        if attention_mask is not None:
            am = attention_mask.to(dev, non_blocking=True)
            am = (am if am.dtype in (torch.bool, torch.uint8) else am != 0).contiguous()
        eos_tag = 2 if self.training else 29871
        check(_lib.lib().mla_splice_index_shared(
            ops._p(input_ids.contiguous()), ops._p(am), C.c_int32(B), C.c_int32(Lt), C.c_int32(F), C.c_int32(n_x),
            C.c_int32(R), C.c_int64(eos_tag), C.c_int32(0), C.c_int32(B * Lt), C.c_int32(B * Lt + B * F),
            C.c_int32(B * Lt + B * F + B), C.c_int32(B * Lt + B * F + B + B * R), ops._p(src_idx), ops._p(mask),
            ops._p(rope_pos), ops._p(prefix_len), ops._p(lti), ops._p(head_rows), ops._p(self._err_flag), ops._stream()))
        # the EOS embedding row feeds every group (and the filler rows all point at one row): the gather is not
        # injective here, its backward accumulates (fp32) instead of permuting
        from .contrastive import _GatherDupFn
        embeds = _GatherDupFn.apply(table, src_idx.view(-1))                                  # [B*S', h]
        return SimpleNamespace(embeds=embeds, mask=mask, lti=lti, head_rows=head_rows, B=B, S=Sp, F=F, h=h, n_x=n_x,
                               prefix_len=prefix_len, rope_pos=rope_pos.view(-1), group=n_x + 2, fused=fused, dev=dev)

    def forward_shared_prefix(self, x, t, proprio, input_ids, attention_mask, images, camera_name, repeats: int):
        """Training forward of the diffusion head with the R copies of every sample packed behind one shared prefix
        (see _fused_sequence_shared).  Same returns as forward() in train + diff mode; `output.hidden_states` are
        [B, S', h] (packed rows), `noise_pred` is [B*R, T+1, action_dim] in the copy order of MLA.forward."""
        if not (self.use_diff and self.training):
            raise RuntimeError("forward_shared_prefix is the training path of the diffusion head")
        if self.use_pointcloud or self.use_tactile or self.use_contrastive or self.use_generation:
            raise NotImplementedError(
                "shared-prefix training is exact only without per-copy randomness in front of the suffix: the point "
                "tokenizer draws a fresh FPS start per copy (Point_PN.py:10), and the contrastive / generation heads "
                "read per-copy rows")
        if self.z_embedder.dropout_prob > 0:
            raise NotImplementedError("class-dropout draws one mask per copy: the prefix is not shared")
        if self.llm_backbone.llm.compute_lm_loss:
            raise NotImplementedError("the vocabulary CE over packed rows is not defined; it is discarded in diffusion mode")
        q = self._fused_sequence_shared(x, t, proprio, input_ids, attention_mask, images, camera_name, repeats)
        output: CausalLMOutputWithPast = self.llm_backbone(
            input_ids=None, attention_mask=q.mask, position_ids=None, past_key_values=None,
            inputs_embeds=q.embeds.view(q.B, q.S, q.h), labels=None, use_cache=None, output_attentions=None,
            output_hidden_states=True, return_dict=True, prefix_len=q.prefix_len, suffix_group=q.group,
            rope_pos=q.rope_pos)
        output.last_true_indices = q.lti
        self._front_px = None
        last = output.hidden_states[-1].reshape(q.B * q.S, q.h)
        rows = ops.GatherRowsFn.apply(last, q.head_rows.view(-1))                             # [B*R*(T+1), h]
        noise_pred = self.final_layer(rows).reshape(q.B * repeats, q.n_x, self.action_dim)
        return output, noise_pred, {}, {}

    # ------------------------------------------------------------------ forward
    def forward(self, x=None, t=None, z=None, proprio=None, gripper_xyz=None, input_ids=None, attention_mask=None,
                images=None, camera_name=None, point_cloud=None, tactile=None, labels=None, inputs_embeds=None,
                past_key_values=None, use_cache=None, output_attentions=None, output_hidden_states=True,
                return_dict=None, multimodal_indices=None, gen_discret_action=None, use_diff=None, next_images=None,
                next_point_cloud=None, next_tactile=None, image_repeat: int = 1, **kwargs):
        if use_diff is not None:
            self.use_diff = use_diff
        if past_key_values is not None or (input_ids is not None and input_ids.shape[1] == 1):
            raise NotImplementedError("cached single-token decoding is inference-only (out of the hot-path scope)")
        if images is None:
            raise RuntimeError("Invalid `forward()` call!")
        if multimodal_indices is not None and len(multimodal_indices) != len(input_ids):
            raise NotImplementedError("unimodal/mixed batches (multimodal_indices) are not part of the VLA training path")
        q = self._fused_sequence(x, t, proprio, input_ids, attention_mask, labels, images, point_cloud, tactile,
                                 gripper_xyz, camera_name, image_repeat)
        embeds, mask, fused_labels, lti, head_rows = q.embeds, q.mask, q.fused_labels, q.lti, q.head_rows
        B, S, h, n_x, dev, fused = q.B, q.S, q.h, q.n_x, q.dev, q.fused
        patch_indices, valid_mask, pos_pc, lin_img, N_pc, N_img = (q.patch_indices, q.valid_mask, q.pos_pc, q.lin_img,
                                                                   q.N_pc, q.N_img)
        pc_idx = (1, 1 + N_pc)
        img_idx = (1 + N_pc, 1 + N_pc + N_img)
        tac_idx = (img_idx[1], img_idx[1] + self.action_dim // 7) if self.use_tactile else None
        output: CausalLMOutputWithPast = self.llm_backbone(
            input_ids=None, attention_mask=mask if attention_mask is not None else None, position_ids=None,
            past_key_values=None, inputs_embeds=embeds.view(B, S, h), labels=fused_labels, use_cache=use_cache,
            output_attentions=output_attentions, output_hidden_states=True, return_dict=True,
            pc_token_indices=pc_idx, img_token_indices=img_idx, tac_token_indices=tac_idx,
            patch_correspondence_indices=patch_indices, correspondence_valid_mask=valid_mask,
            positive_pc_indices_for_tac=pos_pc, linear_positive_img_indices_for_tac=lin_img,
            compute_token_contrastive_loss=self.use_contrastive,
            compute_tactile_contrastive_loss=(self.use_contrastive and self.use_tactile))
        output.last_true_indices = lti
        generation_outputs: Dict[str, torch.Tensor] = {}
        generation_losses: Dict[str, torch.Tensor] = {}
        if (self.use_generation and (self.gen_image or self.gen_pointcloud or self.gen_tactile) and self.training):
            # prismatic.py:1075-1113.  next_* arrive un-repeated (sample b of the B_eff = B*R rows uses entry b % B,
            # which is what MLA.forward's .repeat(R, ...) of the reference produces)
            hidden16 = output.hidden_states[-1].reshape(B * S, h)
            img_feat = cur_px = nxt_px = roi = None
            if self.gen_image:
                assert next_images is not None
                img_feat = fused[:, N_pc:N_pc + N_img, :].reshape(B * N_img, h)
                cur_px = self._front_px.float()
                nxt_px = next_images.to(dev, non_blocking=True).float()
                if self.use_roi:
                    roi = roi_mask(patch_indices, 16, self.roi_dilation_kernel_size)
            if self.gen_pointcloud:
                assert next_point_cloud is not None
            if self.gen_tactile:
                assert next_tactile is not None
            generation_outputs = self.generation_manager.run(hidden16, B, S, img_feat, cur_px, nxt_px, roi)
            generation_losses = compute_generation_losses(
                generation_outputs, self.gen_image, self.gen_pointcloud, self.gen_tactile,
                next_point_cloud.to(dev, non_blocking=True) if next_point_cloud is not None else None,
                next_tactile.to(dev, non_blocking=True) if next_tactile is not None else None)
        self._front_px = None
        if self.use_diff:
            last = output.hidden_states[-1].reshape(B * S, h)
            rows = ops.GatherRowsFn.apply(last, head_rows.view(-1))                           # [B*(T+1), h]
            noise_pred = self.final_layer(rows).reshape(B, n_x, self.action_dim)
            if self.training:
                return output, noise_pred, generation_outputs, generation_losses
            return output, noise_pred
        if self.training:
            return output, generation_outputs, generation_losses
        return output

    # ------------------------------------------------------------------ inference: KV-cached denoise loop
    @torch.no_grad()
    def denoise_prefill(self, input_ids, images, point_cloud=None, proprio=None, camera_name=None, tactile=None,
                        gripper_xyz=None, n_x: Optional[int] = None, embeds_only: bool = False) -> SimpleNamespace:
        """Everything of the eval-mode forward that does not depend on the DDIM step: tokenizers, projectors, the
        splice, and the decoder over the prefix [BOS | fused | text.. | proprio] (prismatic.py:983-992 puts
        [proprio | t | x..] in front of the last tag token 29871).  Under the causal mask the rows behind x (the tag
        token itself) cannot influence the noise prediction read at the x rows (:1121-1124), so they are dropped.
        Returns the per-layer K/V caches and the sizes `denoise_step` needs."""
        if self.training:
            raise RuntimeError("denoise_prefill is the inference path: call .eval() first (predict_action_diff does)")
        if not self.use_diff:
            raise RuntimeError("the denoise loop needs the diffusion head (use_diff=True)")
        n_x = n_x if n_x is not None else self.future_action_window_size + 1
        dev = self.llm_backbone.llm.lm_head.weight.device
        B = input_ids.shape[0]
        x0 = torch.zeros((B, n_x, self.action_dim), dtype=torch.float32, device=dev)
        t0 = torch.zeros((B,), dtype=torch.long, device=dev)
        q = self._fused_sequence(x0, t0, proprio, input_ids, None, None, images, point_cloud, tactile, gripper_xyz,
                                 camera_name)
        self._front_px = None
        self.check_errors()
        lti = q.lti.tolist()                     # one host read per action (the per-step loop has none)
        if len(set(lti)) != 1:
            raise NotImplementedError("KV-cached denoising needs the tag token at the same position in every sample "
                                      "of the batch (use use_kv_cache=False for ragged prompts)")
        P = lti[0] + 1                           # prefix rows: everything up to and including the proprio token
        prefix = q.embeds.view(B, q.S, q.h)[:, :P].reshape(B * P, q.h).contiguous()
        if embeds_only:
            return SimpleNamespace(prefix=prefix, B=B, P=P, n_x=n_x, h=q.h, dev=dev)
        caches = self.llm_backbone.llm.model.prefill(prefix, B, P, 1 + n_x)
        return SimpleNamespace(caches=caches, B=B, P=P, n_x=n_x, h=q.h, dev=dev)

    @torch.no_grad()
    def denoise_session(self, B: int, P: int, n_x: int, sampler) -> SimpleNamespace:
        """Static buffers + two CUDA graphs for one (batch, prefix length, action rows, DDIM schedule): `g_prefill`
        (prefix embeddings -> per-layer K/V caches) and `g_loop` (the whole DDIM loop: noise -> sample).  A denoise
        step is ~320 launches of a few microseconds each over 2-34 rows; issued from Python one by one it is
        host-bound, replayed as a graph it runs at the weight-streaming rate.  Built on first use, replayed afterwards
        (weights are read through the layers' persistent bf16 copies, refreshed in place when the masters change)."""
        key = (B, P, n_x, tuple(sampler.timestep_map))
        sessions = self.__dict__.setdefault("_denoise_sessions", {})
        model = self.llm_backbone.llm.model
        dev = self.llm_backbone.llm.lm_head.weight.device
        fp = sum(p._version for p in self.parameters())
        sess = sessions.get(key)
        if sess is not None and sess.fp == fp:
            return sess
        h = self.token_size
        if sess is None:
            sess = SimpleNamespace(prefix=torch.zeros((B * P, h), dtype=torch.bfloat16, device=dev),
                                   noise=torch.zeros((B, n_x, self.action_dim), dtype=torch.float32, device=dev),
                                   caches=[torch.empty((B, 2, model.heads, P, h // model.heads), dtype=torch.bfloat16,
                                                       device=dev) for _ in model.layers],
                                   g_prefill=None, g_loop=None, out=None, fp=None)
        st = SimpleNamespace(caches=sess.caches, B=B, P=P, n_x=n_x, h=h, dev=dev)

        def run_prefill():
            model.prefill(sess.prefix, B, P, 1 + n_x, caches=sess.caches)

        def run_loop():
            return sampler.ddim_sample_loop(lambda x, t: self.denoise_step(st, x, t), sess.noise.shape, sess.noise,
                                            clip_denoised=False, eta=0.0)

        # eager pass on a side stream: lazy initialisation (bf16 weight copies, RoPE / timestep tables, kernel
        # attributes) happens here, never inside a capture; it also refreshes the copies after a weight update
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            run_prefill()
            run_loop()
        cur.wait_stream(side)
        if sess.g_prefill is None:
            sess.g_prefill = torch.cuda.CUDAGraph()
            with torch.cuda.graph(sess.g_prefill):
                run_prefill()
            sess.g_loop = torch.cuda.CUDAGraph()
            with torch.cuda.graph(sess.g_loop, pool=sess.g_prefill.pool()):
                sess.out = run_loop()
        sess.fp = fp
        sessions[key] = sess
        return sess

    @torch.no_grad()
    def denoise_step(self, st: SimpleNamespace, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """Noise prediction for x_t [B, n_x, action_dim] at timestep t [B] from the cached prefix: only the
        [t | x_0..x_T] rows run through the decoder."""
        dev, B, n = st.dev, st.B, 1 + st.n_x
        xe = self.x_embedder(x.to(dev).to(torch.bfloat16))                                    # [B, n_x, h]
        te = self.t_embedder(t.to(dev)).unsqueeze(1)                                          # [B, 1, h]
        rows = torch.cat([te, xe], dim=1).reshape(B * n, st.h).contiguous()
        hid = self.llm_backbone.llm.model.decode(rows, st.caches, B, st.P, n)                 # [B*n, h], final norm
        xrows = hid.view(B, n, st.h)[:, 1:].reshape(B * st.n_x, st.h).contiguous()
        return self.final_layer(xrows).reshape(B, st.n_x, self.action_dim)

    def check_errors(self) -> None:
        """Deferred device-side error check (one sync): raises what the reference would have raised eagerly."""
        if self._err_flag is not None and int(self._err_flag.item()) != 0:
            self._err_flag.zero_()
            raise IndexError("a sample has no EOS/tag token to splice the action tokens before "
                             "(models/vlm/prismatic.py:983 would raise IndexError)")
