"""LLM backbone: the modified LlamaForCausalLM of the reference and its thin Prismatic wrapper.

  LlamaForCausalLM   transformers/models/llama/modeling_llama.py:1130-1317 (reference modification: the two
                     contrastive modules live inside the LM and the losses are computed from hidden_states[8])
  LLMBackbone API    models/backbones/llm/base_llm.py:41-241, llama2.py:51-104 (embed_input_ids, forward kwargs,
                     transformer_layer_cls, half_precision_dtype, get_fsdp_wrapping_policy, ...)

HuggingFace hub loading (base_llm.py:122-154) is checkpoint plumbing and out of scope: backbones are built from a
`LlamaConfig` (random init, as BASELINE's synthetic benchmark requires) or from a state_dict with reference keys.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from functools import partial
from typing import Any, Callable, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from .contrastive import CoordinateAwareContrastiveLoss, TactileContrastiveLoss
from .llama import LlamaDecoderLayer, LlamaModel


@dataclass
class LlamaConfig:
    vocab_size: int = 32064          # 32000 + <PAD> padded to a multiple of 64 (llama2.py:75-77)
    hidden_size: int = 4096
    intermediate_size: int = 11008
    num_hidden_layers: int = 32
    num_attention_heads: int = 32
    num_key_value_heads: Optional[int] = None
    rms_norm_eps: float = 1e-5       # meta-llama/Llama-2-7b-hf config.json (HF's class default is 1e-6)
    rope_theta: float = 10000.0
    max_position_embeddings: int = 2048
    initializer_range: float = 0.02
    pad_token_id: Optional[int] = 32000
    bos_token_id: int = 1
    eos_token_id: int = 2
    use_cache: bool = False
    pretraining_tp: int = 1
    output_hidden_states: bool = False
    use_return_dict: bool = True

    def __post_init__(self):
        if self.num_key_value_heads is None:
            self.num_key_value_heads = self.num_attention_heads
        if self.num_key_value_heads != self.num_attention_heads:
            raise NotImplementedError("grouped-query attention is not on the MLA hot path (Llama-2-7B has kv=32)")


@dataclass
class CausalLMOutputWithPast:
    """transformers/modeling_outputs.py:706-713 (the reference's three extra fields included)."""
    loss: Optional[torch.Tensor] = None
    img_pc_contrastive_loss: Optional[torch.Tensor] = None
    tactile_contrastive_loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    all_logits_for_action: Optional[torch.Tensor] = None
    past_key_values: Any = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    attentions: Any = None


class LlamaForCausalLM(nn.Module):
    def __init__(self, config: LlamaConfig, use_token_contrastive_loss: bool = True,
                 use_tactile_contrastive_loss: bool = True, contrastive_projection_dim: int = 256):
        super().__init__()
        self.config = config
        self.model = LlamaModel(config.vocab_size, config.hidden_size, config.intermediate_size,
                                config.num_hidden_layers, config.num_attention_heads, config.rms_norm_eps,
                                config.rope_theta, config.pad_token_id)
        self.vocab_size = config.vocab_size
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.use_token_contrastive_loss = use_token_contrastive_loss
        if use_token_contrastive_loss:
            self.coordinate_aware_contrastive_loss_module = CoordinateAwareContrastiveLoss(
                feature_dim=config.hidden_size, projection_dim=contrastive_projection_dim)
        self.use_tactile_contrastive_loss = use_tactile_contrastive_loss
        if use_tactile_contrastive_loss:
            self.tactile_contrastive_loss_module = TactileContrastiveLoss(
                feature_dim=config.hidden_size, projection_dim=contrastive_projection_dim)
        self.compute_lm_loss = False   # CE over the vocabulary is discarded in diffusion mode (model_mla.py:215-217)
        self._init_weights()

    def _init_weights(self):
        """HF `_init_weights` (modeling_llama.py:894-904): N(0, initializer_range) for Linear/Embedding, zero bias,
        zero padding row; norm weights stay 1."""
        std = self.config.initializer_range
        for m in self.modules():
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(mean=0.0, std=std)
                if m.bias is not None:
                    m.bias.data.zero_()
            elif isinstance(m, nn.Embedding):
                m.weight.data.normal_(mean=0.0, std=std)
                if m.padding_idx is not None:
                    m.weight.data[m.padding_idx].zero_()
            elif hasattr(m, "in_features") and hasattr(m, "weight") and m.weight.dim() == 2:
                m.weight.data.normal_(mean=0.0, std=std)

    def get_input_embeddings(self):
        return self.model.embed_tokens

    @property
    def generation_config(self):
        """What VLM.generation_config hands to GenerationMixin (models/vlm/base_vlm.py:49): the special-token ids."""
        from types import SimpleNamespace
        c = self.config
        return SimpleNamespace(bos_token_id=c.bos_token_id, eos_token_id=c.eos_token_id, pad_token_id=c.pad_token_id)

    def resize_token_embeddings(self, new_num_tokens: Optional[int] = None, pad_to_multiple_of: Optional[int] = None):
        """transformers' PreTrainedModel.resize_token_embeddings as scripts/train.py:143-144 and llama2.py:75-77 use it:
        grow (or shrink) embed_tokens and lm_head to `new_num_tokens` rounded up to `pad_to_multiple_of`, keeping the
        existing rows; new rows ~ N(0, initializer_range) (train.py:145-155 then overwrites them with the mean row).
        Updates config.vocab_size and returns the input embedding module."""
        emb, head = self.model.embed_tokens, self.lm_head
        if new_num_tokens is None:
            return emb
        if pad_to_multiple_of is not None:
            if not isinstance(pad_to_multiple_of, int) or pad_to_multiple_of <= 0:
                raise ValueError(f"pad_to_multiple_of must be a positive integer, got {pad_to_multiple_of}")
            new_num_tokens = (new_num_tokens + pad_to_multiple_of - 1) // pad_to_multiple_of * pad_to_multiple_of
        old = emb.weight.shape[0]
        if new_num_tokens == old:
            return emb
        keep = min(old, new_num_tokens)
        std = self.config.initializer_range
        new_emb = nn.Embedding(new_num_tokens, emb.weight.shape[1], emb.padding_idx, device=emb.weight.device,
                               dtype=emb.weight.dtype)
        new_head = nn.Linear(head.in_features, new_num_tokens, bias=False, device=head.weight.device,
                             dtype=head.weight.dtype)
        with torch.no_grad():
            new_emb.weight.normal_(mean=0.0, std=std)
            new_head.weight.normal_(mean=0.0, std=std)
            if new_emb.padding_idx is not None and new_emb.padding_idx < new_num_tokens:
                new_emb.weight[new_emb.padding_idx].zero_()
            new_emb.weight[:keep] = emb.weight[:keep]
            new_head.weight[:keep] = head.weight[:keep]
        new_emb.weight.requires_grad_(emb.weight.requires_grad)
        new_head.weight.requires_grad_(head.weight.requires_grad)
        self.model.embed_tokens, self.lm_head = new_emb, new_head
        self.config.vocab_size = self.vocab_size = new_num_tokens
        return new_emb

    def get_output_embeddings(self):
        return self.lm_head

    def embed(self, input_ids: torch.Tensor) -> torch.Tensor:
        """nn.Embedding lookup -> bf16 [numel, h] (modeling_llama.py:971)."""
        emb = self.model.embed_tokens
        return ops.EmbeddingFn.apply(input_ids, emb.weight, emb.padding_idx)

    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, labels=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                return_dict=None, cache_position=None, pc_token_indices=None, img_token_indices=None,
                tac_token_indices=None, patch_correspondence_indices=None, correspondence_valid_mask=None,
                positive_pc_indices_for_tac=None, linear_positive_img_indices_for_tac=None,
                compute_token_contrastive_loss: bool = False, compute_tactile_contrastive_loss: bool = False,
                prefix_len=None, suffix_group: int = 0, rope_pos=None):
        """prefix_len / suffix_group / rope_pos (ours, not in the reference signature): the shared-prefix layout of
        MLA.forward(share_diffusion_prefix) — see llama.LayerShape."""
        if past_key_values is not None or use_cache:
            raise NotImplementedError("KV-cache decoding is inference-only (out of the hot-path scope)")
        if (input_ids is None) == (inputs_embeds is None):
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time, and must specify either one")
        if inputs_embeds is None:
            B, S = input_ids.shape
            x = self.embed(input_ids)
        else:
            B, S, _ = inputs_embeds.shape
            x = inputs_embeds.reshape(B * S, -1)
            if x.dtype != torch.bfloat16:
                x = x.to(torch.bfloat16)
        mask = None
        if attention_mask is not None:
            mask = attention_mask if attention_mask.dtype in (torch.bool, torch.uint8) else attention_mask != 0
            mask = mask.contiguous()
        hs2d = self.model.run_layers(x.contiguous(), B, S, mask, prefix_len, suffix_group, rope_pos)
        h = self.config.hidden_size
        hidden_states = tuple(t.view(B, S, h) for t in hs2d)

        loss = logits = None
        if self.compute_lm_loss:
            logits_bf16 = ops.linear(hs2d[-1], self.lm_head.weight)                   # [B*S, V] bf16 (nn.Linear output)
            if labels is not None:
                loss = ops.CrossEntropyFn.apply(logits_bf16, labels.to(logits_bf16.device))
            logits = logits_bf16.view(B, S, -1).float()                               # `logits.float()` (:1256)

        img_pc_loss = None
        if self.training and compute_token_contrastive_loss:
            hs8 = hidden_states[8]
            img_pc_loss = self.coordinate_aware_contrastive_loss_module(
                image_features=hs8[:, img_token_indices[0]:img_token_indices[1], :],
                pointcloud_features=hs8[:, pc_token_indices[0]:pc_token_indices[1], :],
                patch_indices=patch_correspondence_indices, valid_mask=correspondence_valid_mask)
            loss = img_pc_loss if loss is None else loss + img_pc_loss
        tac_loss = None
        if self.training and compute_tactile_contrastive_loss:
            hs8 = hidden_states[8]
            tac_loss = self.tactile_contrastive_loss_module(
                tac_features=hs8[:, tac_token_indices[0]:tac_token_indices[1], :],
                pc_features=hs8[:, pc_token_indices[0]:pc_token_indices[1], :],
                img_features=hs8[:, img_token_indices[0]:img_token_indices[1], :],
                positive_pc_indices=positive_pc_indices_for_tac,
                linear_positive_img_indices=linear_positive_img_indices_for_tac)
            loss = tac_loss if loss is None else loss + tac_loss
        return CausalLMOutputWithPast(loss=loss, logits=logits, img_pc_contrastive_loss=img_pc_loss,
                                      tactile_contrastive_loss=tac_loss, hidden_states=hidden_states)


class _OfflineTokenizer:
    """Stand-in for the HF tokenizer attributes PrismaticVLM touches (prismatic.py:208-212, train.py:143-155) when
    no tokenizer files are available (there is no network here)."""
    vocab_size = 32000
    pad_token_id = 32000
    bos_token_id = 1
    eos_token_id = 2
    padding_side = "right"
    model_max_length = 2048

    def __len__(self):
        return 32001

    def encode(self, s: str, add_special_tokens: bool = False) -> List[int]:
        return [sum(map(ord, s)) % 20000 + 100]


class PurePromptBuilder:
    """`In: <message>\nOut: <reply></s>` turns (models/backbones/llm/prompting/base_prompter.py:27-77): the prompt of
    predict_action_diff is one human turn."""

    def __init__(self, model_family: str, system_prompt: Optional[str] = None) -> None:
        self.model_family, self.system_prompt = model_family, system_prompt
        self.bos, self.eos = "<s>", "</s>"
        self.prompt, self.turn_count = "", 0

    def add_turn(self, role: str, message: str) -> str:
        human = self.turn_count % 2 == 0
        assert role == ("human" if human else "gpt")
        message = message.replace("<image>", "").strip()
        wrapped = f"In: {message}\nOut: " if human else f"{message}{self.eos}"
        self.prompt += wrapped
        self.turn_count += 1
        return wrapped

    def get_potential_prompt(self, message: str) -> str:
        return (self.prompt + f"In: {message}\nOut: ").removeprefix(self.bos).rstrip()

    def get_prompt(self) -> str:
        return self.prompt.removeprefix(self.bos).rstrip()


class LLMBackbone(nn.Module):
    """models/backbones/llm/base_llm.py LLMBackbone + HFCausalLLMBackbone surface on our LlamaForCausalLM."""

    def __init__(self, llm_backbone_id: str = "llama2-7b-pure", config: Optional[LlamaConfig] = None,
                 tokenizer=None, llm_max_length: int = 2048, inference_mode: bool = False,
                 use_flash_attention_2: bool = True, llm_vision_layers: int = 1, **_unused):
        super().__init__()
        self.identifier = llm_backbone_id
        self.llm_family = "llama2"
        self.llm_max_length = llm_max_length
        self.inference_mode = inference_mode
        self.llm = LlamaForCausalLM(config or LlamaConfig())
        self.llm.config.use_cache = False if not inference_mode else True
        self.tokenizer = tokenizer if tokenizer is not None else _OfflineTokenizer()

    def get_tokenizer(self):
        return self.tokenizer

    def get_fsdp_wrapping_policy(self) -> Callable:
        from torch.distributed.fsdp.wrap import transformer_auto_wrap_policy
        return partial(transformer_auto_wrap_policy, transformer_layer_cls={self.transformer_layer_cls})

    def enable_gradient_checkpointing(self) -> None:
        """The reference turns on HF gradient checkpointing; here it selects full-layer recompute (save only the
        layer input), the same memory/FLOP trade."""
        self.llm.model.set_save_levels("layer")

    def embed_input_ids(self, input_ids: torch.LongTensor) -> torch.Tensor:
        B, L = input_ids.shape
        return self.llm.embed(input_ids).view(B, L, -1)

    def forward(self, **kwargs) -> CausalLMOutputWithPast:
        return self.llm(**kwargs)

    @property
    def prompt_builder_fn(self):
        """llama2-*-pure (the backbone of every MLA recipe) uses the "pure" template (llama2.py:81-83); the chat /
        vicuna templates belong to other LLM families (out of scope)."""
        return PurePromptBuilder

    @property
    def transformer_layer_cls(self):
        return LlamaDecoderLayer

    @property
    def half_precision_dtype(self) -> torch.dtype:
        return torch.bfloat16

    @property
    def last_layer_finetune_modules(self):
        return (self.llm.model.embed_tokens, self.llm.model.layers[-1], self.llm.lm_head)

    @property
    def embed_dim(self) -> int:
        return self.llm.config.hidden_size

    @property
    def pad_token_id(self) -> int:
        return self.tokenizer.pad_token_id


LLaMa2LLMBackbone = LLMBackbone
