"""Builds and calls the UNMODIFIED reference (through oracle/ref_shim.py) at Llama-2-7B width — test / measurement
infrastructure shared by tools/ref_gpu.py (R-GPU: the reference on one B200) and bench.py's reference arm / cpu_baseline
leg (the reference's own CPU path).  None of our modules or kernels are on this path.  Nothing under mla_b200/ imports it.
"""
from __future__ import annotations

import torch

H, F, L, HEADS, VOCAB = 4096, 11008, 32, 32, 32064


def build_reference_7b(workload: str, layers: int = L, param_dtype=torch.bfloat16, device="cuda", attn="flash_attention_2"):
    from oracle import ref_shim
    ns = ref_shim.load()
    import transformers.models.llama.modeling_llama as ML
    use_pc = workload in ("cfg3", "cfg4")
    flags = dict(use_diff=True, use_pointcloud=use_pc, use_tactile=use_pc, use_contrastive=use_pc, use_generation=False)
    torch.manual_seed(0)
    with torch.device(device):
        cfg = ns.LlamaConfig(vocab_size=VOCAB, hidden_size=H, intermediate_size=F, num_hidden_layers=layers,
                             num_attention_heads=HEADS, num_key_value_heads=HEADS, max_position_embeddings=2048,
                             rms_norm_eps=1e-5)
        cfg._attn_implementation = "sdpa"
        vlm = ns.PrismaticVLM("mla-7b-synthetic", ns.TinyBackbone(cfg), token_size=H, action_dim=7, **flags)
        if attn == "flash_attention_2":
            llm = vlm.llm_backbone.llm
            llm.config._attn_implementation = "flash_attention_2"
            for i, layer in enumerate(llm.model.layers):
                new = ML.LlamaFlashAttention2(config=llm.config, layer_idx=i)
                new.load_state_dict(layer.self_attn.state_dict())
                layer.self_attn = new
        mla = ns.MLA(vlm, ns.ActionTokenizer(ns.FakeTok()), token_size=H, action_dim=7, future_action_window_size=0, **flags)
    with torch.no_grad():
        mla.vlm.final_layer.mlp.fc2.weight.normal_(std=0.02)      # zero-initialised head (prismatic.py:320)
    mla.to(param_dtype).to(device).train()
    mla.vlm.freeze_backbones("finetune")
    return mla, ns


def apply_checkpointing(mla, ns):
    """What FSDPStrategy.run_setup does (training/strategies/fsdp.py:217-223), minus the FSDP wrap itself."""
    from functools import partial
    from torch.distributed.algorithms._checkpoint.checkpoint_wrapper import (CheckpointImpl, apply_activation_checkpointing,
                                                                             checkpoint_wrapper)
    wrapper = partial(checkpoint_wrapper, checkpoint_impl=CheckpointImpl.NO_REENTRANT)
    apply_activation_checkpointing(mla, checkpoint_wrapper_fn=wrapper, check_fn=lambda m: isinstance(m, ns.LlamaDecoderLayer))


def ref_call(mla, b, device_type="cuda", repeats=4):
    with torch.autocast(device_type, dtype=torch.bfloat16):
        loss_dict, _ = mla(input_ids=b["input_ids"], attention_mask=b["attention_mask"], labels=b["labels"],
                           actions=b["actions"], images=b["images"], point_cloud=b.get("point_cloud"),
                           tactile=b.get("tactile"), proprio=b["proprio"], gripper_xyz=b.get("gripper_xyz"),
                           action_masks=b["action_masks"], camera_name="rlbench_front",
                           repeated_diffusion_steps=repeats, use_diff=True)
    return loss_dict["total_loss"]
