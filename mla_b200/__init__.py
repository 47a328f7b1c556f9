"""mla_b200 — B200-native (sm_100a) implementation of the MLA training-step hot path behind the reference's module API.

Public surface mirrors the reference (models/mla, models/vlm, models/backbones/llm): MLA, PrismaticVLM,
LLaMa2LLMBackbone / LLMBackbone, LlamaConfig, LlamaForCausalLM, plus DataParallelTrainer; `create_diffusion` (training
schedule and the DDIM sampler of inference) and `clip_preprocess` (the reference's CPU-side CLIPImageProcessor on the GPU).
The CUDA library is loaded lazily by `mla_b200._lib.lib()`; importing the package never needs a GPU, computing always does.
"""
from .backbone import CausalLMOutputWithPast, LLaMa2LLMBackbone, LLMBackbone, LlamaConfig, LlamaForCausalLM  # noqa: F401
from .mla import MLA  # noqa: F401
from .modules import create_diffusion  # noqa: F401
from .preprocess import clip_preprocess  # noqa: F401
from .trainer import DataParallelTrainer, plan_save_levels  # noqa: F401
from .vlm import PrismaticVLM  # noqa: F401

__all__ = ["MLA", "PrismaticVLM", "LLMBackbone", "LLaMa2LLMBackbone", "LlamaConfig", "LlamaForCausalLM",
           "CausalLMOutputWithPast", "DataParallelTrainer", "plan_save_levels", "create_diffusion", "clip_preprocess"]
