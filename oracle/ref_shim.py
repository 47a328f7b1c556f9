"""Import the UNMODIFIED reference (ZhuoyangLiu2005/MLA at /root/reference) offline — test infrastructure only.

Used by tests/golden/make_golden.py (to produce the committed golden vectors) and by the optional CPU tests that
compare the oracle with the live reference when /root/reference exists (it does not on the GPU box).  Nothing under
mla_b200/ imports this.

What it works around: the vendored transformers 4.40.1 refuses the installed tokenizers
version; timm / draccus / tensorflow / accelerate / matplotlib ... are not installed; `models/__init__` and
`vla/__init__` pull in TensorFlow; HF-hub loading needs a network.  Stand-ins are provided ONLY for code that is
not under /root/reference (timm's Mlp / RmsNorm): their arithmetic is restated from timm's published definition and
is therefore unpinned.
"""
from __future__ import annotations

import importlib.machinery as _im
import importlib.metadata as _md
import os
import sys
import types

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# /root/reference in the build container; on the GPU box (where that path does not exist) the git-ignored install of
# the same unmodified sources under baseline/_ref (tools/install_reference.py)
REF_ROOT = os.environ.get("MLA_REFERENCE_ROOT") or next(
    (p for p in ("/root/reference", os.path.join(_HERE, "baseline", "_ref")) if os.path.isdir(os.path.join(p, "models", "mla"))),
    "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models", "mla"))


_loaded = None


def load(rmsnorm_variance_mode: bool = False):
    """Returns a namespace with the reference classes (MLA, PrismaticVLM, LlamaForCausalLM, LlamaConfig, ...)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch
    import torch.nn as nn

    pins = {"tokenizers": "0.19.1", "huggingface-hub": "0.23.0", "huggingface_hub": "0.23.0",
            "safetensors": "0.4.3", "numpy": "1.26.4", "accelerate": "0.25.0"}
    orig_version = _md.version
    _md.version = lambda name: pins.get(name, orig_version(name))

    class _AnyMeta(type):
        def __getattr__(cls, n):
            if n.startswith("__"):
                raise AttributeError(n)
            return _Any

    class _Any(metaclass=_AnyMeta):
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Any()

        def __getattr__(self, n):
            if n.startswith("__"):
                raise AttributeError(n)
            return _Any()

        def __iter__(self):
            return iter(())

    class _Stub(types.ModuleType):
        def __getattr__(self, n):
            if n.startswith("__"):
                raise AttributeError(n)
            return _Any

    def stub(name):
        parts = name.split(".")
        for i in range(1, len(parts) + 1):
            n = ".".join(parts[:i])
            if n not in sys.modules:
                m = _Stub(n)
                m.__path__ = []
                m.__spec__ = _im.ModuleSpec(n, None)
                sys.modules[n] = m
                if i > 1:
                    setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], m)

    for s in ("torch_geometric.nn.pool torch_scatter matplotlib.pyplot draccus jsonlines easydict ipdb "
              "mpl_toolkits.mplot3d dlimp tensorflow tensorflow_datasets tensorflow_graphics peft accelerate "
              "timm timm.models timm.models.vision_transformer timm.models.layers timm.data").split():
        stub(s)

    # --- stand-ins for timm 0.9.10 (not vendored by the reference; restated from timm's definition)
    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None,
                     bias=True, drop=0., use_conv=False):
            super().__init__()
            out_features = out_features or in_features
            hidden_features = hidden_features or in_features
            self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
            self.act = act_layer()
            self.drop1 = nn.Dropout(drop)
            self.norm = nn.Identity()
            self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
            self.drop2 = nn.Dropout(drop)

        def forward(self, x):
            return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))

    class RmsNorm(nn.Module):
        def __init__(self, channels, eps=1e-6, affine=True):
            super().__init__()
            self.eps = eps
            self.weight = nn.Parameter(torch.ones(channels))

        def forward(self, x):
            xf = x.float()
            v = xf.var(dim=-1, keepdim=True) if rmsnorm_variance_mode else xf.pow(2).mean(-1, keepdim=True)
            return (xf * torch.rsqrt(v + self.eps)).to(x.dtype) * self.weight

    vt = sys.modules["timm.models.vision_transformer"]
    vt.Mlp, vt.RmsNorm, vt.Attention = Mlp, RmsNorm, nn.Identity
    tl = sys.modules["timm.models.layers"]
    tl.DropPath, tl.trunc_normal_ = nn.Identity, nn.init.trunc_normal_

    sys.path.insert(0, REF_ROOT)
    # the pip `transformers` may already be imported by torch/others: make sure the vendored tree wins
    for name in [n for n in sys.modules if n == "transformers" or n.startswith("transformers.")]:
        del sys.modules[name]
    for n, p in (("models", os.path.join(REF_ROOT, "models")), ("vla", os.path.join(REF_ROOT, "vla"))):
        pkg = types.ModuleType(n)
        pkg.__path__ = [p]
        pkg.__spec__ = _im.ModuleSpec(n, None, is_package=True)
        sys.modules[n] = pkg
    from vla.action_tokenizer import ActionTokenizer
    sys.modules["vla"].ActionTokenizer = ActionTokenizer
    import models.mla  # noqa: F401  (must precede anything importing modeling_llama: breaks the import cycle)
    from models.mla.model_mla import MLA
    from models.vlm.prismatic import PrismaticVLM
    from models.backbones.llm.base_llm import LLMBackbone
    from models.backbones.llm.prompting import PurePromptBuilder
    from models.mla.pointcloud.backbone.Point_PN import Point_PN_scan
    from transformers.models.llama.modeling_llama import LlamaConfig, LlamaDecoderLayer, LlamaForCausalLM
    import models.vlm.prismatic as _pm
    _pm.visualize_generation_simple = lambda *a, **k: None      # otherwise every training forward writes to /media/...
    PrismaticVLM.tactile_dim = 12

    class FakeTok:
        vocab_size = 32000
        pad_token_id = 32000
        bos_token_id = 1

        def encode(self, s, add_special_tokens=False):
            return [sum(map(ord, s)) % 1000 + 10]

    class TinyBackbone(LLMBackbone):
        """Offline LLMBackbone: the reference's LlamaForCausalLM built from a config (no HF hub)."""

        def __init__(self, cfg):
            super().__init__("llama2-7b-pure")
            self.llm = LlamaForCausalLM._from_config(cfg)
            self.tokenizer = FakeTok()
            self.llm.config.use_cache = False

        def get_fsdp_wrapping_policy(self):
            return None

        def enable_gradient_checkpointing(self):
            pass

        def embed_input_ids(self, ids):
            return self.llm.get_input_embeddings()(ids)

        def forward(self, **kw):
            return self.llm(**kw)

        prompt_builder_fn = property(lambda s: PurePromptBuilder)
        transformer_layer_cls = property(lambda s: LlamaDecoderLayer)
        half_precision_dtype = property(lambda s: torch.bfloat16)
        last_layer_finetune_modules = property(lambda s: ())

    ns = types.SimpleNamespace(MLA=MLA, PrismaticVLM=PrismaticVLM, LlamaForCausalLM=LlamaForCausalLM,
                               LlamaConfig=LlamaConfig, LlamaDecoderLayer=LlamaDecoderLayer, TinyBackbone=TinyBackbone,
                               ActionTokenizer=ActionTokenizer, FakeTok=FakeTok, Point_PN_scan=Point_PN_scan)
    _loaded = ns
    return ns
