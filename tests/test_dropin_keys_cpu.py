"""Drop-in contract with the reference's training strategy: for every golden case our module tree has the
reference's state_dict keys with the same shapes and dtypes, `freeze_backbones(stage)` leaves exactly the same
parameters trainable, and `all_module_keys` / `trainable_module_keys` are the reference's lists.  The expectations in
tests/golden/state_dict_keys.json were recorded from the unmodified reference (make_golden_keys.py).  CPU only: module
construction runs no kernel."""
import json
import os

import pytest
import torch

from test_oracle_vs_golden import GOLD, build_state_dict, case_cfg

REC = json.load(open(os.path.join(GOLD, "state_dict_keys.json")))
SIGS = REC.pop("__signatures__")


@pytest.mark.parametrize("name", sorted(REC))
def test_state_dict_and_stage_contract(name):
    c = case_cfg(name)
    mla, _ = build_state_dict(c, dtype=torch.float32)
    ref = REC[name]
    ours = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in mla.state_dict().items()}
    missing = sorted(set(ref["state_dict"]) - set(ours))
    extra = sorted(set(ours) - set(ref["state_dict"]))
    assert not missing, ("reference keys we lack", missing[:8])
    assert not extra, ("keys the reference does not have", extra[:8])
    wrong = [(k, ours[k], v) for k, v in ref["state_dict"].items() if ours[k] != v]
    assert not wrong, wrong[:5]
    assert list(mla.all_module_keys) == ref["all_module_keys"]
    for stage, want in ref["stages"].items():
        mla.freeze_backbones(stage)
        got = sorted(k for k, p in mla.named_parameters() if p.requires_grad)
        assert got == want["requires_grad"], (stage, sorted(set(got) ^ set(want["requires_grad"]))[:8])
        assert list(mla.trainable_module_keys) == want["trainable_module_keys"], stage


# parameters we add on purpose (always after the reference's, keyword-only in practice)
EXTRAS = {"PrismaticVLM.__init__": {"image_hidden_dim"}, "PrismaticVLM.forward": {"image_repeat"},
          "MLA.predict_action_diff": {"camera_name", "use_kv_cache"}}
# defaults we relax on purpose: the benches build MLA without an action tokenizer
RELAXED = {("MLA.__init__", "action_tokenizer")}


@pytest.mark.parametrize("what", sorted(SIGS))
def test_call_signatures_match_reference(what):
    """Same parameter names, same positional order, same defaults as the reference's MLA / PrismaticVLM entry points
    (a caller that passes positionally or by keyword lands on the same parameter); our few extras come last."""
    import inspect
    from mla_b200 import MLA, PrismaticVLM
    fn = {"MLA.__init__": MLA.__init__, "MLA.forward": MLA.forward, "PrismaticVLM.__init__": PrismaticVLM.__init__,
          "PrismaticVLM.forward": PrismaticVLM.forward, "MLA.predict_action_diff": MLA.predict_action_diff,
          "MLA.create_ddim": MLA.create_ddim}[what]
    ours = inspect.signature(fn).parameters
    ref = SIGS[what]
    ref_names = [r[0] for r in ref if r[1] not in ("VAR_KEYWORD", "VAR_POSITIONAL")]
    our_names = [n for n, p in ours.items() if p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]
    extras = [n for n in our_names if n not in ref_names]
    assert set(extras) <= EXTRAS.get(what, set()), extras
    assert our_names[:len(ref_names)] == ref_names                       # identical positional order, extras after
    for name, _kind, has_default, default in ref:
        if name not in ours or (what, name) in RELAXED:
            continue
        p = ours[name]
        assert (p.default is not inspect.Parameter.empty) == has_default, (what, name)
        if has_default and isinstance(default, (int, float, bool, str, type(None))) and not isinstance(p.default, type(inspect)):
            if isinstance(p.default, (int, float, bool, str, type(None))):
                assert p.default == default, (what, name, p.default, default)


def test_output_type_fields_match_reference():
    """CausalLMOutputWithPast with the reference's three extra fields, in its order (transformers/modeling_outputs.py:706-713)."""
    import dataclasses
    from mla_b200 import CausalLMOutputWithPast
    assert [f.name for f in dataclasses.fields(CausalLMOutputWithPast)] == [
        "loss", "img_pc_contrastive_loss", "tactile_contrastive_loss", "logits", "all_logits_for_action",
        "past_key_values", "hidden_states", "attentions"]


def test_embedding_resize_like_train_py():
    """scripts/train.py:132-155 (smart_tokenizer_and_embedding_resize): add <BOD>/<EOD>, resize with
    pad_to_multiple_of=64, then overwrite the new rows with the mean row — on our LlamaForCausalLM."""
    from mla_b200 import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=32001, hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=4)
    llm = LlamaForCausalLM(cfg)
    with torch.no_grad():
        llm.model.embed_tokens.weight.normal_()
        llm.lm_head.weight.normal_()
    e0, h0 = llm.get_input_embeddings().weight.detach().clone(), llm.get_output_embeddings().weight.detach().clone()
    out = llm.resize_token_embeddings(32003, pad_to_multiple_of=64)
    assert out is llm.get_input_embeddings()
    assert llm.get_input_embeddings().weight.shape == (32064, 32) and llm.lm_head.weight.shape == (32064, 32)
    assert llm.config.vocab_size == 32064 and llm.lm_head.in_features == 32
    assert torch.equal(llm.get_input_embeddings().weight[:32001], e0) and torch.equal(llm.lm_head.weight[:32001], h0)
    assert llm.get_input_embeddings().weight.requires_grad and llm.get_input_embeddings().weight.dtype == torch.float32
    n_new = 2
    ie, oe = llm.get_input_embeddings().weight.data, llm.get_output_embeddings().weight.data
    ie[-n_new:] = ie[:-n_new].mean(dim=0, keepdim=True)
    oe[-n_new:] = oe[:-n_new].mean(dim=0, keepdim=True)
    assert llm.resize_token_embeddings(32064) is llm.get_input_embeddings()            # no-op at the same size
    assert llm.resize_token_embeddings(None) is llm.get_input_embeddings()
    g = llm.generation_config
    assert (g.bos_token_id, g.eos_token_id, g.pad_token_id) == (1, 2, cfg.pad_token_id)


def test_vlm_device_property():
    c = case_cfg("tiny_img")
    mla, _ = build_state_dict(c, dtype=torch.float32)
    assert mla.vlm.device == torch.device("cpu")


def test_initialize_weights_reaches_decoder_projections():
    """prismatic.py:299-321: `self.apply(_basic_init)` xavier-initialises EVERY nn.Linear, the LLM's q/k/v/o/gate/up/down
    included (they are `_Proj` holders here) — bounded by sqrt(6 / (fan_in + fan_out)), unlike the HF N(0, 0.02) init."""
    import math
    import torch
    from mla_b200.backbone import LLMBackbone, LlamaConfig
    from mla_b200.vlm import PrismaticVLM
    torch.manual_seed(0)
    cfg = LlamaConfig(vocab_size=32064, hidden_size=128, intermediate_size=352, num_hidden_layers=2, num_attention_heads=4)
    vlm = PrismaticVLM("tiny", LLMBackbone(config=cfg), token_size=128, action_dim=7, use_diff=True, use_pointcloud=False,
                       use_tactile=False, use_contrastive=False, use_generation=False)
    vlm.initialize_weights()
    for name, (fi, fo) in (("self_attn.q_proj", (128, 128)), ("mlp.gate_proj", (128, 352)), ("mlp.down_proj", (352, 128))):
        w = vlm.llm_backbone.llm.model.layers[1].get_submodule(name).weight
        bound = math.sqrt(6.0 / (fi + fo))
        assert float(w.abs().max()) <= bound + 1e-6, name
        assert float(w.std()) == pytest.approx(bound / math.sqrt(3.0), rel=0.05), name       # uniform(-b, b)
