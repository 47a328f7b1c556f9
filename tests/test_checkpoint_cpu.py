"""Checkpoint format of the data-parallel trainer: {"model": {module_key: state_dict}} with the reference's module
keys and parameter names (training/strategies/fsdp.py:100-140), readable back; optimizer moments resume exactly.
Host logic only (no kernel runs): CPU."""
import torch

from test_oracle_vs_golden import build_state_dict, case_cfg


def _model():
    c = case_cfg("tiny_pc")
    mla, sd = build_state_dict(c, dtype=torch.float32)
    mla.load_state_dict(sd)
    mla.freeze_backbones("finetune")
    return mla


def test_checkpoint_layout_and_round_trip(tmp_path):
    from mla_b200.trainer import DataParallelTrainer
    mla = _model()
    tr = DataParallelTrainer(mla, lr=1e-4)
    # fake optimizer state for two parameters + a step count
    named = dict(mla.named_parameters())
    k1, k2 = "vlm.projector_2d.mlp.0.weight", "vlm.llm_backbone.llm.model.layers.1.mlp.down_proj.weight"
    for k in (k1, k2):
        tr.state[id(named[k])] = (torch.randn_like(named[k]), torch.rand_like(named[k]))
    tr.step_count = 17
    path = tr.save_checkpoint(tmp_path, global_step=17, epoch=2, train_loss=0.12345)
    assert path.name == "step-000017-epoch-02-loss=0.1235.pt"
    blob = torch.load(path, weights_only=True)
    assert set(blob) == {"model"}
    # reference layout: trainable module keys of the finetune stage without the "vlm." prefix (prismatic.py:470-478)
    assert set(blob["model"]) == {"llm_backbone", "projector_2d", "proprio_embedder", "x_embedder", "t_embedder",
                                  "final_layer", "projector_3d"}
    assert "llm.model.layers.0.self_attn.q_proj.weight" in blob["model"]["llm_backbone"]
    assert "llm.lm_head.weight" in blob["model"]["llm_backbone"]
    assert "mlp.0.weight" in blob["model"]["projector_2d"]
    for sub in blob["model"].values():
        assert all(v.device.type == "cpu" for v in sub.values())
    want = {k: v.detach().clone() for k, v in mla.state_dict().items()}

    # a second model with different weights resumes from it: parameters, Adam moments and the step counter
    other = _model()
    with torch.no_grad():
        for p in other.parameters():
            p.add_(1.0)
    tr2 = DataParallelTrainer(other, lr=1e-4)
    sched = tr2.load_checkpoint(path)
    assert sched == {"epoch": 2, "global_step": 17} and tr2.step_count == 17
    got = other.state_dict()
    saved_prefixes = tuple("vlm." + k + "." for k in blob["model"])
    for k, v in want.items():
        if k.startswith(saved_prefixes):
            assert torch.equal(got[k], v), k
    n2 = dict(other.named_parameters())
    for k in (k1, k2):
        m, v = tr.state[id(named[k])]
        m2, v2 = tr2.state[id(n2[k])]
        assert torch.equal(m, m2) and torch.equal(v, v2)
    # the frozen tokenizers were not in the file (only_trainable): untouched by the load
    assert torch.equal(got["vlm.vision_tower_2d.patch_embedding.weight"],
                       want["vlm.vision_tower_2d.patch_embedding.weight"] + 1.0)


def test_full_checkpoint_has_all_module_keys(tmp_path):
    from mla_b200.trainer import DataParallelTrainer
    mla = _model()
    tr = DataParallelTrainer(mla)
    path = tr.save_checkpoint(tmp_path, 0, 0, None, only_trainable=False, save_optimizer=False)
    assert path.name == "step-000000-epoch-00-loss=inf.pt" and not path.with_suffix(".optimizer").exists()
    blob = torch.load(path, weights_only=True)["model"]
    assert {"vision_tower_2d", "vision_tower_3d", "llm_backbone"} <= set(blob)
