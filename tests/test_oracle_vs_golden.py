"""Pins the oracle (oracle/mla.py, oracle/llama.py) to the reference: replays the committed golden vectors that
tests/golden/make_golden.py recorded from the UNMODIFIED reference run on CPU (bf16 parameters + autocast)."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_case(name):
    path = os.path.join(GOLD, name + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} missing")
    z = np.load(path, allow_pickle=False)
    batch = {"images": {}}
    for k in z.files:
        if k.startswith("batch.images."):
            batch["images"][k[len("batch.images."):]] = torch.from_numpy(z[k].astype(np.float32))
        elif k.startswith("batch."):
            v = torch.from_numpy(z[k])
            batch[k[len("batch."):]] = v.float() if v.dtype == torch.float16 else v
    return z, batch


def case_cfg(name):
    from golden.make_golden import CASES
    return CASES[name]


def build_state_dict(c, dtype=torch.bfloat16):
    """Our module tree (CPU construction only — no kernels run) supplies keys/shapes; values from fill_state_dict."""
    from mla_b200.backbone import LLMBackbone, LlamaConfig
    from mla_b200.mla import MLA
    from mla_b200.vlm import PrismaticVLM
    from mla_b200 import pointcloud
    from oracle import fixtures
    cfg = LlamaConfig(vocab_size=32064, hidden_size=c["h"], intermediate_size=c["f"], num_hidden_layers=c["L"],
                      num_attention_heads=c["heads"], rms_norm_eps=1e-5)
    flags = dict(use_diff=True, use_pointcloud=c["use_pointcloud"], use_tactile=c["use_tactile"],
                 use_contrastive=c["use_contrastive"], use_generation=False)
    flags.update(c.get("gen", {}))
    vlm = PrismaticVLM("tiny", LLMBackbone(config=cfg), token_size=c["h"], action_dim=7, **flags,
                       **c.get("gen_kwargs", {}))
    if c["use_pointcloud"] and c.get("n_points", 1024) != 1024:
        vlm.vision_tower_3d.patch_embed = pointcloud.Point_PN_scan(input_points=c["n_points"], k_neighbors=c["k"])
    mla = MLA(vlm, None, token_size=c["h"], action_dim=7, future_action_window_size=c["T"], **flags)
    sd = fixtures.fill_state_dict(mla.state_dict(), seed=7)
    return mla, {k: (v.to(dtype) if torch.is_floating_point(v) else v) for k, v in sd.items()}


def oracle_cfg(c):
    return dict(n_heads=c["heads"], rms_eps=1e-5, future_action_window_size=c["T"], repeated_diffusion_steps=c["R"],
                use_pointcloud=c["use_pointcloud"], use_tactile=c["use_tactile"], use_contrastive=c["use_contrastive"],
                camera_name="rlbench_front", k_neighbors=c.get("k", 81), **c.get("gen", {}))


def draws_of(z):
    d = dict(noise=torch.from_numpy(z["noise"]), timestep=torch.from_numpy(z["timestep"]))
    starts = [torch.from_numpy(z[k]) for k in sorted(f for f in z.files if f.startswith("fps_start_"))]
    if starts:
        d["fps_starts"] = starts
    knn = [torch.from_numpy(z[k].astype(np.int64)) for k in sorted(f for f in z.files if f.startswith("knn_idx_"))]
    if knn:
        d["knn_idx"] = knn
    return d


@pytest.mark.parametrize("name", ["tiny_img", "tiny_pc", "align"])
def test_oracle_reproduces_reference(name):
    from oracle import mla as O
    z, batch = load_case(name)
    c = case_cfg(name)
    _, sd = build_state_dict(c)
    with torch.no_grad():
        out = O.forward(sd, batch, oracle_cfg(c), draws_of(z), compute_dtype=torch.bfloat16, flavor="cpu")
    assert np.array_equal(out["mask"].numpy(), z["fused_attention_mask"])
    # the multimodal sequence fed to the decoder
    e = rel_err(out["embeds"], torch.from_numpy(z["hidden_first"]))
    assert e < 2e-3, ("fused embeddings", e)     # tokenizers, projectors, embedders, splice: pinned tight
    valid = torch.from_numpy(z["fused_attention_mask"]).bool()
    e = rel_err(out["hidden_states"][-1][valid], torch.from_numpy(z["hidden_last"])[valid])
    assert e < 2e-2, ("last hidden", e)          # decoder: the CPU reference runs SDPA, not flash-attn
    e = rel_err(out["noise_pred"], torch.from_numpy(z["noise_pred"]))
    assert e < 2e-2, ("noise_pred", e)
    tol = 2e-3
    assert abs(float(out["total_loss"]) - float(z["total_loss"])) <= tol * abs(float(z["total_loss"])), \
        (float(out["total_loss"]), float(z["total_loss"]))
    for k in ("img_pc_contrastive_loss", "tactile_contrastive_loss"):
        if k in z.files:
            assert abs(float(out[k]) - float(z[k])) <= 3e-3 * abs(float(z[k])), (k, float(out[k]), float(z[k]))


def test_oracle_fp32_is_close_to_bf16_reference():
    """The fp32 run of the oracle is the truth the bf16 implementations are judged against: it must sit within bf16
    noise of the reference's own bf16 result."""
    from oracle import mla as O
    z, batch = load_case("tiny_img")
    c = case_cfg("tiny_img")
    _, sd = build_state_dict(c)          # bf16-rounded weights, as the reference saw them
    sd32 = {k: (v.float() if torch.is_floating_point(v) else v) for k, v in sd.items()}
    with torch.no_grad():
        out = O.forward(sd32, batch, oracle_cfg(c), draws_of(z), compute_dtype=torch.float32)
    valid = torch.from_numpy(z["fused_attention_mask"]).bool()
    assert rel_err(out["hidden_states"][-1][valid], torch.from_numpy(z["hidden_last"])[valid]) < 2e-2
    assert abs(float(out["total_loss"]) - float(z["total_loss"])) < 2e-2 * abs(float(z["total_loss"]))


def test_oracle_reproduces_reference_generation_heads():
    """Post-training heads: the oracle replays the reference's CPU run of the image / point-cloud /
    tactile generation modules and their losses (dropout off on both sides)."""
    from golden.make_golden import GEN_PATCH_ROWS
    from oracle import mla as O
    z, batch = load_case("gen")
    c = case_cfg("gen")
    mla, sd = build_state_dict(c)
    # key-for-key parity of the generation manager with the reference module tree (names recorded with the gradients)
    ours = set(mla.state_dict())
    for k in z.files:
        if k.startswith("grad.vlm.generation_manager."):
            assert k[len("grad."):] in ours, k
    with torch.no_grad():
        out = O.forward(sd, batch, oracle_cfg(c), draws_of(z), compute_dtype=torch.bfloat16, flavor="cpu")
    assert np.array_equal(out["generation_roi_mask"].numpy(), z["generation_roi_mask"])
    for k, tol in (("image_gen_loss", 3e-3), ("point_cloud_gen_loss", 3e-3), ("tactile_gen_loss", 5e-3), ("total_loss", 3e-3)):
        assert abs(float(out[k]) - float(z[k])) <= tol * abs(float(z[k])), (k, float(out[k]), float(z[k]))
    it = out["image_loss_terms"]
    for k in ("image_roi_generation_loss", "bg_consistency_loss", "delta_magnitude_reward"):
        assert abs(float(it[k]) - float(z[k])) <= 1e-2 * abs(float(z[k])), (k, float(it[k]), float(z[k]))
    assert rel_err(out["image_generation"][:, GEN_PATCH_ROWS], torch.from_numpy(z["image_generation_rows"])) < 2e-2
    assert rel_err(out["alpha_all"], torch.from_numpy(z["alpha_all"])) < 2e-2
    assert rel_err(out["pointcloud_coord_generation"], torch.from_numpy(z["pointcloud_coord_generation"])) < 3e-2
    assert rel_err(out["tactile_generation"], torch.from_numpy(z["tactile_generation"])) < 3e-2


@pytest.mark.parametrize("name", ["tiny_img", "tiny_pc"])
def test_oracle_autograd_reproduces_reference_tokenizer_gradients(name):
    """Stage "pretrain" (prismatic.py:427-434): the gradients the unmodified reference gives the two tokenizers
    (tests/golden/pretrain_*.npz, recorded by make_golden_pretrain.py on the forward of the goldens above) against
    autograd through the oracle in the same bf16-on-CPU arithmetic.  This pins the oracle as the gradient checker of
    tests/test_tower_bwd_gpu.py."""
    from oracle import mla as O
    g = np.load(os.path.join(GOLD, "pretrain_" + name + ".npz"))
    z, batch = load_case(name)
    c = case_cfg(name)
    _, sd = build_state_dict(c)
    keys = [k[len("gradnorm."):] for k in g.files if k.startswith("gradnorm.")]
    assert keys and all(k.startswith(("vlm.vision_tower_2d.", "vlm.vision_tower_3d.")) for k in keys)
    sd = dict(sd)
    for k in keys:
        sd[k] = sd[k].clone().requires_grad_(True)
    out = O.forward(sd, batch, oracle_cfg(c), draws_of(z), compute_dtype=torch.bfloat16, flavor="cpu")
    assert abs(float(out["total_loss"]) - float(g["total_loss"])) <= 2e-3 * abs(float(g["total_loss"]))
    out["total_loss"].backward()
    gmax = max(float(g["gradnorm." + k]) for k in keys)
    worst = 0.0
    for k in keys:
        ref_norm = float(g["gradnorm." + k])
        got = sd[k].grad
        assert got is not None, k
        if ref_norm < 1e-4 * gmax:            # conv bias in front of a train-mode BatchNorm: identically zero gradient
            assert float(got.float().norm()) < 2e-3 * gmax, k
            continue
        assert abs(float(got.float().norm()) - ref_norm) <= 0.1 * ref_norm, (k, float(got.float().norm()), ref_norm)
        if "grad." + k in g.files:
            e = rel_err(got.float(), torch.from_numpy(g["grad." + k]))
            worst = max(worst, e)
            assert e < 0.15, (k, e)           # two bf16 backward passes through the same graph (different op kernels)
    # parameters the reference leaves without a gradient (GlobalAttention, class/split embeddings, cls_token, pos_embed)
    for k in g["params_without_grad"].tolist():
        assert k not in keys
