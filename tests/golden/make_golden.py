"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

    python tests/golden/make_golden.py            # needs /root/reference (not present on the GPU box)

Each case: build the reference MLA with an offline tiny Llama backbone, overwrite every weight with
oracle.fixtures.fill_state_dict (a pure function of key name + seed, so the weights are not stored), cast the
parameters to bf16 and run under autocast(bf16) — the arithmetic FSDP MixedPrecision(param_dtype=bf16) gives the
reference in training (training/strategies/fsdp.py:185-187) — in train mode, forward + backward.  The random
draws the reference makes (noise, timesteps, FPS starts) are recorded so that the oracle and the CUDA path can be
fed the same values.  Stored: the batch, the draws, boundary tensors, losses and a probe set of gradients.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fixtures, ref_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # BASELINE.json configs[0]: Tiny-MLA (2-layer/128-dim Llama, 4x4 image patches, 8 text toks) bs=2, image-only
    "tiny_img": dict(h=128, f=352, L=2, heads=4, B=2, R=2, T=0, Lt=8, hw=168, pad_last=2,
                     use_pointcloud=False, use_tactile=False, use_contrastive=False),
    # configs[0] with the 64-point cloud (Point_PN_scan(64, k=8) -> 16 point tokens)
    "tiny_pc": dict(h=128, f=352, L=2, heads=4, B=2, R=2, T=3, Lt=8, hw=168, pad_last=0, n_points=64, k=8,
                    use_pointcloud=True, use_tactile=False, use_contrastive=False),
    # configs[2] shapes at small width: 256 image + 256 point tokens, tactile, both InfoNCE losses on hidden[8]
    "align": dict(h=128, f=352, L=9, heads=4, B=1, R=2, T=0, Lt=8, hw=672, pad_last=0, n_points=1024, k=81,
                  use_pointcloud=True, use_tactile=True, use_contrastive=True),
    # configs[4] (post-training) at small width: image + point-cloud + tactile generation heads, ROI from the projected
    # point cloud; every dropout / DropPath at p = 0 (their masks cannot be shared between implementations)
    "gen": dict(h=128, f=352, L=2, heads=4, B=1, R=2, T=0, Lt=8, hw=672, pad_last=0, n_points=1024, k=81,
                use_pointcloud=True, use_tactile=False, use_contrastive=False, stage="post-training",
                gen=dict(use_generation=True, gen_image=True, gen_pointcloud=True, gen_tactile=True, use_roi=True),
                gen_kwargs=dict(num_image_gen_queries=32, pointcloud_trans_dim=256)),
}
PROBE_GRADS = [
    "vlm.llm_backbone.llm.model.layers.0.self_attn.q_proj.weight",
    "vlm.llm_backbone.llm.model.layers.1.mlp.down_proj.weight",
    "vlm.llm_backbone.llm.model.layers.0.input_layernorm.weight",
    "vlm.llm_backbone.llm.model.norm.weight",
    "vlm.projector_2d.mlp.2.weight",
    "vlm.x_embedder.mlp.fc1.weight",
    "vlm.t_embedder.mlp.0.bias",
    "vlm.final_layer.mlp.fc2.weight",
    "vlm.final_layer.norm_final.weight",
    "vlm.projector_3d.projector.0.weight",
    "vlm.llm_backbone.llm.coordinate_aware_contrastive_loss_module.image_projection_head.2.weight",
    "vlm.llm_backbone.llm.tactile_contrastive_loss_module.tactile_projection_head.0.weight",
    "vlm.tactile_embedder.mlp.fc2.bias",
    "vlm.generation_manager.image_gen_module.image_gen_queries",
    "vlm.generation_manager.image_gen_module.mae_mask_token",
    "vlm.generation_manager.image_gen_module.mae_pos_embed",
    "vlm.generation_manager.image_gen_module.intent_decoder.layers.0.multihead_attn.in_proj_weight",
    "vlm.generation_manager.image_gen_module.mae_decoder.layers.2.self_attn.out_proj.weight",
    "vlm.generation_manager.image_gen_module.mae_decoder.layers.0.linear1.weight",
    "vlm.generation_manager.image_gen_module.mae_decoder.layers.1.norm2.weight",
    "vlm.generation_manager.image_gen_module.mae_delta_head.weight",
    "vlm.generation_manager.image_gen_module.mae_alpha_head.weight",
    "vlm.generation_manager.image_gen_module.mae_offset_head.weight",
    "vlm.generation_manager.pointcloud_gen_module.feature_projector.weight",
    "vlm.generation_manager.pointcloud_gen_module.pos_embed",
    "vlm.generation_manager.pointcloud_gen_module.decoder_blocks.1.attn.in_proj_weight",
    "vlm.generation_manager.pointcloud_gen_module.future_predictor.1.weight",
    "vlm.generation_manager.pointcloud_gen_module.future_predictor.3.weight",
    "vlm.generation_manager.tactile_gen_module.tactile_query",
    "vlm.generation_manager.tactile_gen_module.decoder.layers.1.multihead_attn.out_proj.weight",
    "vlm.generation_manager.tactile_gen_module.output_head.weight",
]


GEN_PATCH_ROWS = [0, 17, 100, 119, 136, 255]


def build_reference(ns, c):
    cfg = ns.LlamaConfig(vocab_size=32064, hidden_size=c["h"], intermediate_size=c["f"], num_hidden_layers=c["L"],
                         num_attention_heads=c["heads"], num_key_value_heads=c["heads"], max_position_embeddings=2048,
                         rms_norm_eps=1e-5)
    cfg._attn_implementation = "sdpa"      # flash-attn needs a GPU; same math (causal softmax attention)
    flags = dict(use_diff=True, use_pointcloud=c["use_pointcloud"], use_tactile=c["use_tactile"],
                 use_contrastive=c["use_contrastive"], use_generation=False)
    flags.update(c.get("gen", {}))
    vlm = ns.PrismaticVLM("tiny", ns.TinyBackbone(cfg), token_size=c["h"], action_dim=7, **flags,
                          **c.get("gen_kwargs", {}))
    if c["use_pointcloud"] and c.get("n_points", 1024) != 1024:
        vlm.vision_tower_3d.patch_embed = ns.Point_PN_scan(input_points=c["n_points"], k_neighbors=c["k"])
    mla = ns.MLA(vlm, ns.ActionTokenizer(ns.FakeTok()), token_size=c["h"], action_dim=7,
                 future_action_window_size=c["T"], **flags)
    return mla


def run_case(name, c):
    ns = ref_shim.load()
    torch.manual_seed(0)
    mla = build_reference(ns, c)
    fixtures.fill_state_dict(mla.state_dict(), seed=7)
    mla.to(torch.bfloat16).train()
    mla.vlm.freeze_backbones(c.get("stage", "finetune"))
    gen = c.get("gen", {})
    if gen:     # dropout / DropPath off (timm's DropPath is an Identity stand-in in the shim already)
        for m in mla.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
            if isinstance(m, torch.nn.MultiheadAttention):
                m.dropout = 0.0
    batch = fixtures.synthetic_batch(c["B"], c["Lt"], c["T"], c["hw"], c.get("n_points", 1024), seed=1234,
                                     use_pointcloud=c["use_pointcloud"], use_tactile=c["use_tactile"],
                                     pad_last=c["pad_last"], generation=bool(gen))
    if gen:     # images are stored as fp16: run the reference on exactly the stored values
        batch["images"] = {k: v.half().float() for k, v in batch["images"].items()}
        batch["next_images"] = batch["next_images"].half().float()
        # shrink the cloud about its centre so that the ROI (projected + dilated patches) covers only part of the grid
        ctr = batch["point_cloud"].mean(dim=1, keepdim=True)
        batch["point_cloud"] = ctr + 0.35 * (batch["point_cloud"] - ctr)
    # record the reference's random draws
    draws = {"randint": [], "randn_like": []}
    o_randint, o_randn_like = torch.randint, torch.randn_like

    def rec_randint(*a, **k):
        v = o_randint(*a, **k)
        draws["randint"].append(v.clone())
        return v

    def rec_randn_like(*a, **k):
        v = o_randn_like(*a, **k)
        draws["randn_like"].append(v.clone())
        return v

    captured = {}
    mla.vlm.register_forward_hook(lambda m, a, o: captured.update(noise_pred=o[1].detach().float(), gen_out=o[2],
                                                                   gen_losses=o[3]))
    mla.vlm.llm_backbone.register_forward_pre_hook(
        lambda m, a, kw: captured.__setitem__("mask", kw["attention_mask"].clone()), with_kwargs=True)
    import models.mla.pointcloud.backbone.Point_PN as PPN
    knn_rec = []
    o_knn = PPN.knn_point

    def rec_knn(n, xyz, new_xyz):
        r = o_knn(n, xyz, new_xyz)
        knn_rec.append(r.clone())
        return r

    PPN.knn_point = rec_knn
    torch.randint, torch.randn_like = rec_randint, rec_randn_like
    torch.manual_seed(99)
    try:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            loss_dict, out = mla(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                                 labels=batch["labels"], actions=batch["actions"], images=batch["images"],
                                 camera_name=batch["camera_name"], point_cloud=batch.get("point_cloud"),
                                 tactile=batch.get("tactile"), proprio=batch["proprio"],
                                 gripper_xyz=batch.get("gripper_xyz"), action_masks=batch["action_masks"],
                                 next_images=batch.get("next_images"), next_point_cloud=batch.get("next_point_cloud"),
                                 next_tactile=batch.get("next_tactile"),
                                 output_hidden_states=True, repeated_diffusion_steps=c["R"], use_diff=True)
    finally:
        torch.randint, torch.randn_like = o_randint, o_randn_like
        PPN.knn_point = o_knn
    loss_dict["total_loss"].backward()
    # draw order: randn_like(actions_future), randint(0,100,(B_eff,)), then one randint per FPS stage (Point_PN.py:10)
    save = {"noise": draws["randn_like"][0].float().numpy(), "timestep": draws["randint"][0].numpy()}
    for i, v in enumerate(draws["randint"][1:]):
        save[f"fps_start_{i}"] = v.numpy()
    for i, v in enumerate(knn_rec):     # torch.topk tie-breaking is implementation-defined: keep the sets it chose
        save[f"knn_idx_{i}"] = v.numpy().astype(np.int16)
    f32 = lambda t: t.detach().float().numpy()
    for k, v in batch.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                save[f"batch.images.{kk}"] = vv.numpy().astype(np.float16) if c["hw"] > 200 else vv.numpy()
        elif torch.is_tensor(v):
            save["batch." + k] = v.numpy().astype(np.float16) if k == "next_images" else v.numpy()
    hs = out.hidden_states
    save["noise_pred"] = captured["noise_pred"].numpy()
    save["fused_attention_mask"] = captured["mask"].numpy()
    save["hidden_first"] = f32(hs[0])
    save["hidden_last"] = f32(hs[-1])
    if len(hs) > 8:
        save["hidden_8"] = f32(hs[8])
    save["total_loss"] = f32(loss_dict["total_loss"])
    if c["use_contrastive"]:
        save["img_pc_contrastive_loss"] = f32(loss_dict["img_pc_contrastive_loss"])
        if c["use_tactile"]:
            save["tactile_contrastive_loss"] = f32(loss_dict["tactile_contrastive_loss"])
    if gen:
        go, gl = captured["gen_out"], captured["gen_losses"]
        for k in ("image_gen_loss", "point_cloud_gen_loss", "tactile_gen_loss"):
            save[k] = f32(loss_dict[k])
        for k in ("image_roi_generation_loss", "bg_consistency_loss", "delta_magnitude_reward"):
            if k in gl:
                save[k] = f32(gl[k])
        save["generation_roi_mask"] = go["generation_roi_mask"].numpy()
        ig = go["image_generation"].detach().float()
        save["image_generation_rows"] = ig[:, GEN_PATCH_ROWS].numpy()        # a probe set of patches (full = 10.8 MB)
        save["image_generation_absmean"] = np.array(ig.abs().mean().item(), dtype=np.float32)
        save["alpha_all"], save["offset_all"] = f32(go["alpha_all"]), f32(go["offset_all"])
        save["delta_all_rows"] = f32(go["delta_all"])[:, GEN_PATCH_ROWS]
        save["pointcloud_coord_generation"] = f32(go["pointcloud_coord_generation"])
        save["tactile_generation"] = f32(go["tactile_generation"])
    named = dict(mla.named_parameters())
    for k in PROBE_GRADS:
        if k in named and named[k].grad is not None:
            g = named[k].grad.float()
            save["grad." + k] = g.numpy() if g.numel() <= 70000 else g.flatten()[:70000].numpy()
            save["gradnorm." + k] = np.array(g.norm().item(), dtype=np.float32)
    no_grad = sorted(k for k, p in named.items() if p.requires_grad and p.grad is None)
    save["params_without_grad"] = np.array(no_grad)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **save)
    print(name, {k: float(v) for k, v in loss_dict.items() if torch.is_tensor(v) and v.numel() == 1},
          "S =", hs[0].shape[1], "->", path, f"{os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    which = sys.argv[1:] or list(CASES)
    for n in which:
        run_case(n, CASES[n])
