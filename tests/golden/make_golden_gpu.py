"""Generates tests/golden/*_gpu.npz: the UNMODIFIED reference run ON A B200 through its own flash-attn path
(`_attn_implementation="flash_attention_2"`, flash-attn 2.8.3, bf16 parameters + autocast) — the arithmetic
BASELINE.json's north_star names as the parity target.  The second golden source next to the CPU/SDPA goldens of
make_golden.py.

    gpurun -- python tests/golden/make_golden_gpu.py --out gpurun_out/golden_gpu     # then copy *.npz to tests/golden/

Needs the reference at baseline/_ref (tools/install_reference.py; /root/reference does not exist on the GPU box).

* tiny_img / tiny_pc / align: same models, weights, batches AND random draws as the CPU goldens (noise, timesteps, FPS
  starts and kNN sets of <case>.npz are replayed), so CPU-SDPA reference, GPU-flash-attn reference and the CUDA path
  are all fed identical inputs.  Stored: boundary tensors, losses, probe gradients (outputs only — inputs are in
  <case>.npz).
* layer7b: ONE LlamaDecoderLayer at full Llama-2-7B width (h 4096, 32 heads x 128, ffn 11008) on 2 x 548 tokens, the
  second sequence right-padded by 5 (-> flash_attn_varlen_func): forward + backward.  Weights =
  fixtures.fill_state_dict(seed 7) on the layer's own state dict, input / upstream gradient seeded — nothing but the
  results is stored: a probe set of output rows, input-gradient rows, a slice of every weight gradient + all norms.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import fixtures, ref_shim  # noqa: E402
from make_golden import CASES, PROBE_GRADS, build_reference  # noqa: E402

LAYER7B = dict(h=4096, f=11008, heads=32, B=2, S=548, pad=5, eps=1e-5)
LAYER7B_ROWS = [0, 1, 63, 64, 127, 128, 300, 511, 542, 547, 548, 549, 800, 1000, 1090]   # row = b*S + s (valid rows)


def layer7b_inputs(device="cpu"):
    """x [B,S,h] bf16, upstream gradient g [B,S,h] bf16 (zero on padded rows), mask bool [B,S] — pure functions of seeds."""
    c = LAYER7B
    g = torch.Generator().manual_seed(4242)
    x = (torch.randn(c["B"], c["S"], c["h"], generator=g) * 0.7).to(torch.bfloat16)
    dy = (torch.randn(c["B"], c["S"], c["h"], generator=g) * 0.1).to(torch.bfloat16)
    mask = torch.ones(c["B"], c["S"], dtype=torch.bool)
    mask[1, c["S"] - c["pad"]:] = False
    dy = dy * mask[..., None]
    return x.to(device), dy.to(device), mask.to(device)


def _load_batch(name):
    from test_oracle_vs_golden import load_case
    return load_case(name)


def run_case(name, out_dir):
    ns = ref_shim.load()
    c = CASES[name]
    z, batch = _load_batch(name)
    torch.manual_seed(0)
    mla = build_reference(ns, c)
    mla.vlm.llm_backbone.llm.config._attn_implementation = "flash_attention_2"
    # the attention class is chosen at construction (LLAMA_ATTENTION_CLASSES[config._attn_implementation]): rebuild
    import transformers.models.llama.modeling_llama as ML
    for i, layer in enumerate(mla.vlm.llm_backbone.llm.model.layers):
        old = layer.self_attn
        new = ML.LlamaFlashAttention2(config=mla.vlm.llm_backbone.llm.config, layer_idx=i)
        new.load_state_dict(old.state_dict())
        layer.self_attn = new
    fixtures.fill_state_dict(mla.state_dict(), seed=7)
    mla.to(torch.bfloat16).cuda().train()
    mla.vlm.freeze_backbones(c.get("stage", "finetune"))
    dev = lambda t: t.cuda() if torch.is_tensor(t) else t
    batch = {k: ({kk: dev(vv) for kk, vv in v.items()} if isinstance(v, dict) else dev(v)) for k, v in batch.items()}
    ints = [torch.from_numpy(z["timestep"]).cuda()] + [torch.from_numpy(z[k]).cuda() for k in
                                                       sorted(f for f in z.files if f.startswith("fps_start_"))]
    knn = [torch.from_numpy(z[k].astype(np.int64)).cuda() for k in sorted(f for f in z.files if f.startswith("knn_idx_"))]
    import models.mla.pointcloud.backbone.Point_PN as PPN
    o_randint, o_randn_like, o_knn = torch.randint, torch.randn_like, PPN.knn_point
    captured = {}
    mla.vlm.register_forward_hook(lambda m, a, o: captured.update(noise_pred=o[1].detach().float()))
    torch.randint = lambda *a, **k: ints.pop(0)
    torch.randn_like = lambda x, *a, **k: torch.from_numpy(z["noise"]).to(x.device, x.dtype)
    if knn:
        PPN.knn_point = lambda n, xyz, new_xyz: knn.pop(0)
    try:
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss_dict, out = mla(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                                 labels=batch["labels"], actions=batch["actions"], images=batch["images"],
                                 camera_name="rlbench_front", point_cloud=batch.get("point_cloud"),
                                 tactile=batch.get("tactile"), proprio=batch["proprio"],
                                 gripper_xyz=batch.get("gripper_xyz"), action_masks=batch["action_masks"],
                                 output_hidden_states=True, repeated_diffusion_steps=c["R"], use_diff=True)
    finally:
        torch.randint, torch.randn_like, PPN.knn_point = o_randint, o_randn_like, o_knn
    loss_dict["total_loss"].backward()
    f32 = lambda t: t.detach().float().cpu().numpy()
    hs = out.hidden_states
    save = {"noise_pred": captured["noise_pred"].cpu().numpy(), "hidden_first": f32(hs[0]), "hidden_last": f32(hs[-1]),
            "total_loss": f32(loss_dict["total_loss"]),
            "flash_attn_version": np.array(__import__("flash_attn").__version__),
            "device": np.array(torch.cuda.get_device_name(0))}
    if len(hs) > 8:
        save["hidden_8"] = f32(hs[8])
    for k in ("img_pc_contrastive_loss", "tactile_contrastive_loss"):
        if c["use_contrastive"] and k in loss_dict and torch.is_tensor(loss_dict[k]):
            save[k] = f32(loss_dict[k])
    named = dict(mla.named_parameters())
    for k in PROBE_GRADS:
        if k in named and named[k].grad is not None:
            g = named[k].grad.float().cpu()
            save["grad." + k] = g.numpy() if g.numel() <= 70000 else g.flatten()[:70000].numpy()
            save["gradnorm." + k] = np.array(g.norm().item(), dtype=np.float32)
    path = os.path.join(out_dir, name + "_gpu.npz")
    np.savez_compressed(path, **save)
    print(name, "loss gpu/flash-attn", float(loss_dict["total_loss"]), "cpu/sdpa golden", float(z["total_loss"]), "->", path,
          f"{os.path.getsize(path) / 1e6:.2f} MB", flush=True)


def run_layer7b(out_dir):
    ns = ref_shim.load()
    c = LAYER7B
    cfg = ns.LlamaConfig(vocab_size=32064, hidden_size=c["h"], intermediate_size=c["f"], num_hidden_layers=1,
                         num_attention_heads=c["heads"], num_key_value_heads=c["heads"], max_position_embeddings=2048,
                         rms_norm_eps=c["eps"])
    cfg._attn_implementation = "flash_attention_2"
    layer = ns.LlamaDecoderLayer(cfg, 0)
    fixtures.fill_state_dict(layer.state_dict(), seed=7)
    layer.to(torch.bfloat16).cuda().train()
    x, dy, mask = layer7b_inputs("cuda")
    x.requires_grad_(True)
    pos = torch.arange(c["S"], device="cuda")[None]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = layer(x, attention_mask=mask.long(), position_ids=pos)[0]
    y.backward(dy)
    rows = torch.tensor(LAYER7B_ROWS, device="cuda")
    f32 = lambda t: t.detach().float().cpu().numpy()
    y2, dx2 = y.reshape(-1, c["h"]), x.grad.reshape(-1, c["h"])
    valid = mask.reshape(-1)
    save = {"rows": np.array(LAYER7B_ROWS), "y_rows": f32(y2[rows]), "dx_rows": f32(dx2[rows]),
            "y_norm": np.array(y2[valid].float().norm().item(), np.float32),
            "dx_norm": np.array(dx2[valid].float().norm().item(), np.float32),
            "y_colsum": f32(y2[valid].float().sum(0)), "dx_colsum": f32(dx2[valid].float().sum(0)),
            "flash_attn_version": np.array(__import__("flash_attn").__version__),
            "device": np.array(torch.cuda.get_device_name(0))}
    for k, p in layer.named_parameters():
        g = p.grad.float()
        save["gradnorm." + k] = np.array(g.norm().item(), np.float32)
        save["grad." + k] = f32(g.flatten()[:65536])
    path = os.path.join(out_dir, "layer7b_gpu.npz")
    np.savez_compressed(path, **save)
    print("layer7b |y|", float(save["y_norm"]), "|dx|", float(save["dx_norm"]), "->", path,
          f"{os.path.getsize(path) / 1e6:.2f} MB", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=HERE)
    ap.add_argument("cases", nargs="*", default=["tiny_img", "tiny_pc", "align", "layer7b"])
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for n in a.cases:
        if n == "layer7b":
            run_layer7b(a.out)
        else:
            run_case(n, a.out)
