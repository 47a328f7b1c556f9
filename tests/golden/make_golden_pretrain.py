"""Generates tests/golden/pretrain_*.npz: the gradients the UNMODIFIED reference gives the two tokenizers in stage
"pretrain" (models/vlm/prismatic.py:427-434 trains vision_tower_2d / vision_tower_3d), on CPU.

    python tests/golden/make_golden_pretrain.py       # needs /root/reference (not present on the GPU box)

Same Tiny-MLA models, weights, batches AND random draws as the forward goldens of make_golden.py (the recorded noise /
timesteps / FPS starts / neighbour sets of tiny_img.npz / tiny_pc.npz are fed back through the same patches), so the
forward is the recorded one; only `freeze_backbones("pretrain")` differs.  Stored per tokenizer parameter: the gradient
norm, and the full gradient when it is small (<= 20k elements).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import fixtures, ref_shim  # noqa: E402
from make_golden import CASES, build_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def run_case(base):
    ns = ref_shim.load()
    c = CASES[base]
    z = np.load(os.path.join(OUT, base + ".npz"))
    torch.manual_seed(0)
    mla = build_reference(ns, c)
    fixtures.fill_state_dict(mla.state_dict(), seed=7)
    mla.to(torch.bfloat16).train()
    mla.vlm.freeze_backbones("pretrain")
    batch = fixtures.synthetic_batch(c["B"], c["Lt"], c["T"], c["hw"], c.get("n_points", 1024), seed=1234,
                                     use_pointcloud=c["use_pointcloud"], use_tactile=c["use_tactile"],
                                     pad_last=c["pad_last"])
    # replay the recorded draws: randn_like -> noise, randint -> timestep then one FPS start per stage
    ints = [torch.from_numpy(z["timestep"])] + [torch.from_numpy(z[k]) for k in
                                                sorted(f for f in z.files if f.startswith("fps_start_"))]
    knn = [torch.from_numpy(z[k].astype(np.int64)) for k in sorted(f for f in z.files if f.startswith("knn_idx_"))]
    import models.mla.pointcloud.backbone.Point_PN as PPN
    o_randint, o_randn_like, o_knn = torch.randint, torch.randn_like, PPN.knn_point
    torch.randint = lambda *a, **k: ints.pop(0)
    torch.randn_like = lambda x, *a, **k: torch.from_numpy(z["noise"]).to(x.dtype)
    if knn:
        PPN.knn_point = lambda n, xyz, new_xyz: knn.pop(0)
    try:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            loss_dict, _ = mla(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"],
                               labels=batch["labels"], actions=batch["actions"], images=batch["images"],
                               camera_name=batch["camera_name"], point_cloud=batch.get("point_cloud"),
                               proprio=batch["proprio"], action_masks=batch["action_masks"],
                               output_hidden_states=True, repeated_diffusion_steps=c["R"], use_diff=True)
    finally:
        torch.randint, torch.randn_like, PPN.knn_point = o_randint, o_randn_like, o_knn
    assert abs(float(loss_dict["total_loss"]) - float(z["total_loss"])) < 1e-6 * abs(float(z["total_loss"])), \
        "the replayed forward must be the recorded one"
    loss_dict["total_loss"].backward()
    save = {"total_loss": np.array(float(loss_dict["total_loss"]), np.float32)}
    no_grad = []
    for k, p in mla.named_parameters():
        if not k.startswith(("vlm.vision_tower_2d.", "vlm.vision_tower_3d.")):
            continue
        if p.grad is None:
            no_grad.append(k)
            continue
        g = p.grad.float()
        save["gradnorm." + k] = np.array(g.norm().item(), np.float32)
        if g.numel() <= 20000:
            save["grad." + k] = g.numpy()
    save["params_without_grad"] = np.array(sorted(no_grad))
    path = os.path.join(OUT, "pretrain_" + base + ".npz")
    np.savez_compressed(path, **save)
    print(base, "loss", float(loss_dict["total_loss"]), "params with grad",
          sum(1 for k in save if k.startswith("gradnorm.")), "without", len(no_grad), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    for b in (sys.argv[1:] or ["tiny_img", "tiny_pc"]):
        run_case(b)
