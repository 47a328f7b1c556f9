"""Generates tests/golden/ddim_*.npz: the reference's inference denoise loop (the core of MLA.predict_action_diff,
models/mla/model_mla.py:709-772) run by the UNMODIFIED reference on CPU.

    python tests/golden/make_golden_ddim.py       # needs /root/reference (not present on the GPU box)

Same Tiny-MLA models / weights / batches as make_golden.py, in eval mode, bf16 parameters under autocast(bf16):
`create_ddim(ddim_step)` (:1166-1173) then `ddim_diffusion.ddim_sample_loop(vlm.forward, noise.shape, noise,
clip_denoised=False, model_kwargs={input_ids, images, point_cloud, proprio}, eta=0.0)` (:746-755).  input_ids end
with the empty token 29871, the inference tag the action tokens are spliced in front of (prismatic.py:886,:983).
Recorded: the batch, the starting noise, and per DDIM step the model timestep, the step input x_t and the model's
noise prediction; the final sample; the respaced schedule (timestep_map, alphas_cumprod).
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
DDIM_CASES = {"ddim_tiny_img": ("tiny_img", 8), "ddim_tiny_pc": ("tiny_pc", 4)}


def run_case(name, base, ddim_steps):
    ns = ref_shim.load()
    c = CASES[base]
    torch.manual_seed(0)
    mla = build_reference(ns, c)
    fixtures.fill_state_dict(mla.state_dict(), seed=7)
    mla.to(torch.bfloat16).eval()
    batch = fixtures.synthetic_batch(c["B"], c["Lt"], c["T"], c["hw"], c.get("n_points", 1024), seed=4321,
                                     use_pointcloud=c["use_pointcloud"], use_tactile=False, pad_last=0)
    ids = batch["input_ids"].clone()
    ids[:, -1] = 29871                         # predict_action_diff: input_ids[:, :-3] ends with the empty token
    g = torch.Generator().manual_seed(11)
    noise = torch.randn(c["B"], c["T"] + 1, 7, generator=g)
    mla.create_ddim(ddim_step=ddim_steps)
    dd = mla.ddim_diffusion
    steps = []
    fwd = mla.vlm.forward

    def model(x, t, **kw):
        out = fwd(x, t, **kw)
        steps.append((t.clone(), x.detach().float().clone(), out[1].detach().float().clone()))
        return out

    # BatchNorm of the point tokenizer runs on its running statistics in eval mode: make them non-trivial
    for n, b in mla.named_buffers():
        if n.endswith("running_mean"):
            b.copy_(0.05 * torch.randn(b.shape, generator=g))
        elif n.endswith("running_var"):
            b.copy_(1.0 + 0.2 * torch.rand(b.shape, generator=g))
    bn_stats = {n: b.detach().float().numpy().copy() for n, b in mla.named_buffers() if "running_" in n}
    starts, knn_rec = [], []
    import models.mla.pointcloud.backbone.Point_PN as PPN
    o_randint, o_knn = torch.randint, PPN.knn_point

    def rec_randint(*a, **k):
        v = o_randint(*a, **k)
        starts.append(v.clone())
        return v

    def rec_knn(n, xyz, new_xyz):
        r = o_knn(n, xyz, new_xyz)
        knn_rec.append(r.clone())
        return r

    torch.randint, PPN.knn_point = rec_randint, rec_knn
    # model_kwargs of predict_action_diff (:735-740) + camera_name: get_fused_tokens looks the camera up
    # unconditionally (prismatic.py:604), so the published call without it raises ValueError
    kwargs = dict(input_ids=ids, images=batch["images"], point_cloud=batch.get("point_cloud"),
                  proprio=batch["proprio"], camera_name="rlbench_front")
    try:
        with torch.inference_mode(), torch.autocast("cpu", dtype=torch.bfloat16):
            sample = dd.ddim_sample_loop(model, noise.shape, noise, clip_denoised=False, model_kwargs=kwargs,
                                         progress=False, device="cpu", eta=0.0)
    finally:
        torch.randint, PPN.knn_point = o_randint, o_knn
    save = {"noise": noise.numpy(), "sample": sample.float().numpy(), "ddim_steps": np.array(ddim_steps),
            "timestep_map": np.array(dd.timestep_map), "alphas_cumprod": np.asarray(dd.alphas_cumprod, dtype=np.float64),
            "input_ids": ids.numpy(), "proprio": batch["proprio"].numpy(),
            "front_image": batch["images"]["front_image"].numpy()}
    if c["use_pointcloud"]:
        save["point_cloud"] = batch["point_cloud"].numpy()
    n_stage = len(starts) // ddim_steps if starts else 0
    for s, (t, x, eps) in enumerate(steps):
        save[f"step{s}.t"], save[f"step{s}.x"], save[f"step{s}.eps"] = t.numpy(), x.numpy(), eps.numpy()
        for j in range(n_stage):       # the reference redraws the FPS starts (and re-runs the tokenizers) every step
            save[f"step{s}.fps_start_{j}"] = starts[s * n_stage + j].numpy()
            save[f"step{s}.knn_idx_{j}"] = knn_rec[s * n_stage + j].numpy().astype(np.int16)
    for n, v in bn_stats.items():
        save["buf." + n] = v
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(name, "steps", len(steps), "timestep_map", dd.timestep_map, "sample", sample.float().flatten()[:4].tolist(),
          "size", os.path.getsize(os.path.join(OUT, name + ".npz")))


if __name__ == "__main__":
    for n, (base, k) in DDIM_CASES.items():
        if len(sys.argv) < 2 or n in sys.argv[1:]:
            run_case(n, base, k)
