"""GPU timeline of real training steps (power-capped clocks, real dependencies) with torch.profiler/CUPTI:
per-kernel time inside the step, busy time vs wall time, and the idle gaps.  -> gpurun_out/timeline.json"""
import json
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mla_b200.synthetic import make_batch, map_tensors  # noqa: E402
from mla_b200.trainer import DataParallelTrainer, plan_save_levels  # noqa: E402


def main():
    workload = os.environ.get("WORKLOAD", "cfg2")
    B = int(os.environ.get("B", 8))
    use_pc, use_tac, _, _ = bench.WORKLOADS[workload]
    mla = bench.build_model(workload)
    mla.share_diffusion_prefix = os.environ.get("SHARE", "0") == "1"       # SURVEY 8 f2 (cfg2 only)
    trainer = DataParallelTrainer(mla)
    tokens = B * (548 + 3 * 3) if mla.share_diffusion_prefix else B * 4 * 548
    mla.vlm.llm_backbone.llm.model.set_save_levels(plan_save_levels(mla, tokens))
    b = map_tensors(make_batch(B, 32, 0, use_pointcloud=use_pc, use_tactile=use_tac), lambda t: t.cuda())

    def step():
        ld, _ = mla(input_ids=b["input_ids"], attention_mask=b["attention_mask"], labels=b["labels"],
                    actions=b["actions"], images=b["images"], point_cloud=b.get("point_cloud"), tactile=b.get("tactile"),
                    proprio=b["proprio"], gripper_xyz=b.get("gripper_xyz"), action_masks=b["action_masks"],
                    camera_name="rlbench_front", repeated_diffusion_steps=4, use_diff=True)
        ld["total_loss"].backward()
        trainer.step()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    n_steps = 3
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n_steps):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda t: t[0])
    t0, t1 = ks[0][0], max(k[1] for k in ks)
    busy, gaps, last_end = 0.0, [], ks[0][0]
    agg = defaultdict(lambda: [0, 0.0])
    for s, e, n in ks:
        agg[n.split("(")[0][:70]][0] += 1
        agg[n.split("(")[0][:70]][1] += (e - s)
        if s > last_end:
            gaps.append((s - last_end, n.split("(")[0][:50]))
        busy += max(0.0, e - max(s, last_end))
        last_end = max(last_end, e)
    wall = t1 - t0
    gap_by = defaultdict(lambda: [0, 0.0])
    for g, n in gaps:
        gap_by[n][0] += 1
        gap_by[n][1] += g
    out = {"steps": n_steps, "wall_ms_per_step": wall / 1e3 / n_steps, "busy_ms_per_step": busy / 1e3 / n_steps,
           "idle_ms_per_step": (wall - busy) / 1e3 / n_steps,
           "kernels_ms_per_step": {k: [v[0] // n_steps, round(v[1] / 1e3 / n_steps, 3)]
                                   for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]},
           "idle_before_kernel_ms_per_step": {k: [v[0] // n_steps, round(v[1] / 1e3 / n_steps, 3)]
                                              for k, v in sorted(gap_by.items(), key=lambda kv: -kv[1][1])[:15]}}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(os.environ.get("OUT", "gpurun_out/timeline.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
