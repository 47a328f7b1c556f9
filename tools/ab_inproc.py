"""In-process A/B of runtime switches on the real training step: the model is built once and the configurations are
interleaved round-robin (boxes and even minutes differ by several % in power-capped clocks, so separate runs do not
compare).  usage: python tools/ab_inproc.py name=switch:val,switch:val ...   switches: dyn, rope, gemm (0/1/2)"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mla_b200 import _lib, llama, ops  # noqa: E402
from mla_b200.synthetic import make_batch, map_tensors  # noqa: E402
from mla_b200.trainer import DataParallelTrainer, plan_save_levels  # noqa: E402


def apply(cfg):
    ops.DYNAMIC_TILES["on"] = bool(int(cfg.get("dyn", 0)))
    llama.FUSE_ROPE["on"] = bool(int(cfg.get("rope", 1)))
    _lib.lib().mla_gemm_set_mode(C.c_int32(int(cfg.get("gemm", 1))))
    ops.ATTN_IMPL["fwd"] = ops.ATTN_IMPL["bwd"] = cfg.get("attn", "sm100")


def main():
    specs = sys.argv[1:] or ["base=", "dyn=dyn:1"]
    cfgs = {}
    for s in specs:
        name, _, rest = s.partition("=")
        cfgs[name] = dict(kv.split(":") for kv in rest.split(",") if kv)
    workload = os.environ.get("WORKLOAD", "cfg2")
    use_pc, use_tac, _, _ = bench.WORKLOADS[workload]
    mla = bench.build_model(workload)
    trainer = DataParallelTrainer(mla)
    tokens = 8 * 4 * 548
    mla.vlm.llm_backbone.llm.model.set_save_levels(plan_save_levels(mla, tokens))
    b = map_tensors(make_batch(8, 32, 0, use_pointcloud=use_pc, use_tactile=use_tac), lambda t: t.cuda())

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
    times = {n: [] for n in cfgs}
    for rnd in range(int(os.environ.get("ROUNDS", 4))):
        for n, c in cfgs.items():
            apply(c)
            step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                step()
            e1.record()
            torch.cuda.synchronize()
            times[n].append(e0.elapsed_time(e1) / 3)
    out = {n: {"mean_ms": round(sum(t) / len(t), 2), "min_ms": round(min(t), 2), "all": [round(x, 1) for x in t]}
           for n, t in times.items()}
    for n, v in out.items():
        print(f"{n:14s} mean {v['mean_ms']:8.2f}  min {v['min_ms']:8.2f}  {v['all']}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/ab_inproc.json", "w"), indent=1)


if __name__ == "__main__":
    main()
