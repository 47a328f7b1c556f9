"""One profiled training step at the benchmark shapes (for ncu): warm-up steps run outside the capture range,
then cudaProfilerStart ... one step ... cudaProfilerStop.  Use with `ncu --profile-from-start off`."""
import os
import sys

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
    trainer = DataParallelTrainer(mla)
    tokens = B * 4 * 548
    mla.vlm.llm_backbone.llm.model.set_save_levels(plan_save_levels(mla, tokens))
    b = map_tensors(make_batch(B, 32, 0, use_pointcloud=use_pc, use_tactile=use_tac), lambda t: t.cuda())

    def step():
        ld, _ = mla(input_ids=b["input_ids"], attention_mask=b["attention_mask"], labels=b["labels"],
                    actions=b["actions"], images=b["images"], point_cloud=b.get("point_cloud"), tactile=b.get("tactile"),
                    proprio=b["proprio"], gripper_xyz=b.get("gripper_xyz"), action_masks=b["action_masks"],
                    camera_name="rlbench_front", repeated_diffusion_steps=4, use_diff=True)
        ld["total_loss"].backward()
        trainer.step()

    for _ in range(int(os.environ.get("WARM", 2))):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
