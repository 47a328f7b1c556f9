"""Latency of one action prediction (the DDIM denoise loop of MLA.predict_action_diff, 8 steps, batch 1) at
Llama-2-7B size: the reference's schedule (whole eval forward per step) against the K/V-cached schedule
(prefix once + suffix rows per step), and the achieved weight-streaming bandwidth of the per-step decode.

    python tools/bench_denoise.py [--T 0] [--steps 8]     -> gpurun_out/denoise_T<T>.json

CUDA events on the launching stream, 2 warm-ups, 5 timed repetitions; inputs resident on the device.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mla_b200 import _lib  # noqa: E402
from mla_b200.synthetic import make_batch, map_tensors  # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _lib.launch_count()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (_lib.launch_count() - n0) // n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=0, help="future_action_window_size (0 in the scripts, 15 by default)")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"])
    args = ap.parse_args()
    torch.cuda.set_device(0)
    mla = bench.build_model(args.workload, args.T)
    mla.requires_grad_(False)
    mla.eval()
    use_pc, use_tac, _, _ = bench.WORKLOADS[args.workload]
    b = map_tensors(make_batch(1, 32, args.T, 672, 1024, seed=1234, use_pointcloud=use_pc, use_tactile=use_tac),
                    lambda t: t.cuda())
    ids = b["input_ids"].clone()
    ids[:, -1] = 29871                                   # inference tag token (prismatic.py:886)
    kw = dict(point_cloud=b.get("point_cloud"), proprio=b["proprio"], tactile=b.get("tactile"),
              gripper_xyz=b.get("gripper_xyz"), num_ddim_steps=args.steps)
    noise = torch.randn(1, args.T + 1, 7, device="cuda")
    full = lambda: mla.denoise_actions(ids, b["images"], noise=noise, use_kv_cache=False, **kw)
    cached = lambda: mla.denoise_actions(ids, b["images"], noise=noise, use_kv_cache=True, use_cuda_graph=False, **kw)
    graphed = lambda: mla.denoise_actions(ids, b["images"], noise=noise, use_kv_cache=True, use_cuda_graph=True, **kw)
    a, c, g = full(), cached(), graphed()
    dev_rel = float((a - c).norm() / a.norm())
    graph_equal = bool(torch.equal(c, g))
    ms_full, n_full = timeit(full)
    ms_cached, n_cached = timeit(cached)
    ms_graph, _ = timeit(graphed)
    sess = mla.vlm.denoise_session(1, mla.vlm.denoise_prefill(ids, b["images"], point_cloud=b.get("point_cloud"),
                                                             proprio=b["proprio"], camera_name="rlbench_front",
                                                             tactile=b.get("tactile"), gripper_xyz=b.get("gripper_xyz"),
                                                             n_x=args.T + 1, embeds_only=True).P, args.T + 1,
                                   mla.ddim_diffusion)
    ms_gloop, _ = timeit(lambda: sess.g_loop.replay(), n=10, warm=2)
    ms_gprefill, _ = timeit(lambda: sess.g_prefill.replay(), n=10, warm=2)
    st = mla.vlm.denoise_prefill(ids, b["images"], point_cloud=b.get("point_cloud"), proprio=b["proprio"],
                                 camera_name="rlbench_front", tactile=b.get("tactile"), gripper_xyz=b.get("gripper_xyz"),
                                 n_x=args.T + 1)
    t = torch.full((1,), 50, dtype=torch.long, device="cuda")
    ms_step, n_step = timeit(lambda: mla.vlm.denoise_step(st, noise, t), n=20, warm=3)
    ms_prefill, _ = timeit(lambda: mla.vlm.denoise_prefill(ids, b["images"], point_cloud=b.get("point_cloud"),
                                                           proprio=b["proprio"], camera_name="rlbench_front",
                                                           tactile=b.get("tactile"), gripper_xyz=b.get("gripper_xyz"),
                                                           n_x=args.T + 1))
    H, F, L = bench.H, bench.F, bench.L
    wbytes = L * (4 * H * H + 3 * F * H) * 2 + 2 * (H * H + H * 8) * 2      # decoder weights + FinalLayer
    hbm, _, _, src = bench.peaks()
    out = {"workload": f"{args.workload}, batch 1, T={args.T} ({args.T + 2} suffix rows), {args.steps} DDIM steps, "
                       f"prefix {st.P} tokens", "full_forward_per_step_ms": round(ms_full, 3),
           "kv_cached_eager_ms": round(ms_cached, 3), "kv_cached_cuda_graph_ms": round(ms_graph, 3),
           "speedup_vs_full": round(ms_full / ms_graph, 2), "graph_equals_eager_bitwise": graph_equal,
           "graph_prefill_ms": round(ms_gprefill, 3), "graph_ddim_loop_ms": round(ms_gloop, 3),
           "graph_decode_step_ms": round(ms_gloop / args.steps, 4),
           "eager_prefill_ms": round(ms_prefill, 3), "eager_decode_step_ms": round(ms_step, 4),
           "launches": {"full": n_full, "cached": n_cached, "decode_step": n_step},
           "decode_roofline": {"bound": "hbm", "algorithmic_bytes": wbytes,
                               "achieved": round(wbytes / (ms_gloop / args.steps) / 1e6, 1), "peak": hbm, "unit": "GB/s",
                               "frac": round(wbytes / (ms_gloop / args.steps) / 1e6 / hbm, 3), "peak_source": src,
                               "timed": "whole DDIM loop replayed as one CUDA graph / steps"},
           "cached_vs_full_rel_diff": round(dev_rel, 5)}
    print(json.dumps(out), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/denoise_T{args.T}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
