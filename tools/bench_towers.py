"""Forward / forward+backward timing of the two tokenizers at the benchmark sizes (stage "pretrain" trains them):
image tokenizer on 8 distinct 672x672 images, point tokenizer on 32 clouds of 1024 points (K = 81 neighbours).

    python tools/bench_towers.py        -> gpurun_out/towers.json

CUDA events on the launching stream, 2 warm-ups, 5 timed iterations.  Also checks that every parameter the forward
reads receives a finite, non-zero gradient at these sizes (the tiny-size parity is tests/test_tower_bwd_gpu.py).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mla_b200 import _lib, pointcloud_impl  # noqa: E402
from mla_b200.pointcloud import PointTokenizer  # noqa: E402
from mla_b200.vision import VisionTokenizer  # noqa: E402


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


def grads_ok(params):
    bad = []
    for name, p in params:
        if p.grad is None or not torch.isfinite(p.grad).all() or float(p.grad.abs().max()) == 0.0:
            bad.append(name)
    return bad


def main():
    torch.manual_seed(0)
    out = {}
    vt = VisionTokenizer(1024).cuda().train()
    px = torch.randn(8, 4, 672, 672, device="cuda")
    px[:, 3] = 1.0

    def v_fwd():
        with torch.no_grad():
            vt.pooled_features(px)

    def v_fb():
        pooled, _, _ = vt.pooled_features(px)
        pooled.float().square().mean().backward()

    vt.requires_grad_(False)
    ms_f, n_f = timeit(v_fwd)
    vt.requires_grad_(True)
    ms_fb, n_fb = timeit(v_fb)
    used = [(n, p) for n, p in vt.named_parameters() if any(p is q for q in vt._tower_params())]
    out["vision_tokenizer"] = {"images": 8, "fwd_ms": round(ms_f, 3), "fwd_bwd_ms": round(ms_fb, 3),
                               "launches_fwd": n_f, "launches_fwd_bwd": n_fb, "params_with_bad_grad": grads_ok(used)}
    print(out["vision_tokenizer"], flush=True)
    del vt, px
    torch.cuda.empty_cache()

    pt = PointTokenizer().cuda().train()
    pc = torch.rand(32, 1024, 3, device="cuda")

    def p_fwd():
        with torch.no_grad():
            pt(pc)

    def p_fb():
        tok, _ = pt(pc)
        tok.float().square().mean().backward()

    pt.requires_grad_(False)
    ms_f, n_f = timeit(p_fwd)
    pt.requires_grad_(True)
    torch.cuda.reset_peak_memory_stats()
    ms_fb, n_fb = timeit(p_fb)
    tp = pointcloud_impl._tower_params(pt)
    used = [(n, p) for n, p in pt.named_parameters() if any(p is q for q in tp) and not n.endswith("net1.0.bias")
            and not n.endswith("net2.0.bias")]       # conv biases before a train-mode BN: gradient is exactly ~0
    out["point_tokenizer"] = {"clouds": 32, "points": 1024, "k": 81, "fwd_ms": round(ms_f, 3),
                              "fwd_bwd_ms": round(ms_fb, 3), "launches_fwd": n_f, "launches_fwd_bwd": n_fb,
                              "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
                              "params_with_bad_grad": grads_ok(used)}
    print(out["point_tokenizer"], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/towers.json", "w"), indent=1)


if __name__ == "__main__":
    main()
