"""Rasterisation sweep of the CTA-pair GEMM: time each decoder-layer GEMM shape for several group_m values.
    python tools/bench_gemm_raster.py > gpurun_out/r02_gemm_raster.json"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mla_b200 import _lib, ops  # noqa: E402

H, F, T = 4096, 11008, 17536
bf = torch.bfloat16


def main():
    lib = _lib.lib()
    x = torch.randn(T, H, device="cuda").to(bf)
    xf = torch.randn(T, F, device="cuda").to(bf)
    wqkv, wgu, wd = (torch.randn(3 * H, H, device="cuda").to(bf), torch.randn(2 * F, H, device="cuda").to(bf),
                     torch.randn(H, F, device="cuda").to(bf))
    dqkv, dgu = torch.randn(T, 3 * H, device="cuda").to(bf), torch.randn(T, 2 * F, device="cuda").to(bf)
    g = [torch.empty_like(w, dtype=torch.float32) for w in (wqkv, wgu, wd)]
    shapes = {
        "fwd qkv (T,3h,h)": (lambda: ops.gemm(x, wqkv), 2.0 * T * 3 * H * H),
        "fwd gate|up (T,2f,h)": (lambda: ops.gemm(x, wgu), 2.0 * T * 2 * F * H),
        "fwd down (T,h,f)": (lambda: ops.gemm(xf, wd), 2.0 * T * H * F),
        "dgrad qkv (T,h,3h)": (lambda: ops.gemm(dqkv, wqkv, b_mn=True), 2.0 * T * 3 * H * H),
        "dgrad gate|up (T,h,2f)": (lambda: ops.gemm(dgu, wgu, b_mn=True), 2.0 * T * 2 * F * H),
        "dgrad down (T,f,h)": (lambda: ops.gemm(x, wd, b_mn=True), 2.0 * T * H * F),
        "wgrad qkv (3h,h,T)": (lambda: ops.gemm(dqkv, x, a_mn=True, b_mn=True, out=g[0]), 2.0 * T * 3 * H * H),
        "wgrad gate|up (2f,h,T)": (lambda: ops.gemm(dgu, x, a_mn=True, b_mn=True, out=g[1]), 2.0 * T * 2 * F * H),
        "wgrad down (h,f,T)": (lambda: ops.gemm(x, xf, a_mn=True, b_mn=True, out=g[2]), 2.0 * T * H * F),
    }
    res = {}
    gms = (0, 2, 3, 4, 6, 8, 12, 16, 24, 32, 69)
    # the part is power-capped and its clock wanders over ~100 ms: every setting is timed over 25 launches, the settings are
    # interleaved and the whole sweep is repeated 3 times; the median of the three is reported
    for _ in range(20):
        ops.gemm(x, wgu)                      # bring the chip to its sustained clocks first
    for name, (fn, fl) in shapes.items():
        samples = {gm: [] for gm in gms}
        for rnd in range(3):
            for gm in gms:
                lib.mla_gemm_set_group_m(C.c_int32(gm))
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(25):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                samples[gm].append(e0.elapsed_time(e1) / 25)
        row = {}
        for gm in gms:
            ms = sorted(samples[gm])[1]
            row["heuristic" if gm == 0 else gm] = {"ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1),
                                                   "spread": round(max(samples[gm]) / min(samples[gm]), 3)}
        lib.mla_gemm_set_group_m(C.c_int32(0))
        res[name] = row
        print(name, {k: v["tflops"] for k, v in row.items()}, file=sys.stderr, flush=True)
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
