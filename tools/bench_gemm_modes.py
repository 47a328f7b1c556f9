"""Per-shape timing of the 12 GEMMs of a decoder layer (17,536 tokens) with the one-CTA kernel and the CTA-pair kernel
in the same process (mla_gemm_set_mode), interleaved so that both see the same clocks.  -> gpurun_out/gemm_modes.json"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mla_b200 import _lib, ops  # noqa: E402

T, H, F = 17536, 4096, 11008
bf, dev = torch.bfloat16, "cuda"
x = torch.randn(T, H, device=dev).to(bf)
xf = torch.randn(T, F, device=dev).to(bf)
wqkv, wo = torch.randn(3 * H, H, device=dev).to(bf), torch.randn(H, H, device=dev).to(bf)
wgu, wd = torch.randn(2 * F, H, device=dev).to(bf), torch.randn(H, F, device=dev).to(bf)
dqkv, dgu = torch.randn(T, 3 * H, device=dev).to(bf), torch.randn(T, 2 * F, device=dev).to(bf)
g = [torch.empty_like(w, dtype=torch.float32) for w in (wqkv, wo, wgu, wd)]
names = ["fwd qkv", "fwd o", "fwd gate|up", "fwd down", "dgrad qkv", "dgrad o", "dgrad gate|up", "dgrad down",
         "wgrad qkv", "wgrad o", "wgrad gate|up", "wgrad down"]
calls = [
    lambda: ops.gemm(x, wqkv), lambda: ops.gemm(x, wo), lambda: ops.gemm(x, wgu), lambda: ops.gemm(xf, wd),
    lambda: ops.gemm(dqkv, wqkv, b_mn=True), lambda: ops.gemm(x, wo, b_mn=True),
    lambda: ops.gemm(dgu, wgu, b_mn=True), lambda: ops.gemm(x, wd, b_mn=True),
    lambda: ops.gemm(dqkv, x, a_mn=True, b_mn=True, out=g[0]), lambda: ops.gemm(x, x, a_mn=True, b_mn=True, out=g[1]),
    lambda: ops.gemm(dgu, x, a_mn=True, b_mn=True, out=g[2]), lambda: ops.gemm(x, xf, a_mn=True, b_mn=True, out=g[3]),
]
shapes = bench.GEMM_SHAPES(T)
lib = _lib.lib()
res = {}
for rnd in range(3):
    for mode in (0, 1):
        lib.mla_gemm_set_mode(C.c_int32(mode))
        for i, c in enumerate(calls):
            c()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(8):
                c()
            e1.record()
            torch.cuda.synchronize()
            res.setdefault((mode, i), []).append(e0.elapsed_time(e1) / 8)
out = {}
tot = {0: 0.0, 1: 0.0}
for i, n in enumerate(names):
    m, nn, k = shapes[i]
    fl = 2.0 * m * nn * k
    t0, t1 = min(res[(0, i)]), min(res[(1, i)])
    tot[0] += t0
    tot[1] += t1
    out[n] = {"one_cta_ms": round(t0, 4), "pair_ms": round(t1, 4), "one_cta_tflops": round(fl / t0 / 1e9, 1),
              "pair_tflops": round(fl / t1 / 1e9, 1)}
    print(f"{n:16s} 1cta {t0:7.4f} ms {fl / t0 / 1e9:7.1f} TF/s | pair {t1:7.4f} ms {fl / t1 / 1e9:7.1f} TF/s | x{t0 / t1:.3f}")
fl = sum(2.0 * m * n * k for m, n, k in shapes)
print(f"total            1cta {tot[0]:7.3f} ms {fl / tot[0] / 1e9:7.1f} | pair {tot[1]:7.3f} ms {fl / tot[1] / 1e9:7.1f}")
# sustained: the 12 shapes back to back for ~2 s each mode
for mode in (0, 1, 0, 1):
    lib.mla_gemm_set_mode(C.c_int32(mode))
    r = bench.gemm_roofline(T, 1.0, iters=20)
    print("sustained mode", mode, r["achieved"], "TFLOP/s")
    out[f"sustained_mode{mode}"] = r["achieved"]
lib.mla_gemm_set_mode(C.c_int32(0))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gemm_modes.json", "w"), indent=1)
