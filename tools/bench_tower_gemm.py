"""The point tokenizer's LGA GEMMs (1.3 M rows x 96..384 columns: HBM-bound shapes) through the two tcgen05 GEMM kernels:
ms per call and achieved GB/s of the algorithmic bytes (A + C; the weights are L2-resident)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mla_b200 import _lib, ops  # noqa: E402


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    lib = _lib.lib()
    out = {}
    shapes = [(1327104, 96, 192), (1327104, 192, 96), (663552, 192, 384), (663552, 384, 192)]
    for M, N, K in shapes:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = torch.randn(N, K, device="cuda").bfloat16()
        bias = torch.randn(N, device="cuda").bfloat16()
        c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        rec = {}
        for mode, name in ((0, "one_cta"), (2, "pair")):
            lib.mla_gemm_set_mode(C.c_int32(mode))
            ms = timeit(lambda: ops.gemm(a, w, bias=bias, out=c))
            rec[name] = {"ms": round(ms, 4), "GBs": round((M * K + M * N) * 2 / ms / 1e6, 1)}
        lib.mla_gemm_set_mode(C.c_int32(1))
        ms = timeit(lambda: torch.addmm(bias, a, w.t(), out=c))
        rec["cublas"] = {"ms": round(ms, 4), "GBs": round((M * K + M * N) * 2 / ms / 1e6, 1)}
        out[f"{M}x{N}x{K}"] = rec
        print(M, N, K, rec, flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/r02_tower_gemm.json", "w"), indent=1)


if __name__ == "__main__":
    main()
