"""The 12 GEMM launches of one decoder layer (forward, dgrad, wgrad at 17,536 tokens) inside a profiler range, for
    ncu --set full --profile-from-start off --clock-control none -k regex:gemm_bf16 python tools/ncu_gemm.py
One warm-up pass runs outside the range."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mla_b200 import ops  # noqa: E402

T, H, F = 17536, 4096, 11008
bf, dev = torch.bfloat16, "cuda"
x = torch.randn(T, H, device=dev).to(bf)
xf = torch.randn(T, F, device=dev).to(bf)
wqkv, wo = torch.randn(3 * H, H, device=dev).to(bf), torch.randn(H, H, device=dev).to(bf)
wgu, wd = torch.randn(2 * F, H, device=dev).to(bf), torch.randn(H, F, device=dev).to(bf)
dqkv, dgu = torch.randn(T, 3 * H, device=dev).to(bf), torch.randn(T, 2 * F, device=dev).to(bf)
g = [torch.empty_like(w, dtype=torch.float32) for w in (wqkv, wo, wgu, wd)]
calls = [
    lambda: ops.gemm(x, wqkv), lambda: ops.gemm(x, wo), lambda: ops.gemm(x, wgu), lambda: ops.gemm(xf, wd),
    lambda: ops.gemm(dqkv, wqkv, b_mn=True), lambda: ops.gemm(x, wo, b_mn=True),
    lambda: ops.gemm(dgu, wgu, b_mn=True), lambda: ops.gemm(x, wd, b_mn=True),
    lambda: ops.gemm(dqkv, x, a_mn=True, b_mn=True, out=g[0]), lambda: ops.gemm(x, x, a_mn=True, b_mn=True, out=g[1]),
    lambda: ops.gemm(dgu, x, a_mn=True, b_mn=True, out=g[2]), lambda: ops.gemm(x, xf, a_mn=True, b_mn=True, out=g[3]),
]
for c in calls:
    c()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for c in calls:
    c()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
