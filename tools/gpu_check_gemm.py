"""GPU bring-up check for mla_gemm_bf16: all four operand layouts, tails, epilogues, timing.

Run on a B200:  python tools/gpu_check_gemm.py [--quick]    (writes gpurun_out/gemm_check.json)
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mla_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
results = []


def ref_gemm(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float() if b_mn else b.float().t()
    return A @ B


def check(name, M, N, K, a_mn, b_mn, **kw):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).mul_(0.5).bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).mul_(0.5).bfloat16()
    ref = ref_gemm(a, b, a_mn, b_mn)
    extra = {}
    out_dtype = kw.pop("out_dtype", torch.bfloat16)
    bias = res = pre = None
    if kw.get("bias"):
        bias = torch.randn(N, device=dev).bfloat16()
        ref = ref + bias.float()
    ref = ref * kw.get("alpha", 1.0) if not kw.get("bias") else ref
    if out_dtype == torch.bfloat16:
        ref = ref.bfloat16().float()
    pre_ref = ref.clone()
    act = kw.get("act", 0)
    if act == ops.ACT_RELU:
        ref = torch.relu(ref)
    elif act == ops.ACT_GELU_ERF:
        ref = torch.nn.functional.gelu(ref)
    elif act == ops.ACT_GELU_TANH:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    elif act == ops.ACT_SILU:
        ref = torch.nn.functional.silu(ref)
    if act:
        ref = ref.bfloat16().float()
    if kw.get("residual"):
        res = torch.randn(M, N, device=dev).bfloat16()
        ref = (ref + res.float()).bfloat16().float()
    if kw.get("pre_act"):
        pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out = None
    if kw.get("accumulate"):
        out = torch.randn(M, N, device=dev, dtype=torch.float32)
        ref = ref + out
    got = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out=out, out_dtype=out_dtype, bias=bias, act=act, residual=res,
                   pre_act=pre, alpha=kw.get("alpha", 1.0), accumulate=bool(kw.get("accumulate")))
    torch.cuda.synchronize()
    err = (got.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    rel = (got.float() - ref).norm().item() / (ref.norm().item() + 1e-30)
    ok = rel < 4e-3 and err <= 0.02 * scale + 1e-2
    if pre is not None:
        perr = (pre.float() - pre_ref).abs().max().item()
        extra["pre_act_err"] = perr
        ok = ok and perr <= 0.02 * scale + 1e-2
    rec = dict(name=name, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, max_err=err, rel=rel, scale=scale, ok=bool(ok), **extra)
    results.append(rec)
    print(("PASS " if ok else "FAIL ") + json.dumps(rec), flush=True)
    return ok


def bench(name, M, N, K, a_mn, b_mn, out_dtype=torch.bfloat16, iters=20):
    a = torch.randn((K, M) if a_mn else (M, K), device=dev).bfloat16()
    b = torch.randn((K, N) if b_mn else (N, K), device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=out_dtype)
    for _ in range(3):
        ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS comparison (library baseline, not on the product path)
    A = a.t() if a_mn else a
    B = b if b_mn else b.t()
    for _ in range(3):
        torch.matmul(A, B)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        torch.matmul(A, B)
    e1.record()
    torch.cuda.synchronize()
    ms_ref = e0.elapsed_time(e1) / iters
    rec = dict(name=name, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, ms=ms, tflops=tf, cublas_ms=ms_ref,
               cublas_tflops=2.0 * M * N * K / ms_ref / 1e9)
    results.append(rec)
    print("BENCH " + json.dumps(rec), flush=True)


def main():
    quick = "--quick" in sys.argv
    ok = True
    # smallest case first: one tile, one k-block
    ok &= check("1tile_kk", 128, 256, 64, False, False)
    ok &= check("1tile_k4", 128, 256, 256, False, False)
    ok &= check("kk_multi", 512, 512, 1024, False, False)
    ok &= check("k_mn", 256, 512, 512, False, True)
    ok &= check("mn_k", 256, 512, 512, True, False)
    ok &= check("mn_mn", 256, 512, 512, True, True)
    ok &= check("tails_kk", 200, 328, 136, False, False)
    ok &= check("tails_mnmn", 200, 328, 136, True, True)
    ok &= check("tails_kmn", 77, 1000, 4104, False, True)
    ok &= check("many_tiles", 1280, 2304, 320, False, False)  # > 148 tiles exercises the persistent loop + both TMEM stages
    ok &= check("bias_gelu", 300, 512, 256, False, False, bias=True, act=ops.ACT_GELU_ERF, pre_act=True)
    ok &= check("bias_gelu_tanh", 300, 512, 256, False, False, bias=True, act=ops.ACT_GELU_TANH)
    ok &= check("silu_res", 300, 512, 256, False, False, act=ops.ACT_SILU, residual=True)
    ok &= check("relu", 300, 512, 256, False, False, bias=True, act=ops.ACT_RELU)
    ok &= check("f32_out", 256, 512, 512, True, True, out_dtype=torch.float32)
    ok &= check("f32_acc", 256, 520, 512, True, True, out_dtype=torch.float32, accumulate=True)
    ok &= check("odd_n", 64, 7 * 8 + 4, 128, False, False)
    if not quick and ok:
        T = 17536
        bench("qkv_fwd", T, 12288, 4096, False, False)
        bench("o_fwd", T, 4096, 4096, False, False)
        bench("gateup_fwd", T, 22016, 4096, False, False)
        bench("down_fwd", T, 4096, 11008, False, False)
        bench("down_dgrad", T, 11008, 4096, False, True)
        bench("qkv_wgrad", 12288, 4096, T, True, True, out_dtype=torch.float32)
        bench("square8k", 8192, 8192, 8192, False, False)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gemm_check.json", "w") as f:
        json.dump(dict(ok=bool(ok), results=results), f, indent=1)
    print("ALL_OK" if ok else "SOME_FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
