"""Isolated timing of the HBM-bound kernels of the training step at the Llama-2-7B benchmark shapes
(17,536 tokens x h=4096, f=11008): achieved GB/s of ALGORITHMIC bytes against the measured HBM peak.

    python tools/bench_kernels.py        -> gpurun_out/kernels.json (+ one line per kernel on stdout)

CUDA events on the launching stream, 3 warm-ups, 20 timed launches over rotating buffers larger than L2 (126 MB).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mla_b200 import ops  # noqa: E402

T, H, F = 17536, 4096, 11008
dev = "cuda"
bf = torch.bfloat16


ONCE = os.environ.get("ONCE") == "1"      # one launch per kernel (for an ncu --set full capture)


def timeit(fn, n=20, warm=3):
    if ONCE:
        n, warm = 1, 0
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    hbm, _, tf_sust, src = bench.peaks()
    out = {"hbm_peak_gbs": hbm, "peak_source": src, "kernels": {}}
    R = 3   # rotating copies: 3 x (>=143 MB) > L2

    def rec(name, ms, nbytes):
        gbs = nbytes / ms / 1e6
        out["kernels"][name] = {"ms": round(ms, 4), "algorithmic_bytes": nbytes, "gbs": round(gbs, 1),
                                "frac_of_hbm_peak": round(gbs / hbm, 3)}
        print(f"{name:22s} {ms:8.4f} ms  {gbs:8.1f} GB/s  {gbs / hbm:6.3f} of peak", flush=True)

    xs = [torch.randn(T, H, device=dev).to(bf) for _ in range(R)]
    dys = [torch.randn(T, H, device=dev).to(bf) for _ in range(R)]
    w = torch.ones(H, device=dev).to(bf)
    outs = [torch.empty_like(x) for x in xs]
    rec("rmsnorm_fwd", timeit(lambda i: ops.rmsnorm_fwd(xs[i % R], w, 1e-5, out=outs[i % R])), 2 * T * H * 2)
    dw = torch.zeros(H, dtype=torch.float32, device=dev)
    rec("rmsnorm_bwd(+dres)", timeit(lambda i: ops.rmsnorm_bwd(dys[i % R], xs[i % R], w, 1e-5, dres=outs[i % R], dw=dw)),
        4 * T * H * 2)
    qkvs = [torch.randn(T, 3 * H, device=dev).to(bf) for _ in range(R)]
    S = 548
    cos = torch.randn(S, 64, device=dev).to(bf)
    sin = torch.randn(S, 64, device=dev).to(bf)
    rec("rope(q,k in place)", timeit(lambda i: ops.rope_(qkvs[i % R], 0, 64, 128, S, cos, sin)), 2 * T * 2 * H * 2)
    del qkvs
    gus = [torch.randn(T, 2 * F, device=dev).to(bf) for _ in range(R)]
    rec("swiglu_fwd", timeit(lambda i: ops.swiglu_fwd(gus[i % R])), 3 * T * F * 2)
    da = [torch.randn(T, F, device=dev).to(bf) for _ in range(R)]
    rec("swiglu_bwd", timeit(lambda i: ops.swiglu_bwd(da[i % R], gus[i % R])), 5 * T * F * 2)
    del gus, da
    # AdamW on one decoder layer's largest tensor ([2f,h] fp32 master + grad + m + v, bf16 copy): 30 B / parameter
    from mla_b200 import _lib
    import ctypes as C
    n = 2 * F * H
    p, g, m, v = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
    v.abs_()
    pb = torch.empty(n, dtype=bf, device=dev)
    scale = torch.ones(2, device=dev)

    def adam(i):
        _lib.check(_lib.lib().mla_adamw_f32(ops._p(p), ops._p(g), ops._p(m), ops._p(v), ops._p(pb), C.c_int64(n),
                                            C.c_float(1e-5), C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8),
                                            C.c_float(0.0), C.c_int64(1 + i), ops._p(scale), ops._stream()))
    rec("adamw (90M params)", timeit(adam), 30 * n)
    ssq = torch.zeros(1, device=dev)
    rec("sumsq (90M grads)", timeit(lambda i: _lib.check(_lib.lib().mla_sumsq_f32(ops._p(g), C.c_int64(n), ops._p(ssq),
                                                                                   ops._stream()))), 4 * n)
    del p, g, m, v, pb
    torch.cuda.empty_cache()
    # attention at the benchmark shape (tensor-bound; reported in TFLOP/s of causal algorithmic FLOPs)
    B = 32
    qkv = (torch.randn(T, 3 * H, device=dev) * 0.5).to(bf)
    ctx, lse = ops.attn_fwd(qkv, B, S, 32, 128)
    dctx = torch.randn_like(ctx)
    fl_f = 4.0 * B * 32 * S * S * 128 / 2
    ms = timeit(lambda i: ops.attn_fwd(qkv, B, S, 32, 128))
    out["kernels"]["attn_fwd_sm100"] = {"ms": round(ms, 4), "tflops": round(fl_f / ms / 1e9, 1)}
    print(f"attn_fwd_sm100         {ms:8.4f} ms  {fl_f / ms / 1e9:8.1f} TFLOP/s (causal algorithmic)", flush=True)
    ms = timeit(lambda i: ops.attn_bwd(dctx, qkv, ctx, lse, B, S, 32, 128))
    out["kernels"]["attn_bwd_sm100"] = {"ms": round(ms, 4), "tflops": round(2.5 * fl_f / ms / 1e9, 1)}
    print(f"attn_bwd_sm100         {ms:8.4f} ms  {2.5 * fl_f / ms / 1e9:8.1f} TFLOP/s (causal algorithmic)", flush=True)
    del qkv, ctx, dctx
    torch.cuda.empty_cache()
    if ONCE:
        return
    roof = bench.gemm_roofline(T, tf_sust)
    out["gemm"] = roof
    print("gemm 12 shapes:", json.dumps(roof), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    tag = os.environ.get("TAG", "")
    with open(f"gpurun_out/kernels{tag}.json", "w") as fo:
        json.dump(out, fo, indent=1)


if __name__ == "__main__":
    main()
