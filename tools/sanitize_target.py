"""compute-sanitizer target: one small invocation of every hot-path kernel family (tcgen05 GEMM 1-CTA / CTA-pair with
each fused epilogue, tcgen05 attention forward + both backward generations, row kernels, one Tiny-MLA training step).

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py     # out-of-bounds / misaligned accesses
    compute-sanitizer --tool racecheck python tools/sanitize_target.py     # shared-memory hazards
    compute-sanitizer --tool initcheck python tools/sanitize_target.py     # reads of uninitialised global memory
(tools/sanitize.sh runs the three and keeps the logs; shapes are small because the sanitizer serialises every launch)
"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mla_b200 import _lib, llama, ops  # noqa: E402


def main():
    torch.manual_seed(0)
    lib = _lib.lib()
    bf = torch.bfloat16
    # ---- GEMMs: both kernels, every operand layout, fused epilogues
    for mode in (0, 2):
        lib.mla_gemm_set_mode(C.c_int32(mode))
        a = torch.randn(300, 256, device="cuda").to(bf)
        w = torch.randn(512, 256, device="cuda").to(bf)
        y = ops.gemm(a, w, residual=torch.randn(300, 512, device="cuda").to(bf))
        ops.gemm(y, w, b_mn=True)
        g = torch.zeros(512, 256, device="cuda")
        ops.gemm(y, a, a_mn=True, b_mn=True, out=g, accumulate=True)
        ang = torch.rand(150, 64, device="cuda")
        ops.gemm(a, torch.randn(768, 256, device="cuda").to(bf), rope=(ang.cos().to(bf), ang.sin().to(bf), 150, 512))
        gu = ops.gemm(a, w)                                                  # [300, 512] = gate | up, f = 256
        act = torch.empty(300, 256, device="cuda", dtype=bf)
        if mode == 2:
            ops.gemm(a, w, swiglu_out=act)
        wd = torch.randn(256, 256, device="cuda").to(bf)                     # W_down [h, f]
        dgu = torch.empty_like(gu)
        ops.gemm(a, wd, b_mn=True, swiglu_bwd=(gu, dgu, act))
    lib.mla_gemm_set_mode(C.c_int32(1))
    # ---- attention (head_dim 128): forward, backward generation 1, pipelined backward (smem and TMEM hand-over, RoPE fused)
    B, S, H, D = 2, 150, 2, 128
    qkv = torch.randn(B * S, 3 * H * D, device="cuda").to(bf)
    mask = torch.ones(B, S, dtype=torch.bool, device="cuda")
    mask[1, S - 7:] = False
    dctx = torch.randn(B * S, H * D, device="cuda").to(bf)
    ang = torch.rand(S, 64, device="cuda")
    cos, sin = ang.cos().to(bf).contiguous(), ang.sin().to(bf).contiguous()
    for m in (None, mask):
        ctx, lse = ops.attn_fwd(qkv, B, S, H, D, m)
        for impl, ts in (("sm100", 1), ("sm100v2", 0), ("sm100v2", 1)):
            ops.ATTN_IMPL["bwd"] = impl
            lib.mla_attn_bwd2_set_ts(C.c_int32(ts))
            ops.attn_bwd(dctx, qkv, ctx, lse, B, S, H, D, m, rope=(cos, sin) if impl == "sm100v2" else None)
    # ---- one decoder layer forward + backward at head_dim 128 (row kernels, fused epilogues in context)
    m = llama.LlamaModel(64, 256, 512, 2, 2, eps=1e-5).cuda()
    for p in m.parameters():
        torch.nn.init.normal_(p, std=0.05)
    x = torch.randn(B * S, 256, device="cuda").to(bf).requires_grad_(True)
    m.set_save_levels("mlp")
    hs = m.run_layers(x, B, S, mask)
    hs[-1].backward(torch.randn_like(hs[-1]))
    torch.cuda.synchronize()
    # ---- large-call row kernels (cp.async-staged RMSNorm backward) and the inference denoise step: per-op gemv path, the
    #      one-launch stack kernel and the tcgen05 skinny linears (both opt-in), at small sizes
    rows, hh = 1300, 1024
    ops.rmsnorm_bwd(torch.randn(rows, hh, device="cuda").to(bf), torch.randn(rows, hh, device="cuda").to(bf),
                    torch.ones(hh, device="cuda").to(bf), 1e-5, dres=torch.randn(rows, hh, device="cuda").to(bf),
                    dw=torch.zeros(hh, device="cuda"))
    dm = llama.LlamaModel(64, 256, 704, 2, 4, eps=1e-5).cuda().eval()
    P, n = 150, 2
    caches = [torch.randn(1, 2, 4, P, 64, device="cuda").to(bf) for _ in dm.layers]
    xr = torch.randn(n, 256, device="cuda").to(bf)
    with torch.no_grad():
        dm.decode(xr, caches, 1, P, n)
        llama.DECODE_STACK = True
        dm.decode(xr, caches, 1, P, n)
        llama.DECODE_STACK = False
        ops.SKINNY["on"] = True
        dm.decode(xr, caches, 1, P, n)
        dm.decode(torch.randn(17, 256, device="cuda").to(bf), caches, 1, P, 17)
        ops.SKINNY["on"] = False
    torch.cuda.synchronize()
    # ---- Tiny-MLA training step (tokenizers, splice, embedders, head, loss, optimizer)
    import __graft_entry__ as ge
    ge.smoke()
    torch.cuda.synchronize()
    print("sanitize target done")


if __name__ == "__main__":
    main()
