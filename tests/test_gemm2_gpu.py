"""The CTA-pair GEMM (tcgen05.mma.cta_group::2, gemm2_sm100.cu) against fp32 matmul: all four operand layouts, ragged
M / N / K tails (a half-empty 256-row pair tile, partial column tile, K not a multiple of 64), fused epilogues, fp32
accumulation — and bit-identical results to the one-CTA kernel (same MMA shape per row, same K order)."""
import ctypes as C

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture()
def pair_mode(cuda_lib):
    assert cuda_lib.mla_gemm_set_mode(C.c_int32(2)) == 0
    yield
    cuda_lib.mla_gemm_set_mode(C.c_int32(1))


def _mk(M, N, K, a_mn, b_mn):
    a = (torch.randn((K, M) if a_mn else (M, K), device="cuda") * 0.5).bfloat16()
    b = (torch.randn((K, N) if b_mn else (N, K), device="cuda") * 0.5).bfloat16()
    ref = (a.float().t() if a_mn else a.float()) @ (b.float() if b_mn else b.float().t())
    return a, b, ref


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (128, 256, 128), (384, 512, 192), (1000, 264, 520), (2048, 1024, 4096),
                                   (552, 4096, 352), (17536, 512, 256)])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
def test_pair_gemm_layouts(pair_mode, cuda_lib, M, N, K, a_mn, b_mn):
    from mla_b200 import ops
    torch.manual_seed(M + N + K)
    a, b, ref = _mk(M, N, K, a_mn, b_mn)
    got = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn)
    assert rel_err(got, ref) < 4e-3
    cuda_lib.mla_gemm_set_mode(C.c_int32(0))
    one = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn)
    cuda_lib.mla_gemm_set_mode(C.c_int32(2))
    assert torch.equal(got, one)


def test_pair_gemm_epilogues(pair_mode, cuda_lib):
    from mla_b200 import ops
    torch.manual_seed(5)
    M, N, K = 700, 520, 256
    a, b, ref = _mk(M, N, K, False, False)
    bias = torch.randn(N, device="cuda").bfloat16()
    res = torch.randn(M, N, device="cuda").bfloat16()
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    got = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU_TANH, residual=res, pre_act=pre)
    lin = (ref + bias.float()).bfloat16()
    want = (torch.nn.functional.gelu(lin.float(), approximate="tanh").bfloat16().float() + res.float()).bfloat16()
    assert rel_err(pre, lin) < 4e-3 and rel_err(got, want) < 6e-3
    # fp32 output with accumulation (weight gradients)
    a, b, ref = _mk(512, 384, 1096, True, True)
    out = torch.randn(512, 384, device="cuda")
    base = out.clone()
    ops.gemm(a, b, a_mn=True, b_mn=True, out=out, accumulate=True)
    assert rel_err(out - base, ref) < 1e-3
    ops.gemm(a, b, a_mn=True, b_mn=True, out=out, alpha=0.5)
    assert rel_err(out, 0.5 * ref) < 1e-3


@pytest.mark.parametrize("mode", [0, 2])
def test_dynamic_tile_claiming(cuda_lib, mode):
    """sched_ws: tiles claimed from a global counter (used under DDP, where NCCL kernels hold SMs) — same bits as the
    static order, and the counters re-arm themselves between launches (several launches back to back)."""
    from mla_b200 import ops
    torch.manual_seed(11)
    cuda_lib.mla_gemm_set_mode(C.c_int32(mode))
    try:
        for (M, N, K, a_mn, b_mn) in [(17536, 1024, 512, False, False), (4096, 2048, 17536, True, True),
                                      (5000, 3072, 1024, False, True)]:
            a, b, ref = _mk(M, N, K, a_mn, b_mn)
            ops.DYNAMIC_TILES["on"] = False
            want = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn)
            ops.DYNAMIC_TILES["on"] = True
            for _ in range(4):
                got = ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn)
                assert torch.equal(got, want)
            assert rel_err(got, ref) < 4e-3
    finally:
        ops.DYNAMIC_TILES["on"] = False
        cuda_lib.mla_gemm_set_mode(C.c_int32(1))


@pytest.mark.parametrize("mode", [0, 2])
def test_fused_rope_epilogue(cuda_lib, mode):
    """RoPE folded into the q|k|v projection's epilogue == projection followed by the in-place RoPE kernel, bit for bit
    (q and k heads rotated, v columns untouched; rows beyond one sequence wrap their position)."""
    from mla_b200 import ops
    torch.manual_seed(21)
    cuda_lib.mla_gemm_set_mode(C.c_int32(mode))
    try:
        B, S, H, D, h = 3, 150, 4, 128, 512
        x = (torch.randn(B * S, h, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(3 * h, h, device="cuda") * 0.05).bfloat16()
        ang = torch.rand(S, D // 2, device="cuda") * 6.28
        cos, sin = ang.cos().bfloat16().contiguous(), ang.sin().bfloat16().contiguous()
        want = ops.gemm(x, w)
        ops.rope_(want, 0, 2 * H, D, S, cos, sin)
        got = ops.gemm(x, w, rope=(cos, sin, S, 2 * H * D))
        assert torch.equal(got, want)
    finally:
        cuda_lib.mla_gemm_set_mode(C.c_int32(1))


@pytest.mark.parametrize("M,f,K", [(1024, 256, 512), (3000, 1408, 1024), (17536, 11008, 4096)])
def test_fused_swiglu_epilogue_matches_unfused(cuda_lib, M, f, K):
    """gate|up projection with SwiGLU in the CTA-pair epilogue == projection followed by swiglu_fwd, bit for bit."""
    import torch
    from mla_b200 import ops
    torch.manual_seed(M)
    x = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(2 * f, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    gu_ref = ops.gemm(x, w)
    act_ref = ops.swiglu_fwd(gu_ref)
    act = torch.empty((M, f), dtype=torch.bfloat16, device="cuda")
    gu = ops.gemm(x, w, swiglu_out=act)
    assert torch.equal(gu, gu_ref)
    assert torch.equal(act, act_ref)
    act2 = torch.empty_like(act)
    assert ops.gemm(x, w, swiglu_out=act2, store_c=False) is None
    assert torch.equal(act2, act_ref)


@pytest.mark.parametrize("M,f,K,mode", [(1024, 256, 512, 2), (3000, 1408, 1024, 2), (17536, 11008, 4096, 1), (300, 352, 128, 0)])
def test_fused_swiglu_backward_epilogue_matches_unfused(cuda_lib, M, f, K, mode):
    """d_act = dy . W_down with the SwiGLU backward in the epilogue == the GEMM followed by swiglu_bwd_act, bit for bit
    (d(gate|up) and the re-materialised act); both the CTA-pair and the one-CTA kernel."""
    from mla_b200 import ops
    torch.manual_seed(M + f)
    cuda_lib.mla_gemm_set_mode(C.c_int32(mode))
    try:
        dy = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
        wd = (torch.randn(K, f, device="cuda") * K ** -0.5).to(torch.bfloat16)        # W_down [h, f]: b_mn operand
        gu = torch.randn(M, 2 * f, device="cuda").to(torch.bfloat16)
        dact = ops.gemm(dy, wd, b_mn=True)
        dgu_ref, act_ref = ops.swiglu_bwd_act(dact, gu)
        dgu = torch.empty_like(gu)
        act = torch.empty((M, f), dtype=torch.bfloat16, device="cuda")
        assert ops.gemm(dy, wd, b_mn=True, swiglu_bwd=(gu, dgu, act)) is None
        torch.cuda.synchronize()
        assert torch.equal(dgu, dgu_ref)
        assert torch.equal(act, act_ref)
        dgu2 = torch.empty_like(gu)
        ops.gemm(dy, wd, b_mn=True, swiglu_bwd=(gu, dgu2, None))
        assert torch.equal(dgu2, dgu_ref)
    finally:
        cuda_lib.mla_gemm_set_mode(C.c_int32(1))


def test_tma_descriptor_cache(cuda_lib):
    """The same operands at the same addresses re-use their encoded TMA descriptors (no driver call on the launch path);
    an operand of a different shape at a re-used address gets its own descriptor."""
    from mla_b200 import _lib, ops
    torch.manual_seed(11)
    a, b, ref = _mk(512, 512, 256, False, False)
    out = torch.empty(512, 512, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, b, out=out)
    h0, m0 = _lib.tmap_cache_stats()
    first = out.clone()
    for _ in range(5):
        ops.gemm(a, b, out=out)
    h1, m1 = _lib.tmap_cache_stats()
    assert m1 == m0 and h1 >= h0 + 10, (h0, m0, h1, m1)
    assert torch.equal(out, first)
    # same base address, different logical shape: must not alias the cached descriptor
    a2 = a.view(-1)[: 256 * 256].view(256, 256)
    got = ops.gemm(a2, b[:, :256].contiguous())
    want = a2.float() @ b[:, :256].float().t()
    assert rel_err(got, want) < 4e-3
