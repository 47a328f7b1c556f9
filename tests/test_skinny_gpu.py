"""Skinny linears of the denoise step on the tensor cores (csrc/skinny_sm100.cu, mla_skinny_gemm): swap-AB tcgen05 GEMM
with the RMSNorm / SwiGLU prologue written into the B operand's swizzled layout and a fixed-order split-K finish.
Against an fp32 reference with the reference's rounding points (bf16 after the norm, after the linear, after the residual
add; modeling_llama.py:85-90,:240) and against the CUDA-core kernel it replaces; ragged N (not a multiple of 128), K
not a multiple of 64, 1..32 rows; repeated launches bit-identical (the split-K sum does not depend on arrival order)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _ref(x, w, residual, norm, swiglu):
    xf = x.float()
    if norm is not None:
        lw, eps = norm
        xf = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).bfloat16().float() * lw.float()
        xf = xf.bfloat16().float()
    if swiglu:
        k = xf.shape[1] // 2
        g, u = xf[:, :k], xf[:, k:]
        xf = (torch.nn.functional.silu(g).bfloat16().float() * u).bfloat16().float()
    y = (xf @ w.float().t()).bfloat16().float()
    if residual is not None:
        y = (y + residual.float()).bfloat16().float()
    return y


@pytest.mark.parametrize("m,n,k,mode", [
    (1, 256, 128, "plain"), (2, 200, 200, "norm"), (2, 768, 256, "norm"), (5, 384, 704, "swiglu"),
    (16, 1000, 520, "plain"), (17, 512, 1024, "norm"), (32, 640, 384, "swiglu"),
    (2, 12288, 4096, "norm"), (2, 4096, 4096, "plain"), (2, 22016, 4096, "norm"), (2, 4096, 11008, "swiglu"),
    (17, 4096, 11008, "swiglu"), (17, 12288, 4096, "norm"),
])
def test_skinny_gemm_matches_reference(cuda_lib, monkeypatch, m, n, k, mode):
    from mla_b200 import ops
    monkeypatch.setitem(ops.SKINNY, "on", True)
    torch.manual_seed(m * 1000 + n + k)
    w = (torch.randn(n, k, device="cuda") * k ** -0.5).bfloat16()
    x = torch.randn(m, 2 * k if mode == "swiglu" else k, device="cuda").bfloat16()
    res = torch.randn(m, n, device="cuda").bfloat16() if mode != "norm" else None
    norm = ((1 + 0.1 * torch.randn(k, device="cuda")).bfloat16(), 1e-5) if mode == "norm" else None
    assert ops.SKINNY["on"]
    got = ops.gemv(x, w, residual=res, norm=norm, swiglu=mode == "swiglu")
    want = _ref(x, w, res, norm, mode == "swiglu")
    assert rel_err(got, want) < 4e-3, rel_err(got, want)
    again = ops.gemv(x, w, residual=res, norm=norm, swiglu=mode == "swiglu")
    assert torch.equal(got, again)
    if m <= 16:
        ops.SKINNY["on"] = False
        try:
            old = ops.gemv(x, w, residual=res, norm=norm, swiglu=mode == "swiglu")
        finally:
            ops.SKINNY["on"] = True
        assert rel_err(got, old) < 4e-3


def test_skinny_gemm_chain_under_pdl(cuda_lib, monkeypatch):
    """Back-to-back launches (programmatic dependent launch: the next kernel's producer starts while this one drains) on
    shared workspaces: a 3-linear chain repeated, identical every time and equal to the launch-by-launch result."""
    from mla_b200 import ops
    monkeypatch.setitem(ops.SKINNY, "on", True)
    torch.manual_seed(3)
    h, f, m = 1024, 2816, 2
    w1 = (torch.randn(2 * f, h, device="cuda") * h ** -0.5).bfloat16()
    w2 = (torch.randn(h, f, device="cuda") * f ** -0.5).bfloat16()
    w3 = (torch.randn(h, h, device="cuda") * h ** -0.5).bfloat16()
    lw = (1 + 0.1 * torch.randn(h, device="cuda")).bfloat16()
    x = torch.randn(m, h, device="cuda").bfloat16()

    def chain(sync):
        gu = ops.gemv(x, w1, norm=(lw, 1e-5))
        if sync:
            torch.cuda.synchronize()
        y = ops.gemv(gu, w2, residual=x, swiglu=True)
        if sync:
            torch.cuda.synchronize()
        return ops.gemv(y, w3, residual=y)
    want = chain(True)
    for _ in range(5):
        assert torch.equal(chain(False), want)
