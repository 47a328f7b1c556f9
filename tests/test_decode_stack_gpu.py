"""The whole decoder stack of a denoise step as ONE persistent launch (csrc/decode_stack.cu, mla_decode_stack) against
the per-op path it replaces (LlamaDecoderLayer.decode: gemv_fused / decode_attn_rope / gemv_fused x3 per layer).
The kernel keeps every thread -> k-chunk mapping, reduction tree and rounding point of the per-op kernels (attention =
the split-K variant), so the comparison is BITWISE; against the default (un-split) attention it is within bf16 noise.
Also: re-launch without re-zeroing the workspace (self re-arming barrier / counters), capture into a CUDA graph
(cooperative launch inside a capture), one-row and two-sample shapes."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _model(h, f, L, H, seed=0):
    from mla_b200.llama import LlamaModel
    torch.manual_seed(seed)
    m = LlamaModel(64, h, f, L, H).cuda().eval()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            else:
                p.copy_(torch.randn_like(p) * (p.shape[1] ** -0.5))
    return m


def _per_op(model, x, caches, B, P, n, split_k):
    """LlamaModel.decode without the final norm, attention split-K on request."""
    from mla_b200 import ops
    cos, sin = model.rope_tables(P + n, x.device)
    cs, sn = cos[P:P + n].contiguous(), sin[P:P + n].contiguous()
    H = model.heads
    D = model.hidden_size // H
    for layer, cache in zip(model.layers, caches):
        wqkv, wo, wgu, wd, l1, l2 = layer.compute_weights()
        qkv = ops.gemv(x, wqkv, norm=(l1, model.eps))
        ctx = ops.decode_attn_rope(qkv, cache, cs, sn, B, H, n, P, D, split_k=split_k)
        x_mid = ops.gemv(ctx, wo, residual=x)
        gu = ops.gemv(x_mid, wgu, norm=(l2, model.eps))
        x = ops.gemv(gu, wd, residual=x_mid, swiglu=True)
    return x


def _stack(model, x, caches, B, P, n):
    cos, sin = model.rope_tables(P + n, x.device)
    return model._decode_stack(x, caches, B, P, n, cos[P:P + n].contiguous(), sin[P:P + n].contiguous())


def _caches(model, B, P, scale=1.0):
    D = model.hidden_size // model.heads
    return [(torch.randn(B, 2, model.heads, P, D, device="cuda") * scale).bfloat16() for _ in model.layers]


@pytest.mark.parametrize("h,f,L,H,B,n,P", [
    (256, 512, 3, 4, 1, 2, 37),          # head_dim 64, one attention split
    (256, 704, 2, 8, 1, 2, 300),         # head_dim 32, three splits
    (512, 1024, 2, 4, 1, 1, 129),        # head_dim 128, one row (MB = 1), split boundary at 128 | 2
    (512, 1024, 2, 4, 2, 1, 200),        # two samples of one row
    (4096, 11008, 2, 32, 1, 2, 545),     # Llama-2-7B width: K = 11008 takes the 3-chunks-per-thread path
])
def test_stack_equals_per_op_path_bitwise(cuda_lib, h, f, L, H, B, n, P):
    model = _model(h, f, L, H, seed=h + P)
    caches = _caches(model, B, P)
    torch.manual_seed(1)
    x = torch.randn(B * n, h, device="cuda").bfloat16()
    x0 = x.clone()
    want = _per_op(model, x, caches, B, P, n, split_k=True)
    got = _stack(model, x, caches, B, P, n)
    torch.cuda.synchronize()
    assert torch.equal(x, x0)                                   # the input is not modified
    assert torch.equal(got, want), rel_err(got, want)
    plain = _per_op(model, x, caches, B, P, n, split_k=False)   # default per-op attention: same math, other merge order
    assert rel_err(got, plain) < 1e-2
    # second launch on the same workspace: the arrival counters and the grid barrier re-armed themselves
    again = _stack(model, x, caches, B, P, n)
    assert torch.equal(again, want)


def test_stack_in_cuda_graph_and_model_decode(cuda_lib):
    from mla_b200 import llama
    h, f, L, H, B, n, P = 512, 1024, 4, 4, 1, 2, 150
    model = _model(h, f, L, H, seed=5)
    caches = _caches(model, B, P)
    x = torch.randn(B * n, h, device="cuda").bfloat16()
    assert llama.DECODE_STACK
    eager = model.decode(x, caches, B, P, n)                    # builds the pointer table / workspace (never in a capture)
    llama.DECODE_STACK = False
    try:
        per_op = model.decode(x, caches, B, P, n)
    finally:
        llama.DECODE_STACK = True
    assert rel_err(eager, per_op) < 1e-2
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        model.decode(x, caches, B, P, n)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = model.decode(x, caches, B, P, n)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, eager)
    # new activations through the same graph
    x.copy_(torch.randn_like(x))
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, model.decode(x, caches, B, P, n))


def test_stack_rejects_more_than_two_rows(cuda_lib):
    from mla_b200 import _lib, ops
    assert not ops.decode_stack_supported(3, 4096, 11008, 128)
    model = _model(256, 512, 1, 4)
    caches = _caches(model, 1, 20)
    x = torch.randn(3, 256, device="cuda").bfloat16()
    with pytest.raises(_lib.MlaError):
        _stack(model, x, caches, 1, 20, 3)
