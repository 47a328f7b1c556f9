"""GPU parity of the individual kernels against the oracle (oracle/llama.py), called through the C ABI."""
import math

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("rows,h", [(4500, 4096), (5000, 256), (4097, 1024)])
def test_rmsnorm_fwd_warp_per_row_matches_block_kernel(cuda_lib, rows, h):
    """From 4096 rows on the forward norm runs one warp per row (no block barrier, more loads in flight): same result as
    the block-per-row kernel that serves smaller calls, up to the fp32 summation order of the mean square (a last-ulp
    difference in rstd flips a bf16 rounding in a handful of elements)."""
    from mla_b200 import ops
    torch.manual_seed(rows)
    x = _bf(torch.randn(rows, h, device="cuda") * 2)
    w = _bf(1 + 0.1 * torch.randn(h, device="cuda"))
    y = ops.rmsnorm_fwd(x, w, 1e-5)
    half = rows // 2                      # < 4096 rows per call -> block-per-row kernel
    y_ref = torch.cat([ops.rmsnorm_fwd(x[:half].contiguous(), w, 1e-5), ops.rmsnorm_fwd(x[half:].contiguous(), w, 1e-5)])
    assert (y != y_ref).float().mean().item() < 2e-3
    assert rel_err(y, y_ref) < 3e-4
    from oracle import llama as O
    assert rel_err(y, O.rmsnorm(x, w, 1e-5)) < 2e-3


@pytest.mark.parametrize("rows,h", [(7, 128), (300, 4096), (33, 1024)])
def test_rmsnorm_fwd_bwd(cuda_lib, rows, h):
    from mla_b200 import ops
    from oracle import llama as O
    torch.manual_seed(1)
    x = _bf(torch.randn(rows, h, device="cuda") * 2)
    w = _bf(1 + 0.1 * torch.randn(h, device="cuda"))
    y = ops.rmsnorm_fwd(x, w, 1e-5)
    y_ref = O.rmsnorm(x, w, 1e-5)
    assert torch.equal(y, y_ref) or rel_err(y, y_ref) < 2e-3, rel_err(y, y_ref)
    # backward against fp32 autograd of the oracle
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    dy = _bf(torch.randn(rows, h, device="cuda"))
    O.rmsnorm(xf, wf, 1e-5).backward(dy.float())
    dres = _bf(torch.randn(rows, h, device="cuda"))
    dw = torch.zeros(h, device="cuda")
    dx = ops.rmsnorm_bwd(dy, x, w, 1e-5, dres=dres, dw=dw)
    assert rel_err(dx, xf.grad + dres.float()) < 6e-3
    assert rel_err(dw, wf.grad) < 6e-3


@pytest.mark.parametrize("rows,h,with_res", [(4096 + 37, 4096, True), (3000, 2048, False), (17536, 4096, True)])
def test_rmsnorm_bwd_staged_kernel_matches_register_kernel(cuda_lib, rows, h, with_res):
    """Large calls stage x / dy / dres through shared memory with cp.async (4 rows ahead per CTA); the arithmetic per row is
    the register-prefetch kernel's, so dx is bit-identical to what small calls (below the row threshold) produce, and the
    weight gradient agrees with fp32 autograd."""
    from mla_b200 import ops
    from oracle import llama as O
    torch.manual_seed(rows)
    x = _bf(torch.randn(rows, h, device="cuda") * 2)
    w = _bf(1 + 0.1 * torch.randn(h, device="cuda"))
    dy = _bf(torch.randn(rows, h, device="cuda"))
    dres = _bf(torch.randn(rows, h, device="cuda")) if with_res else None
    dw = torch.zeros(h, device="cuda")
    dx = ops.rmsnorm_bwd(dy, x, w, 1e-5, dres=dres, dw=dw)
    k = 500                                # < 4 rows per CTA: the register-prefetch kernel
    for lo in (0, rows - k):
        sl = slice(lo, lo + k)
        dx_small = ops.rmsnorm_bwd(dy[sl].contiguous(), x[sl].contiguous(), w, 1e-5,
                                   dres=dres[sl].contiguous() if with_res else None, dw=torch.zeros(h, device="cuda"))
        assert torch.equal(dx[sl], dx_small)
    xf = x[:2048].float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    O.rmsnorm(xf, wf, 1e-5).backward(dy[:2048].float())
    assert rel_err(dx[:2048], xf.grad + (dres[:2048].float() if with_res else 0)) < 6e-3
    dw2 = torch.zeros(h, device="cuda")
    ops.rmsnorm_bwd(dy[:2048].contiguous(), x[:2048].contiguous(), w, 1e-5, dw=dw2)
    assert rel_err(dw2, wf.grad) < 6e-3
    # whole-call weight gradient against a chunked fp32 evaluation
    n = (x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + 1e-5)).bfloat16().float()
    assert rel_err(dw, (dy.float() * n).sum(0)) < 2e-3


@pytest.mark.parametrize("B,S,H,D", [(2, 44, 4, 32), (2, 70, 2, 128)])
def test_rope_matches_reference_rounding(cuda_lib, B, S, H, D):
    from mla_b200 import ops
    from oracle import llama as O
    torch.manual_seed(2)
    h = H * D
    qkv = _bf(torch.randn(B * S, 3 * h, device="cuda"))
    cos, sin = O.rope_tables(S, D, 10000.0, torch.bfloat16)
    cos, sin = cos.cuda(), sin.cuda()
    q = qkv[:, :h].view(B, S, H, D).transpose(1, 2)
    k = qkv[:, h:2 * h].view(B, S, H, D).transpose(1, 2)
    q_ref, k_ref = O.apply_rope(q, k, cos, sin)
    buf = qkv.clone()
    ops.rope_(buf, 0, 2 * H, D, S, cos[:, :D // 2].contiguous(), sin[:, :D // 2].contiguous())
    q_got = buf[:, :h].view(B, S, H, D).transpose(1, 2)
    k_got = buf[:, h:2 * h].view(B, S, H, D).transpose(1, 2)
    assert torch.equal(q_got, q_ref), rel_err(q_got, q_ref)   # bit-exact: same rounding points
    assert torch.equal(k_got, k_ref)
    assert torch.equal(buf[:, 2 * h:], qkv[:, 2 * h:])          # v untouched
    # transpose(rotation) o rotation ~ identity
    ops.rope_(buf, 0, 2 * H, D, S, cos[:, :D // 2].contiguous(), sin[:, :D // 2].contiguous(), transpose=True)
    assert rel_err(buf[:, :2 * h], qkv[:, :2 * h]) < 1e-2


def test_swiglu_fwd_bwd(cuda_lib):
    from mla_b200 import ops
    torch.manual_seed(3)
    rows, f = 37, 352
    gu = _bf(torch.randn(rows, 2 * f, device="cuda") * 2)
    out = ops.swiglu_fwd(gu)
    g, u = gu[:, :f], gu[:, f:]
    ref = torch.nn.functional.silu(g) * u
    assert torch.equal(out, ref), rel_err(out, ref)
    gf = gu.float().requires_grad_(True)
    (torch.nn.functional.silu(gf[:, :f]) * gf[:, f:]).backward(torch.ones(rows, f, device="cuda"))
    d = ops.swiglu_bwd(_bf(torch.ones(rows, f, device="cuda")), gu)
    assert rel_err(d, gf.grad) < 8e-3
    # fused backward + re-materialisation of act: same bits as the two separate kernels
    d2, act2 = ops.swiglu_bwd_act(_bf(torch.ones(rows, f, device="cuda")), gu)
    assert torch.equal(d2, d) and torch.equal(act2, out)


@pytest.mark.parametrize("B,S,H,D,masked", [(2, 44, 4, 32, False), (2, 150, 2, 128, False), (3, 131, 2, 64, True),
                                             (2, 548, 2, 128, True)])
def test_attention_fwd_bwd(cuda_lib, B, S, H, D, masked):
    from mla_b200 import ops
    from oracle import llama as O
    torch.manual_seed(4)
    h = H * D
    qkv = _bf(torch.randn(B * S, 3 * h, device="cuda"))
    mask = None
    if masked:
        mask = torch.ones(B, S, dtype=torch.bool, device="cuda")
        mask[0, S - 7:] = False          # right padding
        mask[1, S // 2: S // 2 + 3] = False  # a hole (general mask semantics)
    ctx, lse = ops.attn_fwd(qkv, B, S, H, D, mask)

    def split(t):
        return [t[:, i * h:(i + 1) * h].reshape(B, S, H, D).transpose(1, 2) for i in range(3)]
    qf = qkv.float().requires_grad_(True)
    q, k, v = split(qf)
    ref = O.attention(q, k, v, mask)              # fp32 truth
    ref_bf = O.attention(*split(qkv), mask)       # bf16 path
    got = ctx.view(B, S, h)
    assert rel_err(got, ref) < 1e-2, rel_err(got, ref)
    assert rel_err(got, ref_bf) < 6e-3, rel_err(got, ref_bf)
    if masked:
        assert got[0, S - 7:].abs().max() == 0
    dctx = _bf(torch.randn(B * S, h, device="cuda"))
    ref.backward(dctx.view(B, S, h).float())
    dqkv = ops.attn_bwd(dctx, qkv, ctx, lse, B, S, H, D, mask)
    for i, name in enumerate("qkv"):
        e = rel_err(dqkv[:, i * h:(i + 1) * h], qf.grad[:, i * h:(i + 1) * h])
        assert e < 1.5e-2, (name, e)


def test_gather_scatter_embedding(cuda_lib):
    from mla_b200 import ops
    torch.manual_seed(5)
    src = _bf(torch.randn(50, 64, device="cuda")).requires_grad_(True)
    idx = torch.randperm(50, device="cuda")[:30].to(torch.int32)
    idx[3] = -1
    out = ops.GatherRowsFn.apply(src, idx)
    ref = src.detach()[idx.clamp(min=0).long()]
    ref[3] = 0
    assert torch.equal(out, ref)
    out.backward(torch.ones_like(out))
    exp = torch.zeros(50, 64, device="cuda")
    exp[idx[idx >= 0].long()] = 1
    assert torch.equal(src.grad.float(), exp)
    w = torch.randn(40, 64, device="cuda", requires_grad=True)
    ids = torch.tensor([[1, 2, 2, 39]], device="cuda")
    e = ops.EmbeddingFn.apply(ids, w, None)
    assert torch.equal(e, _bf(w.detach())[ids.view(-1)])
    e.backward(torch.ones_like(e))
    assert w.grad[2].eq(2).all() and w.grad[1].eq(1).all() and w.grad[0].eq(0).all()


@pytest.mark.parametrize("K,N,act", [(64, 128, 0), (7, 128, 3), (128, 7, 3), (256, 64, 4), (12, 64, 2)])
def test_linear_fn(cuda_lib, K, N, act):
    from mla_b200 import ops
    torch.manual_seed(6)
    M = 45
    x = _bf(torch.randn(M, K, device="cuda")).requires_grad_(True)
    w = (torch.randn(N, K, device="cuda") * 0.2).requires_grad_(True)
    b = (torch.randn(N, device="cuda") * 0.2).requires_grad_(True)
    y = ops.linear(x, w, b, act)
    xf, wf, bf_ = x.detach().float().requires_grad_(True), _bf(w.detach()).float().requires_grad_(True), _bf(b.detach()).float().requires_grad_(True)
    pre = torch.nn.functional.linear(xf, wf, bf_)
    F = torch.nn.functional
    ref = {0: lambda t: t, 1: F.relu, 2: F.gelu, 3: lambda t: F.gelu(t, approximate="tanh"), 4: F.silu}[act](pre)
    assert rel_err(y, ref) < 6e-3
    dy = _bf(torch.randn(M, N, device="cuda"))
    y.backward(dy)
    ref.backward(dy.float())
    assert rel_err(x.grad, xf.grad) < 1.2e-2
    assert rel_err(w.grad, wf.grad) < 1.2e-2
    assert rel_err(b.grad, bf_.grad) < 1.2e-2


def test_mse_and_qsample(cuda_lib):
    from mla_b200 import ops
    torch.manual_seed(7)
    pred = _bf(torch.randn(32, 1, 7, device="cuda")).requires_grad_(True)
    tgt = torch.randn(32, 1, 7, device="cuda")
    loss = ops.MSEFn.apply(pred, tgt)
    ref = ((pred.detach() - tgt) ** 2).mean()
    assert abs(loss.item() - ref.item()) < 1e-6 * max(1, abs(ref.item()))
    (loss * 3).backward()
    gref = 3 * 2 * (pred.detach().float() - tgt) / tgt.numel()
    assert rel_err(pred.grad, gref) < 5e-3
    a, n = torch.randn(32, 1, 7, device="cuda"), torch.randn(32, 1, 7, device="cuda")
    t = torch.randint(0, 100, (32,), device="cuda")
    sa, sb = torch.rand(100, device="cuda"), torch.rand(100, device="cuda")
    x = ops.q_sample(a, n, t, sa, sb)
    assert torch.allclose(x, sa[t].view(-1, 1, 1) * a + sb[t].view(-1, 1, 1) * n, atol=1e-6)


def test_action_tokenizer_bit_exact(cuda_lib):
    """ActionTokenizer bins (integer) must match numpy's digitize exactly, including edge values and out-of-range."""
    import numpy as np
    from mla_b200.action_tokenizer import ActionTokenizer
    from oracle import llama as O

    class Tok:
        vocab_size = 32000
    at = ActionTokenizer(Tok())
    rng = np.random.default_rng(0)
    edges = np.linspace(-1, 1, 256)
    x = np.concatenate([rng.uniform(-1.3, 1.3, 20000), edges, np.nextafter(edges, 2), np.nextafter(edges, -2),
                        [-1.0, 1.0, 0.0, -0.0, 5.0, -5.0]]).astype(np.float32)
    ids = at.encode_ids(x).cpu().numpy()
    assert np.array_equal(ids, O.action_tokenize(x, 32000))
    x64 = x.astype(np.float64) + 1e-12
    assert np.array_equal(at.encode_ids(x64).cpu().numpy(), O.action_tokenize(x64, 32000))
    back = at.decode_token_ids_to_actions(ids)
    assert np.array_equal(back, O.action_detokenize(ids, 32000))
    # round trip lands within half a bin
    inside = np.abs(x) <= 1
    assert np.max(np.abs(back[inside] - x[inside])) <= (2 / 255) + 1e-6


def test_cross_entropy_fwd_bwd(cuda_lib):
    from mla_b200 import ops
    torch.manual_seed(8)
    B, S, V = 3, 17, 1000
    logits = (torch.randn(B * S, V, device="cuda") * 3).to(torch.bfloat16).requires_grad_(True)
    labels = torch.randint(0, V, (B, S), device="cuda")
    labels[0, 3:6] = -100
    labels[2, :] = -100
    loss = ops.CrossEntropyFn.apply(logits, labels)
    lf = logits.detach().float().view(B, S, V).requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lf[:, :-1].reshape(-1, V), labels[:, 1:].reshape(-1))
    assert abs(loss.item() - ref.item()) < 1e-4 * abs(ref.item())
    (loss * 2).backward()
    (ref * 2).backward()
    assert rel_err(logits.grad.view(B, S, V), lf.grad) < 6e-3


def test_lm_loss_path(cuda_lib):
    """LlamaForCausalLM with compute_lm_loss: loss/logits as modeling_llama.py:1254-1269, lm_head receives a gradient."""
    from mla_b200.backbone import LlamaConfig, LlamaForCausalLM
    from oracle import llama as O
    torch.manual_seed(9)
    cfg = LlamaConfig(vocab_size=512, hidden_size=128, intermediate_size=352, num_hidden_layers=1, num_attention_heads=4,
                      pad_token_id=None)
    m = LlamaForCausalLM(cfg).cuda()
    m.compute_lm_loss = True
    ids = torch.randint(0, 512, (2, 12), device="cuda")
    labels = ids.clone()
    labels[0, :4] = -100
    out = m(input_ids=ids, labels=labels)
    hs_last = out.hidden_states[-1]
    w = m.lm_head.weight.detach().to(torch.bfloat16)
    ref_loss, ref_logits = O.lm_loss(hs_last.detach(), w, labels)
    assert abs(out.loss.item() - ref_loss.item()) < 2e-3 * abs(ref_loss.item())
    assert rel_err(out.logits, ref_logits) < 1e-2 and out.logits.dtype == torch.float32
    out.loss.backward()
    assert m.lm_head.weight.grad is not None and float(m.lm_head.weight.grad.abs().sum()) > 0


@pytest.mark.parametrize("M,valid_frac", [(8192, 1.0), (8192, 0.83), (1000, 0.5)])
def test_infonce_kernels_full_size(cuda_lib, M, valid_frac):
    """img<->pc InfoNCE at the size the training step reaches (M = B_eff * 256 = 8192 rows): similarity GEMM
    + masked row/column softmax statistics + both gradients, against torch fp32 autograd of the reference formula
    (contrastive.py:203-215: compaction of the valid rows, logits / 0.07, mean of the two cross-entropies) evaluated in
    the reference's arithmetic (bf16 logits) and in fp32 (truth)."""
    import torch.nn.functional as F
    from mla_b200.contrastive import _InfoNCEFn
    torch.manual_seed(M)
    D, T = 256, 0.07
    a = F.normalize(torch.randn(M, D, device="cuda"), dim=-1)
    b = F.normalize(a + 0.7 * F.normalize(torch.randn(M, D, device="cuda"), dim=-1), dim=-1)   # positives correlate
    valid = torch.rand(M, device="cuda") < valid_frac
    ab, bb = a.to(torch.bfloat16).requires_grad_(True), b.to(torch.bfloat16).requires_grad_(True)
    loss = _InfoNCEFn.apply(ab, bb, valid.to(torch.uint8).contiguous(), T)
    loss.backward()

    def ref(dt):
        x, y = ab.detach().float().requires_grad_(True), bb.detach().float().requires_grad_(True)
        xv, yv = x[valid], y[valid]
        logits = (torch.matmul(xv.to(dt), yv.to(dt).t()) / T).float()
        lab = torch.arange(xv.shape[0], device="cuda")
        l = (F.cross_entropy(logits, lab) + F.cross_entropy(logits.t(), lab)) / 2
        l.backward()
        return l.detach(), x.grad, y.grad
    l32, da32, db32 = ref(torch.float32)
    l16, da16, db16 = ref(torch.bfloat16)
    assert abs(float(loss) - float(l32)) <= 1.5 * abs(float(l16) - float(l32)) + 2e-3 * abs(float(l32)), (float(loss), float(l32), float(l16))
    for got, t32, t16 in ((ab.grad, da32, da16), (bb.grad, db32, db16)):
        e, e_ref = rel_err(got, t32), rel_err(t16, t32)
        assert e < 1.5 * e_ref + 1e-2, (e, e_ref)
        assert float(got[~valid].abs().max() if (~valid).any() else 0.0) == 0.0      # masked rows get no gradient
