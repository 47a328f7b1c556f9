"""Optimizer step and stream-overlap checks on the GPU.

* clip_grad_norm_ + AdamW (training/strategies/fsdp.py:242-257,:310; base_strategy_mla.py:372-379) as done by
  `DataParallelTrainer.step` (mla_sumsq_f32 / mla_clip_coef / mla_adamw_f32) against torch.optim.AdamW +
  torch.nn.utils.clip_grad_norm_ fed with the same gradients: fp32 masters within 1e-6, bf16 compute copies exact.
* The side-stream overlaps (weight-gradient GEMMs, AdamW) must not change a single bit of the gradients or the
  updated weights with respect to the single-stream order.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(seed=0, h=128, f=352, L=3, heads=4):
    from mla_b200 import llama
    torch.manual_seed(seed)
    m = llama.LlamaModel(64, h, f, L, heads, eps=1e-5).cuda()
    for p in m.parameters():
        torch.nn.init.normal_(p, std=0.05)
    for l in m.layers:
        l.input_layernorm.weight.data.add_(1.0)
        l.post_attention_layernorm.weight.data.add_(1.0)
    m.norm.weight.data.add_(1.0)
    m.embed_tokens.requires_grad_(False)
    return m


def _run(steps, wgrad_overlap, adam_overlap, level="none", B=2, S=150, h=128, wd=0.01, record=None):
    from mla_b200 import llama, trainer as T
    saved = (llama.OVERLAP["wgrad"], T.ADAM_OVERLAP["on"])
    llama.OVERLAP["wgrad"], T.ADAM_OVERLAP["on"] = wgrad_overlap, adam_overlap
    try:
        m = _model()
        m.set_save_levels(level)
        tr = T.DataParallelTrainer(m, lr=1e-2, weight_decay=wd, max_grad_norm=0.5)
        torch.manual_seed(3)
        xs = [(torch.randn(B * S, h, device="cuda") * 0.5).to(torch.bfloat16) for _ in range(steps)]
        gs = [torch.randn(B * S, h, device="cuda").to(torch.bfloat16) for _ in range(steps)]
        for x, g in zip(xs, gs):
            hs = m.run_layers(x.clone().requires_grad_(True), B, S, None)
            hs[-1].backward(g)
            if record is not None:
                torch.cuda.synchronize()
                record.append({n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None})
            tr.step()
        torch.cuda.synchronize()
        params = {n: p.detach().clone() for n, p in m.named_parameters()}
        copies = [tuple(c.clone() for c in l.compute_weights()) for l in m.layers]
        return m, params, copies, float(tr.grad_norm())
    finally:
        llama.OVERLAP["wgrad"], T.ADAM_OVERLAP["on"] = saved


def test_adamw_clip_matches_torch(cuda_lib):
    steps, wd = 3, 0.01
    grads = []
    m0 = _model()
    init = {n: p.detach().clone() for n, p in m0.named_parameters()}
    _, params, copies, gnorm = _run(steps, False, False, wd=wd, record=grads)
    # torch reference fed with the recorded gradients
    ref = {n: torch.nn.Parameter(v.clone()) for n, v in init.items() if n in grads[0]}
    decay = [p for n, p in ref.items() if p.ndim > 1]
    no_decay = [p for n, p in ref.items() if p.ndim <= 1]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": wd}, {"params": no_decay, "weight_decay": 0.0}],
                            lr=1e-2, betas=(0.9, 0.999), eps=1e-8)
    last_norm = None
    for g in grads:
        for n, p in ref.items():
            p.grad = g[n].clone()
        last_norm = torch.nn.utils.clip_grad_norm_(list(ref.values()), 0.5)
        opt.step()
    assert abs(gnorm - float(last_norm)) <= 1e-4 * float(last_norm)
    for n, p in ref.items():
        err = (params[n] - p.detach()).abs().max().item()
        assert err <= 2e-6 + 1e-5 * p.detach().abs().max().item(), (n, err)
    # parameters that never got a gradient are untouched
    for n, v in init.items():
        if n not in ref:
            assert torch.equal(params[n], v), n


def test_bf16_copies_track_masters(cuda_lib):
    m, params, copies, _ = _run(2, True, True)
    h, f = 128, 352
    for li, (wqkv, wo, wgu, wd_, l1, l2) in enumerate(copies):
        pre = f"layers.{li}."
        q, k, v = (params[pre + f"self_attn.{n}_proj.weight"] for n in "qkv")
        assert torch.equal(wqkv, torch.cat([q, k, v]).to(torch.bfloat16))
        assert torch.equal(wo, params[pre + "self_attn.o_proj.weight"].to(torch.bfloat16))
        assert torch.equal(wgu, torch.cat([params[pre + "mlp.gate_proj.weight"],
                                           params[pre + "mlp.up_proj.weight"]]).to(torch.bfloat16))
        assert torch.equal(wd_, params[pre + "mlp.down_proj.weight"].to(torch.bfloat16))
        assert torch.equal(l1, params[pre + "input_layernorm.weight"].to(torch.bfloat16))


@pytest.mark.parametrize("level", ["none", "layer"])
def test_stream_overlap_is_bit_identical(cuda_lib, level):
    """One step: every GEMM-produced gradient and every updated weight must be bit-identical with and without the side
    stream (same kernels, same order of accumulation); the RMSNorm weight gradients are accumulated with fp32 atomics
    (order-dependent in the last bit) and are compared to 1e-4.  Three steps: the trajectories stay together."""
    ga, gb = [], []
    _, pa, _, na = _run(1, False, False, level=level, record=ga)
    _, pb, _, nb = _run(1, True, True, level=level, record=gb)
    assert ga[0].keys() == gb[0].keys()
    for n in ga[0]:
        if "layernorm" in n or n == "norm.weight":
            assert torch.allclose(ga[0][n], gb[0][n], rtol=1e-4, atol=1e-6), n
        else:
            assert torch.equal(ga[0][n], gb[0][n]), n
    for n in pa:
        if not ("layernorm" in n or n == "norm.weight"):
            assert torch.allclose(pa[n], pb[n], rtol=0, atol=1e-7), n
    assert abs(na - nb) <= 1e-5 * abs(na)
    _, _, _, n3a = _run(3, False, False, level=level)
    _, _, _, n3b = _run(3, True, True, level=level)
    assert abs(n3a - n3b) <= 2e-3 * abs(n3a)


@pytest.mark.parametrize("accumulate", [False, True])
def test_gradient_norm_from_the_wgrad_epilogues(cuda_lib, accumulate):
    """Single replica: the global gradient norm clip_grad_norm_ needs is accumulated by the weight-gradient GEMM epilogues
    (sum of squares of the values they write) instead of a pass over every gradient — same norm, same update, also when
    a second micro-batch accumulates onto the first.  The update is compared after ONE step: a second step starts from
    fp32 masters that differ in the last bit (the two norms differ by ~1e-7 relative), which flips a few bf16 weight
    roundings and, through Adam's sign-like early updates, moves individual elements by the learning rate — a property
    of the optimiser, not of the norm; for the second step the norm alone is compared."""
    from mla_b200 import llama, trainer as T
    res = {}
    for fused in (False, True):
        llama.FUSE_GRAD_NORM["on"] = fused
        try:
            m = _model()
            m.set_save_levels("none")
            tr = T.DataParallelTrainer(m, lr=1e-2, weight_decay=0.0, max_grad_norm=0.5)
            torch.manual_seed(9)
            B, S, h = 2, 150, 128
            norms, params = [], None
            for step in range(2):
                for _mb in range(2 if accumulate else 1):
                    x = (torch.randn(B * S, h, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
                    g = torch.randn(B * S, h, device="cuda").to(torch.bfloat16)
                    m.run_layers(x, B, S, None)[-1].backward(g)
                assert all(l._gnorm2_valid == fused for l in m.layers)
                tr.step()
                torch.cuda.synchronize()
                norms.append(float(tr.grad_norm()))
                if step == 0:
                    params = {n: p.detach().clone() for n, p in m.named_parameters()}
            res[fused] = (norms, params)
        finally:
            llama.FUSE_GRAD_NORM["on"] = True
    (n0, p0), (n1, p1) = res[False], res[True]
    assert abs(n0[0] - n1[0]) <= 1e-5 * n0[0], (n0, n1)     # fp32 summation order (tile partials + atomics vs one pass)
    assert abs(n0[1] - n1[1]) <= 1e-3 * n0[1], (n0, n1)
    for k in p0:
        assert torch.allclose(p0[k], p1[k], rtol=0, atol=1e-6 + 1e-5 * float(p0[k].abs().max())), k
