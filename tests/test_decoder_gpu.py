"""Decoder-stack parity on the GPU: our per-layer autograd node (CUDA kernels) vs the oracle's restatement of
modeling_llama.py, forward (all hidden states) and backward (input + every weight gradient)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _make(h, f, L, heads, vocab=64, seed=0):
    from mla_b200 import llama
    torch.manual_seed(seed)
    m = llama.LlamaModel(vocab, h, f, L, heads, eps=1e-5).cuda()
    for p in m.parameters():
        torch.nn.init.normal_(p, std=0.05)
    for l in m.layers:
        l.input_layernorm.weight.data.add_(1.0)
        l.post_attention_layernorm.weight.data.add_(1.0)
    m.norm.weight.data.add_(1.0)
    return m


def _oracle_params(m, dtype):
    layers = []
    for l in m.layers:
        a, mm = l.self_attn, l.mlp
        d = dict(q_proj=a.q_proj.weight, k_proj=a.k_proj.weight, v_proj=a.v_proj.weight, o_proj=a.o_proj.weight,
                 gate_proj=mm.gate_proj.weight, up_proj=mm.up_proj.weight, down_proj=mm.down_proj.weight,
                 ln1=l.input_layernorm.weight, ln2=l.post_attention_layernorm.weight)
        layers.append({k: v.detach().to(torch.bfloat16).to(dtype).requires_grad_(True) for k, v in d.items()})
    return layers, m.norm.weight.detach().to(torch.bfloat16).to(dtype).requires_grad_(True)


@pytest.mark.parametrize("h,f,L,heads,B,S,level,masked", [
    (128, 352, 2, 4, 2, 44, "layer", False),     # BASELINE configs[0] decoder shape (Tiny-MLA)
    (128, 352, 2, 4, 2, 44, "none", True),
    (256, 704, 3, 2, 2, 150, "mlp", True),       # head_dim 128 like Llama-2-7B
])
def test_decoder_forward_backward(cuda_lib, h, f, L, heads, B, S, level, masked):
    from oracle import llama as O
    m = _make(h, f, L, heads)
    m.set_save_levels(level)
    torch.manual_seed(10)
    x = (torch.randn(B * S, h, device="cuda") * 0.5).to(torch.bfloat16).requires_grad_(True)
    mask = None
    if masked:
        mask = torch.ones(B, S, dtype=torch.bool, device="cuda")
        mask[1, S - 5:] = False
    hs = m.run_layers(x, B, S, mask)
    # oracle in bf16 (the reference's arithmetic) and fp32 (truth)
    lb, nb = _oracle_params(m, torch.bfloat16)
    lf, nf = _oracle_params(m, torch.float32)
    xb = x.detach().view(B, S, h)
    xf = xb.float().requires_grad_(True)
    with torch.no_grad():
        hs_b = O.decoder(xb, lb, nb, heads, 1e-5, mask)
    hs_f = O.decoder(xf, lf, nf, heads, 1e-5, mask)
    assert len(hs) == L + 1
    valid = slice(None) if mask is None else mask.view(-1)
    for i, (a, b_, c) in enumerate(zip(hs, hs_b, hs_f)):
        a = a[valid]
        e_ref = rel_err(a, b_.reshape(B * S, h)[valid])
        e_tru = rel_err(a, c.reshape(B * S, h)[valid])
        e_ref_tru = rel_err(b_.reshape(B * S, h)[valid], c.reshape(B * S, h)[valid])
        # as close to the truth as the reference's own bf16 path is (within 1.5x), and close to that path itself
        assert e_tru < 1.5 * e_ref_tru + 1e-3, (i, e_tru, e_ref_tru)
        assert e_ref < 8e-3, (i, e_ref)
    # backward: d(sum of last hidden * g)
    g = torch.randn(B * S, h, device="cuda").to(torch.bfloat16)
    if mask is not None:
        g = g * mask.view(-1, 1)
    hs[-1].backward(g)
    hs_f[-1].backward(g.view(B, S, h).float())
    assert rel_err(x.grad[valid], xf.grad.reshape(B * S, h)[valid]) < 2e-2
    names = ["q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj", "ln1", "ln2"]
    for li, layer in enumerate(m.layers):
        ps = layer._masters()
        for n_, p in zip(names, ps):
            assert p.grad is not None, (li, n_)
            e = rel_err(p.grad, lf[li][n_].grad)
            assert e < 2.5e-2, (li, n_, e)
    assert rel_err(m.norm.weight.grad, nf.grad) < 2e-2
    # second backward accumulates into the same arenas
    g0 = m.layers[0].self_attn.q_proj.weight.grad.clone()
    hs2 = m.run_layers(x.detach().requires_grad_(True), B, S, mask)
    hs2[-1].backward(g)
    assert rel_err(m.layers[0].self_attn.q_proj.weight.grad, 2 * g0) < 1e-3
