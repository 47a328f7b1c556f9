"""Parity against the reference's OWN GPU path: goldens recorded from the unmodified reference run on a B200 with
`_attn_implementation="flash_attention_2"` (flash-attn 2.8.3), bf16 parameters + autocast
(tests/golden/make_golden_gpu.py -> tests/golden/*_gpu.npz) — the arithmetic BASELINE.json's north_star names.

* layer7b: ONE decoder layer at full Llama-2-7B width (h 4096, 32 x 128 heads, ffn 11008), 2 x 548 tokens, the second
  sequence right-padded (varlen path): output rows, input-gradient rows, every weight gradient.
* tiny_img / tiny_pc / align: the whole MLA.forward + backward on the same weights, batches and random draws as the CPU
  goldens; losses, boundary tensors and probe gradients against what the reference computed on the GPU.

Tolerances are set from what bf16 allows, not from the 1e-3 the north_star asks of results: two correct bf16
implementations of the same layer differ by ~1e-3..4e-3 relative L2 on activations (every op output is rounded to 8
mantissa bits; the reference's own GPU and CPU runs differ by that much, see profiles/r02_parity_table.md); scalar
losses agree to <= 2e-3.  The measured values are tabulated by tools/parity_table.py.
"""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold(name):
    p = os.path.join(GOLD, name + "_gpu.npz")
    if not os.path.exists(p):
        pytest.skip(f"{p} missing")
    return np.load(p, allow_pickle=False)


def run_layer7b():
    """Our decoder layer on the layer7b fixture.  Returns dict(y, dx, grads{name: tensor}, mask)."""
    from golden.make_golden_gpu import LAYER7B, layer7b_inputs
    from mla_b200 import llama
    from oracle import fixtures
    c = LAYER7B
    m = llama.LlamaModel(8, c["h"], c["f"], 1, c["heads"], eps=c["eps"])
    layer = m.layers[0]
    sd = fixtures.fill_state_dict(layer.state_dict(), seed=7)
    layer.load_state_dict({k: v.to(torch.bfloat16).float() for k, v in sd.items()})     # the reference ran bf16 weights
    m = m.cuda()
    m.set_save_levels("none")
    x, dy, mask = layer7b_inputs("cuda")
    x2 = x.reshape(-1, c["h"]).clone().requires_grad_(True)
    cos, sin = m.rope_tables(c["S"], x.device)
    sh = llama.LayerShape(c["B"], c["S"], c["heads"], c["h"] // c["heads"], mask, cos, sin)
    anchor = torch.zeros(1, device="cuda", requires_grad=True)
    y = llama._LayerFn.apply(x2, anchor, layer, sh)
    y.backward(dy.reshape(-1, c["h"]))
    grads = {k: p.grad for k, p in layer.named_parameters()}
    return dict(y=y.detach(), dx=x2.grad, grads=grads, mask=mask.reshape(-1))


def layer7b_errors(out, z):
    rows = torch.from_numpy(z["rows"]).cuda()
    e = {"y_rows": rel_err(out["y"][rows].cpu(), torch.from_numpy(z["y_rows"])),
         "dx_rows": rel_err(out["dx"][rows].cpu(), torch.from_numpy(z["dx_rows"])),
         "y_norm": abs(float(out["y"][out["mask"]].float().norm()) - float(z["y_norm"])) / float(z["y_norm"]),
         "dx_norm": abs(float(out["dx"][out["mask"]].float().norm()) - float(z["dx_norm"])) / float(z["dx_norm"]),
         "y_colsum": rel_err(out["y"][out["mask"]].float().sum(0).cpu(), torch.from_numpy(z["y_colsum"])),
         "dx_colsum": rel_err(out["dx"][out["mask"]].float().sum(0).cpu(), torch.from_numpy(z["dx_colsum"]))}
    for k, g in out["grads"].items():
        e["grad." + k] = rel_err(g.flatten()[:65536].cpu(), torch.from_numpy(z["grad." + k]))
        e["gradnorm." + k] = abs(float(g.float().norm()) - float(z["gradnorm." + k])) / float(z["gradnorm." + k])
    return e


def layer7b_truth():
    """fp32 evaluation of the same layer (oracle/llama.py on the bf16-rounded weights, torch fp32 on the GPU): the truth
    both bf16 implementations are measured against.  Same keys as layer7b_errors' inputs."""
    from golden.make_golden_gpu import LAYER7B, layer7b_inputs
    from mla_b200 import llama
    from oracle import fixtures, llama as O
    c = LAYER7B
    m = llama.LlamaModel(8, c["h"], c["f"], 1, c["heads"], eps=c["eps"])
    sd = fixtures.fill_state_dict(m.layers[0].state_dict(), seed=7)
    names = dict(q_proj="self_attn.q_proj.weight", k_proj="self_attn.k_proj.weight", v_proj="self_attn.v_proj.weight",
                 o_proj="self_attn.o_proj.weight", gate_proj="mlp.gate_proj.weight", up_proj="mlp.up_proj.weight",
                 down_proj="mlp.down_proj.weight", ln1="input_layernorm.weight", ln2="post_attention_layernorm.weight")
    p = {k: sd[v].to(torch.bfloat16).float().cuda().requires_grad_(True) for k, v in names.items()}
    x, dy, mask = layer7b_inputs("cuda")
    xf = x.float().requires_grad_(True)
    cos, sin = O.rope_tables(c["S"], c["h"] // c["heads"], 10000.0, torch.float32)
    y = O.decoder_layer(xf, p, c["heads"], c["eps"], cos.cuda(), sin.cuda(), mask)
    y.backward(dy.float())
    return dict(y=y.detach().reshape(-1, c["h"]), dx=xf.grad.reshape(-1, c["h"]),
                grads={v: p[k].grad for k, v in names.items()}, mask=mask.reshape(-1))


def test_full_width_layer_matches_reference_gpu(cuda_lib):
    """Ours vs the reference's GPU run of one full-width layer, judged against the fp32 truth: our bf16 result must be as
    close to the truth as the reference's own bf16 (flash-attn) result is (x1.5 + floor), norms must agree to 1e-3, and
    the two bf16 results must sit within twice the reference's own distance to the truth of each other."""
    z = _gold("layer7b")
    ours, truth = run_layer7b(), layer7b_truth()
    e_ours_ref = layer7b_errors(ours, z)
    e_ref_tru = layer7b_errors(truth, z)          # == distance(reference, truth) on the stored probes
    rows = torch.from_numpy(z["rows"]).cuda()
    e_ours_tru = {"y_rows": rel_err(ours["y"][rows], truth["y"][rows]), "dx_rows": rel_err(ours["dx"][rows], truth["dx"][rows])}
    for k, g in ours["grads"].items():
        e_ours_tru["grad." + k] = rel_err(g.flatten()[:65536], truth["grads"][k].flatten()[:65536])
    for k, v in e_ours_tru.items():
        assert v < 1.5 * e_ref_tru[k] + 1e-3, (k, "ours-vs-truth", v, "reference-vs-truth", e_ref_tru[k])
        assert e_ours_ref[k] < 2.0 * e_ref_tru[k] + 2e-3, (k, "ours-vs-reference", e_ours_ref[k], "reference-vs-truth", e_ref_tru[k])
    for k, v in e_ours_ref.items():
        if k.endswith("_norm") or k.startswith("gradnorm."):
            assert v < 1e-3, (k, v)


def e2e_errors(name):
    """Whole MLA.forward + backward on the CUDA path vs the reference's GPU golden (and its CPU golden)."""
    from test_mla_gpu import build_cuda_model, run_cuda
    from test_oracle_vs_golden import case_cfg, load_case
    zc, batch = load_case(name)
    zg = _gold(name)
    c = case_cfg(name)
    mla, _ = build_cuda_model(c)
    loss_dict, out = run_cuda(mla, batch, zc, c)
    loss_dict["total_loss"].backward()
    valid = torch.from_numpy(zc["fused_attention_mask"]).bool()
    hs = out.hidden_states
    e = {}
    for tag, z in (("gpu", zg), ("cpu", zc)):
        e[f"total_loss.{tag}"] = abs(float(loss_dict["total_loss"]) - float(z["total_loss"])) / abs(float(z["total_loss"]))
        for k in ("img_pc_contrastive_loss", "tactile_contrastive_loss"):
            if k in z.files:
                e[f"{k}.{tag}"] = abs(float(loss_dict[k]) - float(z[k])) / abs(float(z[k]))
        e[f"hidden_first.{tag}"] = rel_err(hs[0].cpu(), torch.from_numpy(z["hidden_first"]))
        e[f"hidden_last.{tag}"] = rel_err(hs[-1].cpu()[valid], torch.from_numpy(z["hidden_last"])[valid])
        if "hidden_8" in z.files:
            e[f"hidden_8.{tag}"] = rel_err(hs[8].cpu()[valid], torch.from_numpy(z["hidden_8"])[valid])
        e[f"noise_pred.{tag}"] = rel_err(out.noise_pred.cpu(), torch.from_numpy(z["noise_pred"]))
        named = dict(mla.named_parameters())
        for k in z.files:
            if k.startswith("gradnorm."):
                g = named[k[len("gradnorm."):]].grad
                e[f"{k}.{tag}"] = abs(float(g.float().norm()) - float(z[k])) / float(z[k])
            elif k.startswith("grad."):
                g = named[k[len("grad."):]].grad.float().cpu()
                ref = torch.from_numpy(z[k])
                e[f"{k}.{tag}"] = rel_err(g.flatten()[:ref.numel()].reshape(ref.shape), ref)
    # the reference against itself: its GPU (flash-attn) run vs its CPU (SDPA) run
    e["ref_gpu_vs_ref_cpu.total_loss"] = abs(float(zg["total_loss"]) - float(zc["total_loss"])) / abs(float(zc["total_loss"]))
    e["ref_gpu_vs_ref_cpu.hidden_last"] = rel_err(torch.from_numpy(zg["hidden_last"])[valid], torch.from_numpy(zc["hidden_last"])[valid])
    e["ref_gpu_vs_ref_cpu.noise_pred"] = rel_err(torch.from_numpy(zg["noise_pred"]), torch.from_numpy(zc["noise_pred"]))
    e["ref_gpu_vs_ref_cpu.hidden_first"] = rel_err(torch.from_numpy(zg["hidden_first"]), torch.from_numpy(zc["hidden_first"]))
    if "hidden_8" in zg.files and "hidden_8" in zc.files:
        e["ref_gpu_vs_ref_cpu.hidden_8"] = rel_err(torch.from_numpy(zg["hidden_8"])[valid], torch.from_numpy(zc["hidden_8"])[valid])
    for k in zg.files:
        if k.startswith("gradnorm.") and k in zc.files:
            e["ref_gpu_vs_ref_cpu." + k] = abs(float(zg[k]) - float(zc[k])) / float(zc[k])
        elif k.startswith("grad.") and k in zc.files:
            e["ref_gpu_vs_ref_cpu." + k] = rel_err(torch.from_numpy(zg[k]), torch.from_numpy(zc[k]))
    for k in ("img_pc_contrastive_loss", "tactile_contrastive_loss"):
        if k in zg.files and k in zc.files:
            e["ref_gpu_vs_ref_cpu." + k] = abs(float(zg[k]) - float(zc[k])) / abs(float(zc[k]))
    return e


@pytest.mark.parametrize("name", ["tiny_img", "tiny_pc", "align"])
def test_mla_matches_reference_gpu_golden(cuda_lib, name):
    e = e2e_errors(name)
    # ours must sit as close to the reference's GPU run as the reference's own two backends sit to each other (x2 + floor)
    assert e["total_loss.gpu"] < 2 * e["ref_gpu_vs_ref_cpu.total_loss"] + 2e-3, e
    assert e["hidden_first.gpu"] < 2 * e["ref_gpu_vs_ref_cpu.hidden_first"] + 4e-3, e
    assert e["hidden_last.gpu"] < 2 * e["ref_gpu_vs_ref_cpu.hidden_last"] + 4e-3, e
    assert e["noise_pred.gpu"] < 2 * e["ref_gpu_vs_ref_cpu.noise_pred"] + 4e-3, e
    for k, v in e.items():
        if k.startswith("gradnorm.") and k.endswith(".gpu"):
            assert v < 8e-2, (k, v)
        if k.endswith("contrastive_loss.gpu"):
            assert v < 4e-3, (k, v)
