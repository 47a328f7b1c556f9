"""R-GPU: the UNMODIFIED reference on one B200 + the flash-attn kernel head-to-head.

    gpurun -- python tools/ref_gpu.py step --workload cfg2 --out gpurun_out/r02_ref_gpu_cfg2.json
    gpurun -- python tools/ref_gpu.py attn --out gpurun_out/r02_attn_vs_flash.json

`step`: builds the reference's own `MLA(PrismaticVLM(LlamaForCausalLM))` (baseline/_ref via oracle/ref_shim.py — none of
our modules or kernels on that path) at Llama-2-7B shapes, random init, `_attn_implementation="flash_attention_2"`
(flash-attn 2.8.3 = FA2 compiled for sm_100, what `use_flash_attention_2=True` selects, models/backbones/llm/llama2.py:62),
the same synthetic batch bench.py uses (per-GPU batch 8 x 4 diffusion repeats, S = 548), and times with CUDA events:
  * fwd and fwd+bwd with bf16 parameters + autocast (FSDP MixedPrecision(param_dtype=bf16) arithmetic,
    training/strategies/fsdp.py:185-187), with per-decoder-layer activation checkpointing (fsdp.py:217-223, the
    reference's default) and — if it fits — without;
  * the whole step as the reference's strategy runs it on one GPU: fp32 master parameters + autocast, activation
    checkpointing, `clip_grad_norm_(1.0)`, torch AdamW (fsdp.py:242-257,:310; base_strategy_mla.py:366-379).
`attn`: flash_attn_func fwd / fwd+bwd (and torch SDPA) against mla_attn_fwd_sm100 / mla_attn_bwd_sm100 at
[32,548,32,128] and [4,3876,32,128], causal, bf16.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.ref_model import L, apply_checkpointing, build_reference_7b, ref_call  # noqa: E402


def _time(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def cmd_step(a):
    import contextlib
    import io
    from mla_b200.synthetic import make_batch, map_tensors
    use_pc = a.workload in ("cfg3", "cfg4")
    B, S = a.batch, 548
    host = make_batch(B, 32, 0, 672, 1024, seed=1234, use_pointcloud=use_pc, use_tactile=use_pc)
    devb = map_tensors(host, lambda t: t.cuda())
    res = {"workload": a.workload, "per_gpu_batch": B, "repeats": 4, "seq_len": S, "tokens_per_step": B * 4 * S,
           "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
           "flash_attn": __import__("flash_attn").__version__, "attn_implementation": "flash_attention_2", "runs": []}
    quiet = contextlib.redirect_stdout(io.StringIO())       # model_mla.py:233 prints the loss dict every forward

    def measure(tag, param_dtype, ckpt, full_step):
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
        row = {"variant": tag, "param_dtype": str(param_dtype).replace("torch.", ""), "activation_checkpointing": ckpt}
        try:
            mla, ns = build_reference_7b(a.workload, a.layers, param_dtype)
            if ckpt:
                apply_checkpointing(mla, ns)
            params = [p for p in mla.parameters() if p.requires_grad]
            opt = None
            if full_step:
                decay = [p for p in params if p.ndim > 1]
                no_decay = [p for p in params if p.ndim <= 1]
                opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.0}, {"params": no_decay, "weight_decay": 0.0}],
                                        lr=2e-5)

            def fwd():
                with torch.no_grad(), quiet:
                    ref_call(mla, devb)

            def fwd_bwd():
                with quiet:
                    loss = ref_call(mla, devb)
                loss.backward()
                for p in params:
                    p.grad = None

            def step():
                with quiet:
                    loss = ref_call(mla, devb)
                loss.backward()
                torch.nn.utils.clip_grad_norm_(params, 1.0)
                opt.step()
                opt.zero_grad()

            if full_step:
                row["step_ms"] = round(_time(step, a.warmup, a.steps), 2)
                row["tokens_per_s"] = round(B * 4 * S / row["step_ms"] * 1e3, 1)
            else:
                row["fwd_ms"] = round(_time(fwd, a.warmup, a.steps), 2)
                row["fwd_bwd_ms"] = round(_time(fwd_bwd, a.warmup, a.steps), 2)
                row["tokens_per_s_fwd_bwd"] = round(B * 4 * S / row["fwd_bwd_ms"] * 1e3, 1)
            row["peak_mem_gb"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)
            del mla, params, opt
        except torch.OutOfMemoryError as ex:
            row["error"] = "CUDA out of memory: " + str(ex)[:120]
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        print(json.dumps(row), flush=True)
        res["runs"].append(row)

    measure("bf16 params + autocast, checkpointing (reference default)", torch.bfloat16, True, False)
    if not a.skip_nockpt:
        measure("bf16 params + autocast, no checkpointing", torch.bfloat16, False, False)
    if not a.skip_full:
        measure("full step: fp32 masters + autocast + checkpointing + clip_grad_norm + torch AdamW", torch.float32, True, True)
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


def cmd_attn(a):
    from flash_attn import flash_attn_func
    import torch.nn.functional as Fn
    from mla_b200 import ops
    out = {"gpu": torch.cuda.get_device_name(0), "flash_attn": __import__("flash_attn").__version__, "shapes": []}
    for (B, S, Hh, D) in ((32, 548, 32, 128), (4, 3876, 32, 128), (8, 1060, 32, 128)):
        torch.manual_seed(0)
        qkv = (torch.randn(B * S, 3 * Hh * D, device="cuda") * 0.5).to(torch.bfloat16)
        q, k, v = [t.reshape(B, S, Hh, D) for t in qkv.split(Hh * D, dim=1)]
        qa, ka, va = [t.contiguous().requires_grad_(True) for t in (q, k, v)]
        do = (torch.randn(B, S, Hh, D, device="cuda") * 0.1).to(torch.bfloat16)
        fl_fwd = 4.0 * B * Hh * S * S * D / 2      # causal: half of the 2 products x 2 S^2 d
        fl_bwd = 2.5 * fl_fwd

        def fa_fwd():
            with torch.no_grad():
                return flash_attn_func(qa, ka, va, 0.0, causal=True)

        def fa_fb():
            o = flash_attn_func(qa, ka, va, 0.0, causal=True)
            o.backward(do)
            qa.grad = ka.grad = va.grad = None

        def sd_fwd():
            with torch.no_grad():
                return Fn.scaled_dot_product_attention(qa.transpose(1, 2), ka.transpose(1, 2), va.transpose(1, 2), is_causal=True)

        def sd_fb():
            o = Fn.scaled_dot_product_attention(qa.transpose(1, 2), ka.transpose(1, 2), va.transpose(1, 2), is_causal=True)
            o.backward(do.transpose(1, 2))
            qa.grad = ka.grad = va.grad = None

        ctx, lse = ops.attn_fwd(qkv, B, S, Hh, D, None)
        dctx = do.reshape(B * S, Hh * D).contiguous()

        def our_fwd():
            return ops.attn_fwd(qkv, B, S, Hh, D, None)

        def our_bwd():
            return ops.attn_bwd(dctx, qkv, ctx, lse, B, S, Hh, D, None)

        # numerics of the two kernels against each other (and both against fp32 softmax attention on a slice)
        o_fa = fa_fwd().reshape(B * S, Hh * D)
        err_fwd = float((ctx.float() - o_fa.float()).norm() / o_fa.float().norm())
        o = flash_attn_func(qa, ka, va, 0.0, causal=True)
        o.backward(do)
        dqkv = our_bwd()
        g_fa = torch.cat([qa.grad.reshape(B * S, -1), ka.grad.reshape(B * S, -1), va.grad.reshape(B * S, -1)], 1)
        err_bwd = float((dqkv.float() - g_fa.float()).norm() / g_fa.float().norm())
        qa.grad = ka.grad = va.grad = None
        t = {"fa_fwd": _time(fa_fwd, 3, 20), "fa_fwd_bwd": _time(fa_fb, 3, 20), "sdpa_fwd": _time(sd_fwd, 3, 20),
             "sdpa_fwd_bwd": _time(sd_fb, 3, 20), "ours_fwd": _time(our_fwd, 3, 20), "ours_bwd": _time(our_bwd, 3, 20)}
        row = {"shape": [B, S, Hh, D], "ms": {k_: round(v_, 4) for k_, v_ in t.items()},
               "tflops": {"flash_attn_fwd": round(fl_fwd / t["fa_fwd"] / 1e9, 1),
                          "flash_attn_bwd": round(fl_bwd / (t["fa_fwd_bwd"] - t["fa_fwd"]) / 1e9, 1),
                          "sdpa_fwd": round(fl_fwd / t["sdpa_fwd"] / 1e9, 1),
                          "sdpa_bwd": round(fl_bwd / (t["sdpa_fwd_bwd"] - t["sdpa_fwd"]) / 1e9, 1),
                          "ours_fwd": round(fl_fwd / t["ours_fwd"] / 1e9, 1),
                          "ours_bwd": round(fl_bwd / t["ours_bwd"] / 1e9, 1)},
               "rel_l2_ours_vs_flash_attn": {"fwd": err_fwd, "bwd": err_bwd}}
        print(json.dumps(row), flush=True)
        out["shapes"].append(row)
        del qkv, q, k, v, qa, ka, va, do, ctx, lse, dctx, o, o_fa, dqkv, g_fa
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("step")
    s.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"])
    s.add_argument("--batch", type=int, default=8)
    s.add_argument("--layers", type=int, default=L)
    s.add_argument("--steps", type=int, default=3)
    s.add_argument("--warmup", type=int, default=1)
    s.add_argument("--skip-nockpt", action="store_true")
    s.add_argument("--skip-full", action="store_true")
    s.add_argument("--out", default="gpurun_out/r02_ref_gpu.json")
    s = sub.add_parser("attn")
    s.add_argument("--out", default="gpurun_out/r02_attn_vs_flash.json")
    a = ap.parse_args()
    {"step": cmd_step, "attn": cmd_attn}[a.cmd](a)
