#!/usr/bin/env python
"""Benchmark of the MLA training step (BASELINE.json metric: multimodal tokens/s, Llama-2-7B MLA, bf16).

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path (one process per GPU under torchrun for N>1)
    python bench.py --impl reference ...                     # the reference algorithm on the host CPU (oracle port)

A step = forward + backward + gradient all-reduce (N>1) + clip + AdamW of one synthetic batch:
workload "cfg2" = BASELINE configs[1]: Llama-2-7B MLA, image-only tokens + 32 text tokens, per-GPU batch 8 x 4
diffusion repeats = 32 sequences of 548 tokens (17,536 multimodal tokens per GPU per step), random-init weights.
`value` is timed with the batch already resident in HBM; `e2e` times the same step through the public module call
with the batch in pinned host memory (H2D inside the timed region) and a D2H read of the loss every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, F, L, HEADS, VOCAB = 4096, 11008, 32, 32, 32064
WORKLOADS = {
    # name: (use_pointcloud, use_tactile, use_contrastive, description)
    "cfg2": (False, False, False, "MLA Llama2-7B, image-only (672x672 patchified -> 256 tokens) + 32 text toks, bs=8 x 4 repeats"),
    "cfg3": (True, True, True, "MLA Llama2-7B, image+pointcloud+tactile alignment + contrastive loss, bs=8 x 4 repeats"),
    # BASELINE configs[4]: post-training (gen_img + gen_pc + use_roi), 14 camera views -> 3876 fused tokens per sequence
    "cfg5": (True, True, True, "MLA post-training (gen_image+gen_pointcloud+use_roi), 14 views -> seq 3876, bs=1 x 4 repeats"),
}
GENERATION = {"cfg5": dict(use_generation=True, gen_image=True, gen_pointcloud=True, gen_tactile=False, use_roi=True)}
EXTRA_VIEWS = {"cfg5": 13}
DEFAULT_BATCH = {"cfg5": 1}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def step_flops(tokens: int, S: int) -> float:
    """Algorithmic FLOPs of one step per GPU: decoder only, recompute NOT counted."""
    return 3.0 * tokens * L * (8 * H * H + 6 * H * F + 2 * S * H)


# ----------------------------------------------------------------------------------------------------------- ours
def build_model(workload: str, T: int = 0, stage: str = ""):
    from mla_b200.backbone import LLMBackbone, LlamaConfig
    from mla_b200.mla import MLA
    from mla_b200.vlm import PrismaticVLM
    use_pc, use_tac, use_con, _ = WORKLOADS[workload]
    flags = dict(use_diff=True, use_pointcloud=use_pc, use_tactile=use_tac, use_contrastive=use_con, use_generation=False)
    flags.update(GENERATION.get(workload, {}))
    torch.manual_seed(0)
    with torch.device("cuda"):
        cfg = LlamaConfig(vocab_size=VOCAB, hidden_size=H, intermediate_size=F, num_hidden_layers=L,
                          num_attention_heads=HEADS)
        vlm = PrismaticVLM("mla-7b-synthetic", LLMBackbone(config=cfg), token_size=H, action_dim=7, **flags)
        mla = MLA(vlm, None, token_size=H, action_dim=7, future_action_window_size=T, **flags)
    with torch.no_grad():   # initialize_weights zeroes the head (prismatic.py:320): give it signal so grads flow
        mla.vlm.final_layer.mlp.fc2.weight.normal_(std=0.02)
    mla.train()
    # scripts/{post,sft}_rlbench.sh; --stage pretrain (scripts/pretrain_*.sh) also trains the two tokenizers
    mla.freeze_backbones(stage or ("post-training" if workload in GENERATION else "finetune"))
    return mla


def GEMM_SHAPES(T):
    """(M, N, K) of the 12 launches timed by gemm_roofline: forward, dgrad (bf16 out), wgrad (fp32 out)."""
    return [(T, 3 * H, H), (T, H, H), (T, 2 * F, H), (T, H, F), (T, H, 3 * H), (T, H, H), (T, H, 2 * F), (T, F, H),
            (3 * H, H, T), (H, H, T), (2 * F, H, T), (H, F, T)]


def ncu_gemm_traffic():
    """Mean DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (None if absent)."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_summary.json")))["gemm"]
        return round(sum((r["dram_read_MB"] + r["dram_write_MB"]) * 1e6 for r in d) / len(d))
    except Exception:
        return None


def gemm_roofline(tokens: int, peak_tf: float, iters: int = 5):
    """The dominant kernel is the tcgen05 GEMM (~97 % of the step's FLOPs): time every GEMM shape of a decoder layer
    (forward, dgrad, wgrad) back to back with CUDA events on the launching stream and report algorithmic FLOP/s."""
    from mla_b200 import ops
    T = tokens
    dev = "cuda"
    bf = torch.bfloat16
    x = torch.randn(T, H, device=dev).to(bf)
    xf = torch.randn(T, F, device=dev).to(bf)
    wqkv, wo = torch.randn(3 * H, H, device=dev).to(bf), torch.randn(H, H, device=dev).to(bf)
    wgu, wd = torch.randn(2 * F, H, device=dev).to(bf), torch.randn(H, F, device=dev).to(bf)
    dqkv, dgu = torch.randn(T, 3 * H, device=dev).to(bf), torch.randn(T, 2 * F, device=dev).to(bf)
    g = [torch.empty_like(w, dtype=torch.float32) for w in (wqkv, wo, wgu, wd)]
    calls = [
        (lambda: ops.gemm(x, wqkv), 2.0 * T * 3 * H * H), (lambda: ops.gemm(x, wo), 2.0 * T * H * H),
        (lambda: ops.gemm(x, wgu), 2.0 * T * 2 * F * H), (lambda: ops.gemm(xf, wd), 2.0 * T * H * F),
        (lambda: ops.gemm(dqkv, wqkv, b_mn=True), 2.0 * T * 3 * H * H), (lambda: ops.gemm(x, wo, b_mn=True), 2.0 * T * H * H),
        (lambda: ops.gemm(dgu, wgu, b_mn=True), 2.0 * T * 2 * F * H), (lambda: ops.gemm(x, wd, b_mn=True), 2.0 * T * H * F),
        (lambda: ops.gemm(dqkv, x, a_mn=True, b_mn=True, out=g[0]), 2.0 * T * 3 * H * H),
        (lambda: ops.gemm(x, x, a_mn=True, b_mn=True, out=g[1]), 2.0 * T * H * H),
        (lambda: ops.gemm(dgu, x, a_mn=True, b_mn=True, out=g[2]), 2.0 * T * 2 * F * H),
        (lambda: ops.gemm(x, xf, a_mn=True, b_mn=True, out=g[3]), 2.0 * T * H * F),
    ]
    for fn, _ in calls:
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        for fn, _ in calls:
            fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = sum(f for _, f in calls)
    ach = flops / ms / 1e9
    return {"bound": "tensor", "achieved": round(ach, 1), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(ach / peak_tf, 4),
            "traffic": ncu_gemm_traffic(), "traffic_unit": "bytes per launch (dram read+write, ncu --set full, mean of "
            "the same 12 launches: profiles/r01_ncu_full_summary.json)",
            "algorithmic_bytes_per_launch": round(sum(2.0 * (m * k + n * k) + (2.0 if i < 8 else 4.0) * m * n
                                                      for i, (m, n, k) in enumerate(GEMM_SHAPES(T))) / 12),
            "kernel": "gemm2_bf16_kernel (tcgen05 cta_group::2, 12 GEMM shapes of one decoder layer fwd+bwd)",
            "launch_ms_avg": round(ms / len(calls), 4)}


def cpu_baseline(sample_layers: int = 2, threads: int = 0):
    """The reference algorithm (oracle port) on the host cores: fp32, a `sample_layers`-layer slice of the 7B decoder
    at full width on one sequence (B=1, R=1, S=548), forward+backward; tokens/s extrapolated to 32 layers."""
    from oracle import llama as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    S = 548
    g = torch.Generator().manual_seed(0)
    layers = []
    for _ in range(sample_layers):
        p = dict(q_proj=(H, H), k_proj=(H, H), v_proj=(H, H), o_proj=(H, H), gate_proj=(F, H), up_proj=(F, H), down_proj=(H, F))
        d = {k: (torch.randn(s, generator=g) * 0.02).requires_grad_(True) for k, s in p.items()}
        d["ln1"] = torch.ones(H, requires_grad=True)
        d["ln2"] = torch.ones(H, requires_grad=True)
        layers.append(d)
    norm = torch.ones(H, requires_grad=True)
    x = torch.randn(1, S, H, generator=g) * 0.02
    t0 = time.perf_counter()
    hs = O.decoder(x, layers, norm, HEADS, 1e-5, None)
    hs[-1].square().mean().backward()
    dt = time.perf_counter() - t0
    full = dt * (L / sample_layers)
    return {"value": round(S / full, 3), "unit": "tokens/s", "cores": threads, "kind": "port",
            "sample": f"oracle (fp32 PyTorch restatement) fwd+bwd of a {sample_layers}-layer slice of the 7B decoder, "
                      f"1 sequence x 548 tokens, {dt:.1f} s measured, extrapolated x{L // sample_layers} to 32 layers",
            "seconds_measured": round(dt, 2)}


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mla_b200 import _lib
    from mla_b200.synthetic import batch_bytes, make_batch, map_tensors
    from mla_b200.trainer import DataParallelTrainer, plan_save_levels
    _lib.check(_lib.lib().mla_device_check())

    use_pc, use_tac, _, desc = WORKLOADS[args.workload]
    B, R, Lt, T = (args.batch or DEFAULT_BATCH.get(args.workload, 8)), 4, 32, 0
    views = EXTRA_VIEWS.get(args.workload, 0)
    mla = build_model(args.workload, T, args.stage)
    trainer = DataParallelTrainer(mla, lr=2e-5, weight_decay=0.0, max_grad_norm=1.0)
    S = 1 + 256 + 256 * (1 + views) + 1 + (Lt - 1) + 1 + 1 + (T + 1)
    tokens = B * R * S
    # stage pretrain keeps the tokenizers' pre-BatchNorm activations through the decoder (7.6 GB for 32 clouds); their
    # backward transients come after the decoder's activations are gone
    reserve = 10.0 + (9.0 if args.stage == "pretrain" else 0.0)
    levels = [args.save_level] * L if args.save_level != "auto" else plan_save_levels(mla, tokens, reserve_gb=reserve)
    mla.vlm.llm_backbone.llm.model.set_save_levels(levels)

    host = make_batch(B, Lt, T, 672, 1024, seed=1234 + rank, use_pointcloud=use_pc, use_tactile=use_tac, pin=True,
                      extra_views=views, generation=args.workload in GENERATION)
    devb = map_tensors(host, lambda t: t.cuda(non_blocking=True))
    kw = dict(camera_name="rlbench_front", repeated_diffusion_steps=R, use_diff=True)

    def call(b):
        loss_dict, _ = mla(input_ids=b["input_ids"], attention_mask=b["attention_mask"], labels=b["labels"],
                           actions=b["actions"], images=b["images"], point_cloud=b.get("point_cloud"),
                           tactile=b.get("tactile"), proprio=b["proprio"], gripper_xyz=b.get("gripper_xyz"),
                           action_masks=b["action_masks"], next_images=b.get("next_images"),
                           next_point_cloud=b.get("next_point_cloud"), next_tactile=b.get("next_tactile"), **kw)
        loss = loss_dict["total_loss"]
        loss.backward()
        trainer.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, batch, read_loss):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(n):
            loss = call(batch)
            if read_loss:
                last = float(loss.item())       # D2H read of the step's result
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / n, last

    for _ in range(max(args.warmup, 3)):
        loss = call(devb)
    float(loss.item())
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _lib.launch_count()
    ms_dev, _ = timed(args.steps, devb, False)
    launches = (_lib.launch_count() - n0) // args.steps
    call(host)
    ms_e2e, last_loss = timed(args.steps, host, True)

    # forward-only and forward+backward (no exchange / optimizer) — BASELINE.json quotes "fwd+bwd ms" next to tokens/s
    model_ = mla.vlm.llm_backbone.llm.model
    layer_ids = {id(p) for l in trainer.layers for p in l._masters()}
    small = [p for p in mla.parameters() if p.requires_grad and id(p) not in layer_ids]

    def fwd_only(b):
        with torch.no_grad():
            return fwd_loss(b)

    def fwd_loss(b):
        loss_dict, _ = mla(input_ids=b["input_ids"], attention_mask=b["attention_mask"], labels=b["labels"],
                           actions=b["actions"], images=b["images"], point_cloud=b.get("point_cloud"),
                           tactile=b.get("tactile"), proprio=b["proprio"], gripper_xyz=b.get("gripper_xyz"),
                           action_masks=b["action_masks"], next_images=b.get("next_images"),
                           next_point_cloud=b.get("next_point_cloud"), next_tactile=b.get("next_tactile"), **kw)
        return loss_dict["total_loss"]

    def fwd_bwd(b):
        fwd_loss(b).backward()
        trainer.exchange()                  # N > 1: the gradient all-reduce issued from backward is part of it
        model_.mark_grads_fresh()           # next backward overwrites the arenas (what zero_grad does in the step)
        for p in small:
            p.grad = None

    def time_fn(fn, n):
        fn(devb)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn(devb)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / n

    ms_fwd = ms_fwd_bwd = None
    if world == 1:                          # per-GPU quantities: measured on the single-GPU run only
        try:
            ms_fwd = round(time_fn(fwd_only, max(2, args.steps // 2)), 2)
            ms_fwd_bwd = round(time_fn(fwd_bwd, max(2, args.steps // 2)), 2)
        except Exception as ex:             # secondary numbers must never take the headline line down
            sys.stderr.write(f"fwd / fwd+bwd timing failed: {ex}\n")
            ms_fwd = ms_fwd_bwd = None
    clocks = sampler.stop() if rank == 0 else None
    mla.vlm.check_errors()
    mem_gb = torch.cuda.max_memory_allocated() / 2 ** 30

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, tf_burst, tf_sust, src = peaks()
    # free the training state before the isolated kernel timing
    roof = None
    try:
        del trainer
        torch.cuda.empty_cache()
        roof = gemm_roofline(tokens, tf_sust)
        roof["peak_source"] = f"bf16_tflops_sustained of {src} MEASURED_PEAKS.json (kernel timed back to back, power-capped)"
    except Exception as ex:  # e.g. not enough free memory next to the model
        roof = {"bound": "tensor", "achieved": None, "peak": tf_sust, "unit": "TFLOP/s", "frac": None, "traffic": None,
                "error": str(ex)[:200]}
    fl = step_flops(tokens, S)
    roof["step_algorithmic_tflops"] = round(fl / ms_dev / 1e9, 1)
    roof["step_frac_of_peak"] = round(fl / ms_dev / 1e9 / tf_sust, 4)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(sample_layers=4)
        except Exception as ex:
            cpu = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"[:200]}
    out = {
        "metric": "multimodal_tokens_per_sec", "value": round(world * tokens / ms_dev * 1e3, 1), "unit": "tokens/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_dev, 2),
        "fwd_ms": ms_fwd, "fwd_bwd_ms": ms_fwd_bwd,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "per_gpu_batch": B, "repeated_diffusion_steps": R,
                   "seq_len": S, "tokens_per_gpu_step": tokens, "global_batch": B * world, "parallelism": f"dp{world}",
                   "stage": ("pretrain (vision tokenizers trained)" if args.stage == "pretrain" else
                             "post-training (generation heads on)" if args.workload in GENERATION
                             else "finetune (vision tokenizers frozen)"), "optimizer": "AdamW fp32 master + fp32 grads",
                   "activation_save_levels": {lv: levels.count(lv) for lv in sorted(set(levels))},
                   "l2": "step streams >100 GB of weights/activations (>> 126 MB L2); no explicit flush needed",
                   "peak_mem_gb": round(mem_gb, 1)},
        "e2e": {"value": round(world * tokens / ms_e2e * 1e3, 1), "unit": "tokens/s", "ms_per_step": round(ms_e2e, 2),
                "h2d_bytes_per_step": batch_bytes(host), "d2h_bytes_per_step": 4, "last_loss": last_loss},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
    }
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------ reference
def run_reference(args):
    """The reference's own CPU path for this metric: the oracle port (the reference cannot be pip-installed here:
    its pyproject pins torch 2.5.1 / tensorflow 2.15 / timm and there is no index)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = []
    res = None
    for i in range(args.warmup + args.steps):
        res = cpu_baseline(sample_layers=1)
        if i >= args.warmup:
            times.append(res["seconds_measured"])
    S = 548
    per_step = sum(times) / len(times) * L      # extrapolated full-depth step of one sequence
    value = S / per_step
    use_pc, use_tac, _, desc = WORKLOADS[args.workload]
    out = {"impl": "reference", "metric": "multimodal_tokens_per_sec", "value": round(value, 3), "unit": "tokens/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(per_step * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{args.workload}: {desc}", "note": "CPU: 1 sequence per step, 1-layer slice x32"},
           "cpu_baseline": {"value": round(value, 3), "unit": "tokens/s", "cores": res["cores"], "kind": "port",
                            "sample": res["sample"]},
           "e2e": {"value": round(value, 3), "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: 8; 1 for cfg5)")
    ap.add_argument("--save-level", default="auto", choices=["auto", "layer", "mlp", "none"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stage", default="", choices=["", "pretrain", "finetune", "post-training"],
                    help="freeze_backbones stage (default: finetune; post-training for cfg5)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
