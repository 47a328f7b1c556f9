#!/usr/bin/env python
"""Benchmark of the MLA training step (BASELINE.json metric: multimodal tokens/s, Llama-2-7B MLA, bf16).

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path (one process per GPU under torchrun for N>1)
    python bench.py --impl reference ...                     # the UNMODIFIED reference's own code on the host CPU cores

A step = forward + backward + gradient all-reduce (N>1) + clip + AdamW of one synthetic batch, per-GPU batch 8 x 4
diffusion repeats = 32 sequences of 548 tokens (17,536 multimodal tokens per GPU per step), random-init weights.
Default workload: the FULL configuration at every N, so that the 1/2/4/8-GPU series BASELINE.json's metric names is one
weak-scaling series — N = 1 -> "cfg3" = BASELINE configs[2] (image + point cloud + tactile + both InfoNCE losses +
diffusion head, per-GPU batch 8), N > 1 -> "cfg4" = BASELINE configs[3] (the same flags and per-GPU batch under DDP,
global batch 8 N).  On one GPU the same run also measures BASELINE configs[1] ("cfg2": image-only) and cfg2 with the
shared decoder prefix and reports them under "also".  `--workload` overrides (e.g. `--workload cfg2`).
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
    # BASELINE configs[3]: the full configuration (use_pointcloud + contrastive + diffusion head) under DDP, global bs = 8 N
    "cfg4": (True, True, True, "MLA Llama2-7B full (pointcloud+tactile+contrastive+diffusion head), DDP, per-GPU bs=8 x 4 repeats"),
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
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_gemm12_summary.json")))["gemm"]
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
            "the same 12 launches: profiles/r02_ncu_gemm12_summary.json)",
            "algorithmic_bytes_per_launch": round(sum(2.0 * (m * k + n * k) + (2.0 if i < 8 else 4.0) * m * n
                                                      for i, (m, n, k) in enumerate(GEMM_SHAPES(T))) / 12),
            "kernel": "gemm2_bf16_kernel (tcgen05 cta_group::2, 12 GEMM shapes of one decoder layer fwd+bwd)",
            "launch_ms_avg": round(ms / len(calls), 4)}


def reference_cpu(steps: int, warmup: int, layers_cpu: int = 32, threads: int = 0, workload: str = "cfg2"):
    """The reference's OWN code on the host cores — `MLA.forward` + backward of the UNMODIFIED reference
    (models/mla/model_mla.py:118, imported from baseline/_ref or /root/reference through oracle/ref_shim.py; vendored
    transformers 4.40.1 Llama with SDPA attention: flash-attn has no CPU path), fp32 parameters under
    autocast(cpu, bf16) as scripts/train.py builds them, all host threads, the full 32-layer Llama-2-7B-shaped model.
    Bounded sample of the workload: ONE of the step's 32 sequences (per-GPU batch 1 x 1 diffusion repeat = 548 fused
    tokens) per timed step; every timed step is executed in full and value = 548 / measured seconds (no extrapolation
    at the default depth).  `layers_cpu` < 32 (hosts short of the ~60 GB the fp32 model + gradients need) runs the
    first layers only and normalises the rate by depth, which is then said in `sample`.
    Falls back to the oracle port when no reference tree is present."""
    import contextlib
    import io
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle import ref_shim
    S = 548
    if not ref_shim.available():
        return _port_cpu(steps, warmup, threads)
    from oracle.ref_model import build_reference_7b, ref_call
    from mla_b200.synthetic import make_batch
    use_pc = WORKLOADS[workload][0]
    if use_pc and layers_cpu < 9:
        raise SystemExit("the full configuration's InfoNCE loss reads hidden_states[8] (modeling_llama.py:1274): "
                         "--layers-cpu must be at least 9 (or pass --workload cfg2)")
    quiet = contextlib.redirect_stdout(io.StringIO())        # the reference prints its loss dict every forward
    t0 = time.perf_counter()
    with quiet, contextlib.redirect_stderr(io.StringIO()):
        mla, _ = build_reference_7b(workload, layers_cpu, torch.float32, "cpu", "sdpa")
    t_build = time.perf_counter() - t0
    batch = make_batch(1, 32, 0, 672, 1024, seed=1234, use_pointcloud=use_pc, use_tactile=use_pc)
    params = [p for p in mla.parameters() if p.requires_grad]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        with quiet:
            loss = ref_call(mla, batch, "cpu", repeats=1)
        loss.backward()
        for p in params:
            p.grad = None
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    value = S * (layers_cpu / L) / per_step
    depth = ("the full 32-layer model" if layers_cpu == L else
             f"{layers_cpu} of {L} decoder layers at 7B width, rate normalised by depth ({layers_cpu}/{L})")
    return {"value": round(value, 3), "unit": "tokens/s", "cores": threads, "kind": "reference",
            "sample": f"unmodified reference MLA.forward+backward on CPU (fp32 params, autocast bf16, SDPA), 1 sequence x "
                      f"{S} tokens (1 of the step's 32) through {depth}, {len(times)} timed steps of {per_step:.2f} s each",
            "seconds_per_sample_step": round(per_step, 3), "layers_cpu": layers_cpu, "tokens_per_sample_step": S,
            "build_seconds": round(t_build, 1)}


def _port_cpu(steps: int, warmup: int, threads: int, sample_layers: int = 2):
    """Fallback when the reference tree is absent: the oracle port (fp32 restatement) of the decoder."""
    from oracle import llama as O
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
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        hs = O.decoder(x, layers, norm, HEADS, 1e-5, None)
        hs[-1].square().mean().backward()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    return {"value": round(S * (sample_layers / L) / per_step, 3), "unit": "tokens/s", "cores": threads, "kind": "port",
            "sample": f"oracle port (fp32 restatement) of a {sample_layers}-layer slice of the 7B decoder, 1 sequence x {S} "
                      f"tokens, {len(times)} timed steps of {per_step:.2f} s, rate normalised by depth",
            "seconds_per_sample_step": round(per_step, 3), "layers_cpu": sample_layers, "tokens_per_sample_step": S}


def cpu_baseline_subprocess(workload: str):
    """cpu_baseline leg of our arm: the reference arm's sample (1 warm-up + 2 timed full-depth steps, ~20 s of CPU work
    on the GPU box's 16 cores) in a child process, so the import shim of the reference (it shadows `transformers`) never
    shares an interpreter with the product path."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--workload", workload]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="0", WORLD_SIZE="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    for line in reversed(r.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)["cpu_baseline"]
    raise RuntimeError((r.stderr or r.stdout)[-300:])


def reference_gpu_record(workload: str = "cfg3"):
    """R-GPU: the unmodified reference (PyTorch + flash-attn 2.8.3) on one B200 at the same per-GPU workload, measured by
    tools/ref_gpu.py this round and committed under profiles/ — a recorded number, not re-measured in this run."""
    try:
        name = f"r02_ref_gpu_{workload}.json"
        if not os.path.exists(os.path.join(ROOT, "profiles", name)):
            name = "r02_ref_gpu_cfg2.json"
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        runs = {r["variant"].split(":")[0].split(",")[0] + (" ckpt" if r.get("activation_checkpointing") else " no-ckpt"): r
                for r in d["runs"]}
        full = next(r for r in d["runs"] if "step_ms" in r)
        fb = next(r for r in d["runs"] if "fwd_bwd_ms" in r and r["activation_checkpointing"])
        return {"source": f"profiles/{name} (tools/ref_gpu.py step, recorded, same B200 pool)",
                "workload": d["workload"], "attn": f"flash_attn {d['flash_attn']}", "step_ms": full["step_ms"],
                "tokens_per_s": full["tokens_per_s"], "fwd_ms": fb["fwd_ms"], "fwd_bwd_ms_checkpointed": fb["fwd_bwd_ms"],
                "fwd_bwd_ms_no_checkpointing": next((r["fwd_bwd_ms"] for r in d["runs"] if "fwd_bwd_ms" in r
                                                     and not r["activation_checkpointing"]), None)}
    except Exception:
        return None


def measure_ours(args, workload: str, world: int, rank: int, local: int, with_breakdown: bool, share_prefix: bool = False):
    """Builds the drop-in module for `workload`, runs warm-up + timed steps (device-resident and end-to-end), frees it.
    Returns the measurements of this workload (identical on every rank: times are max-reduced over ranks)."""
    import torch.distributed as dist
    from mla_b200 import _lib
    from mla_b200.synthetic import batch_bytes, make_batch, map_tensors
    from mla_b200.trainer import DataParallelTrainer, plan_save_levels

    use_pc, use_tac, _, desc = WORKLOADS[workload]
    B, R, Lt, T = (args.batch or DEFAULT_BATCH.get(workload, 8)), 4, 32, 0
    views = EXTRA_VIEWS.get(workload, 0)
    mla = build_model(workload, T, args.stage)
    mla.share_diffusion_prefix = share_prefix
    trainer = DataParallelTrainer(mla, lr=2e-5, weight_decay=0.0, max_grad_norm=1.0)
    S = 1 + 256 + 256 * (1 + views) + 1 + (Lt - 1) + 1 + 1 + (T + 1)
    tokens = B * R * S
    decoder_rows = B * (S + (R - 1) * (T + 3)) if share_prefix else tokens
    # stage pretrain keeps the tokenizers' pre-BatchNorm activations through the decoder (7.6 GB for 32 clouds); their
    # backward transients come after the decoder's activations are gone
    reserve = 10.0 + (9.0 if args.stage == "pretrain" else 0.0)
    levels = [args.save_level] * L if args.save_level != "auto" else plan_save_levels(mla, decoder_rows, reserve_gb=reserve)
    mla.vlm.llm_backbone.llm.model.set_save_levels(levels)

    host = make_batch(B, Lt, T, 672, 1024, seed=1234 + rank, use_pointcloud=use_pc, use_tactile=use_tac, pin=True,
                      extra_views=views, generation=workload in GENERATION)
    devb = map_tensors(host, lambda t: t.cuda(non_blocking=True))
    kw = dict(camera_name="rlbench_front", repeated_diffusion_steps=R, use_diff=True)

    def fwd_loss(b):
        loss_dict, _ = mla(input_ids=b["input_ids"], attention_mask=b["attention_mask"], labels=b["labels"],
                           actions=b["actions"], images=b["images"], point_cloud=b.get("point_cloud"),
                           tactile=b.get("tactile"), proprio=b["proprio"], gripper_xyz=b.get("gripper_xyz"),
                           action_masks=b["action_masks"], next_images=b.get("next_images"),
                           next_point_cloud=b.get("next_point_cloud"), next_tactile=b.get("next_tactile"), **kw)
        return loss_dict["total_loss"]

    def call(b):
        loss = fwd_loss(b)
        loss.backward()
        trainer.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

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
        return reduce_max(e0.elapsed_time(e1)) / n, last

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
        return reduce_max(e0.elapsed_time(e1)) / n

    ms_fwd = ms_fwd_bwd = None
    if world == 1 and with_breakdown:       # per-GPU quantities: measured on the single-GPU run only
        try:
            ms_fwd = round(time_fn(fwd_only, max(2, args.steps // 2)), 2)
            ms_fwd_bwd = round(time_fn(fwd_bwd, max(2, args.steps // 2)), 2)
        except Exception as ex:             # secondary numbers must never take the headline line down
            sys.stderr.write(f"fwd / fwd+bwd timing failed: {ex}\n")
            ms_fwd = ms_fwd_bwd = None
    clocks = sampler.stop() if rank == 0 else None
    mla.vlm.check_errors()
    exch = trainer.exchange_stats() if hasattr(trainer, "exchange_stats") else None
    res = dict(workload=workload, desc=desc, B=B, R=R, S=S, tokens=tokens, decoder_rows=decoder_rows, levels=levels,
               ms_dev=ms_dev, ms_e2e=ms_e2e,
               last_loss=last_loss, launches=int(launches), ms_fwd=ms_fwd, ms_fwd_bwd=ms_fwd_bwd, clocks=clocks,
               mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30, h2d=batch_bytes(host), exchange=exch,
               unused_params=sum(1 for p in mla.parameters() if p.requires_grad and id(p) not in layer_ids
                                 and id(p) not in trainer.state))
    # free the training state (the next workload / the isolated kernel timing need the memory)
    del trainer, mla, model_, small, devb, host, loss
    import gc
    gc.collect()
    from mla_b200 import ops
    ops.clear_bf16_cache()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
    return res


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mla_b200 import _lib
    _lib.check(_lib.lib().mla_device_check())

    workload = args.workload or ("cfg3" if world == 1 else "cfg4")
    m = measure_ours(args, workload, world, rank, local, with_breakdown=True, share_prefix=args.share_prefix)
    also = None
    if world == 1 and not args.workload and not args.no_also:
        try:        # BASELINE configs[1] on the same GPU in the same run (secondary: must never take the headline down)
            a = measure_ours(args, "cfg2", world, rank, local, with_breakdown=False)
            also = {"cfg2": {"workload": f"cfg2: {a['desc']}", "ms_per_step": round(a["ms_dev"], 2),
                             "value": round(a["tokens"] / a["ms_dev"] * 1e3, 1), "unit": "tokens/s",
                             "e2e_value": round(a["tokens"] / a["ms_e2e"] * 1e3, 1), "e2e_ms_per_step": round(a["ms_e2e"], 2),
                             "h2d_bytes_per_step": a["h2d"], "gpu_launches": a["launches"], "last_loss": a["last_loss"],
                             "peak_mem_gb": round(a["mem_gb"], 1),
                             "step_frac_of_peak": round(step_flops(a["tokens"], a["S"]) / a["ms_dev"] / 1e9 / peaks()[2], 4)}}
        except Exception as ex:
            sys.stderr.write(f"cfg2 measurement failed: {ex}\n")
            also = {"cfg2": {"error": str(ex)[:200]}}
        try:        # the same cfg2 step with the 4 diffusion copies sharing one decoder prefix per sample
            a = measure_ours(args, "cfg2", world, rank, local, with_breakdown=False, share_prefix=True)
            also["cfg2_shared_prefix"] = {
                "workload": "cfg2 with MLA.share_diffusion_prefix: same loss and gradients as the repeated batch "
                            "(tests/test_shared_prefix_gpu.py); the decoder runs one prefix per sample + 4 suffix groups",
                "ms_per_step": round(a["ms_dev"], 2), "value": round(a["tokens"] / a["ms_dev"] * 1e3, 1),
                "unit": "tokens/s (the step's 17,536 logical tokens)", "decoder_rows_per_step": a["decoder_rows"],
                "e2e_value": round(a["tokens"] / a["ms_e2e"] * 1e3, 1), "e2e_ms_per_step": round(a["ms_e2e"], 2),
                "gpu_launches": a["launches"], "last_loss": a["last_loss"], "peak_mem_gb": round(a["mem_gb"], 1),
                "note": "reported apart from the headline: the algorithmic-FLOP roofline counts the repeated batch"}
        except Exception as ex:
            sys.stderr.write(f"shared-prefix measurement failed: {ex}\n")
            also["cfg2_shared_prefix"] = {"error": str(ex)[:200]}
    busbw = None
    if world > 1:
        # measured NCCL all-reduce bus bandwidth at the message size the trainer uses (one decoder layer's flat fp32
        # gradient arena, 810 MB), alone on the wire: the reference point for the exchange's share of the step
        try:
            buf = torch.zeros(3 * H * H + H * H + 2 * F * H + H * F + 2 * H, dtype=torch.float32, device="cuda")
            for _ in range(2):
                dist.all_reduce(buf)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                dist.all_reduce(buf)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 5], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            busbw = {"message_bytes": buf.numel() * 4, "ms": round(float(t.item()), 3),
                     "busbw_gbs": round(2.0 * (world - 1) / world * buf.numel() * 4 / float(t.item()) / 1e6, 1)}
            del buf
        except Exception as ex:
            busbw = {"error": str(ex)[:120]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, tf_burst, tf_sust, src = peaks()
    tokens, S, ms_dev, ms_e2e = m["tokens"], m["S"], m["ms_dev"], m["ms_e2e"]
    roof = None
    try:
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
            cpu = cpu_baseline_subprocess("cfg2" if workload == "cfg2" else "cfg3")
        except Exception as ex:
            cpu = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"[:200]}
    levels = m["levels"]
    out = {
        "metric": "multimodal_tokens_per_sec", "value": round(world * tokens / ms_dev * 1e3, 1), "unit": "tokens/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_dev, 2),
        "fwd_ms": m["ms_fwd"], "fwd_bwd_ms": m["ms_fwd_bwd"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{workload}: {m['desc']}" + (" [shared decoder prefix across the diffusion copies]"
                                                             if args.share_prefix else ""),
                   "per_gpu_batch": m["B"], "repeated_diffusion_steps": m["R"], "decoder_rows_per_gpu_step": m["decoder_rows"],
                   "seq_len": S, "tokens_per_gpu_step": tokens, "global_batch": m["B"] * world, "parallelism": f"dp{world}",
                   "stage": ("pretrain (vision tokenizers trained)" if args.stage == "pretrain" else
                             "post-training (generation heads on)" if workload in GENERATION
                             else "finetune (vision tokenizers frozen)"), "optimizer": "AdamW fp32 master + fp32 grads",
                   "activation_save_levels": {lv: levels.count(lv) for lv in sorted(set(levels))},
                   "l2": "step streams >100 GB of weights/activations (>> 126 MB L2); no explicit flush needed",
                   "peak_mem_gb": round(m["mem_gb"], 1),
                   "trainable_params_without_gradient": m["unused_params"]},
        "e2e": {"value": round(world * tokens / ms_e2e * 1e3, 1), "unit": "tokens/s", "ms_per_step": round(ms_e2e, 2),
                "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": 4, "last_loss": m["last_loss"]},
        "gpu_launches": m["launches"], "clocks": m["clocks"], "roofline": roof, "cpu_baseline": cpu,
        "reference_gpu": reference_gpu_record("cfg2" if workload == "cfg2" else "cfg3"),
    }
    if m["exchange"]:
        out["gradient_exchange"] = dict(m["exchange"], allreduce_alone=busbw,
                                        nccl_env={k: v for k, v in os.environ.items() if k.startswith("NCCL_")})
    if also:
        out["also"] = also
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------ reference
def run_reference(args):
    """Reference arm: the UNMODIFIED reference's own CPU execution of the path on the host cores (see reference_cpu).
    Under torchrun only rank 0 runs; the other ranks exit 0 without work."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = args.workload or ("cfg3" if world == 1 else "cfg4")
    res = reference_cpu(args.steps, args.warmup, layers_cpu=args.layers_cpu, workload=workload)
    desc = WORKLOADS[workload][3]
    out = {"impl": "reference", "metric": "multimodal_tokens_per_sec", "value": res["value"], "unit": "tokens/s",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(res["seconds_per_sample_step"] * 1e3, 1), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "bf16 autocast over fp32 params (CPU)", "data": "synthetic",
           "config": {"workload": f"{workload}: {desc}",
                      "sample": f"ms_per_step is the measured time of one SAMPLE step ({res['tokens_per_sample_step']} tokens "
                                f"= 1 of the step's 32 sequences, {res['layers_cpu']}/{L} layers)"},
           "cpu_baseline": res,
           "e2e": {"value": res["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="", choices=[""] + sorted(WORKLOADS),
                    help="default: the full configuration (cfg3 on one GPU, + cfg2 under 'also'; cfg4 under DDP)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: 8; 1 for cfg5)")
    ap.add_argument("--save-level", default="auto", choices=["auto", "layer", "mlp", "none"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="single GPU: skip the secondary cfg3 / shared-prefix measurements")
    ap.add_argument("--share-prefix", action="store_true", help="headline run with MLA.share_diffusion_prefix (cfg2 only)")
    ap.add_argument("--layers-cpu", type=int, default=32, help="reference arm: decoder layers of the CPU sample (32 = full model)")
    ap.add_argument("--stage", default="", choices=["", "pretrain", "finetune", "post-training"],
                    help="freeze_backbones stage (default: finetune; post-training for cfg5)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
