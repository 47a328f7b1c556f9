"""Data-parallel training strategy for the drop-in MLA module: DDP replicas + NCCL gradient all-reduce + fused AdamW.

The reference trains with FSDP only (training/strategies/fsdp.py); north_star asks for plain data-parallel replicas
with gradient all-reduce over NVLink and no parameter all-gather on the hot path.  This strategy keeps the
reference's optimizer semantics — AdamW(lr, weight decay on >=2-D non-bias params only, fsdp.py:242-257),
clip_grad_norm_(max_grad_norm) before the step (:310, base_strategy_mla.py:372), constant or warmup+cosine LR — and
its step order (forward, backward, clip, step, zero_grad), with:
  * ONE NCCL all-reduce per decoder layer (its gradient arenas are views of one flat fp32 buffer), issued from that
    layer's backward so it overlaps the rest of backward; one more for the small modules' gradients (flattened);
    fp32 by default as the reference reduces (`reduce_in_full_precision`, fsdp.py:184-187), bf16 on request;
  * gradient accumulation: `with trainer.no_sync():` around every micro-batch but the last defers the exchange (the
    reference's grad_accumulation_steps, base_strategy_mla.py:100,:366-379);
  * the global norm, clip coefficient, 1/world averaging and the AdamW update computed on the device (no host sync),
    the update refreshing the bf16 compute copies in the same pass.
Samples are independent (per-replica InfoNCE negatives and BatchNorm statistics, as in the reference), so there is no
other exchange.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import check
from .llama import LlamaDecoderLayer, side_stream

# AdamW of the decoder layers on the side stream (MLA_ADAM_STREAM=0 turns it off).  The update is HBM-bound and layer
# k's new weights are first needed by layer k's forward, so it runs next to the NEXT step's forward GEMMs, each layer's
# forward waiting on that layer's event only.  On the power-capped B200 the gain is small — the update's HBM power comes
# out of the GEMMs' clock budget: 647.3 -> 641.0 ms per step on one GPU, 700.2 -> 690.6 ms on two
# (profiles/r02_adam_overlap.json; round 1 measured it slower with the one-CTA GEMM) — but consistent, so it is on.
# A small-footprint launch shape (MLA_ADAM_LEAN=1: one 128-thread CTA per SM, fits beside a persistent GEMM CTA) was
# measured too and is slower than letting the full grid run (645.4).
ADAM_OVERLAP = {"on": __import__("os").environ.get("MLA_ADAM_STREAM", "1") == "1"}
ADAM_LEAN_CTAS = int(__import__("os").environ.get("MLA_ADAM_LEAN", "0"))      # CTAs of 128 threads per SM under overlap (0 = full grid)


class DataParallelTrainer:
    def __init__(self, model: torch.nn.Module, lr: float = 2e-5, weight_decay: float = 0.0,
                 max_grad_norm: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8,
                 lr_scheduler_type: str = "constant", warmup_steps: int = 0, total_steps: Optional[int] = None,
                 process_group=None, reduce_dtype: Optional[torch.dtype] = None, broadcast_from_rank0: bool = True,
                 error_check_interval: int = 50):
        if lr_scheduler_type not in ("constant", "linear-warmup+cosine-decay"):
            # the reference's two schedules (training/strategies/fsdp.py:236,:262,:286)
            raise ValueError(f"Learning Rate Schedule with type `{lr_scheduler_type}` is not supported!")
        if lr_scheduler_type != "constant" and not total_steps:
            raise ValueError("linear-warmup+cosine-decay needs total_steps (the reference derives it from the dataset size)")
        self.model = model
        self.lr, self.weight_decay, self.max_grad_norm = lr, weight_decay, max_grad_norm
        self.betas, self.eps = betas, eps
        self.lr_scheduler_type, self.warmup_steps, self.total_steps = lr_scheduler_type, warmup_steps, total_steps
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.step_count = 0
        self._handles: List = []
        self._sync = True                 # False inside no_sync(): backward accumulates locally, no exchange
        self._reduced: set = set()        # ids of layers whose arenas already hold rank sums (this step)
        self._others_flat = None
        import os as _os
        env_rd = _os.environ.get("MLA_GRAD_REDUCE_DTYPE", "")
        self.reduce_dtype = reduce_dtype or (torch.bfloat16 if env_rd == "bf16" else torch.float32)
        self.error_check_interval = error_check_interval
        self._exch_bytes = 0
        self.layers: List[LlamaDecoderLayer] = [m for m in model.modules() if isinstance(m, LlamaDecoderLayer)]
        layer_params = set()
        for l in self.layers:
            for p in l._masters():
                layer_params.add(id(p))
            if self.world > 1:
                l._grad_ready_cb = self._layer_grads_ready
            else:
                l._want_gnorm2 = True     # single replica: the wgrad GEMMs accumulate the gradient norm themselves
        import os
        if self.world > 1 and torch.cuda.is_available() and os.environ.get("MLA_FORCE_DYN", "1") != "0":
            # the all-reduces share the SMs with backward: let the persistent GEMMs claim tiles dynamically so that
            # CTAs whose SM is held by an NCCL kernel do not stall the whole GEMM
            ops.DYNAMIC_TILES["on"] = True
        # (name, param, decay?) for everything trainable outside the decoder layers
        self.other = [(n, p, not (p.ndim <= 1 or n.endswith(".bias"))) for n, p in model.named_parameters()
                      if p.requires_grad and id(p) not in layer_params]
        self.state: Dict[int, tuple] = {}
        dev = next(model.parameters()).device
        if self.world > 1 and broadcast_from_rank0:
            # what torch DDP does at construction: every replica starts from rank 0's parameters and buffers
            # (BatchNorm running statistics of the point tokenizer included)
            with torch.no_grad():
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t.data, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
                    if isinstance(t, torch.nn.Parameter):
                        torch.autograd.graph.increment_version(t)
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._scale = torch.zeros(2, dtype=torch.float32, device=dev)
        # model.state_dict() (a user's own checkpoint code) must not read weights the side-stream optimizer is still writing
        if hasattr(model, "register_state_dict_pre_hook"):
            model.register_state_dict_pre_hook(lambda *_a, **_k: self.synchronize())

    # ------------------------------------------------------------------ gradient exchange
    def no_sync(self):
        """Context manager for gradient accumulation: backward passes inside it add into the local gradient arenas
        without any exchange; the first backward outside it reduces the accumulated sums (torch DDP's no_sync)."""
        trainer = self

        class _NoSync:
            def __enter__(self_):
                self_.prev, trainer._sync = trainer._sync, False

            def __exit__(self_, *exc):
                trainer._sync = self_.prev
                return False
        return _NoSync()

    def _all_reduce(self, flat: torch.Tensor) -> None:
        """In-place SUM over ranks of a flat fp32 buffer, asynchronously (handle kept)."""
        self._exch_bytes += flat.numel() * (2 if self.reduce_dtype == torch.bfloat16 else 4)
        if self.reduce_dtype == torch.bfloat16:
            # opt-in: halves the bytes on NVLink; the sum is formed in bf16 (the reference's default is fp32)
            lo = ops.cast_bf16(flat) if flat.is_cuda else flat.to(torch.bfloat16)   # (CPU: gloo tests of the host logic)
            h = dist.all_reduce(lo, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self._handles.append((h, lo, flat))
        else:
            self._handles.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True), None, None))

    def _layer_grads_ready(self, layer: LlamaDecoderLayer) -> None:
        """Called at the end of a decoder layer's backward: unless accumulating (no_sync), its arenas are final — reduce
        them (one collective per layer) while earlier layers are still running backward."""
        if id(layer) in self._reduced:
            raise RuntimeError(
                "a decoder layer ran backward again after its gradients were already all-reduced in this step: wrap every "
                "micro-batch but the last in `with trainer.no_sync():` (gradient accumulation) and call step() after the last")
        if not self._sync:
            return
        self._reduced.add(id(layer))
        self._all_reduce(layer._gflat)

    def _reduce_others(self) -> None:
        """The small modules' autograd gradients, flattened into one buffer -> one collective -> scattered back."""
        gs = [p.grad for _, p, _ in self.other if p.grad is not None]
        if not gs:
            return
        flat = torch.cat([g.reshape(-1).float() for g in gs])
        self._all_reduce(flat)
        self._others_flat = (flat, gs)

    def current_lr(self) -> float:
        """LR of the optimizer step being taken.  transformers' get_cosine_schedule_with_warmup evaluates its lambda at
        the number of COMPLETED scheduler steps, so the k-th optimizer step (k = step_count, 1-based) runs at lambda(k-1):
        the very first step has lr 0 when there is a warm-up (fsdp.py:258-260 also zeroes the initial lr)."""
        if self.lr_scheduler_type == "constant":
            return self.lr
        s = max(0, self.step_count - 1)
        if s < self.warmup_steps:
            return self.lr * s / max(1, self.warmup_steps)
        prog = (s - self.warmup_steps) / max(1, self.total_steps - self.warmup_steps)
        return self.lr * max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))

    # ------------------------------------------------------------------ optimizer step
    def _adam(self, param: torch.nn.Parameter, g: torch.Tensor, decay: bool, bf16_dst: Optional[torch.Tensor],
              lr: float):
        p = param.data
        st = self.state.get(id(param))
        if st is None:
            st = (torch.zeros_like(p, memory_format=torch.contiguous_format),
                  torch.zeros_like(p, memory_format=torch.contiguous_format))
            self.state[id(param)] = st
        check(_lib.lib().mla_adamw_f32(
            ops._p(p), ops._p(g), ops._p(st[0]), ops._p(st[1]), ops._p(bf16_dst), C.c_int64(p.numel()), C.c_float(lr),
            C.c_float(self.betas[0]), C.c_float(self.betas[1]), C.c_float(self.eps),
            C.c_float(self.weight_decay if decay else 0.0), C.c_int64(self.step_count), ops._p(self._scale),
            ops._stream()))

    def exchange(self) -> None:
        """Finish the gradient exchange of this step: reduce the small modules' gradients and wait for every
        outstanding all-reduce (decoder layers were issued from their backward).  Gradients hold rank SUMS afterwards;
        the 1/world averaging is folded into the optimizer's gradient scale.  Parameters without a gradient (lm_head
        in diffusion mode, unused tokenizer parameters) are skipped consistently on every rank."""
        if self.world > 1:
            if not self._sync:
                raise RuntimeError("exchange()/step() called inside no_sync(): leave the context for the last micro-batch")
            for l in self.layers:       # layers whose last backward ran under no_sync (or none at all this call)
                if not l._grads_fresh and id(l) not in self._reduced and l._gflat is not None:
                    self._reduced.add(id(l))
                    self._all_reduce(l._gflat)
            self._reduce_others()
            for h, lo, dst in self._handles:
                h.wait()
                if lo is not None:      # bf16 reduce: widen the rank sum back into the fp32 arena
                    dst.copy_(lo)
            self._handles.clear()
            if self._others_flat is not None:
                flat, gs = self._others_flat
                off = 0
                for g in gs:
                    g.copy_(flat[off:off + g.numel()].view_as(g))
                    off += g.numel()
                self._others_flat = None

    def step(self) -> None:
        """clip_grad_norm_ + AdamW.step + zero_grad, after loss.backward()."""
        lib, s = _lib.lib(), ops._stream()
        cuda = self._sumsq.is_cuda
        main = torch.cuda.current_stream() if cuda else None
        side = side_stream(self._sumsq.device) if cuda else None
        self.exchange()
        if cuda:
            main.wait_stream(side)      # weight-gradient GEMMs issued on the side stream (llama.OVERLAP)
        self.step_count += 1
        lr = self.current_lr()
        # global norm over every gradient that exists (params without grad are skipped, as torch does)
        self._sumsq.zero_()
        for l in self.layers:
            if l._grads_fresh:
                continue          # no backward reached this layer since the last step
            g = l._gflat
            if self.world == 1 and l._gnorm2_valid:
                # the four weight-gradient matrices were normed by the GEMM epilogues that wrote them; only the two
                # RMSNorm weight gradients (the arena's tail, accumulated atomically) are left
                self._sumsq.add_(l._gnorm2)
                tail = g[g.numel() - 2 * l.hidden_size:]
                check(lib.mla_sumsq_f32(ops._p(tail), C.c_int64(tail.numel()), ops._p(self._sumsq), s))
            else:
                check(lib.mla_sumsq_f32(ops._p(g), C.c_int64(g.numel()), ops._p(self._sumsq), s))
        for _, p, _ in self.other:
            if p.grad is not None:
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                check(lib.mla_sumsq_f32(ops._p(g), C.c_int64(g.numel()), ops._p(self._sumsq), s))
        check(lib.mla_clip_coef(ops._p(self._sumsq), C.c_float(self.max_grad_norm or 0.0), C.c_float(1.0 / self.world),
                                ops._p(self._scale), s))
        overlap = cuda and ADAM_OVERLAP["on"]
        for l in self.layers:           # Adam state is allocated (first step) on the main stream
            if not l._grads_fresh:
                for p in l._masters():
                    if p.requires_grad and id(p) not in self.state:
                        self.state[id(p)] = (torch.zeros_like(p.data, memory_format=torch.contiguous_format),
                                             torch.zeros_like(p.data, memory_format=torch.contiguous_format))
        if overlap:
            side.wait_stream(main)
            # small-footprint launches: one 128-thread CTA per SM fits beside a persistent GEMM CTA (registers)
            lib.mla_adamw_set_lean(C.c_int32(ADAM_LEAN_CTAS))
        with torch.cuda.stream(side) if overlap else _nullctx():
            for l in self.layers:
                if l._grads_fresh:
                    continue
                h, f = l.hidden_size, l.inter
                wqkv, wo, wgu, wd, l1, l2 = l.compute_weights()
                dsts = [wqkv[:h], wqkv[h:2 * h], wqkv[2 * h:], wo, wgu[:f], wgu[f:], wd, l1, l2]
                for p, g, dst in zip(l._masters(), l._views, dsts):
                    if p.requires_grad:
                        self._adam(p, g, p.ndim > 1, dst, lr)
                l.mark_grads_fresh()
                if overlap:
                    l._weights_ready = side.record_event()
        if overlap:
            lib.mla_adamw_set_lean(C.c_int32(0))
        for _, p, decay in self.other:
            if p.grad is None:
                continue
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            self._adam(p, g, decay, None, lr)
            torch.autograd.graph.increment_version(p)      # invalidates the cached bf16 compute copy
            p.grad = None
        self._reduced.clear()
        if self.error_check_interval and self.step_count % self.error_check_interval == 0:
            # device-side error flags (a sample without EOS/tag token: the reference raises IndexError,
            # prismatic.py:983) are read back every N steps so the hot path stays free of host syncs
            chk = getattr(getattr(self.model, "vlm", None), "check_errors", None)
            if chk is not None:
                chk()

    def exchange_stats(self) -> Optional[dict]:
        """Bytes handed to NCCL per step so far (for the bench's busbw figure); None on a single replica."""
        if self.world == 1 or self.step_count == 0:
            return None
        per_step = self._exch_bytes / max(1, self.step_count)
        return {"collectives_per_step": len(self.layers) + 1, "bytes_per_step": int(per_step),
                "reduce_dtype": str(self.reduce_dtype).replace("torch.", ""),
                "busbw_factor": round(2.0 * (self.world - 1) / self.world, 4)}

    def synchronize(self) -> None:
        """Make the current stream wait for optimizer work still running on the side stream (MLA_ADAM_STREAM): call it
        before reading or writing parameters / optimizer state outside the model's own forward (checkpoints do)."""
        if self._sumsq.is_cuda:
            torch.cuda.current_stream().wait_stream(side_stream(self._sumsq.device))

    def grad_norm(self) -> torch.Tensor:
        """Mean-gradient global norm of the last step (device scalar)."""
        return self._scale[1]

    # ------------------------------------------------------------------ checkpoints (training/strategies/fsdp.py:100-160)
    def save_checkpoint(self, run_dir, global_step: int, epoch: int, train_loss: Optional[float] = None,
                        only_trainable: bool = True, save_optimizer: bool = True):
        """Writes `<run_dir>/checkpoints/step-XXXXXX-epoch-XX-loss=Y.pt` = {"model": {module_key: state_dict}} with the
        reference's module keys ("vlm." stripped, fsdp.py:133-136) and parameter names, so MLA.from_pretrained /
        load_from_checkpoint of the reference read it.  Every data-parallel replica holds the full model: rank 0 writes
        its own copy, no gather.  With save_optimizer the AdamW moments and the step counter go to the `.optimizer`
        file next to it (the path the reference reserves, :158-160, but never writes): resume is exact."""
        self.synchronize()
        from collections import OrderedDict
        from pathlib import Path
        model = self.model
        keys = model.trainable_module_keys if only_trainable else model.all_module_keys
        full = model.state_dict()
        out = {k: OrderedDict() for k in keys}
        for name, v in full.items():
            for k in keys:
                if name.startswith(k + "."):
                    out[k][name[len(k) + 1:]] = v.detach().to("cpu", copy=True)
        out = {(k[4:] if k.startswith("vlm.") else k): v for k, v in out.items()}
        ckpt_dir = Path(run_dir) / "checkpoints"
        loss = "inf" if train_loss is None else f"{train_loss:.4f}"
        path = ckpt_dir / f"step-{global_step:06d}-epoch-{epoch:02d}-loss={loss}.pt"
        rank0 = self.world == 1 or dist.get_rank(self.pg) == 0
        if rank0:
            ckpt_dir.mkdir(parents=True, exist_ok=True)
            torch.save({"model": out}, path)
            if save_optimizer:
                names = {id(p): n for n, p in model.named_parameters()}
                opt = {names[i]: (m.detach().cpu(), v.detach().cpu()) for i, (m, v) in self.state.items() if i in names}
                torch.save({"optimizer": {"state": opt, "step": self.step_count, "lr": self.lr, "betas": self.betas,
                                          "eps": self.eps, "weight_decay": self.weight_decay},
                            "scheduler": {"epoch": epoch, "global_step": global_step}}, path.with_suffix(".optimizer"))
        if self.world > 1:
            dist.barrier(group=self.pg)
        return path

    def load_checkpoint(self, path, load_optimizer: bool = True) -> dict:
        """Inverse of save_checkpoint (also reads checkpoints written by the reference's FSDPStrategy: same layout).
        Returns the scheduler record ({"epoch", "global_step"}) when an optimizer file was found, else {}."""
        self.synchronize()
        from pathlib import Path
        path = Path(path)
        blob = torch.load(path, map_location="cpu", weights_only=True)["model"]
        sd = {}
        for mkey, sub in blob.items():
            for name, v in sub.items():
                sd[f"vlm.{mkey}.{name}"] = v
        missing, unexpected = self.model.load_state_dict(sd, strict=False)
        if unexpected:
            raise KeyError(f"checkpoint holds parameters this model does not have: {unexpected[:5]}")
        for p in self.model.parameters():            # loaded in place: invalidate the cached bf16 compute copies
            torch.autograd.graph.increment_version(p)
        opt_path = path.with_suffix(".optimizer")
        if not (load_optimizer and opt_path.exists()):
            return {}
        rec = torch.load(opt_path, map_location="cpu", weights_only=True)
        named = dict(self.model.named_parameters())
        self.state.clear()
        for name, (m, v) in rec["optimizer"]["state"].items():
            p = named[name]
            self.state[id(p)] = (m.to(p.device).contiguous(), v.to(p.device).contiguous())
        self.step_count = int(rec["optimizer"]["step"])
        return rec.get("scheduler", {})


class _nullctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def plan_save_levels(model: torch.nn.Module, tokens: int, reserve_gb: float = 10.0) -> List[str]:
    """Pick, per decoder layer, how much to keep for backward so that activations fit next to the parameters,
    gradients and the (not yet allocated) Adam state: "none" (no GEMM recompute) where memory allows, then "mlp",
    falling back to full-layer recompute ("layer", what the reference's activation checkpointing does)."""
    layers = [m for m in model.modules() if isinstance(m, LlamaDecoderLayer)]
    if not layers or not torch.cuda.is_available():
        return ["layer"] * len(layers)
    free, _total = torch.cuda.mem_get_info()
    n_train = sum(p.numel() for p in model.parameters() if p.requires_grad)
    n_layer_params = sum(p.numel() for l in layers for p in l._masters())
    pending = 8 * n_train                       # Adam m, v (fp32)
    pending += (4 + 2) * n_layer_params         # gradient arenas + bf16 compute copies, allocated lazily
    pending += 4 * (n_train - n_layer_params)   # autograd gradients of the small modules
    h, f = layers[0].hidden_size, layers[0].inter
    work = 2 * tokens * (8 * f + 12 * h)        # transient buffers of one layer's backward
    budget = free - pending - work - int(reserve_gb * 2 ** 30)
    cost = {"layer": 2 * tokens * h, "mlp": 2 * tokens * 6 * h, "none": 2 * tokens * (6 * h + 2 * f)}
    levels = ["layer"] * len(layers)
    budget -= cost["layer"] * len(layers)
    for target in ("mlp", "none"):
        prev = "layer" if target == "mlp" else "mlp"
        for i in range(len(layers)):
            extra = cost[target] - cost[prev]
            if levels[i] == prev and budget >= extra:
                levels[i] = target
                budget -= extra
    return levels
