"""Llama-2 decoder stack of the MLA hot path on libmla_b200 kernels.

Mirrors the reference's module tree and parameter names (transformers/models/llama/modeling_llama.py:
LlamaRMSNorm :76, LlamaMLP :211, LlamaFlashAttention2 :405, LlamaDecoderLayer :695, LlamaModel :912) so a
`state_dict()` is interchangeable, but the compute is one autograd node per decoder layer that launches our CUDA
kernels directly:

    rmsnorm -> [q|k|v] GEMM -> RoPE (in place) -> flash attention -> o GEMM (+residual)
            -> rmsnorm -> [gate|up] GEMM -> SwiGLU -> down GEMM (+residual)

Parameters stay fp32 `nn.Parameter`s under their reference names (scripts/train.py:306 asserts fp32); each layer
keeps a bf16 compute copy of its weights, fused as [q;k;v] and [gate;up], refreshed from the fp32 masters when they
change.  Weight gradients are written by the wgrad GEMMs straight into fp32 gradient arenas that `param.grad`
aliases (no autograd-side accumulation buffers).
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops

# Optional stream overlap of the backward pass (MLA_WGRAD_STREAM=1, default OFF).  The four weight-gradient GEMMs of
# a layer do not feed the activation-gradient chain, so they can be issued on a second ("side") stream and the
# HBM-bound kernels of the chain (SwiGLU/RMSNorm backward, RoPE) then run next to a wgrad GEMM on the SMs' spare
# registers.  MEASURED on B200 (profiles/r01_overlap_ab.txt): the step gets SLOWER (726 -> 790 ms) — the whole step is
# power-capped (sw_power_cap, ~1380 of 1965 MHz), so concurrent HBM traffic lowers the clocks the GEMMs run at.  Kept
# as a tested switch (tests/test_trainer_gpu.py proves the two orders are bit-identical), not as the default.
OVERLAP = {"wgrad": os.environ.get("MLA_WGRAD_STREAM", "0") == "1"}
_SIDE_STREAMS: dict = {}
FUSE_ROPE = {"on": os.environ.get("MLA_FUSE_ROPE", "1") == "1"}     # RoPE inside the q|k|v GEMM epilogue (head_dim 128)
# SwiGLU inside the gate|up GEMM epilogue (CTA-pair kernel; inter % 128 == 0): bit-identical to the projection followed
# by swiglu_fwd (tests/test_gemm2_gpu.py, validated on B200 in round 2).  MLA_FUSE_SWIGLU=0 restores the separate pass.
FUSE_SWIGLU = {"on": os.environ.get("MLA_FUSE_SWIGLU", "1") == "1"}
# SwiGLU BACKWARD inside the epilogue of the down projection's input-gradient GEMM (d_act is never written): same bits
# as the GEMM followed by swiglu_bwd_act.  MLA_FUSE_SWIGLU_BWD=0 restores the separate pass.
FUSE_SWIGLU_BWD = {"on": os.environ.get("MLA_FUSE_SWIGLU_BWD", "1") == "1"}
# Sum of squares of a layer's weight gradients accumulated by the weight-gradient GEMM epilogues themselves (single
# replica: the trainer then skips its 27.8 GB norm pass).  MLA_FUSE_GRAD_NORM=0 restores the separate pass.
FUSE_GRAD_NORM = {"on": os.environ.get("MLA_FUSE_GRAD_NORM", "1") == "1"}


# NVTX ranges per decoder layer and phase (MLA_NVTX=1): nsys / ncu --nvtx can then attribute kernels to
# layer.N.fwd / layer.N.bwd.mlp / layer.N.bwd.attn.  Off by default (two host calls per range).
NVTX = {"on": os.environ.get("MLA_NVTX", "0") == "1"}
# inference: all decoder layers of a denoise step in ONE persistent launch (csrc/decode_stack.cu) when batch*rows <= 2.
# Correct and tested, but measured SLOWER than the per-op path it was meant to replace (4.27 vs 3.23 ms per step at 7B:
# profiles/r02_decode_stack_trace.json), so it is opt-in.
DECODE_STACK = os.environ.get("MLA_DECODE_STACK", "0") == "1"


class _Range:
    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if NVTX["on"]:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if NVTX["on"]:
            torch.cuda.nvtx.range_pop()
        return False


def side_stream(device) -> "torch.cuda.Stream":
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=key)
        _SIDE_STREAMS[key] = st
    return st


# What a layer keeps for backward.
#   "layer": only its input; the whole layer is recomputed in backward (what the reference's FSDP activation
#            checkpointing does, training/strategies/fsdp.py:217-223)
#   "mlp"  : input, qkv, attention output, LSE, mid residual; only the gate/up GEMM (+SwiGLU) is recomputed
#   "none" : additionally keeps gate/up; nothing but the cheap norm/SwiGLU kernels is recomputed
SAVE_LEVELS = ("layer", "mlp", "none")


class LlamaRMSNorm(nn.Module):
    def __init__(self, hidden_size: int, eps: float = 1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.variance_epsilon = eps


class _Proj(nn.Module):
    """Parameter holder with nn.Linear's attribute layout (weight [out,in], no bias)."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))


class LlamaAttention(nn.Module):
    def __init__(self, hidden: int, heads: int):
        super().__init__()
        self.q_proj = _Proj(hidden, hidden)
        self.k_proj = _Proj(hidden, hidden)
        self.v_proj = _Proj(hidden, hidden)
        self.o_proj = _Proj(hidden, hidden)


class LlamaMLP(nn.Module):
    def __init__(self, hidden: int, inter: int):
        super().__init__()
        self.gate_proj = _Proj(hidden, inter)
        self.up_proj = _Proj(hidden, inter)
        self.down_proj = _Proj(inter, hidden)


@dataclass
class LayerShape:
    B: int
    S: int
    H: int
    D: int
    mask: Optional[torch.Tensor]      # uint8 [B,S] or None
    cos: torch.Tensor                 # bf16 [S, D/2]
    sin: torch.Tensor
    # shared-prefix layout: rows >= prefix_len[b] are `group`-row suffix groups (one per diffusion repeat)
    # that see the prefix and their own group; rope_pos maps every row to its rotary position
    prefix_len: Optional[torch.Tensor] = None     # int32 [B]
    group: int = 0
    rope_pos: Optional[torch.Tensor] = None       # int32 [B*S]

    @property
    def grouped(self):
        return (self.prefix_len, self.group) if self.group else None


class LlamaDecoderLayer(nn.Module):
    def __init__(self, hidden: int, inter: int, heads: int, eps: float, layer_idx: int):
        super().__init__()
        self.hidden_size, self.inter, self.heads, self.eps, self.layer_idx = hidden, inter, heads, eps, layer_idx
        self.self_attn = LlamaAttention(hidden, heads)
        self.mlp = LlamaMLP(hidden, inter)
        self.input_layernorm = LlamaRMSNorm(hidden, eps)
        self.post_attention_layernorm = LlamaRMSNorm(hidden, eps)
        self.save_level = "layer"
        self._c = None          # bf16 compute copies
        self._versions = None
        self._g = None          # fp32 gradient arenas (views of _gflat)
        self._gflat = None
        self._grads_fresh = True
        self._grad_ready_cb = None   # set by the data-parallel trainer: called when this layer's arenas are final
        self._gnorm2 = None          # f32 [1]: sum of squares of the four weight-gradient matrices (written by the wgrad GEMMs)
        self._gnorm2_valid = False   # True when _gnorm2 describes what the arenas hold now
        self._want_gnorm2 = False    # set by the trainer on a single replica
        self._weights_ready = None   # CUDA event: the optimizer's side-stream update of this layer has landed

    # ------------------------------------------------------------------ parameter plumbing
    def _masters(self) -> List[nn.Parameter]:
        a, m = self.self_attn, self.mlp
        return [a.q_proj.weight, a.k_proj.weight, a.v_proj.weight, a.o_proj.weight, m.gate_proj.weight,
                m.up_proj.weight, m.down_proj.weight, self.input_layernorm.weight,
                self.post_attention_layernorm.weight]

    def compute_weights(self):
        """bf16 [q;k;v], o, [gate;up], down, ln1, ln2 — refreshed when any fp32 master changed."""
        ps = self._masters()
        if self._weights_ready is not None:
            torch.cuda.current_stream().wait_event(self._weights_ready)
            self._weights_ready = None
        vers = tuple(p._version for p in ps) + tuple(p.data_ptr() for p in ps)
        if self._c is None or vers != self._versions:
            h, f = self.hidden_size, self.inter
            dev = ps[0].device
            if self._c is None or self._c[0].device != dev:
                self._c = (torch.empty(3 * h, h, dtype=torch.bfloat16, device=dev),
                           torch.empty(h, h, dtype=torch.bfloat16, device=dev),
                           torch.empty(2 * f, h, dtype=torch.bfloat16, device=dev),
                           torch.empty(h, f, dtype=torch.bfloat16, device=dev),
                           torch.empty(h, dtype=torch.bfloat16, device=dev),
                           torch.empty(h, dtype=torch.bfloat16, device=dev))
            wqkv, wo, wgu, wd, l1, l2 = self._c
            with torch.no_grad():
                ops.cast_bf16(ps[0], wqkv[:h]); ops.cast_bf16(ps[1], wqkv[h:2 * h]); ops.cast_bf16(ps[2], wqkv[2 * h:])
                ops.cast_bf16(ps[3], wo)
                ops.cast_bf16(ps[4], wgu[:f]); ops.cast_bf16(ps[5], wgu[f:])
                ops.cast_bf16(ps[6], wd); ops.cast_bf16(ps[7], l1); ops.cast_bf16(ps[8], l2)
            self._versions = vers
        return self._c

    def grad_arenas(self):
        """fp32 [3h,h], [h,h], [2f,h], [h,f], [h], [h] — views of ONE flat buffer per layer (`_gflat`: the data-parallel
        trainer reduces a layer's gradients with a single collective); param.grad aliases views of these."""
        ps = self._masters()
        dev = ps[0].device
        if self._g is None or self._g[0].device != dev:
            h, f = self.hidden_size, self.inter
            sizes = [3 * h * h, h * h, 2 * f * h, h * f, h, h]
            shapes = [(3 * h, h), (h, h), (2 * f, h), (h, f), (h,), (h,)]
            self._gflat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
            off, views = 0, []
            for n, sh in zip(sizes, shapes):
                views.append(self._gflat[off:off + n].view(sh))
                off += n
            self._g = tuple(views)
            self._views = None
        if getattr(self, "_views", None) is None:
            h, f = self.hidden_size, self.inter
            gqkv, go, ggu, gd, g1, g2 = self._g
            self._views = [gqkv[:h], gqkv[h:2 * h], gqkv[2 * h:], go, ggu[:f], ggu[f:], gd, g1, g2]
        return self._g

    def _attach_grads(self) -> bool:
        """Point param.grad at the arenas.  Returns True when the arenas hold live gradients to accumulate into
        (i.e. the caller has not dropped/zeroed them since the last backward)."""
        self.grad_arenas()
        ps = self._masters()
        live = all(p.grad is v for p, v in zip(ps, self._views)) and not self._grads_fresh
        if not live:
            for p, v in zip(ps, self._views):
                if p.requires_grad:
                    p.grad = v
        return live

    def mark_grads_fresh(self):
        """Next backward overwrites the arenas instead of accumulating (cheaper than zeroing 0.8 GB per layer)."""
        self._grads_fresh = True
        self._gnorm2_valid = False

    # ------------------------------------------------------------------ compute
    def _attn_half(self, x: torch.Tensor, sh: LayerShape, keep: bool):
        wqkv, wo, _, _, l1, _ = self.compute_weights()
        n1 = ops.rmsnorm_fwd(x, l1, self.eps)
        if sh.D == 128 and FUSE_ROPE["on"]:
            # RoPE in the projection's epilogue (q and k heads = the leading 2*H*D columns), no extra HBM pass
            qkv = ops.gemm(n1, wqkv, rope=(sh.cos, sh.sin, sh.S, 2 * sh.H * sh.D, sh.rope_pos))
        else:
            if sh.group:
                raise ops._lib.MlaError("the shared-prefix layout needs head_dim 128 with RoPE fused into the projection")
            qkv = ops.gemm(n1, wqkv)
            ops.rope_(qkv, 0, 2 * sh.H, sh.D, sh.S, sh.cos, sh.sin)      # q and k heads are contiguous
        ctx, lse = ops.attn_fwd(qkv, sh.B, sh.S, sh.H, sh.D, sh.mask, grouped=sh.grouped)
        x_mid = ops.gemm(ctx, wo, residual=x)
        return n1, qkv, ctx, lse, x_mid

    def _mlp_half(self, x_mid: torch.Tensor):
        _, _, wgu, wd, _, l2 = self.compute_weights()
        n2 = ops.rmsnorm_fwd(x_mid, l2, self.eps)
        if FUSE_SWIGLU["on"] and self.inter % 128 == 0 and n2.shape[0] >= 1024:
            act = torch.empty((n2.shape[0], self.inter), dtype=torch.bfloat16, device=n2.device)
            gu = ops.gemm(n2, wgu, swiglu_out=act)
        else:
            gu = ops.gemm(n2, wgu)
            act = ops.swiglu_fwd(gu)
        y = ops.gemm(act, wd, residual=x_mid)
        return n2, gu, act, y

    def forward_impl(self, x: torch.Tensor, sh: LayerShape, save_level: str):
        with _Range(f"layer.{self.layer_idx}.fwd"):
            n1, qkv, ctx, lse, x_mid = self._attn_half(x, sh, True)
            n2, gu, act, y = self._mlp_half(x_mid)
        if save_level == "layer":
            saved = (x,)
        elif save_level == "mlp":
            saved = (x, qkv, ctx, lse, x_mid)
        else:
            saved = (x, qkv, ctx, lse, x_mid, gu)
        return y, saved

    def backward_impl(self, dy: torch.Tensor, saved: Tuple[torch.Tensor, ...], sh: LayerShape, save_level: str):
        with _Range(f"layer.{self.layer_idx}.bwd"):
            return self._backward_impl(dy, saved, sh, save_level)

    def _backward_impl(self, dy: torch.Tensor, saved: Tuple[torch.Tensor, ...], sh: LayerShape, save_level: str):
        wqkv, wo, wgu, wd, l1, l2 = self.compute_weights()
        x = saved[0]
        if save_level == "layer":
            n1, qkv, ctx, lse, x_mid = self._attn_half(x, sh, True)
            n2, gu, act, _ = self._mlp_half_no_down(x_mid)
        else:
            qkv, ctx, lse, x_mid = saved[1:5]
            n1 = ops.rmsnorm_fwd(x, l1, self.eps)
            n2 = ops.rmsnorm_fwd(x_mid, l2, self.eps)
            gu = saved[5] if save_level == "none" else ops.gemm(n2, wgu)
            act = None          # re-materialised by the fused SwiGLU backward below
        acc = self._attach_grads()
        gqkv, go, ggu, gd, g1, g2 = self._g
        if not acc:
            g1.zero_(); g2.zero_()   # the norm-weight kernels accumulate atomically
        overlap = OVERLAP["wgrad"] and dy.is_cuda
        main = torch.cuda.current_stream() if overlap else None
        side = side_stream(dy.device) if overlap else None
        ss = None
        if self._want_gnorm2 and FUSE_GRAD_NORM["on"] and not overlap:
            # the four weight-gradient GEMMs below leave the sum of squares of what they write (the final values, also
            # when accumulating onto an earlier micro-batch) in _gnorm2
            if self._gnorm2 is None or self._gnorm2.device != dy.device:
                self._gnorm2 = torch.zeros(1, dtype=torch.float32, device=dy.device)
            else:
                self._gnorm2.zero_()
            ss = self._gnorm2
        self._gnorm2_valid = False

        def wgrad(d, a, out):
            """out (+)= d^T a.  With overlap: on the side stream, ordered after everything issued so far."""
            if not overlap:
                ops.gemm(d, a, a_mn=True, b_mn=True, out=out, accumulate=acc, sumsq=ss)
                return
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ops.gemm(d, a, a_mn=True, b_mn=True, out=out, accumulate=acc)
            d.record_stream(side)    # the caching allocator must not recycle these before the side GEMM has read them
            a.record_stream(side)

        # ---- MLP half
        if FUSE_SWIGLU_BWD["on"] and self.inter % 32 == 0:
            # d_act = dy Wd never reaches HBM: the GEMM's epilogue turns it into d(gate|up) and re-materialises act
            dgu = torch.empty_like(gu)
            act = torch.empty((gu.shape[0], self.inter), dtype=torch.bfloat16, device=gu.device)
            ops.gemm(dy, wd, b_mn=True, swiglu_bwd=(gu, dgu, act))
        else:
            dact = ops.gemm(dy, wd, b_mn=True)                                       # dact = dy Wd
            dgu, act = ops.swiglu_bwd_act(dact, gu)                                  # + act = swiglu(gu), one pass
            del dact
        del gu
        wgrad(dy, act, gd)                                                           # dWd  = dy^T act
        del act
        wgrad(dgu, n2, ggu)                                                          # dWgu = dgu^T n2
        dn2 = ops.gemm(dgu, wgu, b_mn=True)                                          # dn2  = dgu Wgu
        del dgu, n2
        dx_mid = ops.rmsnorm_bwd(dn2, x_mid, l2, self.eps, dres=dy, dw=g2)
        del dn2
        # ---- attention half
        wgrad(dx_mid, ctx, go)                                                       # dWo  = dx_mid^T ctx
        dctx = ops.gemm(dx_mid, wo, b_mn=True)
        if ops.attn_bwd_fuses_rope(sh.D) and FUSE_ROPE["on"]:
            # the backward of RoPE rides in the attention kernel's epilogue (no extra pass over dq | dk)
            dqkv = ops.attn_bwd(dctx, qkv, ctx, lse, sh.B, sh.S, sh.H, sh.D, sh.mask, rope=(sh.cos, sh.sin),
                                grouped=sh.grouped, rope_pos=sh.rope_pos)
        else:
            if sh.group:
                raise ops._lib.MlaError("the shared-prefix layout needs the pipelined attention backward (head_dim 128)")
            dqkv = ops.attn_bwd(dctx, qkv, ctx, lse, sh.B, sh.S, sh.H, sh.D, sh.mask)
            ops.rope_(dqkv, 0, 2 * sh.H, sh.D, sh.S, sh.cos, sh.sin, transpose=True)
        del dctx, ctx, qkv
        wgrad(dqkv, n1, gqkv)                                                        # dWqkv = dqkv^T n1
        dn1 = ops.gemm(dqkv, wqkv, b_mn=True)
        del dqkv, n1
        dx = ops.rmsnorm_bwd(dn1, x, l1, self.eps, dres=dx_mid, dw=g1)
        self._grads_fresh = False
        self._gnorm2_valid = ss is not None
        if self._grad_ready_cb is not None:
            if overlap:
                # the arenas are final once BOTH streams are done: issue the exchange behind the side stream
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    self._grad_ready_cb(self)
            else:
                self._grad_ready_cb(self)
        return dx

    # ------------------------------------------------------------------ inference: prefix once, suffix per DDIM step
    def prefill(self, x: torch.Tensor, sh: LayerShape, cache: torch.Tensor) -> torch.Tensor:
        """Training-path forward of the prefix rows x [B*P, h] (no autograd); the post-RoPE keys | values of every
        prefix position go to the head-major cache [B, 2, H, P, D] (each head's rows contiguous: the decode attention
        streams them instead of gathering 256-byte pieces at a 16 KB stride)."""
        _, qkv, _, _, x_mid = self._attn_half(x, sh, False)
        cache.copy_(qkv.view(sh.B, sh.S, 3, sh.H, sh.D)[:, :, 1:].permute(0, 2, 3, 1, 4))
        del qkv
        return self._mlp_half(x_mid)[3]

    def decode(self, x: torch.Tensor, cache: torch.Tensor, B: int, P: int, n: int, cos: torch.Tensor,
               sin: torch.Tensor) -> torch.Tensor:
        """The n suffix rows per sample, x [B*n, h], at positions P..P+n-1 against the cached prefix K/V.  Every
        linear is a weight-streaming skinny GEMM (ops.gemv); cos/sin are the RoPE table rows P..P+n-1."""
        wqkv, wo, wgu, wd, l1, l2 = self.compute_weights()
        h, H = self.hidden_size, self.heads
        D = h // H
        # 5 launches: RMSNorm and SwiGLU ride in the prologues of the skinny GEMMs (ops.gemv); RoPE of q and of the
        # new keys happens inside the attention kernel, which reads them straight from the projection — the suffix
        # K/V are never appended to the cache (the next DDIM step recomputes them from the next x_t)
        qkv = ops.gemv(x, wqkv, norm=(l1, self.eps))
        ctx = ops.decode_attn_rope(qkv, cache, cos, sin, B, H, n, P, D)
        x_mid = ops.gemv(ctx, wo, residual=x)
        gu = ops.gemv(x_mid, wgu, norm=(l2, self.eps))
        return ops.gemv(gu, wd, residual=x_mid, swiglu=True)

    def _mlp_half_no_down(self, x_mid: torch.Tensor):
        _, _, wgu, _, _, l2 = self.compute_weights()
        n2 = ops.rmsnorm_fwd(x_mid, l2, self.eps)
        gu = ops.gemm(n2, wgu)
        return n2, gu, None, None


class _LayerFn(torch.autograd.Function):
    """One autograd node per decoder layer.  `anchor` is a dummy that keeps the node alive when the layer input
    itself does not require grad (e.g. everything upstream frozen)."""

    @staticmethod
    def forward(ctx, x, anchor, layer: LlamaDecoderLayer, sh: LayerShape):
        level = layer.save_level
        y, saved = layer.forward_impl(x, sh, level)
        ctx.layer, ctx.sh, ctx.level = layer, sh, level
        ctx.save_for_backward(*saved)
        return y

    @staticmethod
    def backward(ctx, dy):
        dx = ctx.layer.backward_impl(dy.contiguous(), ctx.saved_tensors, ctx.sh, ctx.level)
        return dx, None, None, None


class LlamaModel(nn.Module):
    """embed_tokens / layers / norm — modeling_llama.py:912-940."""

    def __init__(self, vocab_size: int, hidden: int, inter: int, n_layers: int, heads: int, eps: float = 1e-5,
                 rope_theta: float = 10000.0, padding_idx: Optional[int] = None):
        super().__init__()
        self.hidden_size, self.heads, self.eps, self.rope_theta = hidden, heads, eps, rope_theta
        self.embed_tokens = nn.Embedding(vocab_size, hidden, padding_idx)
        self.layers = nn.ModuleList(LlamaDecoderLayer(hidden, inter, heads, eps, i) for i in range(n_layers))
        self.norm = LlamaRMSNorm(hidden, eps)
        self._rope_cache = {}
        self._anchor = None

    def rope_tables(self, S: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        """bf16 cos/sin [S, D/2], computed as modeling_llama.py:132-145 does (fp32 outer product, cast to the
        activation dtype); constant per (S, D, theta) so it is built once on the host and cached."""
        D = self.hidden_size // self.heads
        key = (S, D, str(device))
        if key not in self._rope_cache:
            inv_freq = 1.0 / (self.rope_theta ** (torch.arange(0, D, 2, dtype=torch.int64).float() / D))
            pos = torch.arange(S, dtype=torch.int64).float()
            freqs = (inv_freq[None, :, None] @ pos[None, None, :]).transpose(1, 2)[0]
            self._rope_cache[key] = (freqs.cos().to(torch.bfloat16).to(device).contiguous(),
                                     freqs.sin().to(torch.bfloat16).to(device).contiguous())
        return self._rope_cache[key]

    def set_save_levels(self, levels) -> None:
        if isinstance(levels, str):
            levels = [levels] * len(self.layers)
        for l, lv in zip(self.layers, levels):
            assert lv in SAVE_LEVELS, lv
            l.save_level = lv

    def mark_grads_fresh(self):
        for l in self.layers:
            l.mark_grads_fresh()

    def run_layers(self, x: torch.Tensor, B: int, S: int, mask: Optional[torch.Tensor],
                   prefix_len: Optional[torch.Tensor] = None, group: int = 0,
                   rope_pos: Optional[torch.Tensor] = None) -> List[torch.Tensor]:
        """x: bf16 [B*S, h].  Returns the list of hidden states (layer inputs + final normed), each [B*S, h].
        prefix_len / group / rope_pos: shared-prefix layout (see LayerShape)."""
        D = self.hidden_size // self.heads
        cos, sin = self.rope_tables(S, x.device)
        sh = LayerShape(B, S, self.heads, D, mask, cos, sin, prefix_len, group, rope_pos)
        if self._anchor is None or self._anchor.device != x.device:
            self._anchor = torch.zeros(1, device=x.device, requires_grad=True)
        hs = []
        for layer in self.layers:
            hs.append(x)
            x = _LayerFn.apply(x, self._anchor, layer, sh)
        hs.append(ops.RMSNormFn.apply(x, self.norm.weight, self.eps))
        return hs

    # ------------------------------------------------------------------ inference (KV-cached denoise loop)
    @torch.no_grad()
    def prefill(self, x: torch.Tensor, B: int, P: int, extra: int,
                caches: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
        """Run the P prefix rows per sample (x bf16 [B*P, h], no padding) through every layer once.  Returns one
        key | value cache per layer, bf16 [B, 2, H, P, D] (written into `caches` when given: static buffers of a
        CUDA-graph session); `extra` = the number of suffix positions that will follow (sizes the RoPE table)."""
        D = self.hidden_size // self.heads
        cos, sin = self.rope_tables(P + extra, x.device)
        sh = LayerShape(B, P, self.heads, D, None, cos[:P].contiguous(), sin[:P].contiguous())
        if caches is None:
            caches = [torch.empty((B, 2, self.heads, P, D), dtype=torch.bfloat16, device=x.device) for _ in self.layers]
        for layer, cache in zip(self.layers, caches):
            x = layer.prefill(x, sh, cache)
        return caches

    @torch.no_grad()
    def decode(self, x: torch.Tensor, caches: List[torch.Tensor], B: int, P: int, n: int) -> torch.Tensor:
        """n suffix rows per sample (x bf16 [B*n, h]) at positions P..P+n-1 -> final-norm hidden states [B*n, h]."""
        cos, sin = self.rope_tables(P + n, x.device)
        cs, sn = cos[P:P + n].contiguous(), sin[P:P + n].contiguous()
        D = self.hidden_size // self.heads
        f = self.layers[0].inter
        if DECODE_STACK and ops.decode_stack_supported(B * n, self.hidden_size, f, D):
            x = self._decode_stack(x, caches, B, P, n, cs, sn)
        else:
            for layer, cache in zip(self.layers, caches):
                x = layer.decode(x, cache, B, P, n, cs, sn)
        return ops.rmsnorm_fwd(x, ops.bf16_of(self.norm.weight), self.eps)

    def _decode_stack(self, x, caches, B, P, n, cs, sn):
        """One persistent launch for all layers (csrc/decode_stack.cu).  The per-layer pointer table and the workspace
        are built on the first (eager) call for a set of weight / cache buffers and re-used afterwards — a CUDA-graph
        capture of the DDIM loop therefore has to be preceded by an eager pass, which VLM.denoise_session does.  A
        captured graph holds the table's address: entries seen during a capture are pinned, the others are evicted
        least-recently-used (eager calls bring a fresh set of caches every time)."""
        ws = [layer.compute_weights() for layer in self.layers]       # refreshes the bf16 copies if a master changed
        ptrs = tuple(t.data_ptr() for w in ws for t in w) + tuple(c.data_ptr() for c in caches)
        tables = self.__dict__.setdefault("_stack_tables", {})
        st = tables.pop(ptrs, None)
        capturing = torch.cuda.is_current_stream_capturing()
        if st is None:
            if capturing:
                raise RuntimeError("decode: run the denoise step once eagerly before capturing it into a CUDA graph "
                                   "(the per-layer pointer table is built on the first call)")
            rows = [[w[i].data_ptr() for w in ws] for i in range(6)] + [[c.data_ptr() for c in caches]]
            st = SimpleNamespace(table=torch.tensor(rows, dtype=torch.int64).to(x.device), ws={}, pinned=False)
            loose = [k for k, v in tables.items() if not v.pinned]
            for k in loose[:max(0, len(loose) - 3)]:
                del tables[k]
        tables[ptrs] = st                                             # most recently used last
        st.pinned = st.pinned or capturing
        key = (B, n, P)
        if key not in st.ws:
            if capturing:
                raise RuntimeError("decode: workspace for this shape was not built before the capture")
            st.ws[key] = ops.decode_stack_workspace(B, n, P, self.heads, self.hidden_size // self.heads, x.device)
        return ops.decode_stack(x, st.table, cs, sn, st.ws[key], B, n, P, self.heads, self.hidden_size // self.heads,
                                self.layers[0].inter, self.eps)
