"""Oracle (test infrastructure): the modified Llama decoder of the reference, restated op by op.

Follows /root/reference/transformers/models/llama/modeling_llama.py (vendored HF 4.40.1, modified).  Every
function takes plain tensors and runs in the dtype it is given: fp32 tensors give the "truth", bf16 tensors give
the reference's autocast/bf16-parameter arithmetic (each op rounds its output to bf16), which is what the CUDA
kernels reproduce.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """LlamaRMSNorm.forward — modeling_llama.py:85-90."""
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return w * xf.to(dt)


def rope_tables(seq: int, dim: int, base: float = 10000.0, dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """LlamaRotaryEmbedding.__init__/forward — modeling_llama.py:100-101,:132-145 with position_ids = arange(seq)
    (:985-990).  Returns cos, sin of shape [seq, dim] in `dtype` (the reference casts to x.dtype)."""
    inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.int64).float() / dim))
    pos = torch.arange(seq, dtype=torch.int64).float()
    freqs = (inv_freq[None, :, None].float() @ pos[None, None, :].float()).transpose(1, 2)[0]  # [seq, dim/2]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def rotate_half(x: torch.Tensor) -> torch.Tensor:
    """modeling_llama.py:177-181."""
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def apply_rope(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor):
    """apply_rotary_pos_emb — modeling_llama.py:184-208.  q,k: [B,H,S,D]; cos,sin: [S,D]."""
    cos = cos[None, None]
    sin = sin[None, None]
    return (q * cos) + (rotate_half(q) * sin), (k * cos) + (rotate_half(k) * sin)


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """LlamaFlashAttention2._flash_attention_forward — modeling_llama.py:524-559.

    flash-attn itself is an external dependency (README pins 2.5.5, pyproject.toml:35): restated from its
    definition — softmax(QK^T/sqrt(d)) V, causal, fp32 softmax statistics, P rounded to the input dtype before PV.
    With a padding mask the reference unpads (keys with mask 0 disappear, queries keep causal order), runs the
    varlen kernel and `pad_input`s zeros back (:553-557): key j visible to query i iff j <= i and mask[j]; rows
    with mask 0 are zero.  q,k,v: [B,H,S,D] -> [B,S,H*D]."""
    B, H, S, D = q.shape
    scores = torch.matmul(q.float(), k.float().transpose(-1, -2)) / math.sqrt(D)
    vis = torch.ones(S, S, dtype=torch.bool, device=q.device).tril()[None, None]
    if mask is not None:
        vis = vis & mask.bool()[:, None, None, :]
    scores = scores.masked_fill(~vis, float("-inf"))
    m = scores.max(-1, keepdim=True).values
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)
    e = torch.exp(scores - m)
    l = e.sum(-1, keepdim=True)
    pv = torch.matmul(e.to(q.dtype).float(), v.float())
    out = pv / torch.where(l > 0, l, torch.ones_like(l))
    out = out.to(q.dtype)
    if mask is not None:
        out = out * mask.bool()[:, None, :, None].to(out.dtype)
    return out.transpose(1, 2).reshape(B, S, H * D)


def mlp(x: torch.Tensor, wg: torch.Tensor, wu: torch.Tensor, wd: torch.Tensor) -> torch.Tensor:
    """LlamaMLP.forward — modeling_llama.py:240 (pretraining_tp == 1)."""
    return F.linear(F.silu(F.linear(x, wg)) * F.linear(x, wu), wd)


def decoder_layer(x: torch.Tensor, p: Dict[str, torch.Tensor], n_heads: int, eps: float,
                  cos: torch.Tensor, sin: torch.Tensor, mask: Optional[torch.Tensor]) -> torch.Tensor:
    """LlamaDecoderLayer.forward — modeling_llama.py:738-756 with LlamaFlashAttention2.forward :420-500.
    p holds q_proj,k_proj,v_proj,o_proj,gate_proj,up_proj,down_proj ([out,in]) and ln1, ln2 ([h])."""
    B, S, h = x.shape
    D = h // n_heads
    res = x
    hn = rmsnorm(x, p["ln1"], eps)
    q = F.linear(hn, p["q_proj"]).view(B, S, n_heads, D).transpose(1, 2)
    k = F.linear(hn, p["k_proj"]).view(B, S, n_heads, D).transpose(1, 2)
    v = F.linear(hn, p["v_proj"]).view(B, S, n_heads, D).transpose(1, 2)
    q, k = apply_rope(q, k, cos, sin)
    a = attention(q, k, v, mask)
    x = res + F.linear(a, p["o_proj"])
    res = x
    hn = rmsnorm(x, p["ln2"], eps)
    return res + mlp(hn, p["gate_proj"], p["up_proj"], p["down_proj"])


def decoder(x: torch.Tensor, layers: List[Dict[str, torch.Tensor]], final_norm: torch.Tensor, n_heads: int,
            eps: float, mask: Optional[torch.Tensor], rope_base: float = 10000.0) -> List[torch.Tensor]:
    """LlamaModel.forward — modeling_llama.py:996-1040: returns all hidden states (inputs of every layer, then the
    final-normed output), as `output_hidden_states=True` does.  The flash path drops an all-ones mask (:1068-1071)."""
    B, S, h = x.shape
    cos, sin = rope_tables(S, h // n_heads, rope_base, x.dtype)
    cos, sin = cos.to(x.device), sin.to(x.device)
    if mask is not None and bool(mask.bool().all()):
        mask = None
    hs = []
    for p in layers:
        hs.append(x)
        x = decoder_layer(x, p, n_heads, eps, cos, sin, mask)
    hs.append(rmsnorm(x, final_norm, eps))
    return hs


def lm_loss(hidden: torch.Tensor, lm_head: torch.Tensor, labels: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """LlamaForCausalLM.forward — modeling_llama.py:1254-1269: fp32 logits, shifted CE, ignore_index -100."""
    logits = F.linear(hidden, lm_head).float()
    V = logits.shape[-1]
    loss = F.cross_entropy(logits[..., :-1, :].reshape(-1, V), labels[..., 1:].reshape(-1))
    return loss, logits


def action_tokenize(action, vocab_size: int, bins: int = 256, lo: float = -1.0, hi: float = 1.0):
    """ActionTokenizer.__call__ up to the token ids — vla/action_tokenizer.py:43-46 (numpy, float64 edges)."""
    import numpy as np
    a = np.clip(action, a_min=float(lo), a_max=float(hi))
    return vocab_size - np.digitize(a, np.linspace(lo, hi, bins))


def action_detokenize(ids, vocab_size: int, bins: int = 256, lo: float = -1.0, hi: float = 1.0):
    """ActionTokenizer.decode_token_ids_to_actions — vla/action_tokenizer.py:54-71."""
    import numpy as np
    edges = np.linspace(lo, hi, bins)
    centers = (edges[:-1] + edges[1:]) / 2.0
    d = np.clip(vocab_size - ids - 1, a_min=0, a_max=centers.shape[0] - 1)
    return centers[d]
