"""One decoder layer of the K/V-cached denoise step at Llama-2-7B size (2 suffix rows, 546-token prefix) inside a
profiler range, for
    ncu --set full --profile-from-start off --clock-control none -k regex:"gemv|decode_attn|rope_cache" python tools/ncu_decode.py
Eight distinct weight sets rotate so that no launch finds its weights in the 126 MB L2 (one layer = 405 MB)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mla_b200 import ops  # noqa: E402

H, F, HEADS, P, N = 4096, 11008, 32, 546, 2
bf, dev = torch.bfloat16, "cuda"
R = 3
W = [dict(qkv=torch.randn(3 * H, H, device=dev).to(bf) * 0.02, o=torch.randn(H, H, device=dev).to(bf) * 0.02,
          gu=torch.randn(2 * F, H, device=dev).to(bf) * 0.02, d=torch.randn(H, F, device=dev).to(bf) * 0.02)
     for _ in range(R)]
x = torch.randn(N, H, device=dev).to(bf)
cache = torch.randn(1, 2, HEADS, P, H // HEADS, device=dev).to(bf)       # head-major prefix K | V


lnw = torch.ones(H, device=dev).to(bf)
inv = 1.0 / (10000 ** (torch.arange(0, H // HEADS, 2, device=dev).float() / (H // HEADS)))
fr = torch.arange(P, P + N, device=dev).float()[:, None] * inv[None]
cos, sin = fr.cos().to(bf).contiguous(), fr.sin().to(bf).contiguous()


def layer(w):
    """Same launch sequence as LlamaDecoderLayer.decode (mla_b200/llama.py)."""
    qkv = ops.gemv(x, w["qkv"], norm=(lnw, 1e-5))
    ctx = ops.decode_attn_rope(qkv, cache, cos, sin, 1, HEADS, N, P, H // HEADS)
    mid = ops.gemv(ctx, w["o"], residual=x)
    gu = ops.gemv(mid, w["gu"], norm=(lnw, 1e-5))
    return ops.gemv(gu, w["d"], residual=mid, swiglu=True)


for w in W:
    layer(w)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for w in W:
    layer(w)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
