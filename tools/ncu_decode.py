"""One decoder layer of the K/V-cached denoise step at Llama-2-7B size (2 suffix rows, 546-token prefix) inside a
profiler range, for
    ncu --set full --profile-from-start off --clock-control none -k regex:"gemv|decode_attn" python tools/ncu_decode.py
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
cache = torch.randn(P + N, 2 * H, device=dev).to(bf)


def layer(w):
    qkv = ops.gemv(x, w["qkv"])
    ctx = ops.decode_attn(qkv, cache, 1, HEADS, N, P + N, H // HEADS)
    mid = ops.gemv(ctx, w["o"], residual=x)
    act = ops.swiglu_fwd(ops.gemv(mid, w["gu"]))
    return ops.gemv(act, w["d"], residual=mid)


for w in W:
    layer(w)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for w in W:
    layer(w)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
