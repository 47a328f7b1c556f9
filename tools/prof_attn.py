"""ncu target: a few launches of the attention kernels at the benchmark shape (forward, pipelined backward).
    ncu --set full --clock-control none --import-source on -k regex:attn_ -c 6 -o gpurun_out/r02_attn python tools/prof_attn.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mla_b200 import ops  # noqa: E402

B, S, H, D = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (32, 548, 32, 128)))
qkv = (torch.randn(B * S, 3 * H * D, device="cuda") * 0.5).to(torch.bfloat16)
dctx = (torch.randn(B * S, H * D, device="cuda") * 0.1).to(torch.bfloat16)
for _ in range(2):
    ctx, lse = ops.attn_fwd(qkv, B, S, H, D, None)
    ops.attn_bwd(dctx, qkv, ctx, lse, B, S, H, D, None)
torch.cuda.synchronize()
