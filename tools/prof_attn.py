import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mla_b200 import ops
B, S, H, D = 32, 548, 32, 128
qkv = torch.randn(B * S, 3 * H * D, device="cuda").to(torch.bfloat16)
for _ in range(3):
    ops.attn_fwd(qkv, B, S, H, D, None)
torch.cuda.synchronize()
