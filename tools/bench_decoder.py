"""Decoder-only fwd+bwd timing at Llama-2-7B shapes (bring-up tool; bench.py is the contract benchmark)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mla_b200 import llama, _lib  # noqa: E402


def main():
    L = int(os.environ.get("L", 32))
    B, S, h, f, H = int(os.environ.get("B", 32)), int(os.environ.get("S", 548)), 4096, 11008, 32
    level = os.environ.get("LEVEL", "layer")
    torch.manual_seed(0)
    with torch.device("cuda"):
        m = llama.LlamaModel(32064, h, f, L, H, eps=1e-5)
    for p in m.parameters():
        p.data.normal_(std=0.02)
    for l in m.layers:
        l.input_layernorm.weight.data.fill_(1.0)
        l.post_attention_layernorm.weight.data.fill_(1.0)
    m.norm.weight.data.fill_(1.0)
    m.set_save_levels(level)
    x = (torch.randn(B * S, h, device="cuda") * 0.02).to(torch.bfloat16).requires_grad_(True)
    g = torch.randn(B * S, h, device="cuda").to(torch.bfloat16) * 1e-3
    out = {}
    for it in range(4):
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        m.mark_grads_fresh()
        n0 = _lib.launch_count()
        e[0].record()
        hs = m.run_layers(x, B, S, None)
        e[1].record()
        hs[-1].backward(g)
        e[2].record()
        torch.cuda.synchronize()
        fwd, bwd = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        tok = B * S
        flops = 3 * tok * L * (8 * h * h + 6 * h * f + 2 * S * h)
        out = dict(L=L, B=B, S=S, level=level, fwd_ms=fwd, bwd_ms=bwd, tok_per_s=tok / (fwd + bwd) * 1e3,
                   alg_tflops=flops / (fwd + bwd) / 1e9, launches=_lib.launch_count() - n0,
                   mem_gb=torch.cuda.max_memory_allocated() / 2**30)
        print(json.dumps(out), flush=True)
        del hs
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/bench_decoder_{level}_L{L}.json", "w") as fo:
        json.dump(out, fo)


if __name__ == "__main__":
    main()
