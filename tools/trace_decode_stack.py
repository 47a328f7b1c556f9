"""Where the time of the one-launch decoder stack (csrc/decode_stack.cu) goes: per-phase globaltimer stamps of every CTA
(phase entry / work done / barrier passed) at Llama-2-7B size, 2 suffix rows, 545 cached positions.

    python tools/trace_decode_stack.py [--layers 32] [--out gpurun_out/decode_stack_trace.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mla_b200 import ops  # noqa: E402
from mla_b200.llama import LlamaModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--P", type=int, default=545)
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--out", default="gpurun_out/decode_stack_trace.json")
    ap.add_argument("--ahead", type=int, nargs="*", default=[0])
    ap.add_argument("--rpi-big", dest="rpi_big", type=int, nargs="*", default=[2])
    ap.add_argument("--ring", type=int, nargs="*", default=[0], help="ring size caps in KB to sweep (0 = default)")
    ap.add_argument("--dbg", type=int, nargs="*", default=[0],
                    help="mla_decode_stack_set_debug flags to sweep: 1 no math, 2 no grid barriers, 4 no attention")
    a = ap.parse_args()
    h, f, H, L, B, P, n = 4096, 11008, 32, a.layers, 1, a.P, a.n
    D = h // H
    with torch.device("cuda"):
        model = LlamaModel(64, h, f, L, H).eval()
    caches = [torch.randn(B, 2, H, P, D, device="cuda").bfloat16() for _ in range(L)]
    x = torch.randn(B * n, h, device="cuda").bfloat16()
    cos, sin = model.rope_tables(P + n, x.device)
    cs, sn = cos[P:P + n].contiguous(), sin[P:P + n].contiguous()
    model._decode_stack(x, caches, B, P, n, cs, sn)
    st = list(model._stack_tables.values())[-1]
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    trace = torch.zeros(sms * L * 15 + sms * 4, dtype=torch.int64, device="cuda")
    run = lambda tr=None: ops.decode_stack(x, st.table, cs, sn, st.ws[(B, n, P)], B, n, P, H, D, f, model.eps, trace=tr)
    import ctypes as C
    from mla_b200 import _lib
    results = {}

    def per_op():
        xx = x
        for layer, cache in zip(model.layers, caches):
            xx = layer.decode(xx, cache, B, P, n, cs, sn)
        return xx
    g = torch.cuda.CUDAGraph()
    per_op()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        per_op()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    results["per_op_graph_ms_per_step"] = round(e0.elapsed_time(e1) / 10, 4)      # the pod's speed reference
    for ring in a.ring:
      _lib.lib().mla_decode_stack_set_ring_kb(C.c_int32(ring))
      for rb in a.rpi_big:
        _lib.lib().mla_decode_stack_set_rows_per_slot_big(C.c_int32(rb))
        for ahead in a.ahead:
            _lib.lib().mla_decode_stack_set_ahead(C.c_int32(ahead))
            for dbg in a.dbg:
                _lib.lib().mla_decode_stack_set_debug(C.c_int32(dbg))
                results[f"ring{ring}_rb{rb}_ahead{ahead}_dbg{dbg}"] = measure(run, trace, L, h, f)
    _lib.lib().mla_decode_stack_set_ring_kb(C.c_int32(0))
    _lib.lib().mla_decode_stack_set_rows_per_slot_big(C.c_int32(2))
    _lib.lib().mla_decode_stack_set_debug(C.c_int32(0))
    print(json.dumps(results, indent=1))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(results, open(a.out, "w"), indent=1)


def measure(run, trace, L, h, f):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    trace.zero_()
    run(trace)
    torch.cuda.synchronize()
    sms = (trace.numel()) // (L * 15 + 4)
    stats = trace[sms * L * 15:].view(sms, 4).cpu().double()
    t = trace[:sms * L * 15].view(sms, L, 5, 3).cpu().double()                      # [sm, L, phase, 3]
    work = (t[..., 1] - t[..., 0]) / 1e3          # us
    wait = (t[..., 2] - t[..., 1]) / 1e3
    names = ["rmsnorm+qkv", "attention", "o_proj", "rmsnorm+gate|up", "swiglu+down"]
    wbytes = [3 * h * h * 2, 0, h * h * 2, 2 * f * h * 2, h * f * 2]
    out = {"ms_per_step": round(ms, 4), "us_per_layer": round(ms * 1e3 / L, 2),
           "producer_wait_frac": round(float((stats[:, 0] / stats[:, 1]).mean()), 3),
           "producer_cycles_per_slot_busy": round(float(((stats[:, 1] - stats[:, 0]) / stats[:, 2]).mean()), 1),
           "producer_cycles_per_slot_total": round(float((stats[:, 1] / stats[:, 2]).mean()), 1),
           "consumer_wait_frac_of_producer_total": round(float((stats[:, 3] / stats[:, 1]).mean()), 3), "phases": {}}
    for i, nm in enumerate(names):
        # phase wall time = from the moment the FIRST CTA entered to the moment the last one left the barrier
        wall = (t[:, :, i, 2].max(0).values - t[:, :, i, 0].min(0).values) / 1e3
        out["phases"][nm] = {"work_us_mean": round(float(work[:, :, i].mean()), 2),
                             "work_us_max_over_ctas": round(float(work[:, :, i].max(0).values.mean()), 2),
                             "work_us_min_over_ctas": round(float(work[:, :, i].min(0).values.mean()), 2),
                             "barrier_wait_us_min_over_ctas": round(float(wait[:, :, i].min(0).values.mean()), 2),
                             "wall_us": round(float(wall.mean()), 2),
                             "weights_GBs_at_work_mean": round(wbytes[i] / float(work[:, :, i].mean()) / 1e3, 1) if wbytes[i] else None}
    return out


if __name__ == "__main__":
    main()
