#!/bin/bash
O=gpurun_out
mkdir -p $O
MLA_GEMV2=1 timeout 900 python -m pytest tests/test_denoise_gpu.py -x -q 2>&1 | tail -3
MLA_GEMV2=0 timeout 600 python tools/bench_denoise.py > $O/a.log 2>&1; tail -1 $O/a.log | python -c "import json,sys;d=json.loads(sys.stdin.read());print('gemv1',d['graph_decode_step_ms'],d['decode_roofline']['frac'],d['kv_cached_cuda_graph_ms'],d['cached_vs_full_rel_diff'])"
MLA_GEMV2=1 timeout 600 python tools/bench_denoise.py > $O/b.log 2>&1; tail -1 $O/b.log | python -c "import json,sys;d=json.loads(sys.stdin.read());print('gemv2',d['graph_decode_step_ms'],d['decode_roofline']['frac'],d['kv_cached_cuda_graph_ms'],d['cached_vs_full_rel_diff'])"; cp $O/denoise_T0.json $O/r02_denoise_T0_gemv2.json
MLA_GEMV2=0 timeout 600 python tools/bench_denoise.py > $O/a.log 2>&1; tail -1 $O/a.log | python -c "import json,sys;d=json.loads(sys.stdin.read());print('gemv1',d['graph_decode_step_ms'],d['decode_roofline']['frac'],d['kv_cached_cuda_graph_ms'],d['cached_vs_full_rel_diff'])"
