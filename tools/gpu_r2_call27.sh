#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_mla_gpu.py tests/test_tower_bwd_gpu.py tests/test_reference_gpu_golden.py -q 2>&1 | tail -2
WORKLOAD=cfg3 OUT=$O/tl3.json timeout 900 python tools/timeline_step.py > $O/tl.log 2>&1; python -c "
import json;t=json.load(open('$O/tl3.json'));k=t['kernels_ms_per_step']
print(t['wall_ms_per_step']); [print(n[:50],v) for n,v in k.items() if any(s in n for s in ('group_pose','bn_','fps','knn','gemm_bf16_kernel'))]"
