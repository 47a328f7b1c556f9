#!/bin/bash
O=gpurun_out
mkdir -p $O
run() { tag=$1; shift; env "$@" timeout 900 python bench.py --steps 6 --warmup 3 --workload cfg2 --no-also --no-cpu-baseline > $O/r02_cs_$tag.json 2> $O/r02_cs_$tag.err; python -c "import json;d=json.load(open('$O/r02_cs_$tag.json'));print('$tag',d['ms_per_step'],d['e2e']['ms_per_step'],d['value'],d['clocks']['sm_mhz'])" || tail -3 $O/r02_cs_$tag.err; }
timeout 600 python -m pytest tests/test_gemm2_gpu.py tests/test_gemm_gpu.py -q 2>&1 | tail -2
run off MLA_GEMM_CS_STORES=0
run on MLA_GEMM_CS_STORES=1
run off2 MLA_GEMM_CS_STORES=0
run on2 MLA_GEMM_CS_STORES=1
