#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python tools/ref_gpu.py step --workload cfg3 --out $O/r02_ref_gpu_cfg3.json > $O/r02_ref_gpu_cfg3.log 2>&1; grep -v INFO $O/r02_ref_gpu_cfg3.log | tail -4 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; tail -2 $O/r02_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_ref_cpu_cfg3.json 2> $O/r02_ref_cpu_cfg3.err; cat $O/r02_ref_cpu_cfg3.json | cut -c1-700
timeout 1500 python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err; tail -c 2500 $O/r02_bench_default.json; tail -3 $O/r02_bench_default.err
