#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r02_gpu_tests_call9.log 2>&1; echo "rc=$?" >> $O/r02_gpu_tests_call9.log
tail -5 $O/r02_gpu_tests_call9.log
WORKLOAD=cfg2 OUT=$O/r02_timeline_cfg2.json timeout 600 python tools/timeline_step.py > $O/r02_timeline_cfg2.log 2>&1; tail -3 $O/r02_timeline_cfg2.log
WORKLOAD=cfg3 OUT=$O/r02_timeline_cfg3.json timeout 600 python tools/timeline_step.py > $O/r02_timeline_cfg3.log 2>&1; tail -3 $O/r02_timeline_cfg3.log
WORKLOAD=cfg2 SHARE=1 OUT=$O/r02_timeline_cfg2_shared.json timeout 600 python tools/timeline_step.py > $O/r02_timeline_cfg2_shared.log 2>&1; tail -3 $O/r02_timeline_cfg2_shared.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r02_launches_cfg2.csv python tools/profile_step.py > $O/r02_launches_cfg2.log 2>&1; tail -2 $O/r02_launches_cfg2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2 -c 12 -o $O/r02_gemm12 python tools/ncu_gemm.py > $O/r02_ncu_gemm.log 2>&1; tail -2 $O/r02_ncu_gemm.log
timeout 600 python tools/parity_table.py > $O/r02_parity_table.md 2> $O/r02_parity_table.err; echo "parity rc=$?"
timeout 1200 python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err; tail -c 3000 $O/r02_bench_default.json; tail -5 $O/r02_bench_default.err
