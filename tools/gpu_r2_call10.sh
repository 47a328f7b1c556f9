#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py --steps 6 --warmup 3 --workload cfg5 --no-cpu-baseline > $O/r02_bench_cfg5.json 2> $O/r02_bench_cfg5.err; tail -c 1200 $O/r02_bench_cfg5.json; tail -3 $O/r02_bench_cfg5.err
timeout 900 python bench.py --steps 6 --warmup 3 --workload cfg3 --stage pretrain --no-cpu-baseline > $O/r02_bench_cfg3_pretrain.json 2> $O/r02_bench_cfg3_pretrain.err; tail -c 700 $O/r02_bench_cfg3_pretrain.json; tail -3 $O/r02_bench_cfg3_pretrain.err
timeout 600 python tools/bench_denoise.py > $O/r02_denoise_T0.log 2>&1; tail -5 $O/r02_denoise_T0.log
timeout 600 python tools/bench_kernels.py > $O/r02_kernels_hbm.log 2>&1; tail -12 $O/r02_kernels_hbm.log
