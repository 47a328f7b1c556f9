#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/r02_gpu_tests_call11.log 2>&1; echo "rc=$?" >> $O/r02_gpu_tests_call11.log
tail -6 $O/r02_gpu_tests_call11.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_g.json 2> $O/r02_bench_n1_g.err; tail -c 600 $O/r02_bench_n1_g.json; tail -3 $O/r02_bench_n1_g.err
MLA_FUSE_GRAD_NORM=0 timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_g_nonorm.json 2> $O/r02_bench_n1_g_nonorm.err; tail -c 600 $O/r02_bench_n1_g_nonorm.json
timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline --share-prefix > $O/r02_bench_n1_g_shared.json 2> $O/r02_bench_n1_g_shared.err; tail -c 600 $O/r02_bench_n1_g_shared.json
