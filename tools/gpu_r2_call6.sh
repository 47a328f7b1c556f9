#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_attention_sm100_gpu.py -x -q -s -m gpu > $O/r02_attn_tests.log 2>&1; echo "rc=$?" >> $O/r02_attn_tests.log
tail -8 $O/r02_attn_tests.log
timeout 300 python -m pytest tests/test_gemm2_gpu.py tests/test_kernels_gpu.py -x -q -m gpu > $O/r02_gemm2_tests.log 2>&1; echo "rc=$?" >> $O/r02_gemm2_tests.log
tail -4 $O/r02_gemm2_tests.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_attention_sm100_gpu.py --deselect tests/test_gemm2_gpu.py > $O/r02_gpu_tests_call6.log 2>&1; echo "rc=$?" >> $O/r02_gpu_tests_call6.log
tail -5 $O/r02_gpu_tests_call6.log
timeout 600 python tools/ref_gpu.py attn --out $O/r02_attn_vs_flash_v4.json > $O/r02_attn_vs_flash_v4.log 2>&1; tail -4 $O/r02_attn_vs_flash_v4.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_e.json 2> $O/r02_bench_n1_e.err; tail -c 1500 $O/r02_bench_n1_e.json; tail -5 $O/r02_bench_n1_e.err
MLA_GEMM_UNIFORM_ISSUE=0 timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_e_olduni.json 2> $O/r02_bench_n1_e_olduni.err; tail -c 900 $O/r02_bench_n1_e_olduni.json
timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_e2.json 2> $O/r02_bench_n1_e2.err; tail -c 900 $O/r02_bench_n1_e2.json
MLA_GEMM_UNIFORM_ISSUE=0 timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_e_olduni2.json 2> $O/r02_bench_n1_e_olduni2.err; tail -c 900 $O/r02_bench_n1_e_olduni2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 5 -o $O/r02_attn_ncu3 python tools/prof_attn.py > $O/r02_attn_ncu3.log 2>&1; tail -3 $O/r02_attn_ncu3.log
