#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_attention_sm100_gpu.py -x -q -s -m gpu > $O/r02_attn_tests.log 2>&1; echo "rc=$?" >> $O/r02_attn_tests.log
tail -12 $O/r02_attn_tests.log
timeout 300 python -m pytest tests/test_gemm2_gpu.py -x -q -m gpu > $O/r02_gemm2_tests.log 2>&1; echo "rc=$?" >> $O/r02_gemm2_tests.log
tail -6 $O/r02_gemm2_tests.log
if ! grep -q "rc=0" $O/r02_attn_tests.log; then export MLA_ATTN_BWD_TS=0; fi
if ! grep -q "rc=0" $O/r02_gemm2_tests.log; then export MLA_FUSE_SWIGLU_BWD=0; fi
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_attention_sm100_gpu.py --deselect tests/test_gemm2_gpu.py > $O/r02_gpu_tests_call4.log 2>&1; echo "rc=$?" >> $O/r02_gpu_tests_call4.log
tail -8 $O/r02_gpu_tests_call4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 5 -o $O/r02_attn_ncu python tools/prof_attn.py > $O/r02_attn_ncu.log 2>&1; tail -3 $O/r02_attn_ncu.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_c.json 2> $O/r02_bench_n1_c.err; tail -c 1500 $O/r02_bench_n1_c.json; tail -5 $O/r02_bench_n1_c.err
MLA_FUSE_SWIGLU_BWD=0 timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_c_nosb.json 2> $O/r02_bench_n1_c_nosb.err; tail -c 600 $O/r02_bench_n1_c_nosb.json
MLA_ATTN_BWD=sm100 timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_c_attnv1.json 2> $O/r02_bench_n1_c_attnv1.err; tail -c 600 $O/r02_bench_n1_c_attnv1.json
