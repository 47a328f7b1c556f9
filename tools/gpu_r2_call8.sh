#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_attention_sm100_gpu.py tests/test_shared_prefix_gpu.py -x -q -s -m gpu > $O/r02_attn_tests.log 2>&1; echo "rc=$?" >> $O/r02_attn_tests.log
tail -8 $O/r02_attn_tests.log
timeout 600 python tools/bench_gemm_raster.py > $O/r02_gemm_raster.json 2> $O/r02_gemm_raster.err; tail -12 $O/r02_gemm_raster.err
timeout 600 python tools/ref_gpu.py attn --out $O/r02_attn_vs_flash_v5.json > $O/r02_attn_vs_flash_v5.log 2>&1; tail -4 $O/r02_attn_vs_flash_v5.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_attention_sm100_gpu.py --deselect tests/test_shared_prefix_gpu.py > $O/r02_gpu_tests_call8.log 2>&1; echo "rc=$?" >> $O/r02_gpu_tests_call8.log
tail -5 $O/r02_gpu_tests_call8.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_f.json 2> $O/r02_bench_n1_f.err; tail -c 1000 $O/r02_bench_n1_f.json; tail -5 $O/r02_bench_n1_f.err
