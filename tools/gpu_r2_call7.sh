#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_shared_prefix_gpu.py tests/test_attention_sm100_gpu.py -x -q -m gpu > $O/r02_shared_tests.log 2>&1; echo "rc=$?" >> $O/r02_shared_tests.log
tail -30 $O/r02_shared_tests.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_attention_sm100_gpu.py --deselect tests/test_shared_prefix_gpu.py > $O/r02_gpu_tests_call7.log 2>&1; echo "rc=$?" >> $O/r02_gpu_tests_call7.log
tail -5 $O/r02_gpu_tests_call7.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline --share-prefix > $O/r02_bench_n1_shared.json 2> $O/r02_bench_n1_shared.err; tail -c 1500 $O/r02_bench_n1_shared.json; tail -5 $O/r02_bench_n1_shared.err
timeout 600 python tools/bench_gemm_raster.py > $O/r02_gemm_raster.json 2> $O/r02_gemm_raster.err; tail -12 $O/r02_gemm_raster.err
