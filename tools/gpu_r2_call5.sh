#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_attention_sm100_gpu.py -x -q -s -m gpu > $O/r02_attn_tests.log 2>&1; echo "rc=$?" >> $O/r02_attn_tests.log
tail -8 $O/r02_attn_tests.log
timeout 300 python -m pytest tests/test_gemm2_gpu.py -x -q -m gpu > $O/r02_gemm2_tests.log 2>&1; echo "rc=$?" >> $O/r02_gemm2_tests.log
tail -4 $O/r02_gemm2_tests.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_attention_sm100_gpu.py --deselect tests/test_gemm2_gpu.py > $O/r02_gpu_tests_call5.log 2>&1; echo "rc=$?" >> $O/r02_gpu_tests_call5.log
tail -8 $O/r02_gpu_tests_call5.log
timeout 600 python tools/ref_gpu.py attn --out $O/r02_attn_vs_flash_v3.json > $O/r02_attn_vs_flash_v3.log 2>&1; tail -4 $O/r02_attn_vs_flash_v3.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_d.json 2> $O/r02_bench_n1_d.err; tail -c 1500 $O/r02_bench_n1_d.json; tail -5 $O/r02_bench_n1_d.err
MLA_FUSE_SWIGLU_BWD=0 timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_d_nosb.json 2> $O/r02_bench_n1_d_nosb.err; tail -c 600 $O/r02_bench_n1_d_nosb.json
timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_d2.json 2> $O/r02_bench_n1_d2.err; tail -c 600 $O/r02_bench_n1_d2.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 5 -o $O/r02_attn_ncu2 python tools/prof_attn.py > $O/r02_attn_ncu2.log 2>&1; tail -3 $O/r02_attn_ncu2.log
timeout 1500 bash tools/sanitize.sh $O
