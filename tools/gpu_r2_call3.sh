#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_attention_sm100_gpu.py -x -q -s -m gpu > $O/r02_attn_tests.log 2>&1; echo "rc=$?" >> $O/r02_attn_tests.log
tail -15 $O/r02_attn_tests.log
if grep -q "rc=0" $O/r02_attn_tests.log; then export MLA_ATTN_BWD=sm100v2; else export MLA_ATTN_BWD=sm100; fi
echo "MLA_ATTN_BWD=$MLA_ATTN_BWD"
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_attention_sm100_gpu.py > $O/r02_gpu_tests_call3.log 2>&1; echo "rc=$?" >> $O/r02_gpu_tests_call3.log
tail -12 $O/r02_gpu_tests_call3.log
timeout 600 python tools/parity_table.py > $O/r02_parity_table.md 2> $O/r02_parity_table.err; echo "parity rc=$?"
timeout 600 python tools/ref_gpu.py attn --out $O/r02_attn_vs_flash_v2.json > $O/r02_attn_vs_flash_v2.log 2>&1; tail -4 $O/r02_attn_vs_flash_v2.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_b.json 2> $O/r02_bench_n1_b.err; tail -c 1800 $O/r02_bench_n1_b.json; tail -5 $O/r02_bench_n1_b.err
