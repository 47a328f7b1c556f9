#!/bin/bash
# round-2 GPU call 2: pipelined attention backward (tests under a short timeout first), full GPU suite, parity table,
# attention head-to-head, bench (SwiGLU epilogue on / off), the reference's CPU arm incl. one full-depth step
set -x
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_attention_sm100_gpu.py -x -q -s -m gpu > $O/r02_attn_tests.log 2>&1; echo "rc=$?" >> $O/r02_attn_tests.log
tail -15 $O/r02_attn_tests.log
if grep -q "rc=0" $O/r02_attn_tests.log; then export MLA_ATTN_BWD=sm100v2; else export MLA_ATTN_BWD=sm100; fi
echo "MLA_ATTN_BWD=$MLA_ATTN_BWD"
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_attention_sm100_gpu.py > $O/r02_gpu_tests_call2.log 2>&1; echo "rc=$?" >> $O/r02_gpu_tests_call2.log
tail -12 $O/r02_gpu_tests_call2.log
timeout 600 python tools/parity_table.py > $O/r02_parity_table.md 2> $O/r02_parity_table.err; echo "parity rc=$?"
timeout 600 python tools/ref_gpu.py attn --out $O/r02_attn_vs_flash_v2.json > $O/r02_attn_vs_flash_v2.log 2>&1; tail -4 $O/r02_attn_vs_flash_v2.log
timeout 900 python bench.py --steps 8 --warmup 3 > $O/r02_bench_n1_a.json 2> $O/r02_bench_n1_a.err; tail -c 3000 $O/r02_bench_n1_a.json; tail -5 $O/r02_bench_n1_a.err
MLA_FUSE_SWIGLU=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_noswiglu.json 2> $O/r02_bench_n1_noswiglu.err; tail -c 1500 $O/r02_bench_n1_noswiglu.json
MLA_ATTN_BWD=sm100 MLA_FUSE_SWIGLU=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-also --no-cpu-baseline > $O/r02_bench_n1_r1cfg.json 2> $O/r02_bench_n1_r1cfg.err; tail -c 1500 $O/r02_bench_n1_r1cfg.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > $O/r02_ref_cpu_arm.json 2> $O/r02_ref_cpu_arm.err; cat $O/r02_ref_cpu_arm.json
timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 --layers-cpu 32 > $O/r02_ref_cpu_fulldepth.json 2> $O/r02_ref_cpu_fulldepth.err; cat $O/r02_ref_cpu_fulldepth.json
