#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_denoise_gpu.py tests/test_decode_stack_gpu.py tests/test_predict_action_gpu.py -x -q > $O/r02_denoise_tests.log 2>&1; tail -6 $O/r02_denoise_tests.log
MLA_DECODE_SKINNY=0 timeout 600 python tools/bench_denoise.py > $O/a.log 2>&1; tail -1 $O/a.log | cut -c1-900; cp $O/denoise_T0.json $O/r02_denoise_T0_gemv.json
MLA_DECODE_SKINNY=1 timeout 600 python tools/bench_denoise.py > $O/b.log 2>&1; tail -1 $O/b.log | cut -c1-900; cp $O/denoise_T0.json $O/r02_denoise_T0_skinny.json
MLA_DECODE_SKINNY=0 timeout 600 python tools/bench_denoise.py --T 15 > $O/c.log 2>&1; tail -1 $O/c.log | cut -c1-900; cp $O/denoise_T15.json $O/r02_denoise_T15_gemv.json
MLA_DECODE_SKINNY=1 timeout 600 python tools/bench_denoise.py --T 15 > $O/d.log 2>&1; tail -1 $O/d.log | cut -c1-900; cp $O/denoise_T15.json $O/r02_denoise_T15_skinny.json
