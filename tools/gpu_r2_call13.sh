#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_decode_stack_gpu.py -x -q > $O/r02_stack_tests.log 2>&1; tail -25 $O/r02_stack_tests.log
MLA_DECODE_STACK=1 timeout 600 python tools/bench_denoise.py > $O/r02_denoise_T0_stack.log 2>&1; tail -2 $O/r02_denoise_T0_stack.log; cp $O/denoise_T0.json $O/r02_denoise_T0_stack.json
