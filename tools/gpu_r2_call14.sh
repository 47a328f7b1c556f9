#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_decode_stack_gpu.py tests/test_denoise_gpu.py -x -q > $O/r02_stack_tests.log 2>&1; tail -3 $O/r02_stack_tests.log
MLA_GEMV_ONE_COPY=0 timeout 600 python tools/bench_denoise.py > $O/r02_denoise_T0_rowcopies.log 2>&1; tail -1 $O/r02_denoise_T0_rowcopies.log | cut -c1-700; cp $O/denoise_T0.json $O/r02_denoise_T0_rowcopies.json
MLA_GEMV_ONE_COPY=1 timeout 600 python tools/bench_denoise.py > $O/r02_denoise_T0_onecopy.log 2>&1; tail -1 $O/r02_denoise_T0_onecopy.log | cut -c1-700; cp $O/denoise_T0.json $O/r02_denoise_T0_onecopy.json
MLA_DECODE_STACK=1 timeout 600 python tools/bench_denoise.py > $O/r02_denoise_T0_stack.log 2>&1; tail -1 $O/r02_denoise_T0_stack.log | cut -c1-700; cp $O/denoise_T0.json $O/r02_denoise_T0_stack.json
