#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_skinny_gpu.py -x -q > $O/r02_skinny_tests.log 2>&1; tail -15 $O/r02_skinny_tests.log
