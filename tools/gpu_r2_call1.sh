#!/bin/bash
# round-2 GPU call 1: box probe, GPU test-suite incl. the opt-in SwiGLU-epilogue tests, reference-on-GPU goldens,
# flash-attn head-to-head, reference step timing
set -x
O=gpurun_out
mkdir -p $O/golden_gpu
{ nproc; free -g; lscpu | grep -iE "model name|amx|avx512_bf16|^CPU\(s\)|Thread|Socket" ; ls /root/reference 2>&1 | head -3; nvidia-smi --query-gpu=name,memory.total,power.limit --format=csv; } > $O/r02_box_probe.txt 2>&1
MLA_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -x -q > $O/r02_gpu_tests_call1.log 2>&1; echo "pytest rc=$?" >> $O/r02_gpu_tests_call1.log
timeout 900 python tests/golden/make_golden_gpu.py --out $O/golden_gpu > $O/r02_golden_gpu.log 2>&1; echo "rc=$?" >> $O/r02_golden_gpu.log
timeout 600 python tools/ref_gpu.py attn --out $O/r02_attn_vs_flash.json > $O/r02_attn_vs_flash.log 2>&1; echo "rc=$?" >> $O/r02_attn_vs_flash.log
timeout 1500 python tools/ref_gpu.py step --workload cfg2 --out $O/r02_ref_gpu_cfg2.json > $O/r02_ref_gpu_cfg2.log 2>&1; echo "rc=$?" >> $O/r02_ref_gpu_cfg2.log
tail -5 $O/r02_gpu_tests_call1.log; tail -8 $O/r02_golden_gpu.log; tail -5 $O/r02_attn_vs_flash.log; tail -6 $O/r02_ref_gpu_cfg2.log
