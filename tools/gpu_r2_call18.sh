#!/bin/bash
O=gpurun_out
mkdir -p $O
for tool in memcheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > $O/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> $O/r02_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|rc=|Invalid|Uninit" $O/r02_sanitizer_$tool.log | head -8
done
