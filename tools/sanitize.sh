#!/bin/bash
# compute-sanitizer over tools/sanitize_target.py; logs under gpurun_out/ (copy the summaries to profiles/)
O=${1:-gpurun_out}
mkdir -p $O
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py > $O/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> $O/r02_sanitizer_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|rc=" $O/r02_sanitizer_$tool.log
done
