#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none -k regex:attn_ -c 5 --csv --page raw --log-file $O/r02_attn_final_raw.csv python tools/prof_attn.py > $O/r02_attn_final.log 2>&1; tail -2 $O/r02_attn_final.log
python tools/ncu_summary.py $O/r02_attn_final_raw.csv $O/r02_ncu_attn_final_summary.json && python -c "
import json
for r in json.load(open('$O/r02_ncu_attn_final_summary.json')): print(r['kernel'][:40], r['time_ms'], r['tensor_pipe_active_pct'], r['grid'], r['regs'], r['dram_GBs'])"
