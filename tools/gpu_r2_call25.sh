#!/bin/bash
O=gpurun_out
mkdir -p $O
cat > /tmp/one.py <<'PY'
import ctypes as C, sys, os, torch
sys.path.insert(0, os.getcwd())
from mla_b200 import _lib, ops
lib=_lib.lib(); lib.mla_gemm_set_mode(C.c_int32(0))
M,N,K=1327104,192,96
a=torch.randn(M,K,device="cuda").bfloat16(); w=torch.randn(N,K,device="cuda").bfloat16(); b=torch.randn(N,device="cuda").bfloat16(); c=torch.empty(M,N,device="cuda",dtype=torch.bfloat16)
for _ in range(3): ops.gemm(a,w,bias=b,out=c)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none -k regex:gemm_bf16 -s 2 -c 1 --csv --page raw --log-file $O/r02_tower_gemm_ncu.csv python /tmp/one.py > $O/ncu_one.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_tower_gemm_ncu.csv')))
i=next(k for k,r in enumerate(rows) if 'Kernel Name' in r)
h=rows[i]; v=rows[i+2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','l1tex__t_bytes.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum','sm__cycles_active.avg','launch__grid_size','lts__t_sectors_op_write.sum','lts__t_sectors_op_read.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
for k in want:
    if k in h: print(k, v[h.index(k)], rows[i+1][h.index(k)])
for k in h:
    if 'issue_stalled' in k and k.endswith('per_issue_active.ratio'):
        try:
            x=float(v[h.index(k)])
            if x>0.3: print(k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), x)
        except: pass
PY
