#!/bin/bash
O=gpurun_out
mkdir -p $O
for hv in 0 1 2; do
  MLA_GEMM_L2_HINTS=$hv timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm2 -c 12 --csv --log-file $O/r02_gemm12_h$hv.csv python tools/ncu_gemm.py > $O/r02_ncu_gemm_h$hv.log 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('$O/r02_gemm12_h$hv.csv')))
i=next(k for k,r in enumerate(rows) if 'Kernel Name' in r)
h=rows[i]; mn=h.index('Metric Name'); mu=h.index('Metric Unit'); mv=h.index('Metric Value'); idc=h.index('ID')
S={'byte':1e-6,'Kbyte':1e-3,'Mbyte':1.0,'Gbyte':1e3}
agg={}
for r in rows[i+1:]:
    if len(r)<=mv: continue
    k=int(r[idc]); agg.setdefault(k,{})
    v=float(r[mv].replace(',',''))
    if r[mn].startswith('dram__bytes'): v*=S.get(r[mu],1.0)
    agg[k][r[mn]]=v
tot_r=sum(a.get('dram__bytes_read.sum',0) for a in agg.values()); tot_w=sum(a.get('dram__bytes_write.sum',0) for a in agg.values())
print('hints=$hv read MB',round(tot_r,1),'write MB',round(tot_w,1),'reads:',[round(agg[k].get('dram__bytes_read.sum',0)) for k in sorted(agg)])
PY
done
run() { tag=$1; shift; env "$@" timeout 900 python bench.py --steps 6 --warmup 3 --workload cfg2 --no-also --no-cpu-baseline > $O/r02_h_$tag.json 2> $O/r02_h_$tag.err; python -c "import json;d=json.load(open('$O/r02_h_$tag.json'));print('$tag',d['ms_per_step'],d['e2e']['ms_per_step'],d['value'],d['clocks']['sm_mhz'])" || tail -3 $O/r02_h_$tag.err; }
run h0 MLA_GEMM_L2_HINTS=0
run h1 MLA_GEMM_L2_HINTS=1
run h2 MLA_GEMM_L2_HINTS=2
run h0b MLA_GEMM_L2_HINTS=0
