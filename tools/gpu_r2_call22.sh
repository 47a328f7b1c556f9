#!/bin/bash
O=gpurun_out
mkdir -p $O
for cs in 0 1; do
  MLA_GEMM_CS_STORES=$cs timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm2 -c 12 --csv --log-file $O/r02_gemm12_cs$cs.csv python tools/ncu_gemm.py > $O/r02_ncu_gemm_cs$cs.log 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('$O/r02_gemm12_cs$cs.csv')))
i=next(k for k,r in enumerate(rows) if 'Kernel Name' in r)
h=rows[i]; kn=h.index('Kernel Name'); mn=h.index('Metric Name'); mu=h.index('Metric Unit'); mv=h.index('Metric Value'); idc=h.index('ID')
S={'byte':1e-6,'Kbyte':1e-3,'Mbyte':1.0,'Gbyte':1e3}
agg={}
for r in rows[i+1:]:
    if len(r)<=mv: continue
    k=int(r[idc]); agg.setdefault(k,{})
    v=float(r[mv].replace(',',''))
    if r[mn].startswith('dram__bytes'): v*=S.get(r[mu],1.0)
    agg[k][r[mn]]=v
tot_r=sum(a.get('dram__bytes_read.sum',0) for a in agg.values()); tot_w=sum(a.get('dram__bytes_write.sum',0) for a in agg.values())
print('cs=$cs launches',len(agg),'read MB',round(tot_r,1),'write MB',round(tot_w,1),'per launch MB',round((tot_r+tot_w)/max(len(agg),1),1))
for k in sorted(agg): print('   ',k, round(agg[k].get('dram__bytes_read.sum',0),1), round(agg[k].get('dram__bytes_write.sum',0),1))
PY
done
