#!/bin/bash
# Same-box A/B of bench.py under different switches (boxes differ by ~5 % in power-capped clocks, so only same-call
# comparisons mean anything).  usage: tools/ab_bench.sh tag1:ENV=V,ENV2=V tag2:...   -> gpurun_out/ab_<tag>.json
mkdir -p gpurun_out
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}; [ "$envs" = "$spec" ] && envs=""
  env $(echo $envs | tr ',' ' ') timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} \
      > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{tag}.json").read().strip().splitlines()[-1])
    print(f"{tag:22s} step {d['ms_per_step']:8.2f} ms  e2e {d['e2e']['ms_per_step']:8.2f} ms  gemm {d['roofline']['achieved']} TF/s  "
          f"frac {d['roofline']['step_frac_of_peak']}  sm {d['clocks']['sm_mhz']} MHz  mem {d['config']['peak_mem_gb']} GB")
except Exception as ex:
    print(tag, "FAILED", ex)
    print(open(f"gpurun_out/ab_{tag}.err").read()[-1500:])
PY
done
