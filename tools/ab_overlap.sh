#!/bin/bash
# A/B of the stream-overlap switches and GEMM rasterisation on ONE box (boxes differ by ~5 % in power-capped clocks,
# so only same-call comparisons mean anything).  Writes gpurun_out/ab_*.json.
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{tag}.json").read().strip().splitlines()[-1])
    print(f"{tag:28s} step {d['ms_per_step']:8.2f} ms  e2e {d['e2e']['ms_per_step']:8.2f} ms  gemm {d['roofline']['achieved']} TF/s  "
          f"sm {d['clocks']['sm_mhz']} MHz  mem {d['config']['peak_mem_gb']} GB")
except Exception as ex:
    print(tag, "FAILED", ex)
    print(open(f"gpurun_out/ab_{tag}.err").read()[-1500:])
PY
}
run base          MLA_WGRAD_STREAM=0 MLA_ADAM_STREAM=0
run wgrad         MLA_WGRAD_STREAM=1 MLA_ADAM_STREAM=0
run wgrad_adam    MLA_WGRAD_STREAM=1 MLA_ADAM_STREAM=1
run wgrad_adam_dyn MLA_WGRAD_STREAM=1 MLA_ADAM_STREAM=1 MLA_DYNAMIC_TILES=1
run base_group16  MLA_WGRAD_STREAM=0 MLA_ADAM_STREAM=0 MLA_GEMM_GROUP_M=16
