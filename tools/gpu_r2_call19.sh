#!/bin/bash
O=gpurun_out
mkdir -p $O
run() { tag=$1; shift; env "$@" timeout 900 python bench.py --steps 6 --warmup 3 --workload cfg2 --no-also --no-cpu-baseline > $O/r02_adam_$tag.json 2> $O/r02_adam_$tag.err; python -c "import json;d=json.load(open('$O/r02_adam_$tag.json'));print('$tag',d['ms_per_step'],d['e2e']['ms_per_step'],d['value'],d['clocks']['sm_mhz'])" || tail -3 $O/r02_adam_$tag.err; }
timeout 300 python -m pytest tests/test_trainer_gpu.py -q 2>&1 | tail -2
run base A=1
run stream_lean1 MLA_ADAM_STREAM=1 MLA_ADAM_LEAN=1
run stream_lean2 MLA_ADAM_STREAM=1 MLA_ADAM_LEAN=2
run stream_full MLA_ADAM_STREAM=1 MLA_ADAM_LEAN=0
run base2 A=1
