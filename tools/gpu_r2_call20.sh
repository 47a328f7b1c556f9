#!/bin/bash
O=gpurun_out
mkdir -p $O
N=2
run() { tag=$1; shift; env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > $O/r02_adam_ddp2_$tag.json 2> $O/r02_adam_ddp2_$tag.err; python -c "import json;d=json.load(open('$O/r02_adam_ddp2_$tag.json'));print('$tag',d['ms_per_step'],d['e2e']['ms_per_step'],d['value'],d['e2e']['last_loss'])" || tail -5 $O/r02_adam_ddp2_$tag.err; }
run base A=1
run stream_full MLA_ADAM_STREAM=1 MLA_ADAM_LEAN=0
run base2 A=1
run stream_full2 MLA_ADAM_STREAM=1 MLA_ADAM_LEAN=0
