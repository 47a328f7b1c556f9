#!/bin/bash
O=gpurun_out
mkdir -p $O
N=4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 6 --warmup 3 > $O/r02_final_ddp4.json 2> $O/r02_final_ddp4.err
tail -1 $O/r02_final_ddp4.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['n_gpus'],d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['gradient_exchange']['allreduce_alone'])"
