#!/bin/bash
O=gpurun_out
mkdir -p $O
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > $O/r02_final_ddp2.json 2> $O/r02_final_ddp2.err
tail -c 1500 $O/r02_final_ddp2.json; tail -2 $O/r02_final_ddp2.err | cut -c1-300
