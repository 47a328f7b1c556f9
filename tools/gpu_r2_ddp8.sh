#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
N=8
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --gpus $N --steps 6 --warmup 3 > $O/r02_ddp8_cfg4.json 2> $O/r02_ddp8_cfg4.err
grep -v "NCCL INFO" $O/r02_ddp8_cfg4.json | tail -c 2500
grep -c "NVLS" $O/r02_ddp8_cfg4.json $O/r02_ddp8_cfg4.err; grep -h "NVLS\|Algo\|algo" $O/r02_ddp8_cfg4.json $O/r02_ddp8_cfg4.err | head -8
tail -3 $O/r02_ddp8_cfg4.err | cut -c1-300
