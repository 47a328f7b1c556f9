#!/bin/bash
# 2-GPU data-parallel experiments (cfg4 = the full configuration): reduce dtype and NCCL CTA budget
set -x
O=gpurun_out
mkdir -p $O
N=${1:-2}
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps 6 --warmup 3 > $O/r02_ddp${N}_$tag.json 2> $O/r02_ddp${N}_$tag.err
  tail -c 1200 $O/r02_ddp${N}_$tag.json; tail -3 $O/r02_ddp${N}_$tag.err
}
timeout 900 python bench.py --steps 6 --warmup 3 --workload cfg4 --no-cpu-baseline > $O/r02_ddp1_cfg4.json 2> $O/r02_ddp1_cfg4.err; tail -c 700 $O/r02_ddp1_cfg4.json
run fp32 A=1
run bf16 MLA_GRAD_REDUCE_DTYPE=bf16
run fp32_cta8 NCCL_MAX_CTAS=8
run fp32_cta4 NCCL_MAX_CTAS=4
run fp32_nvls NCCL_ALGO=NVLS NCCL_DEBUG=WARN
