#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/r02_final_gputests.log 2>&1; tail -4 $O/r02_final_gputests.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_final_refarm.json 2> $O/r02_final_refarm.err; tail -c 700 $O/r02_final_refarm.json
