#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q > $O/r02_final_gputests.log 2>&1; tail -5 $O/r02_final_gputests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_final_smoke.log 2>&1; tail -3 $O/r02_final_smoke.log
timeout 1500 python bench.py > $O/r02_final_bench.json 2> $O/r02_final_bench.err; tail -c 2500 $O/r02_final_bench.json; tail -3 $O/r02_final_bench.err
