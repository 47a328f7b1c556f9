#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/r02_final_gputests.log 2>&1; tail -2 $O/r02_final_gputests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py > $O/r02_final_bench.json 2> $O/r02_final_bench.err; python -c "
import json;d=json.load(open('$O/r02_final_bench.json'));print(d['ms_per_step'],d['value'],d['e2e']['ms_per_step'],d['e2e']['value'],d['roofline']['step_frac_of_peak'],d['clocks']['sm_mhz'],{k:(v['ms_per_step'],v['value']) for k,v in d['also'].items()})"; tail -2 $O/r02_final_bench.err
