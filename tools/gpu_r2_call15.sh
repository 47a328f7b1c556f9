#!/bin/bash
set -x
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_decoder_gpu.py -x -q > $O/r02_kernels_tests.log 2>&1; tail -3 $O/r02_kernels_tests.log
MLA_RMSNORM_BWD_PIPE=0 timeout 300 python tools/bench_kernels.py 2>&1 | grep -i "rmsnorm"
MLA_RMSNORM_BWD_PIPE=1 timeout 300 python tools/bench_kernels.py 2>&1 | grep -i "rmsnorm"
MLA_RMSNORM_BWD_PIPE=0 timeout 900 python bench.py --steps 6 --warmup 3 --workload cfg2 --no-also --no-cpu-baseline > $O/r02_bench_cfg2_rbpipe0.json 2> $O/err0.log; python -c "import json;d=json.load(open('$O/r02_bench_cfg2_rbpipe0.json'));print('pipe0',d['ms_per_step'],d['value'],d['clocks'])"
MLA_RMSNORM_BWD_PIPE=1 timeout 900 python bench.py --steps 6 --warmup 3 --workload cfg2 --no-also --no-cpu-baseline > $O/r02_bench_cfg2_rbpipe1.json 2> $O/err1.log; python -c "import json;d=json.load(open('$O/r02_bench_cfg2_rbpipe1.json'));print('pipe1',d['ms_per_step'],d['value'],d['clocks'])"
MLA_RMSNORM_BWD_PIPE=0 timeout 900 python bench.py --steps 6 --warmup 3 --workload cfg2 --no-also --no-cpu-baseline > $O/r02_bench_cfg2_rbpipe0b.json 2> $O/err0.log; python -c "import json;d=json.load(open('$O/r02_bench_cfg2_rbpipe0b.json'));print('pipe0 again',d['ms_per_step'],d['value'],d['clocks'])"
