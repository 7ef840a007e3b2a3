#!/bin/bash
# parity tests + bench variants on a GPU box; prints a one-line summary per bench
cd /root/repo
VARS=${VARS:-"SB2_DEFAULT=1"}
TESTS=${TESTS:-tests/test_gpu_parity.py}
/usr/local/graft/bin/gpurun --timeout 1500 -- "timeout 900 python -m pytest $TESTS -x -q -m gpu 2>&1 | tail -25 > gpurun_out/s3_parity.log; for v in $VARS; do env \$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_\$v.json 2> gpurun_out/bench_\$v.err; done" 2>&1 | tail -3
tail -5 gpurun_out/s3_parity.log
for v in $VARS; do f=gpurun_out/bench_$v.json; echo "== $v"; python -c "
import json,sys
try:
    d=json.load(open('$f'))
    r=d['roofline']
    print(round(d['ms_per_step'],3), round(r['kernel_ms'],3), r['stage_ms'], round(d['e2e']['value']/1e6,1))
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/bench_$v.err').read()[-600:])
"; done
