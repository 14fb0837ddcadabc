#!/bin/bash
mkdir -p gpurun_out
for wl in spec512 plane4096x3; do
timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl', d['value'], d['ms_per_step'])
for k in d['kernels']: print('   ', k['plan'], k['kernel'], k['n'], round(k['avg_ms'],4), round(k['achieved_gbs']))"
done
