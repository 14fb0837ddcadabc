#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload batch1024 --planes 1024 --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batch1024 (1024 images)', d['value'], d['ms_per_step'])
for k in d['kernels']: print('   ', k['plan'], k['kernel'], k['n'], round(k['avg_ms'],3), round(k['achieved_gbs']))"
