#!/bin/bash
timeout 200 python bench.py --workload plane8192 --steps 20 --warmup 5 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('plane8192', d['value'], d['ms_per_step'])
for k in d['kernels']: print('   ', k['plan'], k['kernel'], k['n'], round(k['avg_ms'],4), round(k['achieved_gbs']))"
