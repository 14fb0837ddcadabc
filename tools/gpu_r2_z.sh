#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_round2.py -m gpu -x -q -k "c5 or motion" > gpurun_out/pytest_z.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_z.log
tail -3 gpurun_out/pytest_z.log
timeout 300 python bench.py --workload motion3d --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('motion3d', d['value'], d['ms_per_step'], d['u8_roundtrip_exact'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
for k in d['passes_Y']: print('   ', k['plan'], k['kernel'], k['n'], round(k['avg_ms'],3))"
