#!/bin/bash
# round 2: generic kernels at 3 CTAs/SM: parity on the generic sizes, motion3d and C4/C3 timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_q.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_q.log
tail -3 gpurun_out/pytest_q.log
timeout 300 python bench.py --workload motion3d --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('motion3d', d['value'], d['ms_per_step'], d['u8_roundtrip_exact'])
for k in d['passes_Y']: print('   ', k['plan'], k['kernel'], k['n'], round(k['avg_ms'],3))"
