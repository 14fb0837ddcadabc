#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -x -q 2>&1 | tail -3
for env in "X=1" "DSP_DCT_RING_PANEL_MB=16" "DSP_DCT_RING_PANEL_MB=32"; do
  echo "== plane8192 $env"
  env $env timeout 300 python bench.py --workload plane8192 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roundtrip_rel_l2'], [ (k['plan'],k['kernel'],round(k['avg_ms'],4), round(k['achieved_gbs'])) for k in d['kernels']])"
done
bash tools/gpu_trace.sh 2>&1 | tail -5
timeout 300 python scratch/specperf.py 2>&1 | tail -12
