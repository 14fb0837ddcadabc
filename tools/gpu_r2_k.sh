#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_round2.py -x -q -k "ring_column" 2>&1 | tail -3
for env in "X=1" "DSP_DCT_NO_COLRING_INV=1"; do
  echo "== plane8192 $env"
  env $env timeout 90 python bench.py --workload plane8192 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roundtrip_rel_l2'], [ (k['plan'],k['kernel'],round(k['avg_ms'],4), round(k['achieved_gbs'])) for k in d['kernels']])"
done
timeout 120 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_round2.py -x -q -k "ring_column_subpasses_small_panels and (shape0 or shape2)" 2>&1 | tail -4
timeout 400 python -m pytest tests -m gpu -x -q -k "not fullsize" 2>&1 | tail -3
timeout 100 python scratch/specperf.py 2>&1 | tail -12
