#!/bin/bash
for env in "X=1" "DSP_DCT_RING_NODISCARD=1"; do
  for rep in 1 2; do
  echo "== plane8192 $env"
  env $env timeout 300 python bench.py --workload plane8192 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roundtrip_rel_l2'], [ (k['plan'],k['kernel'],round(k['avg_ms'],4), round(k['achieved_gbs'])) for k in d['kernels']])"
  done
done
