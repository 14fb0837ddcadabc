#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "motion or reference_tools or draw" 2>&1 | tail -3
for env in "X=1" "DSP_DCT_NO_FAST_OPS=1"; do
echo "== motion3d $env"
env $env timeout 300 python bench.py --workload motion3d --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('motion3d', d['value'], d['ms_per_step'], d['u8_roundtrip_exact'], d['u8_mismatches'], 'e2e', d['e2e']['value']); [print('   ', k['plan'],k['kernel'],k['n'],round(k['avg_ms'],3)) for k in d['passes_Y']]"
done
