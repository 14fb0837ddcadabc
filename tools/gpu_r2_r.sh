#!/bin/bash
mkdir -p gpurun_out
for tc in 8 16; do
echo "== DSP_DCT_TC=$tc (column kernels built for 4 CTAs/SM)"
DSP_DCT_TC=$tc timeout 300 python bench.py --workload motion3d --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('motion3d', d['value'], d['ms_per_step'], d['u8_roundtrip_exact'])
for k in d['passes_Y']: print('   ', k['plan'], k['kernel'], k['n'], round(k['avg_ms'],3))"
done
