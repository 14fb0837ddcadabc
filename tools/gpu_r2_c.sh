#!/bin/bash
# round 2, visit C: ring column pass kernel -- parity (bounded), then timings by panel size, full suite, ncu
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -k "ring_column" > gpurun_out/pytest_cring.log 2>&1; echo "cring pytest rc=$?" >> gpurun_out/pytest_cring.log
tail -15 gpurun_out/pytest_cring.log
if ! grep -q "cring pytest rc=0" gpurun_out/pytest_cring.log; then
  timeout 120 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_round2.py -x -q -k "ring_column_subpasses_small_panels and shape0" 2>&1 | tail -30
  exit 0
fi
for env in "DSP_DCT_RING_PANEL_MB=16" "DSP_DCT_RING_PANEL_MB=32" "DSP_DCT_RING_PANEL_MB=64" "DSP_DCT_NO_COLRING=1"; do
  echo "== plane8192 $env"
  env $env timeout 300 python bench.py --workload plane8192 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roundtrip_rel_l2'], [ (k['plan'],k['kernel'],round(k['avg_ms'],4), round(k['achieved_gbs'])) for k in d['kernels']])"
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 120 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_round2.py -x -q -k "ring_column_subpasses_small_panels and (shape0 or shape2)" 2>&1 | tail -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_col_ring -s 3 -c 1 -f -o gpurun_out/prof_cring python bench.py --workload plane8192 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_cring.log 2>&1
ls -la gpurun_out/prof_cring.ncu-rep
