#!/bin/bash
# round 2, visit B: the ring row kernel -- parity first (bounded: a hang must not cost the box), then A/B timings
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -k "ring" > gpurun_out/pytest_ring.log 2>&1; echo "ring pytest rc=$?" >> gpurun_out/pytest_ring.log
tail -15 gpurun_out/pytest_ring.log
if ! grep -q "ring pytest rc=0" gpurun_out/pytest_ring.log; then
  DSP_DCT_TRACE=1 timeout 60 python -c "
from tests import cases
from dspfun_b200 import capi, REDFT10
cases.check_interleaved_2d(capi.load(), 'f', 2, 8192, 1, REDFT10)" 2>&1 | tail -20
  exit 0
fi
timeout 120 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_round2.py -x -q -k "ring_row_kernel_planar and (shape0 or shape5 or shape9)" 2>&1 | tail -8
for env in "" "DSP_DCT_NO_RING=1"; do
  echo "== plane8192 $env"
  env $env timeout 300 python bench.py --workload plane8192 --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roundtrip_rel_l2'], [ (k['plan'],k['kernel'],round(k['avg_ms'],4), round(k['achieved_gbs'])) for k in d['kernels']])"
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
# ncu: the ring kernels of one step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_row_ring -s 6 -c 2 -f -o gpurun_out/prof_ring python bench.py --workload plane8192 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_ring.log 2>&1
ls -la gpurun_out/prof_ring.ncu-rep
