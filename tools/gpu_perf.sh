#!/bin/bash
# perf iteration: parity subset, device-resident bench, optional ncu full capture of the pass kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "reference_config or device_resident or golden" > gpurun_out/pytest_perf.log 2>&1; tail -3 gpurun_out/pytest_perf.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS} > gpurun_out/bench_perf.json 2> gpurun_out/bench_perf.err; tail -2 gpurun_out/bench_perf.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_perf.json'))
print("value %.1f Gpx/s  ms/step %.3f  rt_frac %.3f" % (d['value'], d['ms_per_step'], d['roofline']['round_trip_frac']))
for k in d['kernels']: print("  %s/%s n=%d grid=%d smem=%d: %.3f ms  %.0f GB/s" % (k['plan'], k['kernel'], k['n'], k['grid'], k['smem_bytes'], k['avg_ms'], k['achieved_gbs']))
PY
if [ -n "$NCU_FULL" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 12 -c 4 -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS} > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof.ncu-rep
fi
