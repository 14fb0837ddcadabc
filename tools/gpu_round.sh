#!/bin/bash
# One GPU-box visit: parity suite, smoke, bench, ncu launch list.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
# launch list of two full timed steps (plane8192: 66 launches per step; the 3 warm-up steps are skipped)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 198 -c 132 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/launches.csv
if [ -n "$NCU_FULL" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 12 -c 4 -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof.ncu-rep
fi
