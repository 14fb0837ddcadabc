#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 42 -c 6 -f -o gpurun_out/prof_motion python bench.py --workload motion3d --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_motion.log 2>&1
ls -la gpurun_out/prof_motion.ncu-rep
