#!/bin/bash
# round 2: ncu of the motion3d passes (generic-radix kernels at 1920 / 1080)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 30 -c 8 -f -o gpurun_out/prof_m3d python bench.py --workload motion3d --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_m3d.log 2>&1
tail -3 gpurun_out/ncu_m3d.log; ls -la gpurun_out/prof_m3d.ncu-rep
