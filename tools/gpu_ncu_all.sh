#!/bin/bash
# ncu --set full of every kernel class of the plane8192 step (row fwd/inv, split sub-passes A/B fwd, A''/B'' inv)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e"
N="ncu --set full --clock-control none --import-source on -f"
timeout 600 $N -k regex:k_pass -s 6 -c 2 -o gpurun_out/prof_row $B > gpurun_out/ncu_row.log 2>&1
timeout 600 $N -k regex:k_split_[fo] -s 96 -c 2 -o gpurun_out/prof_splitfwd $B > gpurun_out/ncu_splitfwd.log 2>&1
timeout 600 $N -k regex:k_split_inv -s 96 -c 2 -o gpurun_out/prof_splitinv $B > gpurun_out/ncu_splitinv.log 2>&1
ls -la gpurun_out/*.ncu-rep
