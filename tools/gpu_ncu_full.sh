#!/bin/bash
# ncu --set full of selected kernels: args: <kernel regex> <skip> <count> [env...]
mkdir -p gpurun_out
K=$1; S=$2; C=$3; shift 3
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -f -o gpurun_out/prof2 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/prof2.ncu-rep
