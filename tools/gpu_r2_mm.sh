#!/bin/bash
# round 2: bring-up of the tcgen05 block DCT
mkdir -p gpurun_out
timeout 300 python tools/gpu_blockmm.py > gpurun_out/blockmm.log 2>&1; echo "rc=$?" >> gpurun_out/blockmm.log
tail -60 gpurun_out/blockmm.log
