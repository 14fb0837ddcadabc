#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)" > gpurun_out/lscpu.txt 2>&1
timeout 120 python tools/gpu_pcie_probe.py > gpurun_out/pcie_1.json 2>/dev/null; cat gpurun_out/pcie_1.json
N=$(nvidia-smi -L | wc -l)
if [ $N -gt 1 ]; then
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/gpu_pcie_probe.py > gpurun_out/pcie_$N.json 2>/dev/null; cat gpurun_out/pcie_$N.json
fi
head -12 gpurun_out/topo.txt
