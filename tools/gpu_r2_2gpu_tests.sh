#!/bin/bash
# two GPUs: the multi-GPU parity tests alone
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_2gpu.log 2>&1; echo "2gpu pytest rc=$?" >> gpurun_out/pytest_2gpu.log
tail -6 gpurun_out/pytest_2gpu.log
