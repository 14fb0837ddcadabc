#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "sweep or dquant" > gpurun_out/pytest_last.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_last.log
tail -4 gpurun_out/pytest_last.log
