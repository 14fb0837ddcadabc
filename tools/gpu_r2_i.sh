#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize.py -x -q > gpurun_out/pytest_fullsize.log 2>&1; echo "fullsize rc=$?" >> gpurun_out/pytest_fullsize.log
tail -12 gpurun_out/pytest_fullsize.log
