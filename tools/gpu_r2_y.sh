#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "tiled or block" > gpurun_out/pytest_y.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_y.log
tail -3 gpurun_out/pytest_y.log
bash tools/gpu_r2_t.sh
