#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "motion_volume" > gpurun_out/pytest_2gpu_b.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_2gpu_b.log
tail -3 gpurun_out/pytest_2gpu_b.log
for f in 0 1; do
if [ $f = 1 ]; then export DSP_DIST_FUSE_COEFF=1; fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$f bench.py --gpus 2 --steps 10 --warmup 3 --workload motion3d --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('motion3d 2 GPUs fused_in_pass=$f', d['value'], d['ms_per_step'], d['u8_roundtrip_exact'], d['config']['exchange'])
for k in d['passes_Y']: print('   ', k['plan'], k['kernel'], k['n'], round(k['avg_ms'],3))"
done
