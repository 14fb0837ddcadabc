#!/bin/bash
# round 2: block GEMM tests + bench record
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "block or tiled" > gpurun_out/pytest_blockmm.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_blockmm.log
tail -5 gpurun_out/pytest_blockmm.log
timeout 300 python bench.py --workload blocks --steps 20 --warmup 3 > gpurun_out/bench_blocks.json 2> gpurun_out/bench_blocks.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_blocks.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roundtrip_rel_l2'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']); print(d['forward_by_block_size'])"
tail -3 gpurun_out/bench_blocks.err
