#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --workload blocks --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_blocks.json 2> gpurun_out/bench_blocks.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_blocks.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], 'e2e', d['e2e']['value']); print(d['motion_tiled_8x8x8_quant'])"
tail -2 gpurun_out/bench_blocks.err
