#!/bin/bash
# final evidence run: parity suite, smoke, default bench, launch list, the other workloads
bash tools/gpu_round.sh
for w in plane8192_f64 batch1024 spec512 plane4096x3 motion3d; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; head -c 400 gpurun_out/bench_$w.json; echo
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
