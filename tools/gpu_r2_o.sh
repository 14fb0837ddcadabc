#!/bin/bash
# round 2: new GPU tests (tiled C session), zoom2x artefact, NVTX smoke
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "tiled or block" > gpurun_out/pytest_o.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_o.log
tail -4 gpurun_out/pytest_o.log
timeout 300 python bench.py --workload zoom2x --steps 5 --warmup 3 > gpurun_out/bench_zoom2x.json 2> gpurun_out/bench_zoom2x.err; echo "zoom rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_zoom2x.json')); print(d['value'], d['ms_per_step'], d['config'], d['create_ms'], d['even_sample_max_abs_err'], d['dense_path_frame'], d['cpu_baseline'])"
tail -3 gpurun_out/bench_zoom2x.err
DSP_DCT_NVTX=1 timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
