#!/bin/bash
# round 2, last visit: whole GPU suite, smoke, reference arm, default bench (C3 headline + C4 / C5 / blocks records), the other
# workloads, launch list, ncu --set full of the ring kernels and of the two tcgen05 kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    r = json.load(open("gpurun_out/bench_reference.json"))
    print("reference", r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"], {k: v["value"] for k, v in r.get("records", {}).items()})
    d = json.load(open("gpurun_out/bench.json"))
    print("plane8192", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["round_trip_frac"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], d["clocks"])
    for k in d["kernels"]: print("   ", k["plan"], k["kernel"], round(k["avg_ms"], 4), round(k["achieved_gbs"]))
    for n, r in d["records"].items():
        if "error" in r: print(n, r); continue
        print(n, r["value"], r["ms_per_step"], r["roofline"].get("round_trip_frac", r["roofline"]["frac"]), "e2e", r["e2e"]["value"], r.get("u8_roundtrip_exact"), "cpu", (r.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("bench parse failed", e)
PY
for wl in plane8192_f64 spec512 plane4096x3; do
  timeout 120 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$wl.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/bench_$wl.json')); print('$wl', d['value'], d['ms_per_step'], d['roofline']['round_trip_frac'], 'e2e', d['e2e']['value'])"
done
timeout 200 python bench.py --workload zoom2x --steps 5 --warmup 3 > gpurun_out/bench_zoom2x.json 2>/dev/null
DSP_ZOOM_NO_TC=1 timeout 200 python bench.py --workload zoom2x --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_zoom2x_simt.json 2>/dev/null
python -c "
import json
for f in ('bench_zoom2x', 'bench_zoom2x_simt'):
    d=json.load(open('gpurun_out/%s.json' % f)); print(f, d['value'], d['ms_per_step'], d['ms_per_frame_pageable_output'], d['create_ms'], d['dense_path_frame'], (d.get('cpu_baseline') or {}).get('value'))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --workload plane8192 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_row_ring|k_col_ring" -s 8 -c 3 -f -o gpurun_out/prof_ring python bench.py --workload plane8192 --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_ring.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_block_mm" -s 4 -c 1 -f -o gpurun_out/prof_blockmm python bench.py --workload blocks --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_blockmm.log 2>&1
cat > /tmp/zprof.py <<'PY'
import sys
sys.path.insert(0, ".")
import numpy as np
from dspfun_b200 import zoom as gz
z = gz.Zoom(np.random.default_rng(3).random((2048, 2048, 3), dtype=np.float32))
for _ in range(2): z.frame(scale=(3, 2), basis="centered", pinned=True)
PY
timeout 300 ncu --set full --clock-control none -k regex:"k_gemm_tf32x3" -s 6 -c 2 -f -o gpurun_out/prof_gemmtc python /tmp/zprof.py > gpurun_out/ncu_gemmtc.log 2>&1
ls -la gpurun_out/*.ncu-rep
