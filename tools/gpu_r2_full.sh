#!/bin/bash
# round 2: the whole GPU suite, smoke, the default bench (headline + C4 + C5 records), the reference arm, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
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
        for k in r.get("kernels", []) + r.get("passes_Y", []): print("   ", {a: (round(b, 4) if isinstance(b, float) else b) for a, b in k.items() if a in ("plan", "kernel", "n", "avg_ms", "achieved_gbs")})
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --workload plane8192 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/launches.csv
