#!/bin/bash
# round 2, visit A: parity suite, smoke, default bench (headline + C4 + C5 records), L2 chunk sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("plane8192", d["value"], d["ms_per_step"], d["roofline"]["round_trip_frac"], "e2e", d["e2e"]["value"])
    for k in d["kernels"]: print("   ", k["plan"], k["kernel"], k["avg_ms"], k["achieved_gbs"])
    for n, r in d["records"].items():
        if "error" in r: print(n, r); continue
        print(n, r["value"], r["ms_per_step"], r["roofline"], "e2e", r["e2e"], r.get("u8_roundtrip_exact"), r.get("u8_mismatches"))
        for k in r.get("kernels", []) + r.get("passes_Y", []): print("   ", k)
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/bench.err
for mb in 0 16 24 32 48 64 96; do
  echo "== L2 chunk $mb MB"
  DSP_DCT_L2_CHUNK_MB=$mb timeout 300 python bench.py --workload batch1024 --planes 512 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('batch1024', d['value'], d['ms_per_step'], [ (k['plan'],k['kernel'],round(k['avg_ms'],3)) for k in d['kernels']])"
  DSP_DCT_L2_CHUNK_MB=$mb timeout 300 python bench.py --workload spec512 --planes 512 --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('spec512', d['value'], d['ms_per_step'])"
  DSP_DCT_L2_CHUNK_MB=$mb timeout 300 python bench.py --workload motion3d --steps 3 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('motion3d', d['value'], d['ms_per_step'], d['u8_roundtrip_exact'], [(k['plan'],k['kernel'],k['n'],round(k['avg_ms'],3)) for k in d['passes_Y']])"
done
