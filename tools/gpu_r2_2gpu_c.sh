#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_2gpu.json"))
print("plane8192", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
for n, r in d["records"].items():
    if "error" in r: print(n, r); continue
    print(n, r["value"], r["ms_per_step"], "e2e", (r.get("e2e") or {}).get("value"), r.get("u8_roundtrip_exact"), r["config"].get("exchange"), r.get("comm_share_Y"))
PY
