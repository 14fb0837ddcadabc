#!/bin/bash
# eight GPUs: the default bench under torchrun (independent planes / images / blocks weak or strong; motion volume with the
# exchange), bounded tightly
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_4gpu.json"))
    print("plane8192", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
    for n, r in d["records"].items():
        if "error" in r: print(n, r); continue
        print(n, r["value"], r["ms_per_step"], "e2e", (r.get("e2e") or {}).get("value"), r.get("u8_roundtrip_exact"), r["config"].get("exchange"), r.get("comm_share_Y"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -5 gpurun_out/bench_4gpu.err
