#!/bin/bash
# two GPUs: the multi-GPU parity tests, then the default bench and the motion volume under torchrun
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_2gpu.log 2>&1; echo "2gpu pytest rc=$?" >> gpurun_out/pytest_2gpu.log
tail -6 gpurun_out/pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_2gpu.json"))
    print("plane8192", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
    for n, r in d["records"].items():
        if "error" in r: print(n, r); continue
        print(n, r["value"], r["ms_per_step"], "e2e", r["e2e"]["value"], r.get("u8_roundtrip_exact"), r["config"].get("exchange"), r.get("comm_share_Y"), r.get("nvlink_bytes_per_gpu_per_step"))
        for k in r.get("passes_Y", []): print("   ", k)
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_2gpu.err
DSP_DIST_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --workload motion3d --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('motion3d nccl', d['value'], d['ms_per_step'], d['u8_roundtrip_exact'], d['config']['exchange'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('reference arm under torchrun', d['value'], d['ms_per_step'], d['cpu_baseline']['cores'])"
