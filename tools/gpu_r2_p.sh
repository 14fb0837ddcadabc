#!/bin/bash
# round 2: tensor-core GEMM for zoom's dense path
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q -k "zoom" > gpurun_out/pytest_p.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_p.log
tail -15 gpurun_out/pytest_p.log
timeout 300 python - <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np
from dspfun_b200 import zoom as gz
px = np.random.default_rng(3).random((2048, 2048, 3), dtype=np.float32)
z = gz.Zoom(px)
for kw in (dict(scale=(3, 2)), dict(scale=(3, 2), basis="centered")):
    z.frame(pinned=True, **kw)
    t0 = time.perf_counter(); o = z.frame(pinned=True, **kw); dt = time.perf_counter() - t0
    print("2048^2 ->", o.shape, z.last_path, "%.1f ms" % (dt * 1e3))
z.destroy()
PY
