#!/bin/bash
# two GPUs, one process: motion's luma volume (256 x 1080 x 1920 float) through a rank-3 host plan on 1 and 2 GPUs
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/mg_timing.log
import ctypes, sys, time
sys.path.insert(0, ".")
import numpy as np
from dspfun_b200 import Plan, REDFT01, REDFT10, capi
lib = capi.load()
D, H, W = 256, 1080, 1920
n = D * H * W
buf = lib.dsp_dct_alloc(n * 4)
x = np.frombuffer((ctypes.c_char * (n * 4)).from_address(buf), dtype=np.float32).reshape(D, H, W)
rng = np.random.default_rng(1)
for z in range(D):
    x[z] = rng.integers(0, 256, (H, W)).astype(np.float32)
ref = x[::37].copy()
for ng in (1, 2):
    lib.dsp_dct_plan_with_ngpus(ng)
    fwd = Plan("f", [D, H, W], [REDFT10] * 3); inv = Plan("f", [D, H, W], [REDFT01] * 3)
    fwd.execute_host(x); inv.execute_host(x); x *= np.float32(1.0 / (8.0 * n))
    t0 = time.perf_counter()
    fwd.execute_host(x)
    t1 = time.perf_counter()
    inv.execute_host(x)
    t2 = time.perf_counter()
    x *= np.float32(1.0 / (8.0 * n))
    err = float(np.abs(x[::37] - ref).max())
    print("ngpus %d (plan reports %d): forward %.1f ms, inverse %.1f ms, round trip %.2f Gpixel/s host to host, max |err| after two round trips %.3g"
          % (ng, lib.dsp_dct_plan_ngpus(fwd._h), (t1 - t0) * 1e3, (t2 - t1) * 1e3, n / (t2 - t0) / 1e9, err))
    fwd.destroy(); inv.destroy()
PY
