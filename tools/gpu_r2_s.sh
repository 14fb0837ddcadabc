#!/bin/bash
mkdir -p gpurun_out
for v in tc simt; do
if [ $v = simt ]; then export DSP_ZOOM_NO_TC=1; fi
timeout 300 python - <<'PY'
import os, sys, time
sys.path.insert(0, ".")
import numpy as np
from dspfun_b200 import zoom as gz
for n in (1024, 2048, 4096):
    z = gz.Zoom(np.random.default_rng(3).random((n, n, 3), dtype=np.float32))
    kw = dict(scale=(3, 2), basis="centered", pinned=True)
    z.frame(**kw)
    t0 = time.perf_counter()
    for _ in range(3): o = z.frame(**kw)
    dt = (time.perf_counter() - t0) / 3
    fl = 3 * 2.0 * (o.shape[1] * n * n + o.shape[0] * o.shape[1] * n)
    print(os.environ.get("DSP_ZOOM_NO_TC", "tensor"), n, "->", o.shape[:2], z.last_path, "%.2f ms  %.1f TFLOP/s (frame incl. copy-out of %.0f MB)" % (dt * 1e3, fl / dt / 1e12, o.nbytes / 1e6))
    z.destroy()
PY
done
