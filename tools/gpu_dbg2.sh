#!/bin/bash
for env in "X=1" "DSP_DCT_RING_NODISCARD=1"; do
echo "== $env"
env $env DSP_DCT_SPLIT_PANEL_MB=1 timeout 120 python - <<'PY' 2>&1 | tail -12
import numpy as np
from dspfun_b200 import capi, REDFT10, Plan
from oracle import dct as od
lib = capi.load()
for shape in [(8192, 96), (8192, 128), (8192, 64), (8192, 192), (4096, 96)]:
    h, w = shape
    x = np.random.default_rng(0).random((h, w, 1)).astype(np.float32)
    p = Plan.interleaved_2d("f", h, w, 1, REDFT10, lib=lib)
    errs = []
    for rep in range(3):
        y = p.execute_host(x.copy())
        ref = od.dctn_fast(x.astype(np.float64), [od.REDFT10] * 2, axes=(0, 1))
        errs.append(od.rel_l2(y, ref))
        # which columns are wrong?
    bad = np.where(np.abs(y[:, :, 0] - ref[:, :, 0]).max(axis=0) > 1e-2 * np.abs(ref).max())[0]
    print(shape, ["%.2e" % e for e in errs], "bad cols:", bad[:8], len(bad))
    p.destroy()
PY
done
