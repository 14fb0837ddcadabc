#!/bin/bash
mkdir -p gpurun_out
DSP_DCT_SPLIT_PANEL_MB=1 DSP_DCT_TRACE=1 timeout 120 python -u - <<'PY' 2>&1 | tail -40
import numpy as np, sys
from dspfun_b200 import capi, REDFT10
from tests import cases
lib = capi.load()
print("loaded", flush=True)
try:
    print(cases.check_interleaved_2d(lib, "f", 4096, 64, 1, REDFT10), flush=True)
except Exception as e:
    print("EXC", repr(e), flush=True)
print("done", flush=True)
PY
echo "rc=$?"
