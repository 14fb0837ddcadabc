#!/bin/bash
# round 2: ncu of the tcgen05 block DCT
mkdir -p gpurun_out
cat > /tmp/mmprof.py <<'PY'
import sys
sys.path.insert(0, ".")
import torch
from dspfun_b200 import capi
lib = capi.load()
P, H, W = 16, 2048, 2048
a = torch.randn(P, H, W, device="cuda"); o = torch.empty_like(a)
for B in (8, 64):
    for _ in range(2):
        lib.dsp_block_dct2d(b"f", a.data_ptr(), o.data_ptr(), P, H, W, B, capi.REDFT10, 1.0, None)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_block_mm -c 4 -f -o gpurun_out/prof_blockmm python /tmp/mmprof.py > gpurun_out/ncu_blockmm.log 2>&1
tail -3 gpurun_out/ncu_blockmm.log
ls -la gpurun_out/prof_blockmm.ncu-rep
