#!/bin/bash
timeout 120 python -m pytest tests/test_gpu_round2.py -x -q -k "ring_column_subpasses_small_panels" 2>&1 | grep -E "Error|error|assert|passed|failed" | head -12
DSP_DCT_RING_TRACE=1 timeout 120 python - <<'PY' 2>&1 | tail -60
import ctypes, re, sys, os, io
import numpy as np, torch
from dspfun_b200 import Plan, REDFT10, capi
lib = capi.load()
n = 8192
x = torch.rand((2, n, n), device="cuda")
p = Plan.interleaved_2d("f", n, n, 1, REDFT10, nbatch=2)
# capture stderr to find the trace buffer address
r, w = os.pipe(); old = os.dup(2); os.dup2(w, 2)
p.execute_dev(x.data_ptr(), x.data_ptr(), None); torch.cuda.synchronize()
os.dup2(old, 2); os.close(w)
txt = os.read(r, 65536).decode()
m = re.search(r"ring trace buffer (0x[0-9a-f]+)", txt)
print(txt.strip()[-200:])
addr = int(m.group(1), 16)
p.execute_dev(x.data_ptr(), x.data_ptr(), None); torch.cuda.synchronize()
buf = (ctypes.c_longlong * (4 * 4096))()
cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaMemcpy(buf, ctypes.c_void_p(addr), ctypes.sizeof(buf), 2)
t = np.array(buf[:], dtype=np.int64).reshape(4096, 4)
rows = t[t[:, 0] != 0]
t0 = rows[:, 0].min()
print("items of CTA 0:", len(rows))
print(" it  kind  wait_begin  ready(+wait)  compute_end(+)  store_done(+)   [us]")
for i, r in enumerate(rows[:36]):
    kind = "B" if r[2] < 0 else "A"
    c2 = abs(r[2])
    print("%3d   %s   %9.2f  %9.2f  %9.2f  %9.2f" % (i, kind, (r[0]-t0)/1e3, (r[1]-r[0])/1e3, (c2-r[1])/1e3, ((r[3]-c2)/1e3 if r[3] else 0)))
A = rows[rows[:, 2] > 0]; B = rows[rows[:, 2] < 0]
print("A: wait %.2f compute %.2f store %.2f us (n=%d)" % (((A[:,1]-A[:,0]).mean())/1e3, ((A[:,2]-A[:,1]).mean())/1e3, ((A[:,3]-A[:,2]).mean())/1e3, len(A)))
print("B: wait %.2f read+release %.2f us (n=%d)" % (((B[:,1]-B[:,0]).mean())/1e3, ((-B[:,2]-B[:,1]).mean())/1e3, len(B)))
print("span %.1f us" % ((rows[:, 1].max() - t0) / 1e3))
PY
