#!/bin/bash
# perf sweep over tuning overrides; prints per-kernel times
run() { echo "== $1"; env $1 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e ${BENCH_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.1f Gpx/s  ms/step %.3f  rt_frac %.3f err %.1e' % (d['value'], d['ms_per_step'], d['roofline']['round_trip_frac'], d['roundtrip_rel_l2']))
for k in d['kernels']: print('  %s/%s n=%d grid=%d smem=%d: %.3f ms  %.0f GB/s' % (k['plan'], k['kernel'], k['n'], k['grid'], k['smem_bytes'], k['avg_ms'], k['achieved_gbs']))
"; }
for cfg in "$@"; do run "$cfg"; done
