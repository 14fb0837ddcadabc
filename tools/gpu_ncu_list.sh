#!/bin/bash
# ncu launch list of a short bench run; args: env assignments
mkdir -p gpurun_out
env "$@" timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:k_ -s 40 -c 80 --csv --log-file gpurun_out/list.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/list.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.DictReader(l for l in open('gpurun_out/list.csv') if l.startswith('"'))]
by = collections.OrderedDict()
for r in rows:
    k = (r['ID'], r['Kernel Name'][:60], r['Grid Size'], r['Block Size'])
    by.setdefault(k, {})[r['Metric Name']] = r['Metric Value']
for (i, name, g, b), m in list(by.items())[:44]:
    print(i, name[10:58], g, "t=%sus rd=%s wr=%s issue=%s warps=%s regs=%s" % (float(m.get('gpu__time_duration.sum','0').replace(',',''))/1000, m.get('dram__bytes_read.sum'), m.get('dram__bytes_write.sum'), m.get('smsp__issue_active.avg.pct_of_peak_sustained_active'), m.get('sm__warps_active.avg.pct_of_peak_sustained_active'), m.get('launch__registers_per_thread')))
PY
