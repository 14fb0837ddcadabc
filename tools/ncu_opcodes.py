#!/usr/bin/env python
"""Opcode histogram (executed warp instructions + stall samples) per kernel from an `ncu --page source --csv` dump."""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
secs = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'hdr': None, 'data': []}; secs.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr'] = r; continue
    if len(r) == len(cur['hdr']): cur['data'].append(r)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 22
for sec in secs:
    hdr, data = sec['hdr'], sec['data']
    ix = {h:i for i,h in enumerate(hdr)}
    S = ix['# Samples']; SRC = ix['Source']; IE = ix['Instructions Executed']
    tot = sum(int(r[S] or 0) for r in data); totx = sum(int(r[IE] or 0) for r in data)
    print("==", sec['name'][:110]); print("static instrs", len(data), "executed", totx, "samples", tot)
    byop = collections.Counter(); cnt = collections.Counter(); ex = collections.Counter()
    for r in data:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[SRC])
        op = m.group(2) if m else '?'
        op = '.'.join(op.split('.')[:3]) if op.startswith(('LDS','STS','LDG','STG','LD.','ST.')) else op.split('.')[0]
        byop[op] += int(r[S] or 0); cnt[op]+=1; ex[op]+=int(r[IE] or 0)
    for op,c in ex.most_common(top): print("  %-16s executed %10d (%4.1f%%)  static %5d  samples %6d (%4.1f%%)" % (op,c,100*c/totx,cnt[op],byop[op],100*byop[op]/max(tot,1)))
