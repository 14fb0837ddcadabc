#!/usr/bin/env python
"""Split each kernel's SASS (ncu --page source --csv) at BAR.SYNC instructions and report, per region, the
executed warp instructions, stall samples and the dominant stall reasons.  Regions follow program order, which
for these kernels is phase order (copy-in | passes ... | copy-out)."""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
secs = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'hdr': None, 'data': []}; secs.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr'] = r; continue
    if len(r) == len(cur['hdr']): cur['data'].append(r)
seen = set()
for sec in secs:
    if sec['name'] in seen: continue
    seen.add(sec['name'])
    hdr, data = sec['hdr'], sec['data']
    ix = {h:i for i,h in enumerate(hdr)}
    S = ix['# Samples']; SRC = ix['Source']; IE = ix['Instructions Executed']
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    print("==", sec['name'][:120])
    reg = []; cur = {'n':0,'ex':0,'s':0,'st':collections.Counter(), 'ops': collections.Counter()}
    for r in data:
        cur['n'] += 1; cur['ex'] += int(r[IE] or 0); cur['s'] += int(r[S] or 0)
        for h in stalls: cur['st'][h] += int(r[ix[h]] or 0)
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[SRC]); op = m.group(2) if m else '?'
        cur['ops'][op.split('.')[0]] += int(r[IE] or 0)
        if 'BAR.SYNC' in r[SRC] or r[SRC].strip().startswith('EXIT'):
            reg.append(cur); cur = {'n':0,'ex':0,'s':0,'st':collections.Counter(), 'ops': collections.Counter()}
    if cur['n']: reg.append(cur)
    tots = sum(x['s'] for x in reg) or 1; totx = sum(x['ex'] for x in reg) or 1
    for i, x in enumerate(reg):
        if x['ex'] == 0 and x['s'] == 0: continue
        top = ", ".join("%s %.0f%%" % (k.replace('stall_',''), 100*v/max(x['s'],1)) for k,v in x['st'].most_common(4))
        ops = ", ".join("%s %.0f%%" % (k, 100*v/max(x['ex'],1)) for k,v in x['ops'].most_common(5))
        print("  region %2d: static %5d  executed %5.1f%%  samples %5.1f%%  | %s | %s" % (i, x['n'], 100*x['ex']/totx, 100*x['s']/tots, top, ops))
