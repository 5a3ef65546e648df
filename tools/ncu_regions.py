#!/usr/bin/env python
"""Group the stall samples of an `ncu --page source --csv` export into code regions (runs of instructions with the same
execution count) -- how the phase breakdowns in profiles/*.md were made.  usage: ncu_regions.py source.csv [lo hi]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cur, groups, tot = None, [], 0
for k, r in enumerate(rows[2:]):
    ie = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']]); tot += s
    st = {h: int(r[ix[h]]) for h in hdr if h.startswith('stall_') and 'Not Issued' not in h}
    if cur is None or abs(ie - cur['ie']) > 0.15 * max(ie, cur['ie'], 1):
        cur = {'ie': ie, 'k0': k, 'k1': k, 's': 0, 'st': {}}; groups.append(cur)
    cur['k1'] = k; cur['s'] += s
    for h, v in st.items():
        cur['st'][h] = cur['st'].get(h, 0) + v
print(rows[0][1][:100])
print('total samples', tot)
for g in groups:
    if g['s'] >= 8:
        top = sorted(g['st'].items(), key=lambda x: -x[1])[:5]
        print('%4d-%4d executed %6d samples %4d (%4.1f%%)' % (g['k0'], g['k1'], g['ie'], g['s'], 100.0 * g['s'] / tot), [(h[6:], v) for h, v in top if v])
if len(sys.argv) >= 4:
    lo, hi = int(sys.argv[2]), int(sys.argv[3])
    for k, r in enumerate(rows[2:]):
        if lo <= k <= hi:
            s = int(r[ix['# Samples']])
            st = {h[6:]: int(r[ix[h]]) for h in hdr if h.startswith('stall_') and 'Not Issued' not in h and int(r[ix[h]]) > 0}
            print(k, r[ix['Source']].strip()[:64].ljust(64), r[ix['Instructions Executed']], s, sorted(st.items(), key=lambda x: -x[1])[:3] if s >= 3 else '')
