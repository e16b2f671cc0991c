#!/usr/bin/env python3
"""Summarise `ncu --page source --csv --print-source sass` output: executed instruction mix per
unit of work, stall samples per code region, hottest instructions."""
import collections
import csv
import re
import sys

path, units = sys.argv[1], float(sys.argv[2])  # units = e.g. number of warp tiles in the launch
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 250
rows = list(csv.reader(open(path)))
hdr, data = None, []
for r in rows:
    if r and r[0] == 'Address':
        if hdr is None:
            hdr = r
            continue
        break
    if hdr and len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
geti = lambda r, k: int(r[ix[k]] or 0)
c, tot = collections.Counter(), 0
for r in data:
    m = re.match(r'\s*(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)', r[ix['Source']])
    op = m.group(1) if m else r[ix['Source']].strip()[:12]
    ex = geti(r, 'Instructions Executed')
    c[op] += ex
    tot += ex
print('static instructions', len(data), ' executed warp-instr per unit', round(tot / units, 1))
print('  ' + '  '.join(f'{op} {n / units:.1f}' for op, n in c.most_common(30)))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tots = sum(geti(r, '# Samples') for r in data)
print('stall totals', {s: sum(geti(r, s) for r in data) for s in stalls if sum(geti(r, s) for r in data) > tots * 0.01})
for cc in range(0, len(data), chunk):
    ch = data[cc:cc + chunk]
    s = sum(geti(r, '# Samples') for r in ch)
    ex = sum(geti(r, 'Instructions Executed') for r in ch)
    d = {st: sum(geti(r, st) for r in ch) for st in stalls}
    top = sorted(d.items(), key=lambda kv: -kv[1])[:3]
    print(f'{cc:5d} {100 * s / max(tots, 1):5.1f}% samples {ex / units:7.0f} exec/unit', top)
for r in sorted(data, key=lambda r: -geti(r, '# Samples'))[:14]:
    print(geti(r, '# Samples'), r[ix['Source']].strip()[:80], {st: geti(r, st) for st in stalls if geti(r, st) > 20})
