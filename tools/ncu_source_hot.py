#!/usr/bin/env python
"""Hot spots of one kernel from `ncu -i X.ncu-rep --page source --csv`: per-instruction stall samples, top N."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    try:
        data.append((int(r[isamp]), r[isrc], int(r[iex]), r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data), 'executed', sum(d[2] for d in data))
agg = {}
for d in data:
    for i in stall_cols:
        v = int(d[3][i]) if d[3][i] else 0
        agg[hdr[i]] = agg.get(hdr[i], 0) + v
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
for d in sorted(data, key=lambda x: -x[0])[:top]:
    st = sorted([(int(d[3][i]) if d[3][i] else 0, hdr[i]) for i in stall_cols], reverse=True)[:2]
    print(f"{d[0]:6d} {100*d[0]/tot:5.1f}% ex={d[2]:8d} {d[1][:72]:72s} {st}")
