#!/usr/bin/env python3
"""Per-function (RET/EXIT-delimited SASS segment) share of executed warp instructions and stall samples.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > x.csv; python tools/ncu_funcs.py x.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], newline='')))
h = rows[1]
si, ii, src = h.index('# Samples'), h.index('Instructions Executed'), h.index('Source')
seg, cur, ti, ts = [], [0, 0, 0], 0, 0
for n, r in enumerate(rows[2:]):
    if len(r) <= ii or not r[ii].isdigit(): continue
    i, s = int(r[ii]), int(r[si] or 0)
    ti += i; ts += s; cur[0] += i; cur[1] += s; cur[2] += 1
    if r[src].strip().startswith('RET') or 'EXIT' in r[src]:
        seg.append(tuple(cur) + (n,)); cur = [0, 0, 0]
seg.append(tuple(cur) + (len(rows),))
print('warp instructions', ti, 'samples', ts)
for g in seg:
    if g[0] > ti * 0.003: print(f"sass rows ..{g[3]:6d} ({g[2]:5d} instrs): {g[0]*100/ti:5.1f}% ins  {g[1]*100/ts:5.1f}% samples")
