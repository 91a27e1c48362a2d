#!/usr/bin/env python3
"""Summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean, total, share.
usage: python tools/ncu_launches.py launches.csv [skip_first_n]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1], newline='')) if len(r) > 5]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = rows[0]
kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1 + skip:]:
    v = float(r[mv].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'nsecond': 1e-3, 'ms': 1e3, 'msecond': 1e3}.get(r[mu], 1.0)
    name = r[kn].split('(')[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} n={c:5d} mean={t/c:10.2f} us total={t/1e3:10.3f} ms share={100*t/tot:5.1f}%")
