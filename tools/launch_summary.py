#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hdr_i]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hdr_i + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0][:70]
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else v * 1e3 if r[ui] == 'ms' else v * 1e6 if r[ui] == 's' else v
    agg.setdefault(name, [0, 0.0])
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:12.1f} us {n:5d} x {t / n:10.1f} us  {100 * t / tot:5.1f}%  {k}")
