#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel over the LAST update of the run
(from the last-but-N gather_kernel launch on). python tools/summarize_launches.py launches.csv [n_gathers_back]"""
import collections, csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
back = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = [(r[ki], float(r[vi].replace(",", "")) / (1000 if r[ui] == "ns" else 1)) for r in rows[1:]]
idx = [i for i, (k, _) in enumerate(launches) if "gather" in k]
last = launches[idx[-back]:] if idx else launches
agg = collections.OrderedDict()
for k, t in last:
    k = k.split("(")[0][:64]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{len(launches)} launches in the file; last window: {len(last)} launches, {tot:.1f} us")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:64s} {c:3d} x {t / c:8.1f} us = {t:9.1f} us {100 * t / tot:5.1f}%")
