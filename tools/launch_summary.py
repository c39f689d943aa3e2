"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv).
usage: python tools/launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v/1e3 if r[ui] == "ns" else v*1e3 if r[ui] == "ms" else v
    a = agg[r[ki][:48]]
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in agg.values())
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
    print(f"{k:50s} {c:5d} {v:12.1f} us {v/c:10.1f} us/launch {100*v/tot:5.1f}%")
