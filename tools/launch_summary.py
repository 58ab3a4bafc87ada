"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the captured window)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [r for r in rows if r and r[0] == "ID"][0]
data = [r for r in rows if r and r[0].isdigit()]
ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = collections.OrderedDict()
for r in data:
    agg.setdefault(r[ik][:86], []).append((float(r[iv].replace(",", "")), r[ig]))
tot = sum(sum(v for v, _ in x) for x in agg.values())
for k, v in agg.items():
    s = sum(x for x, _ in v)
    print(f"{k:88s} n={len(v):3d} avg={s / len(v) / 1e3:9.1f} us share={100 * s / tot:5.1f}% grid={v[0][1]}")
