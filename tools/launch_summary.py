"""Aggregates an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list
per kernel: launches, average duration, share of the captured window and -- when the DRAM counters were collected --
average DRAM bytes per launch and the resulting GB/s."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [r for r in rows if r and r[0] == "ID"][0]
data = [r for r in rows if r and r[0].isdigit()]
ik, iv, ig, im, iu = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size", "Metric Name", "Metric Unit"))
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.OrderedDict()
for r in data:
    e = agg.setdefault(r[ik][:70], {"grid": r[ig], "t": [], "rd": [], "wr": []})
    v = float(r[iv].replace(",", "")) * SCALE.get(r[iu], 1.0)
    key = {"gpu__time_duration.sum": "t", "dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr"}.get(r[im])
    if key:
        e[key].append(v)
tot = sum(sum(e["t"]) for e in agg.values())
for k, e in agg.items():
    n, s = len(e["t"]), sum(e["t"])
    line = f"{k:72s} n={n:4d} avg={s / n:10.1f} us share={100 * s / tot:5.1f}% grid={e['grid']}"
    if e["rd"]:
        b = (sum(e["rd"]) + sum(e["wr"])) / n
        line += f" dram={b / 1e6:9.2f} MB/launch -> {b / (s / n) / 1e3:7.1f} GB/s"
    print(line)
