"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H, data = rows[hdr], rows[hdr + 1:]
ki, vi, gi, bi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Grid Size"), H.index("Block Size")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(cvb::.*", "", r[ki]).replace("void ", "") + " g=" + r[gi].split(",")[0].strip("(") + " b=" + r[bi].split(",")[0].strip("(")
    agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
tot = 0.0
for k, v in agg.items():
    m = sum(v) / len(v) / 1e3
    tot += m
    print(f"{k[:84]:84s} n={len(v):3d} mean={m:9.1f} us first={v[0] / 1e3:8.1f} min={min(v) / 1e3:8.1f}")
print(f"sum of per-kernel means: {tot:.1f} us")
