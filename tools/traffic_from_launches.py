"""DRAM bytes per sweep kernel group from an ncu launch list with dram__bytes_read.sum / dram__bytes_write.sum
(and gpu__time_duration.sum): a group starts at every k_pack_labels launch and ends before k_apply_reduce.
Writes the JSON bench.py reads for `roofline.traffic` (profiles/r02_traffic.json).

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      --cache-control none -k regex:"k_sweep|k_apply|k_pack|k_lm|k_big" -c 200 --csv --log-file launches.csv \
      python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-scale24
  python tools/traffic_from_launches.py launches.csv profiles/r02_traffic.json
"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ii, ki, vi, mi, ui = H.index("ID"), H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name"), H.index("Metric Unit")
launch = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    d = launch.setdefault(int(r[ii]), {"name": re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("cvb::", "")})
    v = float(r[vi].replace(",", ""))
    u = r[ui].lower()
    if "byte" in u:
        v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    d[r[mi]] = v
groups, cur = [], None
for k in sorted(launch):
    L = launch[k]
    if L["name"].startswith("k_pack_labels"):
        cur = {"bytes": 0.0, "kernels": 0, "first": False}
        groups.append(cur)
    if cur is None or L["name"].startswith("k_apply") or L["name"].startswith("k_reduce"):
        if cur is not None and (L["name"].startswith("k_apply") or L["name"].startswith("k_reduce")):
            cur = None
        continue
    cur["bytes"] += L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0)
    cur["kernels"] += 1
    cur["first"] |= "first" in L["name"]
groups = [g for g in groups if g["kernels"] > 2]
first = [g["bytes"] for g in groups if g["first"]]
later = [g["bytes"] for g in groups if not g["first"]]
mean = lambda x: sum(x) / len(x) if x else None
# a phase of the bench graph is 1 table-free sweep + 3 hash sweeps
out = {"source": sys.argv[1], "sweep_groups": len(groups), "first_sweep_dram_bytes": mean(first), "later_sweep_dram_bytes": mean(later),
       "mean_sweep_dram_bytes": (mean(first) + 3 * mean(later)) / 4 if first and later else mean(first + later)}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
