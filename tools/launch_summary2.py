"""Per-kernel mean duration and executed warp-instructions from an ncu CSV with both metrics."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H, data = rows[hdr], rows[hdr + 1:]
ki, vi, mi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(cvb::.*", "", r[ki]).replace("void ", "")[:62]
    agg.setdefault(name, {}).setdefault(r[mi], []).append(float(r[vi].replace(",", "")))
tt = 0
for k, v in agg.items():
    t, i = v["gpu__time_duration.sum"], v.get("smsp__inst_executed.sum", [0])
    tt += sum(t) / len(t) / 1e3
    print(f"{k:62s} {sum(t)/len(t)/1e3:8.1f} us  inst {sum(i)/len(i)/1e6:7.2f}M")
print(f"sum {tt:.1f} us")
