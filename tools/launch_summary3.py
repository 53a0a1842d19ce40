"""Per-kernel means from an ncu CSV: duration, executed warp instructions, L1TEX utilisation."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, mi = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(cvb::.*", "", r[ki]).replace("void ", "")[:44]
    agg.setdefault(name, {}).setdefault(r[mi], []).append(float(r[vi].replace(",", "")))
tot = 0.0
for k, v in agg.items():
    m = lambda n: sum(v[n]) / len(v[n]) if n in v else float("nan")
    t = v["gpu__time_duration.sum"]
    tot += sum(t) / 1e3
    print(f"{k:44s} n={len(t):3d} {m('gpu__time_duration.sum')/1e3:8.1f}us inst {m('smsp__inst_executed.sum')/1e6:7.2f}M "
          f"l1tex {m('l1tex__throughput.avg.pct_of_peak_sustained_active'):5.1f}%")
print(f"sum over all launches {tot:.1f} us")
