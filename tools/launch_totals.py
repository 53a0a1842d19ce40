"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, total and mean
microseconds, sorted by total.  usage: python tools/launch_totals.py launches.csv [filter-substring ...]"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    want = sys.argv[2:]
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1000 if r[mu] == "ns" else v * 1000 if r[mu] == "ms" else v
        name = re.sub(r"\(.*", "", r[kn])[:80]
        if want and not any(w in name for w in want):
            continue
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:10.1f} us  n={n:4d}  mean {t / n:9.1f} us  {k}")
    print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")


if __name__ == "__main__":
    main()
