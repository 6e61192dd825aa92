"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: python profiles/summarize.py <launches.csv> > summary.md"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v
    a = agg.setdefault(r[ki].split("(")[0][:80], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"| kernel | launches | total us | share |\n|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.1f}% |")
print(f"| all | {sum(a[0] for a in agg.values())} | {tot:.1f} | 100% |")
