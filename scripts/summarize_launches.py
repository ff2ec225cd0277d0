"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python scripts/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
H, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
agg = collections.OrderedDict()
for r in data:
    n = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) * scale[r[ui]]
tot = sum(a[1] for a in agg.values())
print(f"# ncu launch list summary: {sys.argv[1]} ({len(data)} launches, {tot:.1f} ms total; cold-cache, serialised)\n")
print("| kernel | launches | total ms | ms / launch | share |")
print("|---|---:|---:|---:|---:|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c} | {t:.3f} | {t / c:.4f} | {100 * t / tot:.1f}% |")
