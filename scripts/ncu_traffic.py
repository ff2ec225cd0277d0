"""Per-kernel DRAM traffic and time from an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`
launch list: python scripts/ncu_traffic.py launches.csv  ->  one line per kernel family (launch count, mean ms, mean GB)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
hdr = rows[0]
iN, iM, iU, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
iID = hdr.index("ID")
SC = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}
per = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*$", "", r[iN]).replace("void ", "")
    fam = re.sub(r"<.*$", "", name)
    d = per.setdefault(fam, collections.defaultdict(float))
    v = float(r[iV].replace(",", "")) * SC.get(r[iU], 1)
    if r[iM] == "gpu__time_duration.sum":
        d["ms"] += v
        d["n"] += 1
    elif r[iM].startswith("dram__bytes"):
        d["bytes"] += v
print("| kernel family | launches | mean ms | mean DRAM GB (read + write) | GB/s |")
print("|---|---:|---:|---:|---:|")
for fam, d in per.items():
    if d["n"]:
        print(f"| `{fam}` | {int(d['n'])} | {d['ms'] / d['n']:.4f} | {d['bytes'] / d['n'] / 1e9:.4f} | {d['bytes'] / max(d['ms'], 1e-9) / 1e6:.0f} |")
