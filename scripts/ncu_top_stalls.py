"""Print per-kernel top stall SASS lines (with context) from `ncu -i X.ncu-rep --page source --csv --print-source sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 12
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 4
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1][:60], "rows": [], "hdr": None}; kern.append(cur); continue
    if cur is None: continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur["hdr"] and len(r) == len(cur["hdr"]): cur["rows"].append(r)
seen = set()
for k in kern:
    if k["name"] in seen: continue
    seen.add(k["name"])
    h = k["hdr"]; si = h.index("Warp Stall Sampling (All Samples)"); so = h.index("Source"); ii = h.index("Instructions Executed")
    R = k["rows"]; tot = sum(int(r[si]) for r in R)
    print("==", k["name"], "samples", tot, "warp-instr", sum(int(r[ii]) for r in R))
    cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    agg = sorted(((sum(int(r[i] or 0) for r in R), h[i]) for i in cols), reverse=True)[:7]
    print("   reasons:", [(n, round(100 * v / tot, 1)) for v, n in agg])
    order = sorted(range(len(R)), key=lambda i: -int(R[i][si]))[:ntop]
    for idx in order:
        print(f"  --- {100 * int(R[idx][si]) / tot:.1f}% of samples")
        for i in range(max(0, idx - ctx), min(len(R), idx + 2)):
            mark = ">>" if i == idx else "  "
            print(f"   {mark} {int(R[i][si]):6d} x{int(R[i][ii]):8d} {R[i][so].strip()[:110]}")
