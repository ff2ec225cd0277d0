"""SASS opcode histogram per kernel of libsm_b200.so (cuobjdump -sass): which Blackwell / async-copy / packed-SIMD
mnemonics each hot kernel contains.   python scripts/sass_histogram.py > profiles/<round>_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mystereomatching_b200", "libsm_b200.so")
KEY = ["UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "LDGDEPBAR", "CREDUX", "REDUX", "FMNMX3", "FMNMX", "VIMNMX", "VIADDMNMX", "VABSDIFF4", "POPC", "SHFL",
       "LDS", "STS", "LDG", "STG", "ATOMG", "ATOMS", "BAR", "ACQBULK", "FADD", "FMUL", "MUFU", "DADD", "DMUL"]
txt = subprocess.check_output(["cuobjdump", "-sass", LIB], text=True)
arch = re.findall(r"arch = (sm_\w+)", txt)
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.check_output(["c++filt", m.group(1)], text=True).strip()
        cur = re.sub(r"\(.*\)$", "", cur).replace("void ", "")
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
    if m and cur:
        op = m.group(1)
        per[cur][op] += 1
        per[cur]["_total"] += 1
        if op == "VIADDMNMX" or op == "VIMNMX":
            per[cur][op + (m.group(2) or "")] += 1
print(f"# SASS opcode histogram of libsm_b200.so ({', '.join(sorted(set(arch)))} cubins only; `cuobjdump -sass`, static instruction counts)\n")
tot = collections.Counter()
for c in per.values():
    tot.update(c)
print("Whole library: " + ", ".join(f"{k} {tot[k]}" for k in KEY if tot[k]) + f"; {tot['_total']} instructions in {len(per)} kernels.  "
      "No UTC*MMA / LDTM / STTM (tcgen05 / TMEM) and no HMMA: the path has no contraction (DESIGN.md section 4).\n")
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else r"k_cost<|k_cbca_pass|k_sgm_group|k_sgm_path|k_sgm_path_u16|k_tf_cta_fast|k_census|k_arms|k_hamming")
print("| kernel | total | " + " | ".join(KEY) + " |")
print("|---|---:|" + "---:|" * len(KEY))
for name, c in per.items():
    if pat.search(name):
        print(f"| `{name[:70]}` | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEY) + " |")
packed = {k: v for k, v in tot.items() if k.startswith(("VIADDMNMX.", "VIMNMX."))}
print("\nPacked-SIMD forms seen: " + ", ".join(f"{k} {v}" for k, v in sorted(packed.items())))
