"""Markdown table of the metrics the roofline discussion uses, from `ncu -i X.ncu-rep --page raw --csv > X.csv`.
   python scripts/ncu_summary.py X.csv [name-filter-regex]
One column per captured launch (kernel name shortened to its template arguments)."""
import csv
import re
import sys

ROWS = [
    ("duration [ms]", "gpu__time_duration.sum", 1e-6),
    ("DRAM read [GB]", "dram__bytes_read.sum", 1e-9),
    ("DRAM write [GB]", "dram__bytes_write.sum", 1e-9),
    ("DRAM throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("L1/TEX hit rate %", "l1tex__t_sector_hit_rate.pct", 1),
    ("L2 hit rate %", "lts__t_sector_hit_rate.pct", 1),
    ("shared-memory wavefronts [M]", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1e-6),
    ("shared bank conflicts [M]", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1e-6),
    ("L1TEX LSU data-pipe wavefronts % of peak", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1),
    ("issue slots used %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed", 1),
    ("warp instructions [M]", "smsp__inst_executed.sum", 1e-6),
    ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    ("registers / thread", "launch__registers_per_thread", 1),
    ("dynamic shared memory / block [KB]", "launch__shared_mem_per_block_dynamic", 1 / 1024),
    ("grid", "launch__grid_size", 1),
    ("block", "launch__block_size", 1),
]
UNIT_SCALE = {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9,
              "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
              "byte/block": 1, "Kbyte/block": 1024, "Mbyte/block": 1024 * 1024}


def short(name):
    m = re.match(r"(?:void )?([\w:]+)(<.*>)?", name)
    return (m.group(1) + (m.group(2) or "")).replace(" ", "")[:48] if m else name[:48]


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    flt = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[hi], rows[hi + 1]
    col = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[hi + 2:] if len(r) == len(hdr) and (flt is None or flt.search(r[col["Kernel Name"]]))]
    print("| metric | " + " | ".join(short(r[col["Kernel Name"]]) for r in data) + " |")
    print("|---|" + "---:|" * len(data))
    for label, key, scale in ROWS:
        if key not in col:
            continue
        i = col[key]
        u = UNIT_SCALE.get(units[i], 1)
        cells = []
        for r in data:
            try:
                v = float(r[i].replace(",", "")) * u * scale
                cells.append(f"{v:.3f}" if abs(v) < 1000 else f"{v:.0f}")
            except ValueError:
                cells.append(r[i])
        print(f"| {label} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
