"""Per-warp row timeline of CTA 70, rows 400..431, from a SM_SGMG_TRACE dump.  usage: sgmg_timeline.py file H nb"""
import sys
import numpy as np
H, nb = int(sys.argv[2]), int(sys.argv[3])
a = np.fromfile(sys.argv[1], dtype=np.uint64).astype(np.int64)
t = a[nb * 2 * H * 4:].reshape(16, 32, 8)
t0 = t[:, :, 0][t[:, :, 0] > 0].min()
for r in range(4, 10):
    print(f"row {400 + r}")
    for w in range(16):
        if t[w, r, 0] == 0: continue
        x = t[w, r, :5] - t0
        extra = ""
        if t[w, r, 5]:   # edge column: near diagonal + publish, wait for the far row, the other two recurrences
            y = t[w, r, 5:8] - t0
            extra = f"   [edge: near+publish {y[1]-y[0]:5d}  far-row wait {y[2]-y[1]:5d}  A+F {x[2]-y[2]:5d}]"
        print(f"  col {w:2d}: start {x[0]:7d}  +tma/ld {x[1]-x[0]:5d}  +compute {x[2]-x[1]:5d}  +sum {x[3]-x[2]:5d}  +barriers {x[4]-x[3]:5d}   = {x[4]-x[0]:5d}{extra}")
