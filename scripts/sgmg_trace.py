"""Read a SM_SGMG_TRACE dump: hand-off latency between neighbouring CTAs of k_sgm_group.  usage: sgmg_trace.py file H"""
import sys
import numpy as np
H = int(sys.argv[2])
a = np.fromfile(sys.argv[1], dtype=np.uint64).astype(np.int64)
nb = (a.size - 16 * 32 * 8) // (2 * H * 4)
t = a[:nb * 2 * H * 4].reshape(nb, 2, H, 4)
print("ctas", nb)
for b in (10, 70, 130):
    # last column of CTA b reads what the first column of CTA b+1 published one row earlier
    L, F = t[b, 1], t[b + 1, 0]
    rows = np.arange(400, 410)
    for r in rows:
        print(f"cta {b} row {r}: period {L[r,1]-L[r-1,1]:6d} ns  request->got {L[r,2]-L[r,1]:6d}  polls {L[r,3]:3d}  "
              f"neighbour published row {r-1} {L[r,1]-F[r-1,0]:6d} ns before request, {L[r,2]-F[r-1,0]:6d} before got; own publish at +{L[r,0]-L[r,1]:5d}")
lat = []
for b in range(nb - 1):
    L, F = t[b, 1], t[b + 1, 0]
    lat.append((L[2:, 2] - F[1:-1, 0]))
lat = np.concatenate(lat)
print("publish -> got (ns): median", np.median(lat), "p10", np.percentile(lat, 10), "p90", np.percentile(lat, 90))
per = np.diff(t[:, 1, :, 1], axis=1)
print("row period (ns): median", np.median(per), "polls mean", t[:, :, 1:, 3].mean())
