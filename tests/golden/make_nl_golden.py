"""Generates tests/golden/nl_ref.npz by RUNNING THE REFERENCE'S OWN NL/ SOURCES
(ctmf.c, qx_mst_kruskals_image.cpp, qx_tree_filter.cpp compiled by oracle/build_ref.sh into
oracle/_ref/libqxref.so) on seeded inputs.  Needs /root/reference; the committed .npz travels.
Run:  python tests/golden/make_nl_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import pyoracle as po  # noqa: E402

assert po.ref_lib() is not None, "run oracle/build_ref.sh first (needs /root/reference)"
rng = np.random.default_rng(424242)
out = {}
# ctmf: r=1 cn=3 (MST guidance), r=2 cn=1 (Yang's disparity post-filter), odd sizes
img3 = rng.integers(0, 256, (21, 27, 3), dtype=np.uint8)
img1 = rng.integers(0, 64, (48, 67), dtype=np.uint8)  # ctmf needs h*w*cn/544 - 2r >= 1 histograms (NL/ctmf.c:410)
out["ctmf_in3"], out["ctmf_r1_cn3"] = img3, po.ref_ctmf(img3, 1)
out["ctmf_in1"], out["ctmf_r2_cn1"] = img1, po.ref_ctmf(img1, 2)
out["ctmf_r1_cn1"] = po.ref_ctmf(img1, 1)
# MST.  (The reference's own 3x3 smoke input, NL/qx_mst_kruskals_image.cpp:283, cannot be run: with
# memsize = h*w*cn = 9 its ctmf call computes zero stripes and divides by zero, NL/ctmf.c:410-411.)
# A single-channel ramp image, cn=1 ...
ramp = ((np.arange(40)[:, None] * 7 + np.arange(50)[None, :] * 3) % 251).astype(np.uint8)
out["mst_ramp_img"] = ramp
for k, v in po.ref_mst(ramp).items():
    out["mst_ramp_" + k] = v
# ... a smooth image with many equal weights (tie-breaking by edge index matters) ...
yy, xx = np.mgrid[0:24, 0:32]
smooth = np.stack([(xx * 3 + yy) // 4, (xx + yy * 2) // 5, (xx * yy) // 40], -1).astype(np.uint8)
out["mst_smooth_img"] = smooth
for k, v in po.ref_mst(smooth).items():
    out["mst_smooth_" + k] = v
# ... and a noisy colour image
noisy = rng.integers(0, 256, (24, 30, 3), dtype=np.uint8)
out["mst_noisy_img"] = noisy
for k, v in po.ref_mst(noisy).items():
    out["mst_noisy_" + k] = v
# tree filter on a float32-representable volume, sigma 0.1 (NLCCA.cpp:33) and 0.05 (update_table(sigma/2))
vol = rng.random((24, 30, 6)).astype(np.float32)
out["tf_vol"] = vol
out["tf_out_sigma0p1"] = po.ref_tree_filter(noisy, vol.astype(np.float64), 0.1)
out["tf_out_sigma0p05"] = po.ref_tree_filter(noisy, vol.astype(np.float64), 0.05)
out["tf_smooth_out"] = po.ref_tree_filter(np.ascontiguousarray(smooth[:24, :30]), vol.astype(np.float64), 0.1)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "nl_ref.npz"), **out)
print("wrote nl_ref.npz:", sorted(out))
