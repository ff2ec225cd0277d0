"""Generates tests/golden/subpixel_ref.npz by RUNNING THE REFERENCE'S OWN subpixelEnhancement (stereoMatching.cpp:6138-6166,
cut from /root/reference by oracle/build_ref_sm.py and compiled into oracle/_ref/libsmref.so) on the reference's own
AD-Census + CBCA + 8-path-free WTA map, on a random map over the same volume (every branch: disp 0, D-1, negative labels)
and on an integer-valued volume with flat parabolas (denom == 0).  Needs /root/reference; the committed .npz travels.
Run:  python tests/golden/make_subpixel_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import pyoracle as po  # noqa: E402
from mystereomatching_b200 import synth  # noqa: E402

assert po.smref_lib() is not None, "run python oracle/build_ref_sm.py first (needs /root/reference)"
H, W, D = 20, 28, 16
p = synth.make_pair(H, W, D, "texture_warped", seed=33)
r = po.SmRef(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D)
r.set("censusFunc", 3)
r.adcensus()
r.arms()
vol, _ = r.cbca(2)
rng = np.random.default_rng(9)
out = {"float_vol": vol}
out["float_disp_wta"] = r.wta(0)
rnd = rng.integers(0, D, (H, W)).astype(np.int16)
rnd[::5, ::3] = -32        # DISP_OCC
rnd[1::5, 1::3] = -48      # DISP_MIS
rnd[0, :4] = [0, D - 1, 1, D - 2]
out["float_disp_rnd"] = rnd
for k in ("wta", "rnd"):
    out[f"float_se_{k}"] = r.subpixel(out[f"float_disp_{k}"])
flat = rng.integers(1, 4, (H, W, D)).astype(np.float32)   # many zero denominators and |diff| >= 1 cases
r.set_vm(0, flat)
out["flat_vol"] = flat
out["flat_disp_rnd"] = rnd
out["flat_se_rnd"] = r.subpixel(rnd)
r.close()
# refine() then runs cv::medianBlur(SE, SE, 3) on the float map (stereoMatching.cpp:1490): cv2 is that function
import cv2  # noqa: E402
for k in ("float_se_wta", "float_se_rnd", "flat_se_rnd"):
    out[k.replace("_se_", "_semed_")] = cv2.medianBlur(out[k], 3)
frac = (rng.random((H, W)) * 40 - 8).astype(np.float32)
frac[3:6, 4:9] = 7.0                                  # equal values
out["frac_map"] = frac
out["frac_med"] = cv2.medianBlur(frac, 3)
for shp in ((1, 1), (1, 9), (7, 1), (2, 2)):          # degenerate shapes: the replicated border does all the work
    m = (rng.random(shp) * 10).astype(np.float32)
    out["tiny%dx%d_map" % shp] = m
    out["tiny%dx%d_med" % shp] = cv2.medianBlur(m, 3)
out["cv2_version"] = np.array(cv2.__version__)
path = os.path.join(os.path.dirname(__file__), "subpixel_ref.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")
