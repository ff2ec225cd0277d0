"""Generates tests/golden/da_ref.npz for discontinuityAdjust (stereoMatching.cpp:6057-6135).
Two sources, both run HERE:
  * cv2 4.13 itself for the three OpenCV calls the function makes (equalizeHist, GaussianBlur(3x3, 4, 4), Canny(20, 60, 3))
    and for the whole edge chain on disparity maps -- pins oracle/opencv_restated.h;
  * the reference's OWN discontinuityAdjust body (cut from /root/reference by oracle/build_ref_sm.py into
    oracle/_ref/libsmref.so, its three OpenCV calls routed to those pinned restatements) on the reference's own
    AD-Census + CBCA + SGM volume with its WTA map, and on synthetic piecewise maps over random volumes.
Needs /root/reference and cv2; the committed .npz travels.   Run:  python tests/golden/make_da_golden.py
"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import pyoracle as po  # noqa: E402
from mystereomatching_b200 import synth  # noqa: E402

assert po.smref_lib() is not None, "run python oracle/build_ref_sm.py first (needs /root/reference)"
rng = np.random.default_rng(61)
out = {"cv2_version": np.array(cv2.__version__)}


def piecewise(H, W, D, n=6):
    """A disparity-like map: rectangles of constant labels, a ramp, a few invalid markers."""
    m = np.full((H, W), D // 3, np.int16)
    for _ in range(n):
        y0, x0 = rng.integers(0, H - 2), rng.integers(0, W - 2)
        y1, x1 = rng.integers(y0 + 1, H + 1), rng.integers(x0 + 1, W + 1)
        m[y0:y1, x0:x1] = rng.integers(0, D)
    m[:, W // 2:] += (np.arange(W - W // 2) // 5).astype(np.int16)[None, :]
    m = np.clip(m, 0, D - 1).astype(np.int16)
    m[rng.random((H, W)) < 0.02] = -32
    m[rng.random((H, W)) < 0.01] = -48
    return m


def edges_cv2(disp):
    e = np.clip(disp, 0, 255).astype(np.uint8)             # convertTo(CV_8U): saturate_cast
    e = cv2.equalizeHist(e)
    e = cv2.GaussianBlur(e, (3, 3), 4, None, 4)
    return cv2.Canny(e, 20, 60, None, 3)


# ---- the OpenCV pieces
imgs = {
    "noise": rng.integers(0, 256, (23, 31), dtype=np.uint8),
    "few": (rng.integers(0, 256, (17, 40), dtype=np.uint8) // 50 * 7).astype(np.uint8),
    "const": np.full((5, 9), 77, np.uint8),
    "steps": np.repeat(np.repeat(rng.integers(0, 256, (6, 8), dtype=np.uint8), 5, 0), 6, 1),
    "t1x1": np.array([[9]], np.uint8), "t1x7": rng.integers(0, 256, (1, 7), dtype=np.uint8),
    "t6x1": rng.integers(0, 256, (6, 1), dtype=np.uint8), "t2x2": rng.integers(0, 256, (2, 2), dtype=np.uint8),
    "t3x3": rng.integers(0, 256, (3, 3), dtype=np.uint8),
}
for k, im in imgs.items():
    out[f"img_{k}"] = im
    out[f"eq_{k}"] = cv2.equalizeHist(im)
    out[f"gb_{k}"] = cv2.GaussianBlur(im, (3, 3), 4, None, 4)
    out[f"cn_{k}"] = cv2.Canny(im, 20, 60, None, 3)
    out[f"cnb_{k}"] = cv2.Canny(out[f"gb_{k}"], 20, 60, None, 3)

# ---- whole function on the reference's own volume
H, W, D = 40, 56, 24
p = synth.make_pair(H, W, D, "texture_warped", seed=41)
r = po.SmRef(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D)
r.set("censusFunc", 3)
r.adcensus()
r.arms()
r.cbca(2)
r.sgm(0, 4)
vol = r.vm(0)
wta = r.wta(0)
out["a_vol"], out["a_disp"] = vol, wta
out["a_edge_cv2"] = edges_cv2(wta)
out["a_out"] = r.disc_adjust(wta)
# ---- piecewise maps over random volumes (long contours -> long raster-order chains in every direction)
for tag, (h, w, d) in (("b", (40, 56, 24)),):
    for i in range(3):
        m = piecewise(h, w, d)
        v = (rng.random((h, w, d)) * 3).astype(np.float32)
        r.set_vm(0, v)
        out[f"{tag}{i}_vol"], out[f"{tag}{i}_disp"] = v, m
        out[f"{tag}{i}_edge_cv2"] = edges_cv2(m)
        out[f"{tag}{i}_out"] = r.disc_adjust(m)
r.close()
for k in list(out):
    if k.endswith("_out"):
        pre = k[:-4]
        print(pre, "edge px", int((out[pre + "_edge_cv2"] > 0).sum()), "changed", int((out[k] != out[pre + "_disp"]).sum()))
path = os.path.join(os.path.dirname(__file__), "da_ref.npz")
np.savez_compressed(path, **out)
print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")
