"""Generates tests/golden/wm_ref.npz by RUNNING THE REFERENCE'S OWN WM (stereoMatching.cpp:7340-7393, the default-off
bilateral weighted median of refine(); cut from /root/reference by oracle/build_ref_sm.py and compiled into
oracle/_ref/libsmref.so) on seeded maps whose labels all lie in [0, D) -- the only inputs on which the reference is
defined (it indexes its histogram with the labels unchecked, :7371).  Needs /root/reference; the committed .npz travels.
Run:  python tests/golden/make_wm_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import pyoracle as po  # noqa: E402
from mystereomatching_b200 import synth  # noqa: E402

assert po.smref_lib() is not None, "run python oracle/build_ref_sm.py first (needs /root/reference)"
out = {}
for tag, (H, W, D, seed) in {"a": (26, 34, 16, 41), "b": (12, 10, 5, 42), "c": (20, 40, 64, 43)}.items():
    p = synth.make_pair(H, W, D, "texture_warped", seed=seed)
    r = po.SmRef(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D)
    rng = np.random.default_rng(seed)
    noisy = rng.integers(0, D, (H, W)).astype(np.int16)
    planes = np.clip((np.arange(W)[None, :] * (D - 1) // W + (np.arange(H)[:, None] > H // 2) * 2), 0, D - 1).astype(np.int16)
    planes[rng.random((H, W)) < 0.1] = 0                           # outliers the filter should pull back
    m_some = (rng.random((H, W)) < 0.3).astype(np.uint8) * 255
    m_all = np.full((H, W), 1, np.uint8)                            # any value > 0 counts
    out[f"{tag}_bgr"], out[f"{tag}_D"] = p["bgrL"], np.array(D)
    for name, disp, mask in (("noisy_some", noisy, m_some), ("planes_some", planes, m_some), ("planes_all", planes, m_all)):
        out[f"{tag}_{name}_in"], out[f"{tag}_{name}_mask"] = disp, mask
        out[f"{tag}_{name}_out"] = r.wm(disp, mask)
    r.close()
path = os.path.join(os.path.dirname(__file__), "wm_ref.npz")
np.savez_compressed(path, **out)
print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")
