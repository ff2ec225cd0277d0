"""Generates tests/golden/top_ref.npz by RUNNING THE REFERENCE'S OWN selectTopCostFromVolumn (stereoMatching.h:2405-2461,
cut from /root/reference by oracle/build_ref_sm.py and compiled into oracle/_ref/libsmref.so) on two seeded volumes:
an AD-Census + CBCA volume of the reference itself (float costs) and an integer-valued volume with many ties (the
lowest-d rule decides).  Needs /root/reference; the committed .npz travels.
Run:  python tests/golden/make_top_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import pyoracle as po  # noqa: E402
from mystereomatching_b200 import synth  # noqa: E402

assert po.smref_lib() is not None, "run python oracle/build_ref_sm.py first (needs /root/reference)"
H, W, D = 20, 28, 16
p = synth.make_pair(H, W, D, "texture_warped", seed=21)
r = po.SmRef(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D)
r.set("censusFunc", 3)
r.adcensus()
r.arms()
vol, _ = r.cbca(2)
out = {"float_vol": vol}
for num, thres in ((6, 1.08), (3, 1.5), (1, 1.08)):
    out[f"float_top_n{num}_t{thres}"] = r.select_top(0, num, thres)
ties = np.random.default_rng(5).integers(1, 5, (H, W, D)).astype(np.float32)   # lamc_-style integer costs, many equal minima
ties[0, 0] = 0.0                                                               # first cost 0: nothing passes cost < 0 * thres
r.set_vm(0, ties)
out["ties_vol"] = ties
for num, thres in ((6, 1.08), (4, 2.5)):
    out[f"ties_top_n{num}_t{thres}"] = r.select_top(0, num, thres)
r.close()
path = os.path.join(os.path.dirname(__file__), "top_ref.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")
