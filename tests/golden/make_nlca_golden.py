"""Generates tests/golden/nlca_ref.npz by RUNNING THE REFERENCE'S OWN qx_nonlocal_cost_aggregation class
(NL/qx_nonlocal_cost_aggregation.cpp + NL/qx_basic.cpp:577-624 + qx_tree_filter + ctmf, compiled by
oracle/build_ref.sh into oracle/_ref/libqxref.so) on a seeded synthetic pair.  Needs /root/reference.
Run:  python tests/golden/make_nlca_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from mystereomatching_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

assert po.ref_lib() is not None, "run oracle/build_ref.sh first (needs /root/reference)"
D = 12
p = synth.make_pair(64, 80, D, "texture_warped", seed=77)   # ctmf needs h*w >= 5*544 for r = 2
L, R = p["bgrL"], p["bgrR"]
a = po.ref_nlca(L, R, D, 0.1, post=False)
b = po.ref_nlca(L, R, D, 0.1, post=True)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "nlca_ref.npz"), left=L, right=R, D=D,
                    grad_left=a["grad_left"], cost=a["cost"].astype(np.float64), cost_right=a["cost_right"],
                    disp=a["disp"], disp_post=b["disp"])
print("wrote nlca_ref.npz")
