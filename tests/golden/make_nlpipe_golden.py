"""tests/golden/nlpipe_ref.npz: aggregation == "NL" through the reference's OWN code -- ADCensusCal, StereoMatching::NL()
(stereoMatching.cpp:4892-4917) over its own NLCCA::aggreCV (NL/NLCCA.cpp) and qx_tree_filter (oracle/_ref/libqxref.so),
then sgm (4 paths), gen_dispFromVm and the refine() sequence -- all compiled into oracle/_ref/libsmref.so by
oracle/build_ref_sm.py.  Inputs are seeded (mystereomatching_b200.synth), so only the outputs are stored.
Run in the build container (needs /root/reference):  python tests/golden/make_nlpipe_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mystereomatching_b200 import synth          # noqa: E402
from oracle import pyoracle as po                # noqa: E402

assert po.smref_lib() is not None
out = {}
for tag, (H, W, D, kind, seed) in {"a": (40, 56, 16, "texture_warped", 11), "b": (33, 47, 24, "random_dot", 12)}.items():
    p = synth.make_pair(H, W, D, kind, seed)
    r = po.SmRef(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D)
    assert r.has_nl()
    wl, wr, rf, vol = r.pipeline_nl(4, True)
    r.close()
    out[tag + "_shape"] = np.array([H, W, D, seed], np.int32)
    out[tag + "_kind"] = np.array(kind)
    out[tag + "_wtaL"], out[tag + "_wtaR"], out[tag + "_refined"], out[tag + "_vol_after_nl"] = wl, wr, rf, vol
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "nlpipe_ref.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
