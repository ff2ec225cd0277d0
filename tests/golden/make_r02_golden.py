"""Golden vectors added in round 2, all from the reference's OWN functions (oracle/_ref/libsmref.so, built by
oracle/build_ref_sm.py from /root/reference).  Run in the build container:  python tests/golden/make_r02_golden.py
 -> tests/golden/r02_ref.npz

  lor1_*   LRConsistencyCheck(D1, D2, errMask, LOR = 1)  (stereoMatching.cpp:2336-2364) on seeded maps
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mystereomatching_b200 import synth          # noqa: E402
from oracle import pyoracle as po                # noqa: E402


def lor1_cases(out):
    for tag, (H, W, D, seed) in {"a": (24, 40, 12, 5), "b": (16, 22, 32, 6), "c": (30, 52, 16, 7)}.items():
        p = synth.make_pair(H, W, D, "texture_warped", seed)
        r = po.SmRef(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D)
        rng = np.random.default_rng(seed)
        gt = np.clip(np.round(p["gt"]), 0, D - 1).astype(np.int16)
        d1 = gt.copy()
        d1[rng.random((H, W)) < 0.15] = rng.integers(0, D)            # mismatches
        d2 = np.zeros_like(d1)                                       # right map: warp of the left one + noise
        for v in range(H):
            for u in range(W):
                x = u - int(gt[v, u])
                if x >= 0:
                    d2[v, x] = gt[v, u]
        d2[rng.random((H, W)) < 0.2] = rng.integers(-1, D, (H, W))[rng.random((H, W)) < 0.2][:1]
        m = np.full((H, W), 7, np.uint8)
        a = np.ascontiguousarray(d1).copy()
        b = np.ascontiguousarray(d2).copy()
        r.L.smref_lrc_label(r.h, a, b, 1, m.ctypes.data)
        r.close()
        out[f"lor1_{tag}_D"] = np.int32(D)
        out[f"lor1_{tag}_d1"], out[f"lor1_{tag}_d2"] = d1, d2
        out[f"lor1_{tag}_d1_after"], out[f"lor1_{tag}_d2_after"], out[f"lor1_{tag}_errmask"] = a, b, m


def main():
    out = {}
    lor1_cases(out)
    path = os.path.join(ROOT, "tests", "golden", "r02_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
