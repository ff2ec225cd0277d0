"""Golden vectors added in round 2, all from the reference's OWN functions (oracle/_ref/libsmref.so, built by
oracle/build_ref_sm.py from /root/reference).  Run in the build container:  python tests/golden/make_r02_golden.py
 -> tests/golden/r02_ref.npz

  lor1_*   LRConsistencyCheck(D1, D2, errMask, LOR = 1)  (stereoMatching.cpp:2336-2364) on seeded maps
  top_*    genDispFromTopCostVm (stereoMatching.h:2466-2545) and genDispFromTopCostVm2 (stereoMatching.cpp:1514-1886,
           methods 0 / 1 / 2, with / without vmTop_hasCir2 and vmTop_cir3_doColorLimit) on (a) the candidate lists the
           reference's own selectTopCostFromVolumn makes from its AD-Census + CBCA + sgm volume, (b) synthetic lists
           with many equal costs and far-apart candidates (the raster-dependent case of method 0 on most pixels)
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


TOP_VARIANTS = [(1, 0, 10, 1, 0), (2, 0, 10, 1, 0), (2, 0, 3, 1, 0), (2, 0, 10, 0, 0), (2, 0, 10, 1, 1), (2, 0, 1, 1, 1),
                (2, 1, 10, 1, 0), (2, 2, 10, 1, 0)]   # (version, method, ts, hasCir2, colorLimit)


def synth_top(rng, H, W, D, num, spread):
    top = np.zeros((H, W, num + 1, 2), np.float32)
    for v in range(H):
        for u in range(W):
            n = int(rng.integers(1, num + 1))
            base = int(rng.integers(0, D))
            ds = []
            while len(ds) < n and len(ds) < D:
                d = int(np.clip(base + rng.integers(-spread, spread + 1), 0, D - 1)) if rng.random() < 0.7 else int(rng.integers(0, D))
                if d not in ds:
                    ds.append(d)
            cs = np.sort(np.round(rng.random(len(ds)) * 8) / 8 + 1.0).astype(np.float32)     # many equal costs
            for k, d in enumerate(ds):
                top[v, u, k] = (d, cs[k])
            top[v, u, num, 0] = len(ds)
    return top


def top_cases(out):
    cases = {"real": (30, 52, 16, 6, 0, 11), "realM2": (24, 40, 12, 2, 0, 12), "syn": (20, 33, 40, 6, 3, 1), "far": (9, 50, 64, 4, 30, 2),
             "row": (1, 12, 8, 3, 2, 4), "col": (12, 1, 8, 3, 2, 5), "tiny": (2, 2, 8, 6, 9, 6)}
    for tag, (H, W, D, num, spread, seed) in cases.items():
        rng = np.random.default_rng(seed)
        if tag.startswith("real"):
            p = synth.make_pair(H, W, D, "texture_warped", seed)
            r = po.SmRef(p["bgrL"], p["bgrR"], p["grayL"], p["grayR"], D)
            r.adcensus(); r.cbca(2); r.sgm(0, 4)
            top = r.select_top(0, num, 1.09)           # main_.cpp: lamc = 109 -> vmTop_thres 1.09
            bgr = p["bgrL"]
        else:
            bgr = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
            bgr[:, : W // 2] //= 32                     # flat half: the colour limit passes there
            gray = np.zeros((H, W), np.uint8)
            r = po.SmRef(bgr, bgr, gray, gray, D)
            top = synth_top(rng, H, W, D, num, spread)
        out[f"top_{tag}_in"], out[f"top_{tag}_bgr"] = top, bgr
        for (ver, m, ts, c2, cl) in TOP_VARIANTS:
            out[f"top_{tag}_out_v{ver}_m{m}_ts{ts}_c{c2}_l{cl}"] = r.disp_from_top(top, ver, m, ts, c2, cl)
        r.close()


def main():
    out = {}
    lor1_cases(out)
    top_cases(out)
    path = os.path.join(ROOT, "tests", "golden", "r02_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
