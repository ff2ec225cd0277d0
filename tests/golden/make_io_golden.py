"""Fixtures for mystereomatching_b200/host/sm_io (the file side of main_.cpp:85-129 and saveDispMap).

Run once in the build container (cv2 4.13 is importable there):  python tests/golden/make_io_golden.py
Writes small PNG / PPM / PGM / PFM files under tests/golden/io/ and tests/golden/io_ref.npz = what cv2.imread (flags 1 and 0,
the two calls of main_.cpp:91-95) returns for each of them, plus a tiny Middlebury-shaped folder ("Aloe": view1 / view5 /
disp1 / all / nonocc, no disc mask).  The tests read only these files, never cv2."""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "io")


def main():
    rng = np.random.default_rng(20260)
    os.makedirs(OUT, exist_ok=True)
    ref = {}
    h, w = 13, 21
    bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    bgr[0, :8] = [[255, 255, 255], [0, 0, 0], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 2, 3], [254, 253, 252], [128, 127, 129]]
    gray = rng.integers(0, 256, (h, w), dtype=np.uint8)
    g16 = rng.integers(0, 65536, (h, w), dtype=np.uint16)
    bgra = np.concatenate([bgr, rng.integers(0, 256, (h, w, 1), dtype=np.uint8)], axis=2)
    files = {
        "rgb8.png": bgr, "gray8.png": gray, "gray16.png": g16, "rgba8.png": bgra,
        "rgb16.png": (bgr.astype(np.uint16) * 257 + 3), "rgb8.ppm": bgr, "gray8.pgm": gray,
    }
    for name, arr in files.items():
        path = os.path.join(OUT, name)
        assert cv2.imwrite(path, arr)
        ref[name + ":1"] = cv2.imread(path, 1)
        ref[name + ":0"] = cv2.imread(path, 0)
    # a smooth picture so that the encoder's adaptive filtering picks Sub / Up / Average / Paeth rows
    big = cv2.resize(bgr, (64, 40), interpolation=cv2.INTER_CUBIC)
    path = os.path.join(OUT, "smooth.png")
    assert cv2.imwrite(path, big, [cv2.IMWRITE_PNG_STRATEGY, cv2.IMWRITE_PNG_STRATEGY_FILTERED, cv2.IMWRITE_PNG_COMPRESSION, 9])
    ref["smooth.png:1"] = cv2.imread(path, 1)
    ref["smooth.png:0"] = cv2.imread(path, 0)
    # PFM (Middlebury 2014 ground truth): cv2 returns it top-down as float32
    pf = rng.uniform(0, 200, (h, w)).astype(np.float32)
    pf[2, 3] = np.inf
    path = os.path.join(OUT, "gt.pfm")
    assert cv2.imwrite(path, pf)
    ref["gt.pfm"] = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    # Middlebury-shaped object folder
    md = os.path.join(OUT, "md", "Aloe")
    os.makedirs(md, exist_ok=True)
    left = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    right = np.roll(left, -2, axis=1)
    disp = rng.integers(0, 256, (h, w), dtype=np.uint8)
    allm = (rng.random((h, w)) > 0.2).astype(np.uint8) * 255
    nonocc = (rng.random((h, w)) > 0.4).astype(np.uint8) * 255
    for n, a in (("view1", left), ("view5", right), ("disp1", disp), ("all", allm), ("nonocc", nonocc)):
        assert cv2.imwrite(os.path.join(md, n + ".png"), a)
    ref["md:I1_c"] = cv2.imread(os.path.join(md, "view1.png"), 1)
    ref["md:I2_c"] = cv2.imread(os.path.join(md, "view5.png"), 1)
    ref["md:I1_g"] = cv2.imread(os.path.join(md, "view1.png"), 0)
    ref["md:I2_g"] = cv2.imread(os.path.join(md, "view5.png"), 0)
    ref["md:all"] = cv2.imread(os.path.join(md, "all.png"), 0)
    ref["md:nonocc"] = cv2.imread(os.path.join(md, "nonocc.png"), 0)
    dt = cv2.imread(os.path.join(md, "disp1.png"), 0)
    # DT.convertTo(DT, CV_32F, 1.0 / disp_reduceCoeffList[i])  (main_.cpp:128-129; Aloe: coefficient 3).  cv2's Python
    # binding has no Mat::convertTo; OpenCV evaluates 8u -> 32f with scale in FLOAT (core/src/convert_scale.simd.hpp,
    # cvt_32f: dst = src * (float)alpha + (float)beta), which is what this line restates.
    ref["md:DT"] = dt.astype(np.float32) * np.float32(1.0 / np.float32(3.0))
    np.savez_compressed(os.path.join(HERE, "io_ref.npz"), **ref)
    print("wrote", len(ref), "arrays")


if __name__ == "__main__":
    main()
