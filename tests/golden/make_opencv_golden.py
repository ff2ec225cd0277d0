"""Generates tests/golden/opencv_semantics.npz with cv2 (4.13 in the build container).

Pins the three OpenCV behaviours the hot path relies on (the reference's C++ cannot be built
here; these are its only third-party arithmetic on the path):
  * copyMakeBorder(..., BORDER_REFLECT_101)       stereoMatching.h:642, 871
  * medianBlur(CV_16S, 3)                          stereoMatching.cpp:1499
  * imread(...,0) / cvtColor(BGR2GRAY) on 8-bit    main_.cpp:95-96
and the two of the caller's cross-scale step:
  * pyrDown on 8-bit images (1 and 3 channels)     main_.cpp:145-148
  * Mat::inv() of the float regularisation matrix  stereoMatching.cpp:2147-2166 (row 0 = invWgt)
Run:  python tests/golden/make_opencv_golden.py
"""
import os

import cv2
import numpy as np

rng = np.random.default_rng(20240607)
gray = rng.integers(0, 256, (13, 17), dtype=np.uint8)
reflect = cv2.copyMakeBorder(gray, 3, 3, 4, 4, cv2.BORDER_REFLECT_101)
tiny = rng.integers(0, 256, (2, 3), dtype=np.uint8)           # smaller than the border
reflect_tiny = cv2.copyMakeBorder(tiny, 1, 1, 2, 2, cv2.BORDER_REFLECT_101)
disp = rng.integers(-48, 64, (19, 23)).astype(np.int16)
disp[rng.random(disp.shape) < 0.2] = -1
med = cv2.medianBlur(disp, 3)
bgr = rng.integers(0, 256, (11, 9, 3), dtype=np.uint8)
g = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
pyr = {}
for k, shp in enumerate([(7, 9), (8, 10, 3), (5, 4), (2, 2), (1, 5), (3, 1), (31, 47, 3), (24, 40, 3)]):
    img = rng.integers(0, 256, shp, dtype=np.uint8)
    pyr[f"pyr_in{k}"] = img
    pyr[f"pyr_out{k}"] = cv2.pyrDown(img).reshape(((shp[0] + 1) // 2, (shp[1] + 1) // 2) + shp[2:])
inv_lams = np.array([0.3, 0.1, 1.0, 0.05, 2.5], np.float32)
inv_rows = np.zeros((len(inv_lams), 7, 7), np.float32)        # [lambda][levels-1][:levels]
for a, lam in enumerate(inv_lams):
    for n in range(1, 8):
        M = np.zeros((n, n), np.float32)
        for s_ in range(n):
            if s_ == 0:
                M[s_, s_] = 1 + lam
                if n > 1:
                    M[s_, s_ + 1] = -lam
            elif s_ == n - 1:
                M[s_, s_] = 1 + lam
                M[s_, s_ - 1] = -lam
            else:
                M[s_, s_] = 1 + 2 * lam
                M[s_, s_ - 1] = -lam
                M[s_, s_ + 1] = -lam
        inv_rows[a, n - 1, :n] = cv2.invert(M)[1][0]
np.savez_compressed(os.path.join(os.path.dirname(__file__), "opencv_semantics.npz"), inv_lams=inv_lams, inv_rows=inv_rows,
                    n_pyr=len(pyr) // 2, **pyr,
                    gray=gray, reflect=reflect, tiny=tiny, reflect_tiny=reflect_tiny,
                    disp=disp, median3=med, bgr=bgr, bgr2gray=g, cv2_version=cv2.__version__)
print("wrote opencv_semantics.npz with cv2", cv2.__version__)
