"""Generates tests/golden/opencv_semantics.npz with cv2 (4.13 in the build container).

Pins the three OpenCV behaviours the hot path relies on (the reference's C++ cannot be built
here; these are its only third-party arithmetic on the path):
  * copyMakeBorder(..., BORDER_REFLECT_101)       stereoMatching.h:642, 871
  * medianBlur(CV_16S, 3)                          stereoMatching.cpp:1499
  * imread(...,0) / cvtColor(BGR2GRAY) on 8-bit    main_.cpp:95-96
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
np.savez_compressed(os.path.join(os.path.dirname(__file__), "opencv_semantics.npz"),
                    gray=gray, reflect=reflect, tiny=tiny, reflect_tiny=reflect_tiny,
                    disp=disp, median3=med, bgr=bgr, bgr2gray=g, cv2_version=cv2.__version__)
print("wrote opencv_semantics.npz with cv2", cv2.__version__)
