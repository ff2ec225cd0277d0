// TEST INFRASTRUCTURE (oracle/_ref build only).  The slice of OpenCV that the reference's hot-path functions touch,
// so that their bodies -- cut from /root/reference by build_ref_sm.py -- compile without OpenCV C++ (absent in this
// image).  cv::Mat is the product's own stand-in (host/cvmat_lite.h: same layout, ptr<T>, create, clone, copyTo,
// `= scalar`); this header adds the free functions.  copyMakeBorder implements BORDER_REFLECT_101 only (the one mode
// the path uses, stereoMatching.h:642, 871; pinned against cv2 by tests/golden/opencv_semantics.npz); everything
// the default parameters never reach aborts.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <vector>
#include "cvmat_lite.h"

namespace cv {
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4 };

inline int reflect101_index(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int mode) {
  if (mode != BORDER_REFLECT_101 || src.dims != 2) { fprintf(stderr, "cv_standin: unsupported copyMakeBorder\n"); abort(); }
  Mat out(src.rows + top + bottom, src.cols + left + right, src.type());
  const size_t es = src.elemSize();
  for (int v = 0; v < out.rows; v++) {
    const int sv = reflect101_index(v - top, src.rows);
    for (int u = 0; u < out.cols; u++) {
      const int su = reflect101_index(u - left, src.cols);
      memcpy(out.data + v * out.step[0] + u * es, src.data + sv * src.step[0] + su * es, es);
    }
  }
  dst = out;
}
inline bool imwrite(const std::string&, const Mat&) { return true; }  // debug dumps of the reference: dropped
namespace ximgproc {
inline void guidedFilter(const Mat&, const Mat&, Mat&, int, double) {
  fprintf(stderr, "cv_standin: ximgproc::guidedFilter is not available (doGF_bef_calArm must stay false)\n");
  abort();
}
}  // namespace ximgproc
}  // namespace cv
using namespace std;
using namespace cv;
#define popcnt64 __builtin_popcountll   // stereoMatching.cpp:9 (the non-MSVC branch)
