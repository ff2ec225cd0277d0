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
#include "opencv_restated.h"

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
// discontinuityAdjust's three imgproc calls (stereoMatching.cpp:6061-6064) -> the cv2-pinned restatements of
// opencv_restated.h; only the argument combinations that function uses are accepted
struct Size { int width, height; Size(int w, int h) : width(w), height(h) {} };
inline void equalizeHist(const Mat& src, Mat& dst) {
  if (src.dims != 2 || src.type() != CV_8UC1) { fprintf(stderr, "cv_standin: unsupported equalizeHist\n"); abort(); }
  Mat out(src.rows, src.cols, CV_8UC1);
  orc_cv::equalize_hist(src.data, out.data, src.rows, src.cols);
  dst = out;
}
inline void GaussianBlur(const Mat& src, Mat& dst, Size k, double sx, double sy) {
  if (src.dims != 2 || src.type() != CV_8UC1 || k.width != 3 || k.height != 3 || sx != 4 || sy != 4) {
    fprintf(stderr, "cv_standin: unsupported GaussianBlur\n"); abort();
  }
  Mat out(src.rows, src.cols, CV_8UC1);
  orc_cv::gauss3_sigma4(src.data, out.data, src.rows, src.cols);
  dst = out;
}
inline void Canny(const Mat& src, Mat& dst, double low, double high, int aperture) {
  if (src.dims != 2 || src.type() != CV_8UC1 || aperture != 3) { fprintf(stderr, "cv_standin: unsupported Canny\n"); abort(); }
  Mat out(src.rows, src.cols, CV_8UC1);
  orc_cv::canny3_l1(src.data, out.data, src.rows, src.cols, (int)std::floor(low), (int)std::floor(high));
  dst = out;
}
inline bool imwrite(const std::string&, const Mat&) { return true; }  // debug dumps of the reference: dropped
// cv::pyrDown on 8-bit images (main_.cpp:145-148): [1 4 6 4 1]/16 separable, BORDER_REFLECT_101, (sum + 128) >> 8,
// dst = ((cols+1)/2, (rows+1)/2); pinned against cv2.pyrDown (tests/golden/opencv_semantics.npz)
inline void pyrDown(const Mat& src, Mat& dst) {
  if (src.dims != 2 || src.depth() != CV_8U) { fprintf(stderr, "cv_standin: unsupported pyrDown\n"); abort(); }
  static const int k[5] = {1, 4, 6, 4, 1};
  const int cn = src.channels(), Ho = (src.rows + 1) / 2, Wo = (src.cols + 1) / 2;
  Mat out(Ho, Wo, src.type());
  for (int y = 0; y < Ho; y++)
    for (int x = 0; x < Wo; x++)
      for (int c = 0; c < cn; c++) {
        int s = 0;
        for (int i = 0; i < 5; i++) {
          const uchar* row = src.data + (size_t)reflect101_index(2 * y + i - 2, src.rows) * src.step[0];
          for (int j = 0; j < 5; j++) s += k[i] * k[j] * row[reflect101_index(2 * x + j - 2, src.cols) * cn + c];
        }
        out.data[(size_t)y * out.step[0] + x * cn + c] = (uchar)((s + 128) >> 8);
      }
  dst = out;
}
// `vm[0] /= wetNL` (StereoMatching::NL, stereoMatching.cpp:4910): element-wise float division of two CV_32F Mats holding
// the same number of values (both are h x w x d there, as 2-D multi-channel Mats)
inline Mat& operator/=(Mat& a, const Mat& b) {
  const size_t n = a.total() * a.channels();
  if (a.depth() != CV_32F || b.depth() != CV_32F || n != b.total() * b.channels()) { fprintf(stderr, "cv_standin: unsupported operator/=\n"); abort(); }
  float* x = (float*)a.data;
  const float* y = (const float*)b.data;
  for (size_t i = 0; i < n; i++) x[i] = x[i] / y[i];
  return a;
}
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
// SolveAll / PrintMat report on stdout with printf (stereoMatching.cpp:2144-2145, 2127-2136); stdout carries
// bench.py's one JSON line, so that chatter is dropped
inline int smref_printf_sink(const char*, ...) { return 0; }
#define printf smref_printf_sink
