// cvmat_lite.h -- the slice of cv::Mat the StereoMatching stage API touches, so the class header compiles without
// OpenCV (absent in the build image).  Same names, same memory layout (row-major, channels interleaved, n-dim
// `size[]`/`step[]`), reference-counted buffer, `ptr<T>(i0[,i1[,i2]])`, `create`, `clone`, `copyTo`, CV_Assert
// raising cv::Exception.  When real OpenCV is available, define SM_USE_OPENCV before including stereoMatching.h
// and this file is skipped.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_CN_MAX 512
#define CV_CN_SHIFT 3
#define CV_DEPTH_MAX (1 << CV_CN_SHIFT)
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH(flags) ((flags) & (CV_DEPTH_MAX - 1))
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_MAT_CN(flags) ((((flags) >> CV_CN_SHIFT) & (CV_CN_MAX - 1)) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16SC1 CV_MAKETYPE(CV_16S, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_32FC(n) CV_MAKETYPE(CV_32F, (n))
#define CV_16UC(n) CV_MAKETYPE(CV_16U, (n))

namespace cv {

class Exception : public std::runtime_error {
 public:
  explicit Exception(const std::string& m) : std::runtime_error(m) {}
};

#define CV_Assert(expr)                                                                                   \
  do {                                                                                                    \
    if (!(expr))                                                                                          \
      throw cv::Exception(std::string("CV_Assert failed: ") + #expr + " at " + __FILE__ + ":" +           \
                          std::to_string(__LINE__));                                                      \
  } while (0)

struct Scalar {
  double val[4];
  Scalar(double v0 = 0, double v1 = 0, double v2 = 0, double v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
  static Scalar all(double v) { return Scalar(v, v, v, v); }
};

class Mat {
 public:
  enum { MAX_DIM = 4 };
  int flags = 0;  // type
  int dims = 0;
  int rows = 0, cols = 0;
  int size[MAX_DIM] = {0, 0, 0, 0};
  size_t step[MAX_DIM] = {0, 0, 0, 0};
  uchar* data = nullptr;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int nd, const int* sz, int type) { create(nd, sz, type); }
  Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); *this = s; }
  Mat(int nd, const int* sz, int type, const Scalar& s) { create(nd, sz, type); *this = s; }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type, Scalar::all(0)); }
  static Mat ones(int r, int c, int type) { return Mat(r, c, type, Scalar(1)); }
  // header over caller-owned memory (no copy), like cv::Mat(rows, cols, type, void*)
  Mat(int r, int c, int type, void* ext) { header2d(r, c, type); data = (uchar*)ext; }

  static size_t depth_bytes(int depth) {
    static const size_t b[7] = {1, 1, 2, 2, 4, 4, 8};
    return b[depth];
  }
  int type() const { return flags; }
  int depth() const { return CV_MAT_DEPTH(flags); }
  int channels() const { return CV_MAT_CN(flags); }
  size_t elemSize1() const { return depth_bytes(depth()); }
  size_t elemSize() const { return elemSize1() * channels(); }
  bool empty() const { return data == nullptr || total() == 0; }
  bool isContinuous() const { return true; }
  size_t total() const {
    if (dims == 0) return 0;
    size_t n = 1;
    for (int i = 0; i < dims; i++) n *= (size_t)size[i];
    return n;
  }
  size_t bytes() const { return total() * elemSize(); }

  void create(int r, int c, int type) {
    if (dims == 2 && rows == r && cols == c && flags == type && data) return;
    header2d(r, c, type);
    alloc();
  }
  void create(int nd, const int* sz, int type) {
    CV_Assert(nd >= 1 && nd <= MAX_DIM);
    bool same = dims == nd && flags == type && data;
    for (int i = 0; same && i < nd; i++) same = size[i] == sz[i];
    if (same) return;
    flags = type; dims = nd;
    for (int i = 0; i < MAX_DIM; i++) size[i] = i < nd ? sz[i] : 0;
    rows = nd >= 1 ? sz[0] : 0; cols = nd >= 2 ? sz[1] : 1;
    if (nd > 2) rows = cols = -1;   // OpenCV convention for n-dim matrices
    size_t s = elemSize();
    for (int i = nd - 1; i >= 0; i--) { step[i] = s; s *= (size_t)sz[i]; }
    alloc();
  }
  void release() { buf_.reset(); data = nullptr; dims = 0; rows = cols = 0; }

  template <typename T> T* ptr(int i0 = 0) { return (T*)(data + (size_t)i0 * step[0]); }
  template <typename T> const T* ptr(int i0 = 0) const { return (const T*)(data + (size_t)i0 * step[0]); }
  template <typename T> T* ptr(int i0, int i1) { return (T*)(data + (size_t)i0 * step[0] + (size_t)i1 * step[1]); }
  template <typename T> const T* ptr(int i0, int i1) const { return (const T*)(data + (size_t)i0 * step[0] + (size_t)i1 * step[1]); }
  template <typename T> T* ptr(int i0, int i1, int i2) {
    return (T*)(data + (size_t)i0 * step[0] + (size_t)i1 * step[1] + (size_t)i2 * step[2]);
  }
  template <typename T> T& at(int i0, int i1) { return *ptr<T>(i0, i1); }
  template <typename T> const T& at(int i0, int i1) const { return *ptr<T>(i0, i1); }
  template <typename T> T& at(int i0, int i1, int i2) { return *ptr<T>(i0, i1, i2); }
  // cv::Mat::inv() (DECOMP_LU) for a square CV_32F matrix, the way cv::invert evaluates it: n = 1, 3 closed forms in
  // double rounded once; n = 2 determinant in double, float products with (float)(1/det) (the SIMD128 branch);
  // n > 3 Gaussian elimination with partial pivoting in float.  Reproduces cv2.invert bit for bit on the matrices
  // SolveAll builds (tests/golden/opencv_semantics.npz); a singular matrix yields zeros, as in OpenCV.
  Mat inv() const {
    CV_Assert(dims == 2 && rows == cols && type() == CV_32FC1);
    const int n = rows;
    Mat out = zeros(n, n, CV_32FC1);
    auto S = [&](int r, int c) { return at<float>(r, c); };
    if (n == 1) {
      const double d = S(0, 0);
      if (d != 0.) out.at<float>(0, 0) = (float)(1. / d);
    } else if (n == 2) {
      double d = (double)S(0, 0) * S(1, 1) - (double)S(0, 1) * S(1, 0);
      if (d != 0.) {
        const float f = (float)(1. / d);
        out.at<float>(1, 1) = S(0, 0) * f; out.at<float>(0, 0) = S(1, 1) * f;
        out.at<float>(0, 1) = -(S(0, 1) * f); out.at<float>(1, 0) = -(S(1, 0) * f);
      }
    } else if (n == 3) {
      auto D = [&](int r, int c) { return (double)at<float>(r, c); };
      double d = D(0, 0) * (D(1, 1) * D(2, 2) - D(1, 2) * D(2, 1)) - D(0, 1) * (D(1, 0) * D(2, 2) - D(1, 2) * D(2, 0)) +
                 D(0, 2) * (D(1, 0) * D(2, 1) - D(1, 1) * D(2, 0));
      if (d != 0.) {
        d = 1. / d;
        out.at<float>(0, 0) = (float)((D(1, 1) * D(2, 2) - D(1, 2) * D(2, 1)) * d);
        out.at<float>(0, 1) = (float)((D(0, 2) * D(2, 1) - D(0, 1) * D(2, 2)) * d);
        out.at<float>(0, 2) = (float)((D(0, 1) * D(1, 2) - D(0, 2) * D(1, 1)) * d);
        out.at<float>(1, 0) = (float)((D(1, 2) * D(2, 0) - D(1, 0) * D(2, 2)) * d);
        out.at<float>(1, 1) = (float)((D(0, 0) * D(2, 2) - D(0, 2) * D(2, 0)) * d);
        out.at<float>(1, 2) = (float)((D(0, 2) * D(1, 0) - D(0, 0) * D(1, 2)) * d);
        out.at<float>(2, 0) = (float)((D(1, 0) * D(2, 1) - D(1, 1) * D(2, 0)) * d);
        out.at<float>(2, 1) = (float)((D(0, 1) * D(2, 0) - D(0, 0) * D(2, 1)) * d);
        out.at<float>(2, 2) = (float)((D(0, 0) * D(1, 1) - D(0, 1) * D(1, 0)) * d);
      }
    } else {
      Mat A = clone();
      for (int i = 0; i < n; i++) out.at<float>(i, i) = 1.f;
      volatile float t;   // every product and sum stays a rounded float
      for (int i = 0; i < n; i++) {
        int k = i;
        for (int j = i + 1; j < n; j++)
          if (std::abs(A.at<float>(j, i)) > std::abs(A.at<float>(k, i))) k = j;
        if (std::abs(A.at<float>(k, i)) < 1.1920929e-06f) return zeros(n, n, CV_32FC1);   // FLT_EPSILON * 10
        if (k != i)
          for (int c = 0; c < n; c++) { std::swap(A.at<float>(i, c), A.at<float>(k, c)); std::swap(out.at<float>(i, c), out.at<float>(k, c)); }
        const float d = -1 / A.at<float>(i, i);
        for (int j = i + 1; j < n; j++) {
          const float alpha = A.at<float>(j, i) * d;
          for (int c = i + 1; c < n; c++) { t = alpha * A.at<float>(i, c); A.at<float>(j, c) += t; }
          for (int c = 0; c < n; c++) { t = alpha * out.at<float>(i, c); out.at<float>(j, c) += t; }
        }
      }
      for (int i = n - 1; i >= 0; i--)
        for (int j = 0; j < n; j++) {
          float sacc = out.at<float>(i, j);
          for (int c = i + 1; c < n; c++) { t = A.at<float>(i, c) * out.at<float>(c, j); sacc -= t; }
          out.at<float>(i, j) = sacc / A.at<float>(i, i);
        }
    }
    return out;
  }

  Mat clone() const { Mat m; copyTo(m); return m; }
  void copyTo(Mat& dst) const {
    if (dims <= 2) dst.create(rows, cols, flags); else dst.create(dims, size, flags);
    if (bytes()) memcpy(dst.data, data, bytes());
  }
  // convertTo without scale, for the one conversion the path uses (discontinuityAdjust, stereoMatching.cpp:6060:
  // CV_16S -> CV_8U with saturate_cast; NL's CV_16S -> CV_32F) plus the identity; anything else is a CV_Assert failure
  void convertTo(Mat& dst, int rtype) const {
    const int ddepth = rtype & 7;
    if (ddepth == depth()) { Mat t = clone(); dst = t; return; }
    CV_Assert(depth() == CV_16S && (ddepth == CV_8U || ddepth == CV_32F) && dims <= 2);
    Mat out(rows, cols, CV_MAKETYPE(ddepth, channels()));
    const size_t n = total() * channels();
    const short* s = (const short*)data;
    if (ddepth == CV_8U) for (size_t i = 0; i < n; i++) out.data[i] = (uchar)(s[i] < 0 ? 0 : s[i] > 255 ? 255 : s[i]);
    else for (size_t i = 0; i < n; i++) ((float*)out.data)[i] = (float)s[i];   // StereoMatching::NL, stereoMatching.cpp:4915
    dst = out;
  }
  Mat& setZero() { if (bytes()) memset(data, 0, bytes()); return *this; }
  // `m = s`: every element's channel c becomes s.val[c] (saturate_cast semantics are not needed by the callers:
  // they only assign 0, 1 and small integers)
  Mat& operator=(const Scalar& s) {
    const int cn = channels();
    const size_t n = total();
    for (size_t i = 0; i < n; i++)
      for (int c = 0; c < cn; c++) {
        const double v = s.val[c < 4 ? c : 3];
        uchar* p = data + (i * cn + c) * elemSize1();
        switch (depth()) {
          case CV_8U: *(uchar*)p = (uchar)v; break;
          case CV_8S: *(signed char*)p = (signed char)v; break;
          case CV_16U: *(ushort*)p = (ushort)v; break;
          case CV_16S: *(short*)p = (short)v; break;
          case CV_32S: *(int*)p = (int)v; break;
          case CV_32F: *(float*)p = (float)v; break;
          default: *(double*)p = v; break;
        }
      }
    return *this;
  }

 private:
  std::shared_ptr<uchar> buf_;
  void header2d(int r, int c, int type) {
    flags = type; dims = 2; rows = r; cols = c;
    size[0] = r; size[1] = c; size[2] = size[3] = 0;
    step[1] = elemSize(); step[0] = step[1] * (size_t)c; step[2] = step[3] = 0;
  }
  void alloc() {
    size_t n = bytes();
    buf_.reset(n ? new uchar[n] : nullptr, std::default_delete<uchar[]>());
    data = buf_.get();
  }
};

template <typename T> struct mat_depth;
template <> struct mat_depth<uchar> { enum { value = CV_8U }; };
template <> struct mat_depth<ushort> { enum { value = CV_16U }; };
template <> struct mat_depth<short> { enum { value = CV_16S }; };
template <> struct mat_depth<int> { enum { value = CV_32S }; };
template <> struct mat_depth<float> { enum { value = CV_32F }; };
template <> struct mat_depth<double> { enum { value = CV_64F }; };

template <typename T>
class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, CV_MAKETYPE(mat_depth<T>::value, 1)) {}
  void create(int nd, const int* sz) { Mat::create(nd, sz, CV_MAKETYPE(mat_depth<T>::value, 1)); }
  void create(int r, int c) { Mat::create(r, c, CV_MAKETYPE(mat_depth<T>::value, 1)); }
};
typedef Mat_<float> Mat1f;

}  // namespace cv
