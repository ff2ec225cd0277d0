// NLCCA.h -- NLCCA::aggreCV (NL/NLCCA.h:21-31, NL/NLCCA.cpp:27-95) over the sm_b200 C ABI.
#pragma once
#ifndef SM_USE_OPENCV
#include "../cvmat_lite.h"
#else
#include <opencv2/core.hpp>
#endif
using cv::Mat;

class NLCCA {
 public:
  NLCCA(void) {}
  ~NLCCA(void) {}
  // lImg: H x W CV_8UC3 guidance; costVol: H x W x maxDis float32 (d fastest), filtered in place
  // (f32 -> f64, MST on the median-filtered left image, tree filter with sigma 0.1, f64 -> f32).
  void aggreCV(const Mat& lImg, const Mat& rImg, const int maxDis, Mat& costVol);
};
