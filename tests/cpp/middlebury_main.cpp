// The reference's main() (main_.cpp:20-200) over the drop-in class and the file side in host/sm_io.h: a Middlebury object
// folder in (colour pair, gray pair, masks, ground truth / disp_reduceCoeff), the refined disparity map out as the
// picture saveDispMap writes, plus the calErr line.  Needs a GPU (the class constructor creates the device context).
//   middlebury_main <root/> <object> [ext=.png] [paths=4] [raw_out.i16]
// Output pictures go to StereoMatching::root + object + "/<cost>-<aggr>-<opt>/20200627_test_so/" below the working directory,
// as in the reference (stereoMatching.h:348) with its root left empty.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

#include "sm_io.h"
#include "stereoMatching.h"

static cv::Mat wrap(smio::Image& im) {
  if (im.empty()) return cv::Mat();
  return cv::Mat(im.h, im.w, im.c == 3 ? CV_8UC3 : CV_8UC1, im.data.data());
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: middlebury_main <root/> <object> [ext] [paths] [raw_out]\n"); return 2; }
  const std::string mdroot = argv[1], object = argv[2], ext = argc > 3 ? argv[3] : ".png";
  const int paths = argc > 4 ? atoi(argv[4]) : 4;
  StereoMatching::costcalculation = "ADCensus";
  StereoMatching::aggregation = "CBCA";
  StereoMatching::optimization = "sgm";
  StereoMatching::object = object;
  const std::string method = StereoMatching::costcalculation + StereoMatching::aggregation + StereoMatching::optimization;
  printf("method: %s\nobject: %s\n", method.c_str(), object.c_str());

  smio::StereoPair pair;
  std::string err;
  if (!smio::load_middlebury(mdroot, object, pair, &err, ext)) {   // main_.cpp:103-107
    std::cout << "can't read original img" << std::endl;
    fprintf(stderr, "%s\n", err.c_str());
    return -1;
  }
  if (pair.all_mask.empty() || pair.nonocc_mask.empty() || pair.disc_mask.empty())   // main_.cpp:108-112: not fatal
    std::cout << "can't read mask img" << std::endl;
  std::cout << "read-in img done" << std::endl;

  cv::Mat I1_c = wrap(pair.I1_c), I2_c = wrap(pair.I2_c), I1 = wrap(pair.I1_g), I2 = wrap(pair.I2_g);
  cv::Mat all_maskM = wrap(pair.all_mask), nonocc_maskM = wrap(pair.nonocc_mask), disc_maskM = wrap(pair.disc_mask);
  cv::Mat DT;
  if (!pair.DT.data.empty()) DT = cv::Mat(pair.DT.h, pair.DT.w, CV_32FC1, pair.DT.data.data());

  try {
    const int lamG = 1, lamCen = 13, M = 2, lamc = 109, ts = 10, disSc = 1, PY_LEV = 1;   // main_.cpp:60-64, 131-132
    StereoMatching** smPsy = new StereoMatching*[PY_LEV];
    StereoMatching::Parameters param(pair.maxdisp, I1_c.rows, I1_c.cols, lamCen, lamG, M, lamc, ts, "", disSc);
    smPsy[0] = new StereoMatching(I1_c, I2_c, I1, I2, DT, all_maskM, nonocc_maskM, disc_maskM, param);
    smPsy[0]->setSgmPaths(paths);
    smPsy[0]->costCalculate();
    SolveAll(smPsy, PY_LEV, 0.3f);   // main_.cpp:157-158
    smPsy[0]->dispOptimize();
    smPsy[0]->refine();
    smPsy[0]->syncToHost(false);
    cv::Mat& dp = smPsy[0]->DP[0];
    const bool haveTruth = !DT.empty() && !all_maskM.empty();
    if (haveTruth) smPsy[0]->calErr<short>(dp, smPsy[0]->DT, "final");
    smPsy[0]->saveDispMap<short>(dp, smPsy[0]->DT, "final", haveTruth);
    printf("saved: %sfinal.png\n", smPsy[0]->param_.savePath.c_str());
    if (argc > 5) {
      std::ofstream f(argv[5], std::ios::binary);
      f.write((const char*)dp.data, (std::streamsize)((size_t)dp.rows * dp.cols * 2));
    }
    delete smPsy[0];
    delete[] smPsy;
  } catch (const cv::Exception& e) {
    fprintf(stderr, "cv::Exception: %s\n", e.what());
    return 1;
  }
  std::cout << "complete " << object << std::endl;
  return 0;
}
