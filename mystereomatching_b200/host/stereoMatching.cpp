// stereoMatching.cpp -- host wrappers of the StereoMatching stage API over the sm_b200 C ABI.
// Each method names the reference body it replaces (file:line under the reference root) in stereoMatching.h.
#include "stereoMatching.h"
#include "sm_io.h"

#include <sys/stat.h>

#include <cmath>
#include <iostream>

#include <algorithm>
#include <cmath>
#include <cstdio>

const std::string StereoMatching::root = "";
std::string StereoMatching::costcalculation = "ADCensus";
std::string StereoMatching::aggregation = "CBCA";
std::string StereoMatching::optimization = "sgm";
std::string StereoMatching::object = "";
bool StereoMatching::Do_refine = true;
bool StereoMatching::Do_LRConsis = true;
bool StereoMatching::Do_regionVote = true;
bool StereoMatching::Do_properIpol = true;
bool StereoMatching::Do_lastMedianBlur = true;
bool StereoMatching::Do_subpixelEnhancement = false;
bool StereoMatching::Do_WM = false;
bool StereoMatching::Do_discontinuityAdjust = false;

// reference defaults: stereoMatching.h:204-350
StereoMatching::Parameters::Parameters(int maxDisp, int h, int w, int lamCen_, int lamG_, int M_, int lamc_, int ts_,
                                       string errCsvName_, int disSc_) {
  (void)h; (void)w;
  W_U = 4; W_V = 3;
  ChooseSmall = true;
  numDisparities = maxDisp + 1;
  LRmaxDiff = 0;
  DISP_INV = -16; DISP_OCC = -2 * 16; DISP_MIS = -3 * 16; DISP_PKR = -4 * 16;
  DISP_SCALE = 16; DISP_SHIFT = 4;
  errorThreshold = 1;
  SD_AD_channel = 3; census_channel = 1;
  sgm_scanNum = 4; sgm_P1 = 10000; sgm_P2 = 10000;
  sgm_corDifThres = 15; sgm_reduCoeffi1 = 4; sgm_reduCoeffi2 = 10;
  censusFunc = 3;
  is_censusNorm = 0; is_adNorm = 0;
  has_initArm = 0; has_calArms = 0;
  cbca_minArmL = 1; cbca_iterationNum = 2; cbca_intersect = true;
  cbca_crossL[0] = 17; cbca_crossL[1] = 17;
  cbca_crossL_out[0] = 34; cbca_crossL_out[1] = 34;
  cbca_cTresh[0] = 20; cbca_cTresh[1] = 20;
  cbca_cTresh_out[0] = 6; cbca_cTresh_out[1] = 6;
  cbca_double_win = false; cbca_armHV = 1; cbca_armTile = 0;
  region_vote_nums = 2; regVote_SThres = 20; regVote_hratioThres = 0.4f;
  Do_vmTop = false;
  vmTop_method = 0; vmTop_Num = M_; vmTop_thres = lamc_ * 0.01f; vmTop_thres_dirNum = 8;
  vmTop_hasCir2 = true; vmTop_cir3_doColorLimit = false;
  lamCen = lamCen_; lamG = lamG_; M = M_; lamc = lamc_; ts = ts_; disSc = disSc_;
  errCsvName = errCsvName_;
  err_ip_dispV = -50; cor_ip_dispV = -100;
  savePath = StereoMatching::root + StereoMatching::object + "/" + StereoMatching::costcalculation + "-" +
             StereoMatching::aggregation + "-" + StereoMatching::optimization + "/" + "20200627_test_so" + "/";
}

// ------------------------------------------------------------------ plumbing
void StereoMatching::check(int rc, const char* what) {
  if (rc != SM_OK) throw cv::Exception(std::string(what) + ": " + sm_last_error());
}
void* StereoMatching::dalloc(size_t bytes) {
  void* p = nullptr;
  check(sm_dev_alloc(ctx_, &p, bytes), "sm_dev_alloc");
  owned_.push_back(p);
  return p;
}
void StereoMatching::upload(void* d, const void* h, size_t bytes) {
  check(sm_memcpy_h2d(ctx_, d, h, bytes), "sm_memcpy_h2d");
  check(sm_ctx_sync(ctx_), "sm_ctx_sync");   // pageable source: finish before the caller may touch it
}
void StereoMatching::download(void* h, const void* d, size_t bytes) {
  check(sm_memcpy_d2h(ctx_, h, d, bytes), "sm_memcpy_d2h");
  check(sm_ctx_sync(ctx_), "sm_ctx_sync");
}

namespace {
struct TmpDev {   // scoped device allocation for the explicit-argument methods
  sm_ctx* c; void* p = nullptr;
  TmpDev(sm_ctx* ctx, size_t bytes) : c(ctx) {
    if (sm_dev_alloc(c, &p, bytes) != SM_OK) throw cv::Exception(std::string("sm_dev_alloc: ") + sm_last_error());
  }
  ~TmpDev() { sm_dev_free(c, p); }
  template <typename T> T* as() { return (T*)p; }
};
void ensure_vol(Mat& m, int h, int w, int d, int type = CV_32F) {
  const bool ok = !m.empty() && m.total() * m.channels() == (size_t)h * w * d && m.depth() == type;
  if (!ok) { int sz[3] = {h, w, d}; m.create(3, sz, CV_MAKETYPE(type, 1)); }
}
}  // namespace

StereoMatching::StereoMatching(cv::Mat& I1_c, cv::Mat& I2_c, cv::Mat& I1_g, cv::Mat& I2_g, cv::Mat& DT_, cv::Mat& all_mask,
                               cv::Mat& nonocc_mask, cv::Mat& disc_mask, const Parameters& param)
    : param_(param) {
  // stereoMatching.cpp:2058-2110
  h_ = I1_c.rows; w_ = I1_c.cols; d_ = param_.numDisparities;
  CV_Assert(h_ > 0 && w_ > 0 && I2_c.rows == h_ && I2_c.cols == w_);
  CV_Assert(I1_c.type() == CV_8UC3 && I2_c.type() == CV_8UC3);
  CV_Assert(I1_g.type() == CV_8UC1 && I2_g.type() == CV_8UC1 && I1_g.rows == h_ && I1_g.cols == w_);
  CV_Assert(d_ >= 1 && d_ <= CV_CN_MAX);   // vm is CV_32FC(D)
  I_c = {I1_c, I2_c}; I_g = {I1_g, I2_g};
  I_mask = {nonocc_mask, all_mask, disc_mask};
  DT = DT_;
  size_vm[0] = h_; size_vm[1] = w_; size_vm[2] = d_;
  vm.resize(2);
  for (int i = 0; i < 2; i++) vm[i].create(h_, w_, CV_32FC(d_));
  HVL_num = 2;
  if (sm_device_count() <= 0) throw cv::Exception("StereoMatching: no CUDA device (sm_b200 has no CPU fallback)");
  check(sm_ctx_create(&ctx_, 0, nullptr), "sm_ctx_create");
  const size_t npix = (size_t)h_ * w_, nvol = npix * d_;
  for (int i = 0; i < 2; i++) {
    d_bgr_[i] = (uint8_t*)dalloc(npix * 3);
    d_gray_[i] = (uint8_t*)dalloc(npix);
    d_cen_[i] = (uint64_t*)dalloc(npix * 16);
    d_arms_[i] = (uint16_t*)dalloc(npix * 10);
    d_disp_[i] = (int16_t*)dalloc(npix * 2);
    upload(d_bgr_[i], I_c[i].data, npix * 3);
    upload(d_gray_[i], I_g[i].data, npix);
  }
  d_tmp16_ = (int16_t*)dalloc(npix * 2);
  for (int i = 0; i < 3; i++) d_vol_[i] = (float*)dalloc(nvol * sizeof(float));
}

StereoMatching::~StereoMatching() {
  if (!ctx_) return;
  for (void* p : owned_) sm_dev_free(ctx_, p);
  sm_ctx_destroy(ctx_);
}

void StereoMatching::uploadVm(int i) {
  if (!vm_dev_fresh_[i]) { upload(d_vol_[i], vm[i].data, (size_t)h_ * w_ * d_ * 4); vm_dev_fresh_[i] = true; }
}

void StereoMatching::syncToHost(bool volumes) {
  const size_t npix = (size_t)h_ * w_;
  for (int i = 0; i < 2; i++) {
    if (volumes && vm_dev_fresh_[i]) download(vm[i].data, d_vol_[i], npix * d_ * 4);
    if (dp_dev_fresh_[i]) { DP[i].create(h_, w_, CV_16SC1); download(DP[i].data, d_disp_[i], npix * 2); }
  }
  if (arms_dev_ && arms_host_stale_ && HVL.size() == 2) {
    for (int i = 0; i < 2; i++) download(HVL[i].data, d_arms_[i], npix * 10);
    arms_host_stale_ = false;
  }
}
Mat& StereoMatching::hostDP(int i) {
  if (dp_dev_fresh_[i]) { DP[i].create(h_, w_, CV_16SC1); download(DP[i].data, d_disp_[i], (size_t)h_ * w_ * 2); }
  return DP[i];
}
Mat& StereoMatching::hostVm(int i) {
  if (vm_dev_fresh_[i]) download(vm[i].data, d_vol_[i], (size_t)h_ * w_ * d_ * 4);
  return vm[i];
}

// ------------------------------------------------------------------ stage entry points
void StereoMatching::pipeline() {
  costCalculate();
  dispOptimize();
  refine();
}

void StereoMatching::costCalculate() {
  if (costcalculation == "ADCensus") ADCensusCal();
  else if (costcalculation == "censusGrad") censusGrad(vm);
  else if (costcalculation == "Census") censusCal(vm, 1);
  else if (costcalculation == "AD") asdCal(vm, "AD", Do_LRConsis ? 2 : 1, 1000);
  else throw cv::Exception("costCalculate: costcalculation \"" + costcalculation + "\" is outside the hot path");
  if (aggregation == "CBCA") CBCA();
  else if (aggregation == "NL") NL();
  else if (!aggregation.empty()) throw cv::Exception("costCalculate: aggregation \"" + aggregation + "\" is outside the hot path");
}

void StereoMatching::dispOptimize() {
  DP[0].create(h_, w_, CV_16SC1);
  DP[1].create(h_, w_, CV_16SC1);
  const int num = (Do_refine && Do_LRConsis) ? 2 : 1;
  if (optimization == "sgm")
    for (int i = 0; i < num; i++) sgm(vm[i], i == 0);
  else if (!optimization.empty())
    throw cv::Exception("dispOptimize: optimization \"" + optimization + "\" is outside the hot path");
  if (param_.Do_vmTop) {   // stereoMatching.cpp:1111-1121
    int sizeVmTop[] = {h_, w_, param_.vmTop_Num + 1, 2};
    Mat topDisp(4, sizeVmTop, CV_32F);
    for (int i = 0; i < num; i++) {
      Mat vm_copy = hostVm(i).clone();
      selectTopCostFromVolumn(vm_copy, topDisp, param_.vmTop_thres);
      genDispFromTopCostVm2(topDisp, DP[i]);
    }
  } else
    for (int i = 0; i < num; i++) gen_dispFromVm(vm[i], DP[i]);
}

void StereoMatching::refine() {
  if (!Do_refine) return;
  if (Do_LRConsis) LRConsistencyCheck_normal(DP[0], DP[1], LRC_Err_Mask);
  if (Do_regionVote) {
    if (!param_.has_initArm) initArm();
    if (!param_.has_calArms)
      calArms<uchar>(I_c, HVL, HVL_INTERSECTION, param_.cbca_crossL[0], param_.cbca_crossL_out[0], param_.cbca_cTresh[0],
                     param_.cbca_cTresh_out[0]);
    const float rv_ratio[] = {0.4f, 0.4f, 0.4f, 0.4f};
    const int rv_s[] = {20, 20, 20, 20};
    for (int i = 0; i < param_.region_vote_nums; i++) regionVote_my(DP[0], rv_ratio[i & 3], rv_s[i & 3]);
  }
  if (Do_properIpol)
    for (int i = 0; i < param_.region_vote_nums; i++) properIpol(DP[0], I_c[0]);
  if (Do_WM) {   // stereoMatching.cpp:1467-1471.  The mask is the labelling check's (LRConsistencyCheck); the check refine()
    // itself runs, LRConsistencyCheck_normal, never creates it -- the reference would read an empty Mat here
    CV_Assert(!LRC_Err_Mask.empty());
    WM(DP[0], LRC_Err_Mask, I_c[0]);
  }
  if (Do_discontinuityAdjust) discontinuityAdjust(DP[0]);   // stereoMatching.cpp:1473-1480
  if (Do_subpixelEnhancement) {   // stereoMatching.cpp:1482-1490: SE from DP[0] and vm[0], then its 3x3 median; DP[0] untouched
    const size_t npix = (size_t)h_ * w_;
    if (!dp_dev_fresh_[0]) { upload(d_disp_[0], DP[0].data, npix * 2); dp_dev_fresh_[0] = true; }
    uploadVm(0);
    TmpDev a(ctx_, npix * 4), b(ctx_, npix * 4);
    check(sm_subpixel_enhancement(ctx_, d_disp_[0], d_vol_[0], h_, w_, d_, a.as<float>()), "sm_subpixel_enhancement");
    check(sm_median3_f32(ctx_, a.as<float>(), b.as<float>(), h_, w_), "sm_median3_f32");
    SE.create(h_, w_, CV_32FC1);
    download(SE.data, b.p, npix * 4);
  }
  if (Do_lastMedianBlur) {   // cv::medianBlur(DP[0], DP[0], 3)
    if (!dp_dev_fresh_[0]) { upload(d_disp_[0], DP[0].data, (size_t)h_ * w_ * 2); dp_dev_fresh_[0] = true; }
    check(sm_median3_i16(ctx_, d_disp_[0], d_tmp16_, h_, w_), "sm_median3_i16");
    std::swap(d_disp_[0], d_tmp16_);
  }
}

// ------------------------------------------------------------------ cost
int StereoMatching::HammingDistance(uint64_t c1, uint64_t c2) { return __builtin_popcountll(c1 ^ c2); }

void StereoMatching::ADCensusCal() {
  const int imgNum = Do_LRConsis ? 2 : 1;
  for (int i = 0; i < 2; i++) check(sm_census(ctx_, d_gray_[i], h_, w_, param_.censusFunc, d_cen_[i]), "sm_census");
  for (int i = 0; i < imgNum; i++) {
    check(sm_cost_adcensus(ctx_, d_bgr_[0], d_bgr_[1], d_cen_[0], d_cen_[1], h_, w_, d_, param_.censusFunc, 1000.f, 10.f,
                           30.f, i, d_vol_[i]), "sm_cost_adcensus");
    vm_dev_fresh_[i] = true;
  }
}

void StereoMatching::gen_ad_sd_vm(Mat& asd_vm, int LOR, int AOS, float trunc) {
  CV_Assert(AOS == 0);   // AD; SD is outside the path
  ensure_vol(asd_vm, h_, w_, d_);
  TmpDev t(ctx_, (size_t)h_ * w_ * d_ * 4);
  check(sm_cost_ad(ctx_, d_bgr_[0], d_bgr_[1], h_, w_, d_, LOR, trunc, t.as<float>()), "sm_cost_ad");
  download(asd_vm.data, t.p, (size_t)h_ * w_ * d_ * 4);
}

void StereoMatching::asdCal(vector<Mat>& vm_asd, string method, int imgNum, float Trunc) {
  CV_Assert(method == "AD");
  CV_Assert((int)vm_asd.size() >= imgNum);
  for (int i = 0; i < imgNum; i++) {
    if (&vm_asd == &vm) {
      check(sm_cost_ad(ctx_, d_bgr_[0], d_bgr_[1], h_, w_, d_, i, Trunc, d_vol_[i]), "sm_cost_ad");
      vm_dev_fresh_[i] = true;
    } else {
      gen_ad_sd_vm(vm_asd[i], i, 0, Trunc);
    }
  }
}

template <typename T>
void StereoMatching::genCensusCode(vector<Mat>& I, vector<Mat>& census, int R_V, int R_U) {
  static_assert(sizeof(T) == 1, "the path uses 8-bit gray images");
  CV_Assert(R_V == 3 && R_U == 4 && I.size() >= 2 && census.size() >= 2);   // censusCal's window, stereoMatching.cpp:815
  for (int i = 0; i < 2; i++) {
    CV_Assert(I[i].type() == CV_8UC1 && I[i].rows == h_ && I[i].cols == w_);
    ensure_vol(census[i], h_, w_, 1, CV_64F);
    TmpDev g(ctx_, (size_t)h_ * w_), c(ctx_, (size_t)h_ * w_ * 8);
    upload(g.p, I[i].data, (size_t)h_ * w_);
    check(sm_census(ctx_, g.as<uint8_t>(), h_, w_, 0, c.as<uint64_t>()), "sm_census");
    download(census[i].data, c.p, (size_t)h_ * w_ * 8);
  }
}
template void StereoMatching::genCensusCode<uchar>(vector<Mat>&, vector<Mat>&, int, int);

void StereoMatching::genCensusCode_NC_Sur(vector<Mat>& I, vector<Mat>& census, int R_V, int R_U) {
  CV_Assert(R_V == 3 && R_U == 4 && I.size() >= 2 && census.size() >= 2);
  for (int i = 0; i < 2; i++) {
    CV_Assert(I[i].type() == CV_8UC1 && I[i].rows == h_ && I[i].cols == w_);
    ensure_vol(census[i], h_, w_, 2, CV_64F);
    TmpDev g(ctx_, (size_t)h_ * w_), c(ctx_, (size_t)h_ * w_ * 16);
    upload(g.p, I[i].data, (size_t)h_ * w_);
    check(sm_census(ctx_, g.as<uint8_t>(), h_, w_, 3, c.as<uint64_t>()), "sm_census");
    download(census[i].data, c.p, (size_t)h_ * w_ * 16);
  }
}

void StereoMatching::gen_cenVM_XOR(vector<Mat>& census, Mat& cenVm, int codeLength, float truncRat, int LOR) {
  CV_Assert(census.size() >= 2 && truncRat == 1.0f && (codeLength == 63 || codeLength == 71));
  const int func = codeLength == 71 ? 3 : 0, nw = sm_census_words(func);
  CV_Assert(census[0].total() == (size_t)h_ * w_ * nw && census[1].total() == (size_t)h_ * w_ * nw);
  ensure_vol(cenVm, h_, w_, d_);
  const size_t cb = (size_t)h_ * w_ * nw * 8, vb = (size_t)h_ * w_ * d_ * 4;
  TmpDev cl(ctx_, cb), cr(ctx_, cb), v(ctx_, vb);
  upload(cl.p, census[0].data, cb);
  upload(cr.p, census[1].data, cb);
  check(sm_cost_hamming(ctx_, cl.as<uint64_t>(), cr.as<uint64_t>(), h_, w_, d_, func, LOR, v.as<float>()), "sm_cost_hamming");
  download(cenVm.data, v.p, vb);
}

void StereoMatching::censusCal(vector<Mat>& vm_census, float truncRatio) {
  CV_Assert(truncRatio == 1.0f);
  const int imgNum = Do_LRConsis ? 2 : 1;
  CV_Assert((int)vm_census.size() >= imgNum);
  for (int i = 0; i < 2; i++) check(sm_census(ctx_, d_gray_[i], h_, w_, param_.censusFunc, d_cen_[i]), "sm_census");
  for (int i = 0; i < imgNum; i++) {
    if (&vm_census == &vm) {
      check(sm_cost_hamming(ctx_, d_cen_[0], d_cen_[1], h_, w_, d_, param_.censusFunc, i, d_vol_[i]), "sm_cost_hamming");
      vm_dev_fresh_[i] = true;
    } else {
      ensure_vol(vm_census[i], h_, w_, d_);
      check(sm_cost_hamming(ctx_, d_cen_[0], d_cen_[1], h_, w_, d_, param_.censusFunc, i, d_vol_[2]), "sm_cost_hamming");
      download(vm_census[i].data, d_vol_[2], (size_t)h_ * w_ * d_ * 4);
    }
  }
}

void StereoMatching::gen_vm_from2vm_exp(cv::Mat& combinedVm, cv::Mat& vm0, cv::Mat& vm1, const float ARU0, const float ARU1,
                                        int LOR) {
  (void)LOR;
  const size_t n = (size_t)h_ * w_ * d_;
  CV_Assert(vm0.total() * vm0.channels() == n && vm1.total() * vm1.channels() == n);
  TmpDev a(ctx_, n * 4), b(ctx_, n * 4), o(ctx_, n * 4);
  upload(a.p, vm0.data, n * 4);
  upload(b.p, vm1.data, n * 4);
  check(sm_combine_exp(ctx_, a.as<float>(), b.as<float>(), n, ARU0, ARU1, o.as<float>()), "sm_combine_exp");
  int idx = &combinedVm == &vm[0] ? 0 : (&combinedVm == &vm[1] ? 1 : -1);
  if (idx < 0) ensure_vol(combinedVm, h_, w_, d_);
  download(combinedVm.data, o.p, n * 4);
  if (idx >= 0) vm_dev_fresh_[idx] = false;
}

void StereoMatching::adCensus(vector<Mat>& vm_ad, vector<Mat>& vm_census) {
  const int imgNum = Do_LRConsis ? 2 : 1;
  for (int i = 0; i < imgNum; i++) gen_vm_from2vm_exp(vm[i], vm_ad[i], vm_census[i], 10, 30, i);   // :5270
}

// ------------------------------------------------------------------ aggregation
// ---- gradient cost family (stereoMatching.cpp:25-48, 271-368, 388-455, 603-656)
void StereoMatching::ensureGrad() {
  for (int i = 0; i < 2; i++) {
    for (int k = 0; k < 2; k++)
      if (!d_grad_[i][k]) d_grad_[i][k] = (float*)dalloc((size_t)h_ * w_ * sizeof(float));
    check(sm_grad_xy(ctx_, d_gray_[i], h_, w_, d_grad_[i][0], d_grad_[i][1]), "sm_grad_xy");
  }
}
void StereoMatching::calGrad(Mat& grad, Mat& img) {
  CV_Assert(img.channels() == 1 && img.rows == h_ && img.cols == w_);   // the path feeds I_g (stereoMatching.cpp:620)
  if (grad.empty()) grad.create(h_, w_, CV_32F);
  const size_t n = (size_t)h_ * w_;
  TmpDev g(ctx_, n), gx(ctx_, n * 4), gy(ctx_, n * 4);
  upload(g.p, img.data, n);
  check(sm_grad_xy(ctx_, g.as<uint8_t>(), h_, w_, gx.as<float>(), gy.as<float>()), "sm_grad_xy");
  download(grad.data, gx.p, n * 4);
}
void StereoMatching::calGrad_y(Mat& grad, Mat& img) {
  CV_Assert(img.channels() == 1 && img.rows == h_ && img.cols == w_);
  if (grad.empty()) grad.create(h_, w_, CV_32F);
  const size_t n = (size_t)h_ * w_;
  TmpDev g(ctx_, n), gx(ctx_, n * 4), gy(ctx_, n * 4);
  upload(g.p, img.data, n);
  check(sm_grad_xy(ctx_, g.as<uint8_t>(), h_, w_, gx.as<float>(), gy.as<float>()), "sm_grad_xy");
  download(grad.data, gy.p, n * 4);
}
void StereoMatching::calgradvm(Mat& vm_, vector<Mat>& grad_, vector<Mat>& grad_y, int num, float Trunc) {
  CV_Assert(grad_.size() >= 2 && grad_y.size() >= 2 && grad_[0].depth() == CV_32F && (num == 0 || num == 1));
  ensureArms();
  ensure_vol(vm_, h_, w_, d_);
  const size_t n = (size_t)h_ * w_;
  TmpDev gx0(ctx_, n * 4), gy0(ctx_, n * 4), gx1(ctx_, n * 4), gy1(ctx_, n * 4), out(ctx_, n * d_ * 4);
  upload(gx0.p, grad_[0].data, n * 4); upload(gx1.p, grad_[1].data, n * 4);
  upload(gy0.p, grad_y[0].data, n * 4); upload(gy1.p, grad_y[1].data, n * 4);
  check(sm_cost_grad(ctx_, gx0.as<float>(), gy0.as<float>(), gx1.as<float>(), gy1.as<float>(), d_arms_[num], h_, w_, d_,
                     Trunc, num, out.as<float>()), "sm_cost_grad");
  download(vm_.data, out.p, n * d_ * 4);
}
void StereoMatching::grad(vector<Mat>& vm_grad, float Trunc) {
  const int imgNum = Do_LRConsis ? 2 : 1;
  CV_Assert((int)vm_grad.size() >= imgNum);
  ensureGrad();
  ensureArms();   // stereoMatching.cpp:628-631
  for (int i = 0; i < imgNum; i++) {
    float* dst = &vm_grad == &vm ? d_vol_[i] : d_vol_[2];
    check(sm_cost_grad(ctx_, d_grad_[0][0], d_grad_[0][1], d_grad_[1][0], d_grad_[1][1], d_arms_[i], h_, w_, d_, Trunc, i,
                       dst), "sm_cost_grad");
    if (&vm_grad == &vm) vm_dev_fresh_[i] = true;
    else { ensure_vol(vm_grad[i], h_, w_, d_); download(vm_grad[i].data, dst, (size_t)h_ * w_ * d_ * 4); }
  }
}
void StereoMatching::censusGrad(vector<Mat>& vm_) {
  const int imgNum = Do_LRConsis ? 2 : 1;
  CV_Assert((int)vm_.size() >= imgNum);
  for (int i = 0; i < 2; i++) check(sm_census(ctx_, d_gray_[i], h_, w_, param_.censusFunc, d_cen_[i]), "sm_census");
  ensureGrad();
  ensureArms();
  for (int i = 0; i < imgNum; i++) {
    float* dst = &vm_ == &vm ? d_vol_[i] : d_vol_[2];
    check(sm_cost_censusgrad(ctx_, d_cen_[0], d_cen_[1], d_grad_[0][0], d_grad_[0][1], d_grad_[1][0], d_grad_[1][1],
                             d_arms_[i], h_, w_, d_, param_.censusFunc, (float)param_.lamCen, (float)param_.lamG, 500.f, i,
                             dst), "sm_cost_censusgrad");
    if (&vm_ == &vm) vm_dev_fresh_[i] = true;
    else { ensure_vol(vm_[i], h_, w_, d_); download(vm_[i].data, dst, (size_t)h_ * w_ * d_ * 4); }
  }
}

void StereoMatching::initArm() {
  HVL_num = 2;
  HVL.resize(2);
  HVL_INTERSECTION.resize(2);   // headers only: the 2 x H*W*D*5 u16 tensors are evaluated on the fly by the kernels
  int sz[3] = {h_, w_, 5};
  for (int i = 0; i < 2; i++) HVL[i].create(3, sz, CV_16UC1);
  param_.has_initArm = 1;
}

template <typename T>
void StereoMatching::calHorVerDis(Mat& I, Mat& cross, int L, int L_out, int C_D, int C_D_out, int minL) {
  static_assert(sizeof(T) == 1, "8-bit images");
  CV_Assert(I.type() == CV_8UC3);
  const size_t npix = (size_t)I.rows * I.cols;
  int sz[3] = {I.rows, I.cols, 5};
  cross.create(3, sz, CV_16UC1);
  TmpDev img(ctx_, npix * 3), a(ctx_, npix * 10);
  upload(img.p, I.data, npix * 3);
  check(sm_arms(ctx_, img.as<uint8_t>(), I.rows, I.cols, L, L_out, C_D, C_D_out, minL, a.as<uint16_t>()), "sm_arms");
  download(cross.data, a.p, npix * 10);
}
template void StereoMatching::calHorVerDis<uchar>(Mat&, Mat&, int, int, int, int, int);

template <typename T>
void StereoMatching::calArms(vector<Mat>& I, vector<Mat>& cross, vector<Mat>& cross_intersec, int L, int L_out, int cTresh,
                             int cTresh_out) {
  (void)cross_intersec;
  const int scale = param_.disSc > 1 ? param_.disSc : 1;
  if (&I == &I_c && &cross == &HVL) {
    for (int i = 0; i < 2; i++)
      check(sm_arms(ctx_, d_bgr_[i], h_, w_, L / scale, L_out / scale, cTresh, cTresh_out, param_.cbca_minArmL, d_arms_[i]),
            "sm_arms");
    arms_dev_ = true; arms_host_stale_ = true;
  } else {
    for (int i = 0; i < 2; i++) calHorVerDis<T>(I[i], cross[i], L / scale, L_out / scale, cTresh, cTresh_out, param_.cbca_minArmL);
  }
  param_.has_calArms = 1;
}
template void StereoMatching::calArms<uchar>(vector<Mat>&, vector<Mat>&, vector<Mat>&, int, int, int, int);

void StereoMatching::ensureArms() {
  if (!param_.has_initArm) initArm();
  if (!param_.has_calArms || !arms_dev_) {
    if (param_.has_calArms && !arms_dev_) {   // arms were computed into caller-visible HVL on the host: upload
      for (int i = 0; i < 2; i++) upload(d_arms_[i], HVL[i].data, (size_t)h_ * w_ * 10);
      arms_dev_ = true;
    } else {
      calArms<uchar>(I_c, HVL, HVL_INTERSECTION, param_.cbca_crossL[0], param_.cbca_crossL_out[0], param_.cbca_cTresh[0],
                     param_.cbca_cTresh_out[0]);
    }
  }
}

void StereoMatching::genTrueHorVerArms(vector<Mat>& HVL_, vector<Mat>& HVL_IS) {
  // materialises the reference's per-(pixel, d) intersection tensors on request (inspection / tests)
  CV_Assert(HVL_.size() >= 2);
  HVL_IS.resize(2);
  const size_t npix = (size_t)h_ * w_;
  TmpDev al(ctx_, npix * 10), ar(ctx_, npix * 10), o(ctx_, npix * d_ * 10);
  if (&HVL_ == &HVL && arms_dev_ && arms_host_stale_) syncToHost(false);
  upload(al.p, HVL_[0].data, npix * 10);
  upload(ar.p, HVL_[1].data, npix * 10);
  int sz[4] = {h_, w_, d_, 5};
  for (int view = 0; view < 2; view++) {
    HVL_IS[view].create(4, sz, CV_16UC1);
    check(sm_arms_intersect(ctx_, al.as<uint16_t>(), ar.as<uint16_t>(), h_, w_, d_, view, o.as<uint16_t>()), "sm_arms_intersect");
    download(HVL_IS[view].data, o.p, npix * d_ * 10);
  }
}

void StereoMatching::gen1DCumu(cv::Mat& vm_, cv::Mat& area, Mat& areaIS, int dv, int du) {
  (void)area;   // cbca_intersect = true: the per-disparity areaIS is the one in use (stereoMatching.h:262)
  const size_t n = (size_t)h_ * w_ * d_;
  CV_Assert(vm_.depth() == CV_32F && vm_.total() * vm_.channels() == n);
  CV_Assert(areaIS.depth() == CV_32S && areaIS.total() * areaIS.channels() == n);
  TmpDev v(ctx_, n * 4), a(ctx_, n * 4);
  upload(v.p, vm_.data, n * 4); upload(a.p, areaIS.data, n * 4);
  check(sm_cumsum_1d(ctx_, v.as<float>(), a.as<int32_t>(), h_, w_, d_, dv, du), "sm_cumsum_1d");
  download(vm_.data, v.p, n * 4); download(areaIS.data, a.p, n * 4);
}
void StereoMatching::cal1DCost(Mat& vm_, cv::Mat& HVL_, cv::Mat& area, Mat& areaIS, Mat& HVL_IS, int dv, int du, int direc) {
  (void)area;
  CV_Assert(HVL_.depth() == CV_16U);   // stereoMatching.h:1647
  const size_t n = (size_t)h_ * w_ * d_;
  CV_Assert(vm_.depth() == CV_32F && vm_.total() * vm_.channels() == n);
  CV_Assert(areaIS.depth() == CV_32S && areaIS.total() * areaIS.channels() == n);
  CV_Assert(HVL_IS.depth() == CV_16U && HVL_IS.total() * HVL_IS.channels() == n * 5);
  TmpDev v(ctx_, n * 4), a(ctx_, n * 4), tv(ctx_, n * 4), ta(ctx_, n * 4), is(ctx_, n * 10);
  upload(v.p, vm_.data, n * 4); upload(a.p, areaIS.data, n * 4); upload(is.p, HVL_IS.data, n * 10);
  check(sm_span_1d(ctx_, v.as<float>(), a.as<int32_t>(), is.as<uint16_t>(), tv.as<float>(), ta.as<int32_t>(), h_, w_, d_, dv,
                   du, direc), "sm_span_1d");
  download(vm_.data, v.p, n * 4); download(areaIS.data, a.p, n * 4);
}
void StereoMatching::genfinalVm_cbca(Mat& vm_, Mat& area, Mat& areaIS, int imgNum) {
  (void)area; (void)imgNum;
  CV_Assert(vm_.depth() == CV_32F);   // stereoMatching.cpp:3971
  const size_t n = (size_t)h_ * w_ * d_;
  CV_Assert(vm_.total() * vm_.channels() == n && areaIS.depth() == CV_32S && areaIS.total() * areaIS.channels() == n);
  TmpDev v(ctx_, n * 4), a(ctx_, n * 4);
  upload(v.p, vm_.data, n * 4); upload(a.p, areaIS.data, n * 4);
  check(sm_div_area(ctx_, v.as<float>(), a.as<int32_t>(), n), "sm_div_area");
  download(vm_.data, v.p, n * 4);
}
template <typename T>
void StereoMatching::updateCost(cv::Mat& Lr, cv::Mat& vm_, int v, int u, int n, int rv, int ru, bool preIsInner, bool leftFirst) {
  static_assert(sizeof(T) == sizeof(float), "the ctor only ever allocates CV_32F volumes (stereoMatching.cpp:2080)");
  const size_t nv = (size_t)h_ * w_ * n;
  CV_Assert(vm_.depth() == CV_32F && Lr.depth() == CV_32F && n == d_);
  CV_Assert(vm_.total() * vm_.channels() == nv && Lr.total() * Lr.channels() == nv);
  TmpDev l(ctx_, nv * 4), c(ctx_, nv * 4);
  upload(l.p, Lr.data, nv * 4); upload(c.p, vm_.data, nv * 4);
  check(sm_update_cost(ctx_, l.as<float>(), c.as<float>(), d_bgr_[leftFirst ? 0 : 1], h_, w_, n, v, u, rv, ru, preIsInner ? 1 : 0,
                       param_.sgm_corDifThres, param_.sgm_reduCoeffi1), "sm_update_cost");
  download(Lr.data, l.p, nv * 4);
}
template void StereoMatching::updateCost<float>(cv::Mat&, cv::Mat&, int, int, int, int, int, bool, bool);

template <typename T>
void StereoMatching::calErr(Mat& DP_, Mat& DT_, string procedure, bool calCSV) {
  (void)calCSV;
  static_assert(sizeof(T) == sizeof(short), "the path evaluates CV_16S disparity maps (stereoMatching.cpp:1128)");
  CV_Assert(DP_.depth() == CV_16S && DT_.depth() == CV_32F && DP_.rows == h_ && DP_.cols == w_ && DT_.rows == h_ && DT_.cols == w_);
  const size_t n = (size_t)h_ * w_;
  TmpDev d(ctx_, n * 2), g(ctx_, n * 4), m(ctx_, n);
  upload(d.p, DP_.data, n * 2); upload(g.p, DT_.data, n * 4);
  static const char* names[3] = {"nonocc", "all", "disc"};
  for (int region = 0; region < 3; region++) {
    lastErr[region] = ErrPair();
    if (region >= (int)I_mask.size() || I_mask[region].empty()) continue;
    CV_Assert(I_mask[region].depth() == CV_8U && I_mask[region].rows == h_ && I_mask[region].cols == w_);
    upload(m.p, I_mask[region].data, n);
    long long sumNum = 0, errorNumer = 0;
    double esum = 0.0;
    check(sm_cal_err(ctx_, d.as<int16_t>(), g.as<float>(), m.as<uint8_t>(), h_, w_, param_.errorThreshold, &sumNum, &errorNumer,
                     &esum), "sm_cal_err");
    lastErr[region].PBM = (float)errorNumer / sumNum;
    lastErr[region].RMS = std::sqrt((float)esum / sumNum);
    lastErr[region].valid = true;
    std::cout << std::endl << names[region] << "\terrorRatio: " << lastErr[region].PBM << " epe: " << lastErr[region].RMS
              << " " + procedure << std::endl;   // stereoMatching.h:1797
  }
}
template void StereoMatching::calErr<short>(Mat&, Mat&, string, bool);

template <typename T, int imgNum>
void StereoMatching::saveDispMap(const cv::Mat& dispM, const Mat& trueM, string method, bool calErr) {
  static_assert(sizeof(T) == sizeof(short), "the path saves CV_16S disparity maps");
  CV_Assert(dispM.depth() == CV_16S && dispM.channels() == 1 && dispM.rows > 0 && dispM.cols > 0);
  const int h = dispM.rows, w = dispM.cols;
  // "IF NOT EXIST path (mkdir path)" (stereoMatching.h:2080): every missing component of savePath
  for (size_t i = 1; i <= param_.savePath.size(); i++)
    if (i == param_.savePath.size() || param_.savePath[i] == '/') ::mkdir(param_.savePath.substr(0, i).c_str(), 0777);
  std::vector<uint8_t> bgr;
  std::string err;
  smio::disp_to_bgr((const int16_t*)dispM.data, h, w, param_.DISP_OCC, param_.DISP_MIS, param_.DISP_PKR, bgr, nullptr, nullptr,
                    param_.err_ip_dispV, param_.cor_ip_dispV);
  if (!smio::write_png(param_.savePath + method + ".png", bgr.data(), h, w, 3, &err)) throw cv::Exception("saveDispMap: " + err);
  if (calErr) {
    CV_Assert(trueM.depth() == CV_32F && trueM.rows == h && trueM.cols == w);
    CV_Assert(I_mask.size() > 1 && !I_mask[1].empty() && I_mask[1].rows == h && I_mask[1].cols == w);
    smio::disp_to_bgr((const int16_t*)dispM.data, h, w, param_.DISP_OCC, param_.DISP_MIS, param_.DISP_PKR, bgr,
                      (const float*)trueM.data, I_mask[1].data, param_.err_ip_dispV, param_.cor_ip_dispV);
    if (!smio::write_png(param_.savePath + method + "_err.png", bgr.data(), h, w, 3, &err)) throw cv::Exception("saveDispMap: " + err);
  }
}
template void StereoMatching::saveDispMap<short, 1>(const cv::Mat&, const Mat&, string, bool);

void StereoMatching::LRConsistencyCheck_new(Mat& errorMask) {
  CV_Assert(errorMask.depth() == CV_8U && errorMask.rows == h_ && errorMask.cols == w_);
  const size_t n = (size_t)h_ * w_;
  for (int i = 0; i < 2; i++)
    if (!dp_dev_fresh_[i]) { CV_Assert(!DP[i].empty()); upload(d_disp_[i], DP[i].data, n * 2); }
  TmpDev m(ctx_, n);
  upload(m.p, errorMask.data, n);
  check(sm_lrc_mask(ctx_, d_disp_[0], d_disp_[1], h_, w_, m.as<uint8_t>()), "sm_lrc_mask");
  download(errorMask.data, m.p, n);
}

void StereoMatching::cbca_core(vector<Mat>& HVL_, vector<Mat>& HVL_IS, vector<Mat>& vm_, int ITNUM) {
  (void)HVL_IS;
  const int imgNum = (Do_refine && Do_LRConsis) ? 2 : 1;
  const size_t npix = (size_t)h_ * w_, vb = npix * d_ * 4;
  if (&HVL_ != &HVL) {   // caller-supplied arms
    for (int i = 0; i < 2; i++) upload(d_arms_[i], HVL_[i].data, npix * 10);
    arms_dev_ = true; arms_host_stale_ = false;
  } else {
    ensureArms();
  }
  for (int LOR = 0; LOR < imgNum; LOR++) {
    if (&vm_ == &vm) {
      uploadVm(LOR);
      check(sm_cbca(ctx_, d_vol_[LOR], d_vol_[2], d_arms_[0], d_arms_[1], h_, w_, d_, ITNUM, LOR), "sm_cbca");
    } else {
      TmpDev v(ctx_, vb);
      upload(v.p, vm_[LOR].data, vb);
      check(sm_cbca(ctx_, v.as<float>(), d_vol_[2], d_arms_[0], d_arms_[1], h_, w_, d_, ITNUM, LOR), "sm_cbca");
      download(vm_[LOR].data, v.p, vb);
    }
  }
}

void StereoMatching::cbca_aggregate(int param_Num, vector<Mat>& vm_) {
  if (!param_.has_initArm) initArm();
  if (!param_.has_calArms)
    calArms<uchar>(I_c, HVL, HVL_INTERSECTION, param_.cbca_crossL[param_Num], param_.cbca_crossL_out[param_Num],
                   param_.cbca_cTresh[param_Num], param_.cbca_cTresh_out[param_Num]);
  cbca_core(HVL, HVL_INTERSECTION, vm_, param_.cbca_iterationNum);
}

void StereoMatching::CBCA() {
  CV_Assert(!param_.cbca_double_win && param_.cbca_armTile == 0 && param_.cbca_intersect);   // the default chain
  cbca_aggregate(0, vm);
}

void StereoMatching::NL() {
  uploadVm(0);
  check(sm_nl(ctx_, d_bgr_[0], d_vol_[0], h_, w_, d_), "sm_nl");
  // gen_dispFromVm(vm[0], ...) -> guideDisp (stereoMatching.cpp:4913)
  check(sm_wta(ctx_, d_vol_[0], h_, w_, d_, d_tmp16_), "sm_wta");
  Mat tmp(h_, w_, CV_16SC1);
  download(tmp.data, d_tmp16_, (size_t)h_ * w_ * 2);
  guideDisp.create(h_, w_, CV_32FC1);
  for (size_t i = 0; i < (size_t)h_ * w_; i++) guideDisp.ptr<float>()[i] = (float)tmp.ptr<short>()[i];
}

void SolveAll(StereoMatching**& smPyr, const int PY_LVL, const float REG_LAMBDA) {
  CV_Assert(smPyr && PY_LVL >= 1 && PY_LVL <= SM_MAX_PYRAMID);
  StereoMatching* s0 = smPyr[0];
  int Hs[SM_MAX_PYRAMID], Ws[SM_MAX_PYRAMID], Ds[SM_MAX_PYRAMID];
  for (int s = 0; s < PY_LVL; s++) {
    CV_Assert(smPyr[s] != nullptr);
    Hs[s] = smPyr[s]->h_; Ws[s] = smPyr[s]->w_; Ds[s] = smPyr[s]->d_;
    for (int i = 0; i < 2; i++) smPyr[s]->uploadVm(i);
    smPyr[s]->check(sm_ctx_sync(smPyr[s]->ctx_), "sm_ctx_sync");   // each level has its own stream: finish before the gather
  }
  const int img_n = StereoMatching::Do_refine ? 2 : 1;   // stereoMatching.cpp:2179
  for (int n = 0; n < img_n; n++) {
    float* vols[SM_MAX_PYRAMID];
    for (int s = 0; s < PY_LVL; s++) vols[s] = smPyr[s]->d_vol_[n];
    s0->check(sm_cross_scale(s0->ctx_, vols, Hs, Ws, Ds, PY_LVL, REG_LAMBDA), "sm_cross_scale");
    s0->vm_dev_fresh_[n] = true;
  }
}

void pyrDown_u8(const cv::Mat& src, cv::Mat& dst) {
  CV_Assert(src.dims == 2 && src.depth() == CV_8U && (src.channels() == 1 || src.channels() == 3) && !src.empty());
  sm_ctx* c = nullptr;
  if (sm_ctx_create(&c, 0, nullptr) != SM_OK) throw cv::Exception(std::string("sm_ctx_create: ") + sm_last_error());
  const int H = src.rows, W = src.cols, cn = src.channels();
  cv::Mat out((H + 1) / 2, (W + 1) / 2, src.type());
  void *ds = nullptr, *dd = nullptr;
  int rc = sm_dev_alloc(c, &ds, (size_t)H * W * cn);
  if (rc == SM_OK) rc = sm_dev_alloc(c, &dd, out.total() * cn);
  if (rc == SM_OK) rc = sm_memcpy_h2d(c, ds, src.data, (size_t)H * W * cn);
  if (rc == SM_OK) rc = sm_pyr_down_u8(c, (const uint8_t*)ds, H, W, cn, (uint8_t*)dd);
  if (rc == SM_OK) rc = sm_memcpy_d2h(c, out.data, dd, out.total() * cn);
  if (rc == SM_OK) rc = sm_ctx_sync(c);
  const std::string err = rc == SM_OK ? "" : sm_last_error();
  sm_dev_free(c, ds); sm_dev_free(c, dd); sm_ctx_destroy(c);
  if (rc != SM_OK) throw cv::Exception("pyrDown_u8: " + err);
  dst = out;
}

void StereoMatching::SolveAll(int PY_LVL, float REG_LAMBDA) {
  CV_Assert(PY_LVL == 1);   // main_.cpp:132: one pyramid level
  for (int i = 0; i < 2; i++) {
    uploadVm(i);
    check(sm_cross_scale_1level(ctx_, d_vol_[i], (size_t)h_ * w_ * d_, REG_LAMBDA), "sm_cross_scale_1level");
  }
}

// ------------------------------------------------------------------ optimisation / selection
static int sgm_path_index(int rv, int ru) {
  static const int RV[8] = {+1, -1, 0, 0, +1, +1, -1, -1}, RU[8] = {0, 0, +1, -1, -1, +1, +1, -1};
  for (int i = 0; i < 8; i++)
    if (RV[i] == rv && RU[i] == ru) return i;
  return -1;
}

void StereoMatching::sgm(cv::Mat& vm_, bool leftFirst) {
  const int idx = &vm_ == &vm[0] ? 0 : (&vm_ == &vm[1] ? 1 : -1);
  const uint8_t* bgr = leftFirst ? d_bgr_[0] : d_bgr_[1];
  const size_t vb = (size_t)h_ * w_ * d_ * 4;
  L.clear();   // the P path volumes L[i] are summed on the fly; costScan() materialises one on request
  if (idx >= 0) {
    uploadVm(idx);
    check(sm_sgm(ctx_, d_vol_[idx], bgr, h_, w_, d_, sgm_paths_, param_.sgm_corDifThres, param_.sgm_reduCoeffi1, d_vol_[2]),
          "sm_sgm");
    std::swap(d_vol_[idx], d_vol_[2]);
  } else {
    CV_Assert(vm_.depth() == CV_32F && vm_.total() * vm_.channels() == (size_t)h_ * w_ * d_);
    TmpDev v(ctx_, vb);
    upload(v.p, vm_.data, vb);
    check(sm_sgm(ctx_, v.as<float>(), bgr, h_, w_, d_, sgm_paths_, param_.sgm_corDifThres, param_.sgm_reduCoeffi1, d_vol_[2]),
          "sm_sgm");
    download(vm_.data, d_vol_[2], vb);
  }
}

void StereoMatching::costScan(cv::Mat& Lr, cv::Mat& vm_, int rv, int ru, bool leftFirst) {
  const int path = sgm_path_index(rv, ru);
  CV_Assert(path >= 0);
  const size_t vb = (size_t)h_ * w_ * d_ * 4;
  ensure_vol(Lr, h_, w_, d_);
  if (vm_.depth() == CV_8U || vm_.depth() == CV_16U) {
    // the integer-cost entries (stereoMatching.cpp:2007-2014): updateCost<uchar> / <ushort> see the cost as a float
    const size_t nel = (size_t)h_ * w_ * d_, eb = vm_.depth() == CV_8U ? 1 : 2;
    CV_Assert(vm_.total() * vm_.channels() == nel);
    TmpDev raw(ctx_, nel * eb), f(ctx_, vb);
    upload(raw.p, vm_.data, nel * eb);
    check(sm_vol_to_f32(ctx_, raw.p, (int)eb, nel, f.as<float>()), "sm_vol_to_f32");
    check(sm_sgm_path(ctx_, f.as<float>(), leftFirst ? d_bgr_[0] : d_bgr_[1], h_, w_, d_, path, param_.sgm_corDifThres,
                      param_.sgm_reduCoeffi1, 0, d_vol_[2]), "sm_sgm_path");
    download(Lr.data, d_vol_[2], vb);
    return;
  }
  if (vm_.depth() != CV_32F) throw cv::Exception("cost volumn's type is not reasonable");   // stereoMatching.cpp:2019

  const int idx = &vm_ == &vm[0] ? 0 : (&vm_ == &vm[1] ? 1 : -1);
  const float* src;
  TmpDev v(ctx_, idx >= 0 ? 16 : vb);
  if (idx >= 0) { uploadVm(idx); src = d_vol_[idx]; }
  else { upload(v.p, vm_.data, vb); src = v.as<float>(); }
  check(sm_sgm_path(ctx_, src, leftFirst ? d_bgr_[0] : d_bgr_[1], h_, w_, d_, path, param_.sgm_corDifThres,
                    param_.sgm_reduCoeffi1, 0, d_vol_[2]), "sm_sgm_path");
  download(Lr.data, d_vol_[2], vb);
}

void StereoMatching::gen_sgm_vm(Mat& vm_, vector<cv::Mat1f>& Lr, int numOfDirec) {
  CV_Assert((int)Lr.size() >= numOfDirec && numOfDirec >= 1);
  const size_t n = (size_t)h_ * w_ * d_;
  TmpDev x(ctx_, n * 4);
  check(sm_memset(ctx_, d_vol_[2], 0, n * 4), "sm_memset");   // float sum = 0
  for (int k = 0; k < numOfDirec; k++) {
    upload(x.p, Lr[k].data, n * 4);
    check(sm_vol_accumulate(ctx_, d_vol_[2], x.as<float>(), n), "sm_vol_accumulate");   // sum += Lr[k]
  }
  const int idx = &vm_ == &vm[0] ? 0 : (&vm_ == &vm[1] ? 1 : -1);
  if (idx >= 0) { std::swap(d_vol_[idx], d_vol_[2]); vm_dev_fresh_[idx] = true; }
  else download(vm_.data, d_vol_[2], n * 4);
}

void StereoMatching::gen_dispFromVm(Mat& vm_, Mat& dispMap) {
  const int idx = &vm_ == &vm[0] ? 0 : (&vm_ == &vm[1] ? 1 : -1);
  const int didx = &dispMap == &DP[0] ? 0 : (&dispMap == &DP[1] ? 1 : -1);
  const size_t vb = (size_t)h_ * w_ * d_ * 4;
  const float* src;
  TmpDev v(ctx_, idx >= 0 ? 16 : vb);
  if (idx >= 0) { uploadVm(idx); src = d_vol_[idx]; }
  else { upload(v.p, vm_.data, vb); src = v.as<float>(); }
  int16_t* dst = didx >= 0 ? d_disp_[didx] : d_tmp16_;
  check(sm_wta(ctx_, src, h_, w_, d_, dst), "sm_wta");
  if (didx >= 0) dp_dev_fresh_[didx] = true;
  else { dispMap.create(h_, w_, CV_16SC1); download(dispMap.data, dst, (size_t)h_ * w_ * 2); }
}

void StereoMatching::wta_Co(cv::Mat& vm_, cv::Mat& D1, cv::Mat& D2) {
  const int idx = &vm_ == &vm[0] ? 0 : (&vm_ == &vm[1] ? 1 : -1);
  const size_t vb = (size_t)h_ * w_ * d_ * 4, pb = (size_t)h_ * w_ * 2;
  const float* src;
  TmpDev v(ctx_, idx >= 0 ? 16 : vb), a(ctx_, pb), b(ctx_, pb);
  if (idx >= 0) { uploadVm(idx); src = d_vol_[idx]; }
  else { upload(v.p, vm_.data, vb); src = v.as<float>(); }
  check(sm_wta_co(ctx_, src, h_, w_, d_, param_.DISP_SCALE, a.as<int16_t>(), b.as<int16_t>()), "sm_wta_co");
  D1.create(h_, w_, CV_16SC1); D2.create(h_, w_, CV_16SC1);
  download(D1.data, a.p, pb);
  download(D2.data, b.p, pb);
}

// subpixelEnhancement (stereoMatching.cpp:6138-6166): reads vm[0] of the object (device copy brought up to date first).
void StereoMatching::subpixelEnhancement(Mat& disparity, Mat& floatDisp) {
  CV_Assert(disparity.type() == CV_16S);
  CV_Assert(floatDisp.type() == CV_32F);
  CV_Assert(disparity.rows == h_ && disparity.cols == w_ && floatDisp.rows == h_ && floatDisp.cols == w_);
  const size_t npix = (size_t)h_ * w_;
  TmpDev d(ctx_, npix * 2), f(ctx_, npix * 4);
  upload(d.p, disparity.data, npix * 2);
  uploadVm(0);
  check(sm_subpixel_enhancement(ctx_, d.as<int16_t>(), d_vol_[0], h_, w_, d_, f.as<float>()), "sm_subpixel_enhancement");
  download(floatDisp.data, f.p, npix * 4);
}

// selectTopCostFromVolumn (stereoMatching.h:2405-2461): the candidates come from sm_select_top_cost; the reference also
// overwrites the entries it takes in the Mat it is handed (its caller passes a clone, stereoMatching.cpp:1118), which
// is replayed on the host copy from the candidate list.
void StereoMatching::selectTopCostFromVolumn(Mat& vm_, Mat& topDisp, float thres) {
  CV_Assert(topDisp.dims == 4);
  CV_Assert(topDisp.type() == CV_32F);
  CV_Assert(topDisp.size[0] == h_ && topDisp.size[1] == w_ && topDisp.size[2] >= 2 && topDisp.size[3] == 2);
  const int num = topDisp.size[2] - 1;
  const int idx = &vm_ == &vm[0] ? 0 : (&vm_ == &vm[1] ? 1 : -1);
  const size_t vb = (size_t)h_ * w_ * d_ * 4, tb = (size_t)h_ * w_ * (num + 1) * 2 * 4;
  const float* src;
  TmpDev v(ctx_, idx >= 0 ? 16 : vb), t(ctx_, tb);
  if (idx >= 0) { uploadVm(idx); src = d_vol_[idx]; }
  else { upload(v.p, vm_.data, vb); src = v.as<float>(); }
  check(sm_select_top_cost(ctx_, src, h_, w_, d_, num, thres, t.as<float>()), "sm_select_top_cost");
  download(topDisp.data, t.p, tb);
  if (idx >= 0) hostVm(idx);   // the host copy is brought up to date before its taken entries are overwritten
  float* c = vm_.ptr<float>();
  const float* o = topDisp.ptr<float>();
  for (size_t i = 0; i < (size_t)h_ * w_; i++, c += d_, o += (size_t)(num + 1) * 2)
    for (int k = 0; k < (int)o[2 * num]; k++) c[(int)o[2 * k]] = std::numeric_limits<float>::max();
  if (idx >= 0) vm_dev_fresh_[idx] = false;
}

// genDispFromTopCostVm (stereoMatching.h:2466-2545) / genDispFromTopCostVm2 (stereoMatching.cpp:1514-1886): the map from
// the candidate lists; disp keeps its content where a pixel has no candidate, as in the reference.
static void top_to_disp(StereoMatching* self, sm_ctx* ctx, Mat& topDisp, Mat& disp, int version, const uint8_t* d_bgr, int h,
                        int w, const StereoMatching::Parameters& P) {
  CV_Assert(topDisp.dims == 4 && topDisp.type() == CV_32F);
  CV_Assert(topDisp.size[0] == h && topDisp.size[1] == w && topDisp.size[2] >= 2 && topDisp.size[3] == 2);
  CV_Assert(disp.type() == CV_16SC1 && disp.rows == h && disp.cols == w);
  (void)self;
  const int num = topDisp.size[2] - 1;
  const size_t tb = (size_t)h * w * (num + 1) * 2 * 4, pb = (size_t)h * w * 2;
  void *t = nullptr, *d = nullptr;
  auto chk = [&](int rc, const char* what) {
    if (rc != SM_OK) {
      if (t) sm_dev_free(ctx, t);
      if (d) sm_dev_free(ctx, d);
      throw cv::Exception(std::string(what) + ": " + sm_last_error());
    }
  };
  chk(sm_dev_alloc(ctx, &t, tb), "sm_dev_alloc");
  chk(sm_dev_alloc(ctx, &d, pb), "sm_dev_alloc");
  chk(sm_memcpy_h2d(ctx, t, topDisp.data, tb), "sm_memcpy_h2d");
  chk(sm_memcpy_h2d(ctx, d, disp.data, pb), "sm_memcpy_h2d");
  if (version == 1) chk(sm_disp_from_top(ctx, (const float*)t, h, w, num, (int16_t*)d), "sm_disp_from_top");
  else
    chk(sm_disp_from_top2(ctx, (const float*)t, d_bgr, h, w, num, P.vmTop_method, P.ts, P.vmTop_hasCir2 ? 1 : 0,
                          P.vmTop_cir3_doColorLimit ? 1 : 0, (int16_t*)d), "sm_disp_from_top2");
  chk(sm_memcpy_d2h(ctx, disp.data, d, pb), "sm_memcpy_d2h");
  chk(sm_ctx_sync(ctx), "sm_ctx_sync");
  sm_dev_free(ctx, t);
  sm_dev_free(ctx, d);
}
void StereoMatching::genDispFromTopCostVm(Mat& topDisp, Mat& disp) {
  const int i = &disp == &DP[0] ? 0 : (&disp == &DP[1] ? 1 : -1);
  if (i >= 0 && dp_dev_fresh_[i]) hostDP(i);
  top_to_disp(this, ctx_, topDisp, disp, 1, d_bgr_[0], h_, w_, param_);
  if (i >= 0) dp_dev_fresh_[i] = false;
}
void StereoMatching::genDispFromTopCostVm2(Mat& topDisp, Mat& disp) {
  const int i = &disp == &DP[0] ? 0 : (&disp == &DP[1] ? 1 : -1);
  if (i >= 0 && dp_dev_fresh_[i]) hostDP(i);
  top_to_disp(this, ctx_, topDisp, disp, 2, d_bgr_[0], h_, w_, param_);
  if (i >= 0) dp_dev_fresh_[i] = false;
}

// ------------------------------------------------------------------ refinement
void StereoMatching::LRConsistencyCheck_normal(cv::Mat& D1, cv::Mat& D2, cv::Mat& errMask, int LOR) {
  (void)errMask;
  if (LOR != 0) return;   // the reference's body is `if (LOR == 0) {...}` and nothing else (stereoMatching.cpp:2262-2282)
  if (&D1 == &DP[0] && &D2 == &DP[1]) {
    for (int i = 0; i < 2; i++)
      if (!dp_dev_fresh_[i]) { upload(d_disp_[i], DP[i].data, (size_t)h_ * w_ * 2); dp_dev_fresh_[i] = true; }
    check(sm_lrc(ctx_, d_disp_[0], d_disp_[1], h_, w_, param_.LRmaxDiff), "sm_lrc");
  } else {
    const size_t pb = (size_t)h_ * w_ * 2;
    TmpDev a(ctx_, pb), b(ctx_, pb);
    upload(a.p, D1.data, pb); upload(b.p, D2.data, pb);
    check(sm_lrc(ctx_, a.as<int16_t>(), b.as<int16_t>(), h_, w_, param_.LRmaxDiff), "sm_lrc");
    download(D1.data, a.p, pb);
  }
}

// LOR 0 labels D1 against D2, LOR 1 labels D2 against D1 (stereoMatching.cpp:2284-2364).  On the LOR 1 branch the
// reference leaves errMask all zero (its flags go to a local mask that is only dumped to LR1.png; no file is written here).
void StereoMatching::LRConsistencyCheck(cv::Mat& D1, cv::Mat& D2, cv::Mat& errMask, int LOR) {
  CV_Assert(LOR == 0 || LOR == 1);
  const size_t pb = (size_t)h_ * w_ * 2;
  TmpDev a(ctx_, pb), b(ctx_, pb), m(ctx_, (size_t)h_ * w_);
  const bool member = &D1 == &DP[0] && &D2 == &DP[1];
  if (member && dp_dev_fresh_[0]) hostDP(0);
  if (member && dp_dev_fresh_[1]) hostDP(1);
  upload(a.p, D1.data, pb); upload(b.p, D2.data, pb);
  check(sm_lrc_label_lor(ctx_, a.as<int16_t>(), b.as<int16_t>(), h_, w_, d_, param_.LRmaxDiff, param_.DISP_OCC,
                         param_.DISP_MIS, LOR, m.as<uint8_t>(), nullptr), "sm_lrc_label_lor");
  if (LOR == 0) download(D1.data, a.p, pb);
  else download(D2.data, b.p, pb);
  errMask.create(h_, w_, CV_8UC1);
  download(errMask.data, m.p, (size_t)h_ * w_);
  if (member) dp_dev_fresh_[LOR] = false;
}

// WM (stereoMatching.cpp:7340-7393): 19x19 bilateral weighted median on the mask > 0 pixels, guidance img (3 x 8 bit).
void StereoMatching::WM(Mat& disp, Mat& mask, Mat& img) {
  CV_Assert(disp.type() == CV_16SC1 && mask.type() == CV_8UC1 && img.type() == CV_8UC3);
  const size_t npix = (size_t)h_ * w_;
  const bool member = &disp == &DP[0];
  if (member && dp_dev_fresh_[0]) hostDP(0);
  TmpDev a(ctx_, npix * 2), t(ctx_, npix * 2), m(ctx_, npix), g(ctx_, npix * 3);
  upload(a.p, disp.data, npix * 2);
  upload(m.p, mask.data, npix);
  const uint8_t* d_img = d_bgr_[0];
  if (&img != &I_c[0]) { upload(g.p, img.data, npix * 3); d_img = g.as<uint8_t>(); }
  check(sm_wm(ctx_, a.as<int16_t>(), t.as<int16_t>(), m.as<uint8_t>(), d_img, h_, w_, d_, nullptr), "sm_wm");
  download(disp.data, a.p, npix * 2);
  if (member) dp_dev_fresh_[0] = false;
}

void StereoMatching::discontinuityAdjust(cv::Mat& disp) {
  CV_Assert(disp.type() == CV_16SC1 && disp.rows == h_ && disp.cols == w_);
  const size_t npix = (size_t)h_ * w_;
  uploadVm(0);
  if (&disp == &DP[0]) {   // the member map stays on the device
    if (!dp_dev_fresh_[0]) { upload(d_disp_[0], DP[0].data, npix * 2); dp_dev_fresh_[0] = true; }
    check(sm_discontinuity_adjust(ctx_, d_disp_[0], d_vol_[0], h_, w_, d_, nullptr), "sm_discontinuity_adjust");
    return;
  }
  TmpDev a(ctx_, npix * 2);
  upload(a.p, disp.data, npix * 2);
  check(sm_discontinuity_adjust(ctx_, a.as<int16_t>(), d_vol_[0], h_, w_, d_, nullptr), "sm_discontinuity_adjust");
  download(disp.data, a.p, npix * 2);
}

void StereoMatching::regionVote_my(cv::Mat& Dp, float rv_ratio, int rv_s) {
  CV_Assert(Dp.empty() || Dp.type() == CV_16SC1);
  ensureArms();
  if (&Dp == &DP[0]) {
    if (!dp_dev_fresh_[0]) { upload(d_disp_[0], DP[0].data, (size_t)h_ * w_ * 2); dp_dev_fresh_[0] = true; }
    check(sm_region_vote(ctx_, d_disp_[0], d_tmp16_, d_arms_[0], h_, w_, d_, rv_ratio, rv_s), "sm_region_vote");
  } else {
    const size_t pb = (size_t)h_ * w_ * 2;
    TmpDev a(ctx_, pb);
    upload(a.p, Dp.data, pb);
    check(sm_region_vote(ctx_, a.as<int16_t>(), d_tmp16_, d_arms_[0], h_, w_, d_, rv_ratio, rv_s), "sm_region_vote");
    download(Dp.data, a.p, pb);
  }
}

void StereoMatching::properIpol(cv::Mat& Dp, cv::Mat& I1_c) {
  CV_Assert(Dp.empty() || Dp.type() == CV_16SC1);
  const bool own_img = &I1_c == &I_c[0];
  const size_t pb = (size_t)h_ * w_ * 2;
  TmpDev img(ctx_, own_img ? 16 : (size_t)h_ * w_ * 3);
  const uint8_t* bgr = d_bgr_[0];
  if (!own_img) { upload(img.p, I1_c.data, (size_t)h_ * w_ * 3); bgr = img.as<uint8_t>(); }
  if (&Dp == &DP[0]) {
    if (!dp_dev_fresh_[0]) { upload(d_disp_[0], DP[0].data, pb); dp_dev_fresh_[0] = true; }
    check(sm_proper_ipol(ctx_, d_disp_[0], d_tmp16_, bgr, h_, w_, param_.DISP_OCC), "sm_proper_ipol");
  } else {
    TmpDev a(ctx_, pb);
    upload(a.p, Dp.data, pb);
    check(sm_proper_ipol(ctx_, a.as<int16_t>(), d_tmp16_, bgr, h_, w_, param_.DISP_OCC), "sm_proper_ipol");
    download(Dp.data, a.p, pb);
  }
}
