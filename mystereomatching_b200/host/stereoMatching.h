// stereoMatching.h -- the reference's stage API (class StereoMatching, reference stereoMatching.h:46-2738) for the
// dense-stereo hot path, with the bodies replaced by calls into the sm_b200 C ABI (include/sm_b200.h).
//
// Same class name, constructor, static algorithm-selecting strings, `Parameters` field names and defaults, method
// names, argument meaning and public state (`vm`, `HVL`, `HVL_INTERSECTION`, `L`, `DP`, `LRC_Err_Mask`,
// `guideDisp`) as the reference, restricted to the methods on the path.  A caller written
// against the reference (main_.cpp:138-166) compiles against this header unchanged.
//
// Execution model.  Every method is a thin host wrapper: the instance owns one sm_ctx and a set of DEVICE buffers
// that mirror the public cv::Mat members.  Stage methods compute on the device and mark the mirrored host Mat
// stale; a host Mat is refreshed from the device only when the caller asks for it (`syncToHost()` or the `host*()`
// accessors), because copying a 2 GB volume over PCIe after every stage would dominate the run time.  Methods that
// take explicit cv::Mat arguments (gen_ad_sd_vm, gen_cenVM_XOR, gen1DCumu-level helpers ...) upload their inputs and
// download their outputs: exact drop-ins, priced accordingly.  There is no CPU fallback: without a CUDA device
// the constructor throws cv::Exception.
//
// Error convention: the reference's methods return void and raise cv::Exception through CV_Assert
// (stereoMatching.h:941-942, stereoMatching.cpp:3931-3932); so do these -- every non-zero sm_status is turned
// into a cv::Exception carrying sm_last_error().
#pragma once
#ifndef SM_USE_OPENCV
#include "cvmat_lite.h"
#else
#include <opencv2/core.hpp>
#endif
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/sm_b200.h"

using cv::Mat;
using std::string;
using std::vector;

class StereoMatching {
 public:
  // algorithm selection lives in mutable statics, as in the reference (stereoMatching.h:50-54, main_.cpp:15-19)
  static const std::string root;
  static std::string costcalculation;  // "ADCensus" | "censusGrad" | "Census" | "AD"
  static std::string aggregation;      // "CBCA" | "NL" | ""
  static std::string optimization;     // "sgm" | ""
  static std::string object;

  // step switches (stereoMatching.h:57-83); only the ones the path honours
  static bool Do_refine, Do_LRConsis, Do_regionVote, Do_properIpol, Do_lastMedianBlur;
  static bool Do_subpixelEnhancement;   // stereoMatching.h:79 (0 in the reference; a run-time switch here)
  static bool Do_WM;                    // stereoMatching.h:74 (0 in the reference; a run-time switch here)
  static bool Do_discontinuityAdjust;   // stereoMatching.h:78 (0 in the reference; a run-time switch here)
  static const bool UniqCk = false, SubIpl = false;

  struct Parameters {  // stereoMatching.h:85-351 (fields read on the hot path; same names, same defaults)
    int W_U, W_V;
    int numDisparities;
    float LRmaxDiff;
    int DISP_INV, DISP_OCC, DISP_MIS, DISP_PKR, DISP_SCALE, DISP_SHIFT;
    bool ChooseSmall;
    int errorThreshold;
    int SD_AD_channel, census_channel;
    int sgm_scanNum;
    float sgm_P1, sgm_P2;
    int sgm_corDifThres, sgm_reduCoeffi1, sgm_reduCoeffi2;
    int censusFunc;
    int is_censusNorm, is_adNorm;
    int has_initArm, has_calArms;
    uchar cbca_minArmL;
    int cbca_iterationNum;
    bool cbca_intersect;
    int cbca_crossL[2], cbca_crossL_out[2], cbca_cTresh[2], cbca_cTresh_out[2];
    bool cbca_double_win;
    int cbca_armHV, cbca_armTile;
    int region_vote_nums;
    int regVote_SThres;
    float regVote_hratioThres;
    bool Do_vmTop;
    int vmTop_method, vmTop_Num;       // stereoMatching.h:183-184 (0, M)
    float vmTop_thres;                 // :185 (lamc * 0.01)
    int vmTop_thres_dirNum;            // :186 (8; read by no live code)
    bool vmTop_hasCir2, vmTop_cir3_doColorLimit;   // :187-188 (true, false)
    int lamCen, lamG, M, lamc, ts, disSc;
    std::string errCsvName;
    int err_ip_dispV, cor_ip_dispV;    // stereoMatching.h:316-317 (-50, -100): marker values saveDispMap colours
    std::string savePath;              // stereoMatching.h:348: root + object + "/" + method names + "/20200627_test_so/"

    Parameters(int maxDisp, int h, int w, int lamCen_, int lamG_, int M_, int lamc_, int ts_, string errCsvName_,
               int disSc_);
  };

  StereoMatching(cv::Mat& I1_c, cv::Mat& I2_c, cv::Mat& I1_g, cv::Mat& I2_g, cv::Mat& DT, cv::Mat& all_mask,
                 cv::Mat& nonocc_mask, cv::Mat& disc_mask, const Parameters& param);
  ~StereoMatching();
  StereoMatching(const StereoMatching&) = delete;
  StereoMatching& operator=(const StereoMatching&) = delete;

  // ---- the four stage entry points (stereoMatching.h:401-410)
  void pipeline();        // stereoMatching.cpp:1950-1981
  void costCalculate();   // stereoMatching.cpp:945-1021
  void dispOptimize();    // stereoMatching.cpp:1046-1136
  void refine();          // stereoMatching.cpp:1364-1506

  // ---- cost computation
  void asdCal(vector<Mat>& vm_asd, string method, int imgNum, float Trunc);             // stereoMatching.cpp:72-88
  void censusCal(vector<Mat>& vm_census, float truncRatio);                             // stereoMatching.cpp:807-892
  void ADCensusCal();                                                                   // stereoMatching.cpp:894-915
  void adCensus(vector<Mat>& vm_ad, vector<Mat>& vm_census);                            // stereoMatching.cpp:5250-5277
  int HammingDistance(uint64_t c1, uint64_t c2);                                        // stereoMatching.cpp:2210
  void gen_ad_sd_vm(Mat& asd_vm, int LOR, int AOS, float trunc = 1000000);              // stereoMatching.cpp:2468-2509
  template <typename T>
  void genCensusCode(vector<Mat>& I, vector<Mat>& census, int R_V, int R_U);            // stereoMatching.h:634-688
  void genCensusCode_NC_Sur(vector<Mat>& I, vector<Mat>& census, int R_V, int R_U);     // stereoMatching.h:867-934
  void gen_cenVM_XOR(vector<Mat>& census, Mat& cenVm, int codeLength, float truncRat, int LOR = 0);  // :936-981
  void gen_vm_from2vm_exp(cv::Mat& combinedVm, cv::Mat& vm0, cv::Mat& vm1, const float ARU0, const float ARU1,
                          int LOR);                                                     // stereoMatching.cpp:3566-3590

  // gradient family ("censusGrad" is the selector main_.cpp:15 compiles in)
  void censusGrad(vector<Mat>& vm);                                                     // stereoMatching.cpp:25-48
  void grad(vector<Mat>& vm_grad, float Trunc);                                         // stereoMatching.cpp:603-656
  void calGrad(Mat& grad, Mat& img);                                                    // stereoMatching.cpp:271-318
  void calGrad_y(Mat& grad, Mat& img);                                                  // stereoMatching.cpp:320-368
  void calgradvm(Mat& vm, vector<Mat>& grad, vector<Mat>& grad_y, int num, float Trunc);  // stereoMatching.cpp:388-455

  // ---- aggregation
  void CBCA();                                                                          // stereoMatching.cpp:4333-4402
  void NL();                                                                            // stereoMatching.cpp:4892-4917
  void cbca_aggregate(int param_Num, vector<Mat>& vm);                                  // stereoMatching.cpp:5668-5690
  // defined and called as (HVL, HVL_INTERSECTION, vm, ITNUM) in the reference (stereoMatching.cpp:5585, 5681)
  void cbca_core(vector<Mat>& HVL, vector<Mat>& HVL_INTERSECTION, vector<Mat>& vm, int ITNUM);
  void initArm();                                                                       // stereoMatching.cpp:5548-5564
  template <typename T>
  void calArms(vector<Mat>& I, vector<Mat>& cross, vector<Mat>& cross_intersec, int L, int L_out, int cTresh,
               int cTresh_out);                                                         // stereoMatching.cpp:5354-5392
  template <typename T>
  void calHorVerDis(Mat& I, Mat& cross, int L, int L_out, int C_D, int C_D_out, int minL);  // :2958-3050
  void genTrueHorVerArms(vector<Mat>& HVL, vector<Mat>& HVL_INTERSECTION);              // stereoMatching.cpp:2794-2845
  // the sub-steps cbca_core composes, on caller-provided Mats (host in, host out; the fused k_cbca_pass is what
  // cbca_core itself runs)
  void gen1DCumu(cv::Mat& vm, cv::Mat& area, Mat& areaIS, int dv, int du);              // stereoMatching.cpp:3896-3926
  void cal1DCost(Mat& vm, cv::Mat& HVL, cv::Mat& area, Mat& areaIS, Mat& HVL_INTERSECTION, int dv, int du,
                 int direc);                                                           // stereoMatching.h:1643-1715
  void genfinalVm_cbca(Mat& vm, Mat& area, Mat& areaIS, int imgNum);                    // stereoMatching.cpp:3969-3992
  void SolveAll(int PY_LVL, float REG_LAMBDA);                                          // member form, 1 level (main_.cpp:158)
  friend void SolveAll(StereoMatching**& smPyr, const int PY_LVL, const float REG_LAMBDA);

  // ---- optimisation / selection
  void sgm(cv::Mat& vm, bool leftFirst = true);                                         // stereoMatching.cpp:6204-6224
  void costScan(cv::Mat& Lr, cv::Mat& vm, int rv, int ru, bool leftFirst);              // stereoMatching.cpp:1983-2029
  void gen_sgm_vm(Mat& vm, vector<cv::Mat1f>& Lr, int numOfDirec);                      // stereoMatching.cpp:2031-2056
  static float min4(float a, float b, float c, float d) { return std::min(std::min(a, b), std::min(c, d)); }
  template <typename T>
  void updateCost(cv::Mat& Lr, cv::Mat& vm, int v, int u, int n, int rv, int ru, bool preIsInner, bool leftFirst);  // stereoMatching.h:2205-2280
  void gen_dispFromVm(Mat& vm, Mat& dispMap);                                           // stereoMatching.cpp:3928-3967
  void wta_Co(cv::Mat& vm, cv::Mat& D1, cv::Mat& D2);                                   // stereoMatching.cpp:2709-2792
  // topDisp: 4-D CV_32F {h, w, num + 1, 2}; like the reference, the Mat handed in has its taken entries set to FLT_MAX
  void selectTopCostFromVolumn(Mat& vm, Mat& topDisp, float thres);                     // stereoMatching.h:2405-2461
  void subpixelEnhancement(Mat& disparity, Mat& floatDisp);                             // stereoMatching.cpp:6138-6166
  void genDispFromTopCostVm(Mat& topDisp, Mat& disp);                                   // stereoMatching.h:2466-2545
  void genDispFromTopCostVm2(Mat& topDisp, Mat& disp);                                  // stereoMatching.cpp:1514-1886

  // ---- refinement
  void LRConsistencyCheck(cv::Mat& D1, cv::Mat& D2, cv::Mat& errMask, int LOR = 0);         // :2284-2364
  void LRConsistencyCheck_normal(cv::Mat& D1, cv::Mat& D2, cv::Mat& errMask, int LOR = 0);  // :2262-2282
  void LRConsistencyCheck_new(Mat& errorMask);                                              // :2367-2382 (reads DP[0], DP[1])
  void regionVote_my(cv::Mat& Dp, float rv_ratio, int rv_s);                            // stereoMatching.cpp:7219-7277
  void properIpol(cv::Mat& Dp, cv::Mat& I1_c);                                          // stereoMatching.cpp:7395-7490
  void WM(Mat& disp, Mat& mask, Mat& img);                                              // stereoMatching.cpp:7340-7393
  void discontinuityAdjust(cv::Mat& ipol);                                              // stereoMatching.cpp:6057-6135 (reads vm[0])

  // ---- evaluation (stereoMatching.h:1748-1825).  Same arithmetic and the same console line per region
  // ("nonocc" / "all" / "disc" = I_mask[0..2], skipped when empty); the reference also appends to
  // param_.savePath + err_name and to the CSV stream -- file output is outside the path, the numbers are kept in
  // lastErr instead (addition): lastErr[region] = {PBM, RMS}.
  template <typename T>
  void calErr(Mat& DP, Mat& DT, string procedure, bool calCSV = false);
  struct ErrPair { float PBM = 0.f, RMS = 0.f; bool valid = false; };
  ErrPair lastErr[3];
  // saveDispMap<short> (stereoMatching.h:2005-2110): the disparity map as an 8-bit BGR picture written to
  // param_.savePath + method + ".png" (valid pixels stretched to [0, 255], DISP_OCC blue, DISP_MIS red, DISP_PKR yellow,
  // err_ip_dispV magenta, cor_ip_dispV cyan), and with calErr the "_err.png" variant (error > 1 on I_mask[1] painted red).
  // PNG encoding: sm_io.h (zlib), since OpenCV's imgcodecs is not linked; the directory is created when missing.
  template <typename T, int imgNum = 1>
  void saveDispMap(const cv::Mat& dispM, const Mat& trueM, string method, bool calErr = false);

  // ---- device <-> host mirroring (additions; everything above is the reference's surface)
  void syncToHost(bool volumes = true);   // refresh vm[], HVL[], DP[] host Mats from the device
  Mat& hostDP(int i);                     // DP[i], refreshed if stale
  Mat& hostVm(int i);                     // vm[i], refreshed if stale
  sm_ctx* ctx() { return ctx_; }
  int sgmPaths() const { return sgm_paths_; }
  void setSgmPaths(int p) { sgm_paths_ = p; }   // 4 is compiled into the reference (stereoMatching.cpp:6214); 8 = same table

  // ---- public state, as in the reference (stereoMatching.h:2701-2737)
  Parameters param_;
  vector<Mat> I_c, I_g, I_mask;
  int h_, w_, d_;
  std::vector<cv::Mat1f> L;
  int size_vm[3];
  int HVL_num;
  vector<Mat> HVL, HVL_INTERSECTION;
  vector<Mat> vm;
  cv::Mat DP[2];
  cv::Mat DT;
  cv::Mat LRC_Err_Mask;
  cv::Mat SE;   // refine()'s sub-pixel map after its 3x3 median (a local the reference drops, stereoMatching.cpp:1484-1490)
  cv::Mat guideDisp;

 private:
  void check(int rc, const char* what);
  void upload(void* d, const void* h, size_t bytes);
  void download(void* h, const void* d, size_t bytes);
  void uploadVm(int i);
  void ensureArms();
  void* dalloc(size_t bytes);

  sm_ctx* ctx_ = nullptr;
  int sgm_paths_ = 4;
  // device mirrors
  uint8_t *d_bgr_[2] = {nullptr, nullptr}, *d_gray_[2] = {nullptr, nullptr};
  uint64_t* d_cen_[2] = {nullptr, nullptr};
  uint16_t* d_arms_[2] = {nullptr, nullptr};
  float* d_grad_[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // gx, gy per image (allocated on first use)
  void ensureGrad();
  float* d_vol_[3] = {nullptr, nullptr, nullptr};   // vm[0], vm[1], scratch
  int16_t *d_disp_[2] = {nullptr, nullptr}, *d_tmp16_ = nullptr;
  bool vm_dev_fresh_[2] = {false, false};   // device copy is the authoritative one (host Mat stale)
  bool dp_dev_fresh_[2] = {false, false};
  bool arms_dev_ = false, arms_host_stale_ = false;
  vector<void*> owned_;
};

// The caller's cross-scale step, as the reference declares it (stereoMatching.h:2740, stereoMatching.cpp:2142-2208):
// smPyr[s] = the StereoMatching of pyramid level s after costCalculate(); level 0's vm[] is replaced by the blend.
void SolveAll(StereoMatching**& smPyr, const int PY_LVL, const float REG_LAMBDA);
// cv::pyrDown for the 8-bit images of the pyramid loop (main_.cpp:145-148), on the GPU (addition: OpenCV's own
// pyrDown serves equally when the build has it).
void pyrDown_u8(const cv::Mat& src, cv::Mat& dst);
