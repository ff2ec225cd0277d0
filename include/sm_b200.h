/* ============================================================================
 * sm_b200.h -- C ABI of the B200-native dense-stereo hot path.
 *
 * Drop-in boundary for the stage API of xinge456/myStereoMatching
 * (class StereoMatching, stereoMatching.h:46-2738, and the NL/ aggregator).
 * Each entry point names the reference interface it replaces (file:line under
 * the reference root).  Plain pointers and sizes only; every function returns
 * 0 on success or a negative sm_status (sm_last_error() gives the text).
 *
 * Conventions
 *  - "d_" pointers are DEVICE pointers (allocate with sm_dev_alloc or any CUDA
 *    allocator on the ctx's device); "h_" pointers are HOST pointers.
 *  - All work of one sm_ctx is ordered on that ctx's CUDA stream.  Calls on a
 *    ctx must come from one host thread at a time; distinct ctxs (one per GPU)
 *    are independent.  Stage calls are asynchronous; sm_ctx_sync() waits.
 *  - Layouts are the reference's: cost volume [H][W][D] float32 with d fastest
 *    (stereoMatching.cpp:2080-2081); disparity [H][W] int16, unscaled
 *    (stereoMatching.cpp:3964); arms [H][W][5] uint16 =
 *    [left,right,up,down,sum] (stereoMatching.cpp:5555-5559); census codes
 *    [H][W][nwords] uint64 (stereoMatching.cpp:834-838); images BGR u8
 *    interleaved / gray u8, row-major, no padding.
 *  - There is no CPU fallback: without a CUDA device every call fails.
 * ========================================================================== */
#ifndef SM_B200_H
#define SM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sm_ctx sm_ctx;

typedef enum sm_status {
  SM_OK = 0,
  SM_ERR_ARG = -1,      /* CV_Assert-style precondition failure                */
  SM_ERR_CUDA = -2,     /* a CUDA runtime call or kernel launch failed         */
  SM_ERR_NOMEM = -3,    /* device allocation failed                            */
  SM_ERR_UNSUPPORTED = -4
} sm_status;

/* POD copy of the StereoMatching::Parameters fields the hot path reads
 * (stereoMatching.h:85-351) plus the static step switches it honours
 * (stereoMatching.h:57-83).  sm_params_default() fills the reference defaults. */
typedef struct sm_params {
  int numDisparities;       /* maxDisp+1, stereoMatching.h:209 (<= 512 = CV_CN_MAX)     */
  int censusFunc;           /* 0 = 9x7 centre census (63 bit), 3 = +8 ring bits (71)  :244 */
  float adTrunc;            /* AD truncation / out-of-range default, 1000  stereoMatching.cpp:905 */
  float lamAD, lamCen;      /* 10, 30                                      stereoMatching.cpp:5270 */
  int cbca_crossL;          /* 17  stereoMatching.h:263 */
  int cbca_crossL_out;      /* 34  :266 */
  int cbca_cTresh;          /* 20  :269 */
  int cbca_cTresh_out;      /* 6   :272 */
  int cbca_minArmL;         /* 1   :259 */
  int cbca_iterationNum;    /* 2   :260 */
  int sgm_paths;            /* 4 (compiled-in, stereoMatching.cpp:6214) or 8 */
  int sgm_corDifThres;      /* 15  :239 */
  int sgm_reduCoeffi1;      /* 4   :240 */
  float LRmaxDiff;          /* 0   :212 */
  int region_vote_nums;     /* 2   :306 */
  int regVote_SThres;       /* 20  stereoMatching.cpp:1402 */
  float regVote_hratioThres;/* 0.4 stereoMatching.cpp:1401 */
  int DISP_OCC;             /* -32 stereoMatching.h:216 */
  int DISP_MIS;             /* -48 :217 */
  int aggregation;          /* 0 none, 1 "CBCA", 2 "NL"  (static string, stereoMatching.h:52) */
  int Do_refine;            /* stereoMatching.h:70; 1 in the benchmark configs */
  int Do_LRConsis;          /* :72 */
  int Do_regionVote;        /* :75 */
  int Do_properIpol;        /* :76 */
  int Do_lastMedianBlur;    /* :80 */
  float crossScaleLambda;   /* <0: skip; >=0: the caller's SolveAll step with REG_LAMBDA = this (0.3 in main_.cpp:157) */
  int sgm_grouped;          /* 8 paths only.  0: add the path volumes in the reference's order L0+L1+...+L7
                             * (bit-exact).  1 (default): sweep the row-wise paths {0,4,5} and {1,6,7} together
                             * (one read of C and one read-modify-write of the sum for three paths); every path
                             * volume is still exact, only the order of the eight additions differs: integer-valued
                             * costs stay bit-exact, float sums agree to a few ulp (see sm_sgm_grouped). */
  int costcalculation;      /* 0 "ADCensus" (BASELINE configs), 1 "censusGrad" (the selector main_.cpp:15 compiles in),
                             * 2 "Census" (censusCal(vm, 1), stereoMatching.cpp:975-976): Hamming volume.  With no
                             * aggregation and a power-of-two sgm_reduCoeffi1 the whole frame runs on uint16 volumes
                             * (sm_sgm_u16: exact, half the bytes); sm_pipeline_buffer(0 / 1) then are uint16 volumes in
                             * fixed point (sgm_reduCoeffi1 x the float values).  Otherwise float32 as the reference. */
  float cg_lamCen, cg_lamG; /* censusGrad: Parameters::lamCen = 13, lamG = 1 (main_.cpp:60-61, stereoMatching.cpp:37-41) */
  float gradTrunc;          /* 500  (censusGrad -> grad(gradVm, 500), stereoMatching.cpp:34) */
  int pyramidLevels;        /* PY_LEV of main_.cpp:132 (1 there).  > 1 with crossScaleLambda >= 0: cost + aggregation run
                             * on every level of the pyrDown pyramid (maxDisp/2+1, arm lengths / 2 per level,
                             * main_.cpp:134-151) and SolveAll blends them into level 0 before SGM */
  /* vmTop (stereoMatching.h:182-188, 320-326; dispOptimize, stereoMatching.cpp:1111-1121): instead of the plain WTA the
   * map comes from the vmTop_Num best candidates per pixel (selectTopCostFromVolumn) through genDispFromTopCostVm2 */
  int Do_vmTop;             /* 0   :320 */
  int vmTop_method;         /* 0   :321 */
  int vmTop_Num;            /* M of the constructor (2 in main_.cpp:62), <= 16   :322 */
  float vmTop_thres;        /* lamc * 0.01 (1.09 in main_.cpp:63)   :323 */
  int vmTop_ts;             /* Parameters::ts (10 in main_.cpp:64): disp_DifThres of genDispFromTopCostVm2 */
  int vmTop_hasCir2;        /* 1   :325 */
  int vmTop_cir3_doColorLimit; /* 0 :326 */
  int keep_right_volume;    /* sm_pipeline only.  0 (default): the last sgm path of the right view feeds gen_dispFromVm
                             * directly and its path sum is not stored -- nothing reads vm[1] after dispOptimize's WTA
                             * (refine() works on DP[] and vm[0]); sm_pipeline_buffer(1) then holds the sum WITHOUT the last
                             * path.  1: vm[1] = the finished sum, as the reference leaves it.  Disparity maps are identical. */
} sm_params;

void sm_params_default(sm_params* p, int maxDisp);

/* ---- context / memory ---------------------------------------------------- */
/* stream: a cudaStream_t to launch on (e.g. the caller's current stream), or
 * NULL to let the ctx create its own non-blocking stream.  To run on the legacy
 * default stream pass cudaStreamLegacy ((void*)0x1), not 0. */
int sm_ctx_create(sm_ctx** out, int device, void* stream);
int sm_ctx_destroy(sm_ctx* ctx);
int sm_ctx_sync(sm_ctx* ctx);
void* sm_ctx_stream(sm_ctx* ctx);
const char* sm_last_error(void);
int sm_device_count(void);
int sm_dev_alloc(sm_ctx* ctx, void** d_ptr, size_t bytes);
int sm_dev_free(sm_ctx* ctx, void* d_ptr);
int sm_host_alloc_pinned(void** h_ptr, size_t bytes);
int sm_host_free_pinned(void* h_ptr);
int sm_memcpy_h2d(sm_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int sm_memcpy_d2h(sm_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
int sm_memset(sm_ctx* ctx, void* d_dst, int byte, size_t bytes);
int sm_memcpy_d2d(sm_ctx* ctx, void* d_dst, const void* d_src, size_t bytes);
/* number of kernels this ctx has launched since creation (bench: gpu_launches) */
long long sm_ctx_launch_count(sm_ctx* ctx);

/* ---- cost computation ---------------------------------------------------- */
/* cv::imread(...,0) gray conversion of a BGR image (main_.cpp:95-96). */
int sm_bgr2gray(sm_ctx* ctx, const uint8_t* d_bgr, int H, int W, uint8_t* d_gray);

/* genCensusCode<uchar> (func 0, stereoMatching.h:634-688) and
 * genCensusCode_NC_Sur (func 3, stereoMatching.h:867-934) with the 7x9 window
 * censusCal fixes (stereoMatching.cpp:815).  d_words: [H][W][nwords] uint64,
 * nwords = 1 (func 0) or 2 (func 3). */
int sm_census(sm_ctx* ctx, const uint8_t* d_gray, int H, int W, int func, uint64_t* d_words);
int sm_census_words(int func);
int sm_census_code_length(int func);

/* gen_cenVM_XOR (stereoMatching.h:936-981): Hamming cost volume as float32.
 * LOR 0 = left-referenced, 1 = right-referenced. */
int sm_cost_hamming(sm_ctx* ctx, const uint64_t* d_cenL, const uint64_t* d_cenR, int H, int W, int D,
                    int func, int LOR, float* d_vol);
/* Same volume stored as uint16 (exact; the integer-cost path of costScan,
 * stereoMatching.cpp:2007-2014). */
int sm_cost_hamming_u16(sm_ctx* ctx, const uint64_t* d_cenL, const uint64_t* d_cenR, int H, int W,
                        int D, int func, int LOR, uint16_t* d_vol);
/* gen_ad_sd_vm with AOS=0 (stereoMatching.cpp:2468-2509). */
int sm_cost_ad(sm_ctx* ctx, const uint8_t* d_bgrL, const uint8_t* d_bgrR, int H, int W, int D,
               int LOR, float trunc, float* d_vol);
/* ADCensusCal (stereoMatching.cpp:894-915) = gen_ad_sd_vm + gen_cenVM_XOR +
 * gen_vm_from2vm_exp (stereoMatching.cpp:3566-3590) fused into one pass:
 * vol = 2 - exp(-AD/lamAD) - exp(-census/lamCen); neither intermediate volume
 * is materialised. */
int sm_cost_adcensus(sm_ctx* ctx, const uint8_t* d_bgrL, const uint8_t* d_bgrR,
                     const uint64_t* d_cenL, const uint64_t* d_cenR, int H, int W, int D, int func,
                     float adTrunc, float lamAD, float lamCen, int LOR, float* d_vol);
/* gen_vm_from2vm_exp on two materialised volumes (stage API completeness). */
int sm_combine_exp(sm_ctx* ctx, const float* d_vm0, const float* d_vm1, size_t n, float aru0,
                   float aru1, float* d_out);

/* ---- gradient cost family ---------------------------- */
/* calGrad + calGrad_y (stereoMatching.cpp:271-368) on a gray image: central differences 0.5*(next-prev), one-sided
 * (not halved) in the border column / row.  H, W >= 2 (the reference reads pixel 1 and H-2 unconditionally). */
int sm_grad_xy(sm_ctx* ctx, const uint8_t* d_gray, int H, int W, float* d_gx, float* d_gy);
/* grad() -> calgradvm (stereoMatching.cpp:603-656, 388-455; gradFuse_adpWgt = 1, grad_use2direc = 1):
 * vol[v][u][d] = a*min(|gx0[u0]-gx1[u1]|,trunc) + (1-a)*min(|gy0[u0]-gy1[u1]|,trunc), a = sH/(sH+sV) from the arms
 * of the view's own image (d_armsView = HVL[LOR]); out of range sqrt(2 trunc^2).  Bit-exact. */
int sm_cost_grad(sm_ctx* ctx, const float* d_gxL, const float* d_gyL, const float* d_gxR, const float* d_gyR,
                 const uint16_t* d_armsView, int H, int W, int D, float trunc, int LOR, float* d_vol);
/* censusGrad (stereoMatching.cpp:25-48) fused: 2 - exp(-hamming/lamCen) - exp(-grad/lamG); neither the census nor
 * the gradient volume is materialised.  Census term from a host libm table (exact), gradient term by the device expf:
 * <= 1e-4 relative (in practice ~1e-7). */
int sm_cost_censusgrad(sm_ctx* ctx, const uint64_t* d_cenL, const uint64_t* d_cenR, const float* d_gxL,
                       const float* d_gyL, const float* d_gxR, const float* d_gyR, const uint16_t* d_armsView,
                       int H, int W, int D, int func, float lamCen, float lamG, float gradTrunc, int LOR,
                       float* d_vol);

/* ---- aggregation: CBCA ---------------------------------------------------- */
/* calHorVerDis<uchar> (stereoMatching.cpp:2958-3050) for one 3-channel image.
 * d_arms: [H][W][5] uint16. */
int sm_arms(sm_ctx* ctx, const uint8_t* d_bgr, int H, int W, int L, int L_out, int cTresh,
            int cTresh_out, int minL, uint16_t* d_arms);
/* genTrueHorVerArms (stereoMatching.cpp:2794-2845), materialised
 * [H][W][D][5] uint16 -- for inspection/tests; sm_cbca never needs it. */
int sm_arms_intersect(sm_ctx* ctx, const uint16_t* d_armsL, const uint16_t* d_armsR, int H, int W,
                      int D, int view, uint16_t* d_out);
/* cbca_core for one view (stereoMatching.cpp:5585-5666) with
 * cbca_intersect=true: `iters` iterations of {cumulate, span} along both axes
 * (H,V on even iterations, V,H on odd) each followed by the area division.
 * In place on d_vol; d_tmp is a scratch volume of the same size. */
int sm_cbca(sm_ctx* ctx, float* d_vol, float* d_tmp, const uint16_t* d_armsL,
            const uint16_t* d_armsR, int H, int W, int D, int iters, int view);

/* ---- aggregation: NL (non-local MST tree filter) -------------------------- */
/* ctmf (NL/ctmf.h:8, NL/ctmf.c:378-433): (2r+1)^2 median, 8-bit, cn
 * interleaved channels, edge-replicated border.  r in {1,2,3}. */
int sm_median_u8(sm_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, int H, int W, int r, int cn);
/* qx_mst_kruskals_image::mst (NL/qx_mst_kruskals_image.cpp:167-277): the MST of
 * the 4-connected grid under the total order (weight, edge enumeration index),
 * rooted at pixel 0.  Outputs, each H*W: parent (root = itself), weight = edge
 * to parent (root 0), rank = depth.  d_order (nullable): nodes sorted by
 * (rank, index) -- a valid parent-before-child order (the reference's BFS order
 * sorts within a level by adjacency; only the fp64 summation order of siblings
 * depends on it). */
int sm_mst_build(sm_ctx* ctx, const uint8_t* d_bgr, int H, int W, int cn, int32_t* d_parent,
                 uint8_t* d_weight, int32_t* d_rank, int32_t* d_order);
/* qx_tree_filter::filter (NL/qx_tree_filter.cpp:61-117) on a float32 volume with
 * fp64 arithmetic inside (NLCCA::aggreCV's f32->f64->f32, NL/NLCCA.cpp:56-90).
 * d_work: scratch of H*W*D doubles. */
int sm_tree_filter(sm_ctx* ctx, float* d_vol, double* d_work, int H, int W, int D,
                   const int32_t* d_parent, const uint8_t* d_weight, const int32_t* d_rank,
                   const int32_t* d_order, double sigma);
/* qx_tree_filter::filter itself (NL/qx_tree_filter.cpp:61-117) on a float64
 * volume [H*W][D], in place -- the entry point a qx_tree_filter drop-in binds
 * (NL/qx_tree_filter.h:24: filter(double* cost, double* cost_backup, int nr_plane)). */
int sm_tree_filter_f64(sm_ctx* ctx, double* d_cost, int H, int W, int D, const int32_t* d_parent,
                       const uint8_t* d_weight, const int32_t* d_rank, const int32_t* d_order, double sigma);
/* StereoMatching::NL (stereoMatching.cpp:4892-4917): aggreCV(vm[0]),
 * aggreCV(ones), divide.  In place on d_vol. */
int sm_nl(sm_ctx* ctx, const uint8_t* d_bgrL, float* d_vol, int H, int W, int D);

/* ---- Yang's own driver: qx_nonlocal_cost_aggregation (float64 volumes [H][W][D]) --- */
/* compute_gradient (NL/qx_nonlocal_cost_aggregation.cpp:219-236). */
int sm_nlca_gradient(sm_ctx* ctx, const uint8_t* d_img, int H, int W, float* d_grad);
/* matching_cost_from_color_and_gradient (NL/qx_nonlocal_cost_aggregation.cpp:190-218):
 * w*min(mean|dRGB|, maxc) + (1-w)*min(|dgrad|, maxg); class defaults 7, 2, 0.11. */
int sm_nlca_cost(sm_ctx* ctx, const uint8_t* d_left, const uint8_t* d_right, int H, int W, int D,
                 double max_color_difference, double max_gradient_difference, double weight_on_color,
                 double* d_vol);
/* qx_stereo_flip_corr_vol (NL/qx_basic.cpp:577-588). */
int sm_nlca_flip(sm_ctx* ctx, const double* d_vol, int H, int W, int D, double* d_vol_right);
/* depth_best_cost / vec_min_pos (NL/qx_basic.cpp:589-602): first minimum as u8 (D <= 256). */
int sm_depth_best_cost(sm_ctx* ctx, const double* d_vol, int H, int W, int D, uint8_t* d_depth);
/* qx_detect_occlusion_left_right (NL/qx_basic.cpp:603-624). */
int sm_nlca_occlusion(sm_ctx* ctx, const uint8_t* d_disp_left, const uint8_t* d_disp_right, int H, int W,
                      uint8_t* d_mask);
/* disparity()'s refinement volume (NL/qx_nonlocal_cost_aggregation.cpp:92-99):
 * 0 on occluded pixels, |disp - d| elsewhere. */
int sm_nlca_refine_cost(sm_ctx* ctx, const uint8_t* d_disp, const uint8_t* d_mask, int H, int W, int D,
                        double* d_vol);

/* ---- optimisation: SGM ---------------------------------------------------- */
/* costScan + updateCost<float> for ONE path (stereoMatching.cpp:1983-2029,
 * stereoMatching.h:2205-2280).  path indexes the reference's direction table
 * rv={+1,-1,0,0,+1,+1,-1,-1}, ru={0,0,+1,-1,-1,+1,+1,-1}
 * (stereoMatching.cpp:6207-6208).  d_bgr = colour image of the view.
 * mode 0: d_out = Lr (the member L[i]); mode 1: d_out += Lr (gen_sgm_vm's
 * running sum, stereoMatching.cpp:2031-2056). */
int sm_sgm_path(sm_ctx* ctx, const float* d_vol, const uint8_t* d_bgr, int H, int W, int D, int path,
                int corDifThres, int reduCoeffi1, int mode, float* d_out);
/* sgm() (stereoMatching.cpp:6204-6224): `paths` sweeps summed in table order
 * into d_sum (must not alias d_vol). */
int sm_sgm(sm_ctx* ctx, const float* d_vol, const uint8_t* d_bgr, int H, int W, int D, int paths,
           int corDifThres, int reduCoeffi1, float* d_sum);

/* sgm() on a 16-bit INTEGER cost volume, natively (costScan's integer entry, stereoMatching.cpp:2007-2014:
 * vm.depth() CV_8U / CV_16U -> updateCost<uchar | ushort>, stereoMatching.h:2205-2280).  d_vol: [H][W][D] uint16 raw
 * costs (e.g. sm_cost_hamming_u16).  With integer costs and a power-of-two reduCoeffi1 every value of updateCost is a
 * multiple of 1 / reduCoeffi1 and the reference's float arithmetic is exact; the kernels keep Lr and the path sum as
 * uint16 fixed point: d_sum = reduCoeffi1 x (the float volume sgm() leaves in vm), exactly, and d_disp (nullable) =
 * gen_dispFromVm of it.  Half the HBM bytes of the float path per pass.  maxCost bounds the raw costs (71 for the
 * 71-bit census); SM_ERR_UNSUPPORTED unless reduCoeffi1 is a power of two and paths*(maxCost+3)*reduCoeffi1 <= 16000
 * (then: sm_vol_to_f32 + sm_sgm). */
int sm_sgm_u16(sm_ctx* ctx, const uint16_t* d_vol, const uint8_t* d_bgr, int H, int W, int D, int paths, int corDifThres,
               int reduCoeffi1, int maxCost, uint16_t* d_sum, int16_t* d_disp);
/* gen_dispFromVm (stereoMatching.cpp:3928-3967) on a uint16 volume: first minimum. */
int sm_wta_u16(sm_ctx* ctx, const uint16_t* d_vol, int H, int W, int D, int16_t* d_disp);

/* gen_sgm_vm's inner statement `sum += Lr[num]` (stereoMatching.cpp:2051) for one
 * materialised path volume: d_acc[i] = d_acc[i] + d_x[i]. */
int sm_vol_accumulate(sm_ctx* ctx, float* d_acc, const float* d_x, size_t n);

/* sgm() with 8 paths where the three upward paths {0,4,5} and the three downward paths {1,6,7} are each computed
 * in ONE sweep over the rows (3 V b of HBM traffic per group instead of 9 V b).  Each path's Lr is computed exactly
 * as updateCost does; the sum is formed as ((L0+L4)+L5) + L1 + L6 + L7 + L2 + L3 instead of L0+L1+...+L7, so it is
 * bit-exact for integer-valued costs and within a few ulp (<= 1e-6 relative) otherwise.  Falls back to sm_sgm's
 * path-by-path order when the shape is unsupported (D % 4 != 0, D <= 64, very wide images). */
int sm_sgm_grouped(sm_ctx* ctx, const float* d_vol, const uint8_t* d_bgr, int H, int W, int D, int corDifThres,
                   int reduCoeffi1, float* d_sum);
/* The same for both views of a frame at once (StereoMatching::dispOptimize runs sgm() on vm[0] and vm[1],
 * stereoMatching.cpp:1051-1089): each row sweep processes the left and the right volume in ONE launch, two thread
 * blocks per SM, which hides the sweep's row-to-row latency.  Results are identical to two sm_sgm_grouped calls. */
int sm_sgm_grouped2(sm_ctx* ctx, const float* d_volL, const float* d_volR, const uint8_t* d_bgrL, const uint8_t* d_bgrR,
                    int H, int W, int D, int corDifThres, int reduCoeffi1, float* d_sumL, float* d_sumR);

/* ---- disparity selection --------------------------------------------------- */
/* gen_dispFromVm (stereoMatching.cpp:3928-3967), ChooseSmall = true. */
int sm_wta(sm_ctx* ctx, const float* d_vol, int H, int W, int D, int16_t* d_disp);
/* wta_Co (stereoMatching.cpp:2709-2792): left map and right map from the LEFT
 * volume's diagonal, both multiplied by `scale` (DISP_SCALE = 16). */
/* selectTopCostFromVolumn (stereoMatching.h:2405-2461; called on a clone of vm,
 * stereoMatching.cpp:1118-1119): per pixel up to `num` candidate disparities in
 * order of increasing cost -- first minimum of what is left, lowest d on ties;
 * candidate 0 always, candidate k > 0 only while cost < firstCost * thres.
 * d_top = float [H][W][num+1][2]: [k] = {d, cost}, [num][0] = candidate count;
 * entries the reference leaves unwritten are 0.  d_vol is not modified. */
int sm_select_top_cost(sm_ctx* ctx, const float* d_vol, int H, int W, int D, int num,
                       float thres, float* d_top);

/* vmTop, second half (param_.Do_vmTop, stereoMatching.cpp:1111-1121): the disparity map from the candidate lists of
 * sm_select_top_cost.  d_top = float [H][W][num+1][2], num <= 16.  Pixels whose candidate count is < 1 keep what
 * d_disp holds on entry, as in the reference.
 * sm_disp_from_top  = genDispFromTopCostVm  (stereoMatching.h:2466-2545): own + left / right neighbours' candidates
 *   vote (count, then summed cost; the reference's `dNum = dispNum && cost_ < cost` assignment is reproduced).
 * sm_disp_from_top2 = genDispFromTopCostVm2 (stereoMatching.cpp:1514-1886), method = param_.vmTop_method:
 *   0: the author's heuristic over the 8-neighbourhood (ts = param_.ts, hasCir2 = vmTop_hasCir2, colorLimit =
 *      vmTop_cir3_doColorLimit); its raster-order dependency (left / up / up-left / up-right results) is honoured;
 *   1, 2: row scans against the previous pixel's result.  d_bgr = I_c[0] (methods 0 with colorLimit, and 2).
 * Bit-exact (same float accumulation order per disparity). */
int sm_disp_from_top(sm_ctx* ctx, const float* d_top, int H, int W, int num, int16_t* d_disp);
int sm_disp_from_top2(sm_ctx* ctx, const float* d_top, const uint8_t* d_bgr, int H, int W, int num, int method, int ts,
                      int hasCir2, int colorLimit, int16_t* d_disp);

/* subpixelEnhancement (stereoMatching.cpp:6138-6166; off by default,
 * Do_subpixelEnhancement, stereoMatching.h:79): for 0 < disp < D-1 the parabola
 * offset diff = (c[+1]-c[-1]) / (2*(c[+1]+c[-1]-2*c[0])) is subtracted when
 * denom != 0 and -1 < diff < 1 -- on the short, as the reference does
 * (truncation toward zero), then converted.  d_floatDisp = float [H][W]. */
int sm_subpixel_enhancement(sm_ctx* ctx, const int16_t* d_disp, const float* d_vol, int H, int W, int D,
                            float* d_floatDisp);
int sm_wta_co(sm_ctx* ctx, const float* d_vol, int H, int W, int D, int scale, int16_t* d_D1,
              int16_t* d_D2);

/* ---- refinement ------------------------------------------------------------ */
/* LRConsistencyCheck_normal (stereoMatching.cpp:2262-2282): in place on d_D1. */
int sm_lrc(sm_ctx* ctx, int16_t* d_D1, const int16_t* d_D2, int H, int W, float LRmaxDiff);
/* LRConsistencyCheck, LOR=0 (stereoMatching.cpp:2284-2335): occlusion/mismatch
 * labelling; d_errMask (nullable) [H][W] u8. */
int sm_lrc_label(sm_ctx* ctx, int16_t* d_D1, const int16_t* d_D2, int H, int W, int D,
                 float LRmaxDiff, int DISP_OCC, int DISP_MIS, uint8_t* d_errMask);
/* LRConsistencyCheck with either LOR (stereoMatching.cpp:2284-2364).  LOR 0: as sm_lrc_label (d_D1 labelled in
 * place, d_errMask = flags, d_errMask1 = 0).  LOR 1 (:2336-2364): the RIGHT map d_D2 is checked against d_D1
 * (u + d inside the image, |d - D1[u+d]| <= LRmaxDiff) and labelled in place; the reference leaves errMask all zero on
 * this branch and puts the flags into a local errMask1 (written to LR1.png only): d_errMask <- 0, d_errMask1 <- flags.
 * Both masks are nullable. */
int sm_lrc_label_lor(sm_ctx* ctx, int16_t* d_D1, int16_t* d_D2, int H, int W, int D, float LRmaxDiff, int DISP_OCC,
                     int DISP_MIS, int LOR, uint8_t* d_errMask, uint8_t* d_errMask1);
/* regionVote_my (stereoMatching.cpp:7219-7277): one Jacobi sweep, in place.
 * d_arms = HVL[0]; d_tmp = scratch [H][W] int16. */
int sm_region_vote(sm_ctx* ctx, int16_t* d_disp, int16_t* d_tmp, const uint16_t* d_arms, int H, int W,
                   int D, float ratio, int S);
/* properIpol (stereoMatching.cpp:7395-7490): one Jacobi sweep, in place. */
int sm_proper_ipol(sm_ctx* ctx, int16_t* d_disp, int16_t* d_tmp, const uint8_t* d_bgr, int H, int W,
                   int DISP_OCC);
/* WM (stereoMatching.cpp:7340-7393; Do_WM, stereoMatching.h:74, off): 19x19 bilateral weighted median on the pixels
 * with d_mask > 0, in place on d_disp (the window reads the map as it was on entry; d_tmp = scratch [H][W] int16).
 * d_bgr = the guidance image (I_c[0]).  Weights use expf exactly as the host libm evaluates it, and every float sum
 * keeps the reference's order: bit-exact.  A label outside [0, D) inside a window is undefined behaviour in the
 * reference (:7371); here it adds to the total weight but casts no vote, and *d_numInvalid (device int, nullable)
 * receives how many such neighbours were seen. */
int sm_wm(sm_ctx* ctx, int16_t* d_disp, int16_t* d_tmp, const uint8_t* d_mask, const uint8_t* d_bgr, int H, int W, int D,
          int* d_numInvalid);
/* discontinuityAdjust (stereoMatching.cpp:6057-6135; Do_discontinuityAdjust, stereoMatching.h:78, off), in place on
 * d_disp: the map as an 8-bit picture (saturating convertTo) -> equalizeHist -> GaussianBlur(3x3, sigma 4) ->
 * Canny(20, 60, 3) (integer kernels, pinned against cv2 4.13), then every interior edge pixel whose 3x3 edge
 * neighbourhood names a direction takes the cheapest (in d_vol = vm[0]) of its own label and the two neighbours' across
 * that direction, with the reference's raster-order semantics (the first neighbour may already be adjusted).
 * d_edge (nullable, [H][W] u8) receives the Canny map.  Labels >= D are never used as indices (undefined in the
 * reference): such a centre stays, such a neighbour is no candidate. */
int sm_discontinuity_adjust(sm_ctx* ctx, int16_t* d_disp, const float* d_vol, int H, int W, int D, uint8_t* d_edge);
/* cv::medianBlur(CV_16S, 3) (stereoMatching.cpp:1499). d_dst != d_src. */
int sm_median3_i16(sm_ctx* ctx, const int16_t* d_src, int16_t* d_dst, int H, int W);
/* cv::medianBlur(SE, SE, 3) on the CV_32F map subpixelEnhancement returns
 * (stereoMatching.cpp:1490): 3x3 median, replicated border, NaN-free input;
 * d_src != d_dst. */
int sm_median3_f32(sm_ctx* ctx, const float* d_src, float* d_dst, int H, int W);
/* SolveAll with one pyramid level (stereoMatching.cpp:2142-2208, main_.cpp:158). */
int sm_cross_scale_1level(sm_ctx* ctx, float* d_vol, size_t n, float lambda);
/* ---- the caller's cross-scale step, any number of levels ---- */
#define SM_MAX_PYRAMID 6
/* cv::pyrDown on an 8-bit image, cn = 1 | 3 (main_.cpp:145-148): [1 4 6 4 1]/16 separable, REFLECT_101,
 * d_dst is ((H+1)/2) x ((W+1)/2).  Bit-exact. */
int sm_pyr_down_u8(sm_ctx* ctx, const uint8_t* d_src, int H, int W, int cn, uint8_t* d_dst);
/* invWgt[0..n) = row 0 of regMat.inv() (stereoMatching.cpp:2147-2170), exactly as cv::invert computes it on
 * CV_32F.  Host-only helper (no device work). */
int sm_cross_scale_weights(int n, float lambda, float* invWgt);
/* SolveAll (stereoMatching.cpp:2142-2208): d_vols[s] is the level-s volume [Hs[s]][Ws[s]][Ds[s]]; d_vols[0] is
 * replaced by sum_s invWgt[s] * vol_s[y>>s][x>>s][d_s], d_{s+1} = (d_s+1)/2.  Bit-exact. */
int sm_cross_scale(sm_ctx* ctx, float* const* d_vols, const int* Hs, const int* Ws, const int* Ds, int levels,
                   float lambda);

/* ---- stage-API sub-steps the pipeline kernels fuse, one by one (stage_parts.cu) ---- */
/* gen1DCumu (stereoMatching.cpp:3896-3926): in-place running sum along -u ((dv,du) = (0,-1)) or -v ((-1,0)) of the
 * volume and, when not NULL, of the int32 area volume.  Sequential float order: bit-exact. */
int sm_cumsum_1d(sm_ctx* ctx, float* d_vol, int32_t* d_areaIS, int H, int W, int D, int dv, int du);
/* cal1DCost (stereoMatching.h:1643-1715, cbca_intersect = true): out = cum[head] - cum[pre_tail] with the
 * materialised intersected arms d_hvl_is [H][W][D][5] (sm_arms_intersect); results replace d_vol / d_areaIS
 * (d_tmp_*: scratch of the same sizes).  direc 0: horizontal span, 1: vertical. */
int sm_span_1d(sm_ctx* ctx, float* d_vol, int32_t* d_areaIS, const uint16_t* d_hvl_is, float* d_tmp_vol,
               int32_t* d_tmp_area, int H, int W, int D, int dv, int du, int direc);
/* genfinalVm_cbca (stereoMatching.cpp:3969-3992): vol[i] /= (float)areaIS[i]. */
int sm_div_area(sm_ctx* ctx, float* d_vol, const int32_t* d_areaIS, size_t n);
/* updateCost<float> (stereoMatching.h:2205-2280) at ONE pixel (v,u) of path (rv,ru); d_bgr = the image of the view
 * (I_c[0] if leftFirst else I_c[1]). */
int sm_update_cost(sm_ctx* ctx, float* d_Lr, const float* d_vm, const uint8_t* d_bgr, int H, int W, int n, int v,
                   int u, int rv, int ru, int preIsInner, int corDifThres, int reduCoeffi1);
/* The integer-cost entry of costScan (vm.depth() CV_8U / CV_16U, stereoMatching.cpp:2007-2021): updateCost<uchar> /
 * <ushort> add `T cost` to a float, i.e. they see the volume converted to float; convert (elem_bytes 1 | 2), then
 * sm_sgm_path / sm_sgm.  Exact. */
int sm_vol_to_f32(sm_ctx* ctx, const void* d_src, int elem_bytes, size_t n, float* d_dst);
/* LRConsistencyCheck_new (stereoMatching.cpp:2367-2382): mask[v][u] = 0 where the left-right check fails (Thres 0). */
int sm_lrc_mask(sm_ctx* ctx, const int16_t* d_D1, const int16_t* d_D2, int H, int W, uint8_t* d_mask);

/* calErr<short> (stereoMatching.h:1748-1825): over mask == 255: *h_sumNum pixels, *h_errorNum of them with
 * |gt - disp| > thres or disp < 0, *h_errorValueSum = sum of dif^2 (2 for an invalid pixel).  PBM = errorNum / sumNum,
 * RMS = sqrt(errorValueSum / sumNum).  Synchronous (returns host values). */
int sm_cal_err(sm_ctx* ctx, const int16_t* d_disp, const float* d_gt, const uint8_t* d_mask, int H, int W, int thres,
               long long* h_sumNum, long long* h_errorNum, double* h_errorValueSum);

/* ---- whole frame ------------------------------------------------------------ */
/* A frame pipeline owns every device buffer a W x H x D frame needs (three
 * volumes, codes, arms, images, disparities) so a stream of frames reuses them.
 * sm_pipeline_run = pipeline() (stereoMatching.cpp:1950-1981): costCalculate
 * -> dispOptimize -> refine, host images in, host disparity out; the copies are
 * part of the call.  h_gray* may be NULL (computed on the device from BGR).
 * h_dispR may be NULL.  sm_pipeline_run_device runs the same stages on images
 * already uploaded with sm_pipeline_upload (no host traffic). */
typedef struct sm_pipeline sm_pipeline;
int sm_pipeline_create(sm_ctx* ctx, int H, int W, const sm_params* p, sm_pipeline** out);
int sm_pipeline_destroy(sm_pipeline* pl);
int sm_pipeline_upload(sm_pipeline* pl, const uint8_t* h_bgrL, const uint8_t* h_bgrR,
                       const uint8_t* h_grayL, const uint8_t* h_grayR);
int sm_pipeline_run_device(sm_pipeline* pl);
int sm_pipeline_download(sm_pipeline* pl, int16_t* h_dispL, int16_t* h_dispR);
int sm_pipeline_run(sm_pipeline* pl, const uint8_t* h_bgrL, const uint8_t* h_bgrR,
                    const uint8_t* h_grayL, const uint8_t* h_grayR, int16_t* h_dispL,
                    int16_t* h_dispR);
/* The next sm_pipeline_run_device reads the pair from caller-owned DEVICE buffers (bgr: [H][W][3] u8, gray: [H][W] u8,
 * gray NULL = computed from BGR); all NULL restores the pipeline's own buffers.  No copy is made. */
int sm_pipeline_bind_inputs(sm_pipeline* pl, const uint8_t* d_bgrL, const uint8_t* d_bgrR, const uint8_t* d_grayL,
                            const uint8_t* d_grayR);
/* device views of the pipeline's buffers (for tests / the C++ class): which =
 * 0 vm[0], 1 vm[1], 2 DP[0], 3 DP[1], 4 HVL[0], 5 HVL[1], 6 census L, 7 census R */
void* sm_pipeline_buffer(sm_pipeline* pl, int which);
/* per-stage device time (ms) of the last sm_pipeline_run_device when timing was
 * enabled: census, cost, arms, aggregation, sgm, wta, refine, total */
int sm_pipeline_enable_timing(sm_pipeline* pl, int on);
int sm_pipeline_stage_ms(sm_pipeline* pl, float* out8);
/* split of the sgm stage of that run: out2[0] = the grouped row sweeps (k_sgm_group), out2[1] = the single-path
 * launches (k_sgm_path*); {0, sgm} when the grouped sweeps did not run */
int sm_pipeline_sgm_split_ms(sm_pipeline* pl, float* out2);

/* ---- frame stream over one or more GPUs (BASELINE config 5) ------------------------------------
 * The reference runs one StereoMatching object per stereo pair (main_.cpp:138-166).  A stream of pairs parallelises BY
 * FRAME (SGM paths do not shard along rows): sm_stream_create starts one worker per listed device -- a host thread
 * with its own sm_ctx, sm_pipeline, copy stream and double-buffered device / pinned staging buffers -- and frame i
 * (the i-th sm_stream_submit) runs on worker i mod n_devices.  Inside a worker the upload of frame i+1 and the
 * download of frame i-1 overlap the compute of frame i.  No collective: the path has no exchange step.
 *   sm_stream_submit  queues one pair (HOST buffers: pinned ones are copied from directly, pageable ones through the
 *                     worker's staging); returns at once unless queue_depth frames are already waiting on that worker.
 *                     The buffers must stay valid until sm_stream_wait(ticket) returns; h_gray* may be NULL.
 *   sm_stream_wait    blocks until that frame's left disparity map is in h_dispL (status of the worker).
 *   sm_stream_drain   waits for everything submitted so far. */
typedef struct sm_stream sm_stream;
int sm_stream_create(const int* devices, int n_devices, int H, int W, const sm_params* p, int queue_depth, sm_stream** out);
int sm_stream_submit(sm_stream* s, const uint8_t* h_bgrL, const uint8_t* h_bgrR, const uint8_t* h_grayL,
                     const uint8_t* h_grayR, int16_t* h_dispL, long long* ticket);
int sm_stream_wait(sm_stream* s, long long ticket);
int sm_stream_drain(sm_stream* s);
int sm_stream_destroy(sm_stream* s);
int sm_stream_device_count(sm_stream* s);
long long sm_stream_frames_done(sm_stream* s, int worker);   /* frames retired by one worker */
long long sm_stream_launch_count(sm_stream* s);              /* kernels launched by all workers */

#ifdef __cplusplus
}
#endif
#endif /* SM_B200_H */
