// ============================================================================
// stereo_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the dense-stereo hot path of xinge456/myStereoMatching
// (AD-Census cost -> CBCA -> SGM -> WTA -> LR check -> refinement), written
// from the reference's arithmetic, loop order and float operation order.
// Each function cites the reference file:line it follows (paths relative to
// the reference root).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.
//
// Pinning status: PINNED.  The reference ships no tests, fixtures or golden
// vectors for stereoMatching.{h,cpp} and those files cannot be compiled as
// files here (OpenCV C++ + ximgproc + a missing util.h), but the hot-path
// function bodies can: oracle/build_ref_sm.py cuts them out of the reference
// where it lies and compiles them against a cv::Mat stand-in into
// oracle/_ref/libsmref.so.  tests/golden/sm_ref.npz holds that library's outputs
// at every stage boundary on three seeded stereo pairs;
// tests/test_oracle_sm_golden.py requires every function in this file to
// reproduce them bit for bit (and to agree with the live library on further
// shapes when it is present).  The three OpenCV semantics on the path
// (BORDER_REFLECT_101, medianBlur on CV_16S, BGR2GRAY) are pinned by cv2 4.13
// golden vectors (tests/golden/opencv_semantics.npz).  The NL/ part (ctmf /
// MST / tree filter, nl_oracle.cpp) is pinned against oracle/_ref/libqxref.so.
//
// Build: see oracle/Makefile (g++ -O2 -ffp-contract=off, optional -fopenmp).
// Float contraction is disabled so a*b+c never becomes an FMA: the reference
// is an MSVC x64 build, which does not contract.
// ============================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#include "opencv_restated.h"
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned char u8;
typedef unsigned short u16;
typedef uint64_t u64;
typedef short i16;

// Thread policy: the reference hot loops are single-threaded (OMP_PARALLEL_FOR
// is commented out on every hot loop, stereoMatching.h:596,648,1661;
// stereoMatching.cpp:2483,2812,3901,3975).  orc_threads==1 reproduces that.
// orc_threads>1 parallelises only loops whose iterations are independent, and
// never changes any result (no reduction order depends on the thread count).
static int orc_threads = 1;
#ifdef _OPENMP
#define ORC_PAR_FOR _Pragma("omp parallel for schedule(static) num_threads(orc_threads)")
#else
#define ORC_PAR_FOR
#endif

extern "C" {

void orc_set_threads(int n) { orc_threads = n < 1 ? 1 : n; }
int orc_get_threads() { return orc_threads; }
int orc_has_openmp() {
#ifdef _OPENMP
  return 1;
#else
  return 0;
#endif
}

// ---------------------------------------------------------------------------
// BGR -> gray as cv::imread(...,0)/cvtColor(BGR2GRAY) does it for 8-bit:
// fixed point, 14 fractional bits (main_.cpp:95-96 loads the gray pair with
// imread flag 0).  Pinned against cv2 in tests/golden/opencv_semantics.npz.
// ---------------------------------------------------------------------------
void orc_bgr2gray(const u8* bgr, int H, int W, u8* gray) {
  for (long i = 0; i < (long)H * W; i++) {
    int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
    gray[i] = (u8)((1868 * b + 9617 * g + 4899 * r + 8192) >> 14);
  }
}

// cv::BORDER_REFLECT_101 index map (stereoMatching.h:642, 871).
static inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) {
    if (p < 0) p = -p;
    else p = 2 * (n - 1) - p;
  }
  return p;
}

// ---------------------------------------------------------------------------
// Census codes.  func 0: genCensusCode<uchar> (stereoMatching.h:634-688);
// func 3: genCensusCode_NC_Sur (stereoMatching.h:867-934).  Window rows
// dv=-RV..RV (outer), cols du=-RU..RU (inner), centre INCLUDED; bit =
// (centre < neighbour), shifted in MSB-first; a word is flushed only when a
// 65th bit arrives (the "step > 63" test).  func 3 appends 8 bits comparing
// consecutive pixels of the inner 3x3 ring clockwise from top-left.
// words: [H][W][nwords], nwords = ceil(codeLength/64) (stereoMatching.cpp:829-836).
// ---------------------------------------------------------------------------
int orc_census_code_length(int func, int RV, int RU) {
  int len = (2 * RV + 1) * (2 * RU + 1);
  if (func == 3) len += 8;
  return len;
}

void orc_census(const u8* gray, int H, int W, int func, int RV, int RU, u64* words) {
  const int codeLen = orc_census_code_length(func, RV, RU);
  const int nwords = (codeLen + 63) / 64;
  static const int dv_sur[9] = {-1, -1, -1, 0, 1, 1, 1, 0, -1};
  static const int du_sur[9] = {-1, 0, 1, 1, 1, 0, -1, -1, -1};
  ORC_PAR_FOR
  for (int v = 0; v < H; v++) {
    for (int u = 0; u < W; u++) {
      u64* out = words + ((long)v * W + u) * nwords;
      for (int k = 0; k < nwords; k++) out[k] = 0;  // stereoMatching.cpp:839
      const int centre = gray[(long)v * W + u];
      u64 cs = 0;
      int step = 0, dep = 0;
      auto push = [&](bool bit) {
        if (step > 63) { out[dep] = cs; cs = 0; step = 0; dep++; }
        cs <<= 1;
        if (bit) cs++;
        step++;
      };
      for (int dv = -RV; dv <= RV; dv++)
        for (int du = -RU; du <= RU; du++) {
          int nb = gray[(long)reflect101(v + dv, H) * W + reflect101(u + du, W)];
          push(centre - nb < 0);
        }
      if (func == 3) {
        for (int i = 0; i < 8; i++) {
          int pre = gray[(long)reflect101(v + dv_sur[i], H) * W + reflect101(u + du_sur[i], W)];
          int aft = gray[(long)reflect101(v + dv_sur[i + 1], H) * W + reflect101(u + du_sur[i + 1], W)];
          push(pre - aft < 0);
        }
      }
      if (step > 0) out[dep] = cs;
    }
  }
}

// ---------------------------------------------------------------------------
// Hamming cost volume: gen_cenVM_XOR (stereoMatching.h:936-981) with
// HammingDistance = popcount64 (stereoMatching.cpp:2210).  is_censusNorm = 0
// (stereoMatching.h:247) so out-of-range = DEFAULT = codeLength*truncRat.
// LOR 0: lp=u, rp=u-d.  LOR 1: lp=u+d, rp=u.
// ---------------------------------------------------------------------------
void orc_hamming_vol(const u64* cenL, const u64* cenR, int H, int W, int D, int nwords,
                     int codeLength, float truncRat, int LOR, float* vol) {
  const float DEFAULT = codeLength * truncRat;
  const int lc = LOR == 1 ? 1 : 0, rc = LOR == 1 ? 0 : 1;
  ORC_PAR_FOR
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      float* o = vol + ((long)v * W + u) * D;
      for (int d = 0; d < D; d++) {
        int lp = u + d * lc, rp = u - d * rc;
        if (lp >= W || rp < 0) { o[d] = DEFAULT; continue; }
        float cost = 0;
        for (int k = 0; k < nwords; k++)
          cost += (float)__builtin_popcountll(cenL[((long)v * W + lp) * nwords + k] ^
                                              cenR[((long)v * W + rp) * nwords + k]);
        o[d] = std::min(cost, DEFAULT);
      }
    }
}

// ---------------------------------------------------------------------------
// AD cost volume: gen_ad_sd_vm with AOS=0 (stereoMatching.cpp:2468-2509),
// 3-channel BGR (SD_AD_channel=3, stereoMatching.h:228), is_adNorm=0.
// ---------------------------------------------------------------------------
void orc_ad_vol(const u8* bgrL, const u8* bgrR, int H, int W, int D, int LOR, float trunc,
                float* vol) {
  const int lc = LOR == 1 ? 1 : 0, rc = LOR == 1 ? 0 : -1;
  ORC_PAR_FOR
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      float* o = vol + ((long)v * W + u) * D;
      for (int d = 0; d < D; d++) {
        int uL = u + d * lc, uR = u + d * rc;
        if (uL >= W || uR < 0) { o[d] = trunc; continue; }
        const u8* l = bgrL + ((long)v * W + uL) * 3;
        const u8* r = bgrR + ((long)v * W + uR) * 3;
        float sum = 0;
        for (int c = 0; c < 3; c++) sum += std::fabs((float)l[c] - (float)r[c]);
        o[d] = std::min(sum / 3, trunc);
      }
    }
}

// gen_vm_from2vm_exp (stereoMatching.cpp:3566-3590): 2 - exp(-a/l0) - exp(-b/l1),
// float exp overload, left-to-right subtraction.
void orc_combine_exp(const float* a, const float* b, long n, float l0, float l1, float* out) {
  ORC_PAR_FOR
  for (long i = 0; i < n; i++) out[i] = 2 - std::exp(-a[i] / l0) - std::exp(-b[i] / l1);
}

// The two 1-D tables the fused GPU kernel indexes: for every possible AD
// numerator k = sum_c |l_c - r_c| (0..765, plus slot 766 = out-of-range) and
// every Hamming count c (0..codeLength) the exact float the reference's libm
// call produces.  Host-side helper of the same arithmetic as orc_ad_vol +
// orc_combine_exp; the product library has its own copy of this logic.
void orc_exp_tables(float trunc, float lamAD, float lamCen, int codeLength, float* tabAD /*767*/,
                    float* tabCen /*codeLength+1*/) {
  for (int k = 0; k <= 765; k++) {
    float ad = std::min((float)k / 3, trunc);
    tabAD[k] = std::exp(-ad / lamAD);
  }
  tabAD[766] = std::exp(-trunc / lamAD);
  for (int c = 0; c <= codeLength; c++) tabCen[c] = std::exp(-(float)c / lamCen);
}

// ---------------------------------------------------------------------------
// Gradient cost family ("censusGrad" is the selector main_.cpp:15 compiles in).
//   calGrad / calGrad_y (stereoMatching.cpp:271-368), gray input: central difference 0.5*(next - prev), one-sided
//     (next - prev of the two border pixels, NOT halved) in the first/last column resp. row.  Needs W,H >= 2.
//   calgradvm (stereoMatching.cpp:388-455), gradFuse_adpWgt = 1, grad_use2direc = 1 (stereoMatching.h:245-246):
//     a = sH / (sH + sV), sH = min(left,right) arm, sV = min(up,down) arm of the VIEW's own image (0 -> 1), read as
//     short; in range: a*min(|gx0[u0]-gx1[u1]|,T) + (1-a)*min(|gy0[u0]-gy1[u1]|,T); out of range:
//     sqrt(pow(T,2)*2) evaluated in double (pow(float,int) promotes), stored as float.
//   censusGrad (stereoMatching.cpp:25-48): gen_vm_from2vm_exp(vm, censusVm, gradVm, lamCen, lamG) with grad(.,500).
// ---------------------------------------------------------------------------
void orc_grad_xy(const u8* gray, int H, int W, float* gx, float* gy) {
  for (int v = 0; v < H; v++) {
    const u8* i = gray + (long)v * W;
    float* g = gx + (long)v * W;
    for (int u = 1; u < W - 1; u++) g[u] = 0.5 * (i[u + 1] - i[u - 1]);
    g[0] = i[1] - i[0];
    g[W - 1] = i[W - 1] - i[W - 2];
  }
  for (int v = 1; v < H - 1; v++)
    for (int u = 0; u < W; u++) gy[(long)v * W + u] = 0.5 * (gray[(long)(v + 1) * W + u] - gray[(long)(v - 1) * W + u]);
  for (int u = 0; u < W; u++) {
    gy[u] = gray[W + u] - gray[u];
    gy[(long)(H - 1) * W + u] = gray[(long)(H - 1) * W + u] - gray[(long)(H - 2) * W + u];
  }
}

void orc_grad_vol(const float* gx0, const float* gx1, const float* gy0, const float* gy1, const u16* armsView,
                  int H, int W, int D, int num, float Trunc, float* vol) {
  const int leftCoe = num == 1 ? 1 : 0, rightCoe = num == 1 ? 0 : -1;
  ORC_PAR_FOR
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      const short* arm = (const short*)(armsView + ((long)v * W + u) * 5);
      float shortestH = 10000, shortestV = 10000;
      for (int dir = 0; dir < 2; dir++) if (arm[dir] < shortestH) shortestH = arm[dir];
      for (int dir = 2; dir < 4; dir++) if (arm[dir] < shortestV) shortestV = arm[dir];
      if (shortestH == 0) shortestH++;
      if (shortestV == 0) shortestV++;
      const float a = shortestH / (shortestH + shortestV);
      float* o = vol + ((long)v * W + u) * D;
      for (int d = 0; d < D; d++) {
        const int u0 = u + leftCoe * d, u1 = u + rightCoe * d;
        if (u0 >= W || u1 < 0) { o[d] = std::sqrt(std::pow((double)Trunc, 2) * 2); continue; }
        const float dx = std::min(std::fabs(gx0[(long)v * W + u0] - gx1[(long)v * W + u1]), Trunc);
        const float dy = std::min(std::fabs(gy0[(long)v * W + u0] - gy1[(long)v * W + u1]), Trunc);
        o[d] = a * dx + (1 - a) * dy;
      }
    }
}

// ---------------------------------------------------------------------------
// Cross arms: calHorVerDis<uchar>(I,cross,L,L_out,C_D,C_D_out,minL)
// (stereoMatching.cpp:2958-3050) with judgeColorDif (stereoMatching.cpp:2847).
// cross: [H][W][5] u16 = [left,right,up,down,sum].  C = channels (3 for I_c).
// ---------------------------------------------------------------------------
static inline bool color_within(const u8* a, const u8* b, int thres, int C) {
  for (int c = 0; c < C; c++)
    if (std::abs((int)a[c] - (int)b[c]) > thres) return false;
  return true;
}

void orc_arms(const u8* img, int H, int W, int C, int L, int L_out, int C_D, int C_D_out, int minL,
              u16* cross) {
  static const int DU[4] = {-1, +1, 0, 0};  // left, right, up, down
  static const int DV[4] = {0, 0, -1, +1};
  ORC_PAR_FOR
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      const u8* p = img + ((long)v * W + u) * C;
      u16* out = cross + ((long)v * W + u) * 5;
      int sum = 0;
      for (int dir = 0; dir < 4; dir++) {
        const int du = DU[dir], dv = DV[dir];
        int arm = 1;
        for (; arm <= L_out; arm++) {
          int va = v + arm * dv, ua = u + arm * du;
          if (va < 0 || va >= H || ua < 0 || ua >= W) break;
          const u8* q = img + ((long)va * W + ua) * C;
          const u8* qp = img + ((long)(v + (arm - 1) * dv) * W + (u + (arm - 1) * du)) * C;
          bool nb_ok = color_within(q, qp, C_D, C);
          bool an_ok = color_within(p, q, arm <= L ? C_D : C_D_out, C);
          if (!nb_ok || !an_ok) break;
        }
        --arm;
        int res = 0;
        if (arm >= minL) res = arm;
        else {
          for (int len = minL; len >= 0; len--)
            if (u + len * du >= 0 && u + len * du <= W - 1 && v + len * dv >= 0 &&
                v + len * dv <= H - 1) { res = len; break; }
        }
        out[dir] = (u16)res;
        sum += res;
      }
      out[4] = (u16)sum;
    }
}

// Per-(pixel,d) intersected arms: genTrueHorVerArms (stereoMatching.cpp:2794-2845).
// view 0: u_l=u, u_r=u-d; view 1: u_l=u+d, u_r=u.  Out-of-range entries stay 0.
static inline void isect_arms(const u16* armsL, const u16* armsR, int W, int v, int u, int d,
                              int view, int a[4]) {
  int ul = view == 1 ? u + d : u, ur = view == 1 ? u : u - d;
  if (ur < 0 || ul >= W) { a[0] = a[1] = a[2] = a[3] = 0; return; }
  const u16* l = armsL + ((long)v * W + ul) * 5;
  const u16* r = armsR + ((long)v * W + ur) * 5;
  for (int k = 0; k < 4; k++) a[k] = std::min(l[k], r[k]);
}

// Materialised form (small sizes only; used by tests): [H][W][D][5] u16.
void orc_arms_intersect(const u16* armsL, const u16* armsR, int H, int W, int D, int view,
                        u16* out) {
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++)
      for (int d = 0; d < D; d++) {
        int a[4];
        isect_arms(armsL, armsR, W, v, u, d, view, a);
        u16* o = out + (((long)v * W + u) * D + d) * 5;
        for (int k = 0; k < 4; k++) o[k] = (u16)a[k];
        o[4] = (u16)(a[0] + a[1] + a[2] + a[3]);
      }
}

// ---------------------------------------------------------------------------
// CBCA: cbca_core (stereoMatching.cpp:5585-5666) for ONE view with
// cbca_intersect=true: per iteration areaIS<-1, then (even iteration) H cumsum,
// H span, V cumsum, V span, (odd iteration) V first then H; then vm /= areaIS.
// gen1DCumu: stereoMatching.cpp:3896-3926 (in-place raster running sum);
// cal1DCost: stereoMatching.h:1643-1715; genfinalVm_cbca: stereoMatching.cpp:3969-3992.
// The reference's H*W*D*5 intersection tensor is evaluated on the fly (same
// values, see isect_arms).  area_out (nullable) receives the last iteration's
// areaIS (int32, [H][W][D]).
// ---------------------------------------------------------------------------
static void cumu_1d(float* vm, int* area, int H, int W, int D, int dv, int du) {
  // Sequential dependence along the scan axis only; rows (du=-1) resp. columns
  // (dv=-1) are independent, which is what the parallel variant exploits.
  if (du == -1) {
    ORC_PAR_FOR
    for (int v = 0; v < H; v++)
      for (int u = 1; u < W; u++) {
        float* c = vm + ((long)v * W + u) * D;
        int* a = area + ((long)v * W + u) * D;
        for (int d = 0; d < D; d++) { c[d] += c[d - D]; a[d] += a[d - D]; }
      }
  } else {
    const long rs = (long)W * D;
    for (int v = 1; v < H; v++) {
      ORC_PAR_FOR
      for (int u = 0; u < W; u++) {
        float* c = vm + ((long)v * W + u) * D;
        int* a = area + ((long)v * W + u) * D;
        for (int d = 0; d < D; d++) { c[d] += c[d - rs]; a[d] += a[d - rs]; }
      }
    }
  }
}

static void span_1d(const float* vm, const int* area, float* vmT, int* areaT, const u16* armsL,
                    const u16* armsR, int H, int W, int D, int view, int dv, int du, int direc) {
  const int head_num = direc * 2 + 1, tail_num = direc * 2;
  ORC_PAR_FOR
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++)
      for (int d = 0; d < D; d++) {
        int a[4];
        isect_arms(armsL, armsR, W, v, u, d, view, a);
        int tail_u = u + du * a[tail_num], tail_v = v + dv * a[tail_num];
        int head_u = u - du * a[head_num], head_v = v - dv * a[head_num];
        int pu = tail_u + du, pv = tail_v + dv;
        bool inner = pu >= 0 && pu < W && pv >= 0 && pv < H;
        long ih = ((long)head_v * W + head_u) * D + d, ip = ((long)pv * W + pu) * D + d;
        long io = ((long)v * W + u) * D + d;
        areaT[io] = inner ? area[ih] - area[ip] : area[ih];
        vmT[io] = inner ? vm[ih] - vm[ip] : vm[ih];
      }
}

void orc_cbca(float* vol, const u16* armsL, const u16* armsR, int H, int W, int D, int iters,
              int view, int* area_out) {
  const long n = (long)H * W * D;
  std::vector<float> vmT(n);
  std::vector<int> area(n), areaT(n);
  for (int it = 0; it < iters; it++) {
    std::fill(area.begin(), area.end(), 1);
    for (int pass = 0; pass < 2; pass++) {
      int direc = (it % 2 == 0) ? pass : 1 - pass;  // 0 = horizontal, 1 = vertical
      int dv = direc == 0 ? 0 : -1, du = direc == 0 ? -1 : 0;
      cumu_1d(vol, area.data(), H, W, D, dv, du);
      span_1d(vol, area.data(), vmT.data(), areaT.data(), armsL, armsR, H, W, D, view, dv, du,
              direc);
      std::memcpy(vol, vmT.data(), n * sizeof(float));
      area.swap(areaT);
    }
    ORC_PAR_FOR
    for (long i = 0; i < n; i++) vol[i] /= area[i];
  }
  if (area_out) std::memcpy(area_out, area.data(), n * sizeof(int));
}

// ---------------------------------------------------------------------------
// SGM.  sgm (stereoMatching.cpp:6204-6224): predecessor offsets
// rv={+1,-1,0,0,+1,+1,-1,-1}, ru={0,0,+1,-1,-1,+1,+1,-1}; costScan
// (stereoMatching.cpp:1983-2029) raster order, reversed when rv>0 or
// (rv==0 && ru>0); updateCost<float> (stereoMatching.h:2205-2280): literal
// P1=1, P2=3, both /= reduCoeffi1 when the current-view colour step D1 >
// corDifThres; P1 -= minC; Lr = C + min4(Lr'[d]-minC, Lr'[d-1]+P1, Lr'[d+1]+P1, P2).
// bgr = colour image of the view being optimised (I_c[0] if leftFirst).
// ---------------------------------------------------------------------------
static const int SGM_RV[8] = {+1, -1, 0, 0, +1, +1, -1, -1};
static const int SGM_RU[8] = {0, 0, +1, -1, -1, +1, +1, -1};

void orc_sgm_path(const float* vol, const u8* bgr, int H, int W, int D, int rv, int ru,
                  int corDifThres, int reduCoeffi1, float* Lr) {
  int v0 = 0, v1 = H, u0 = 0, u1 = W, dv = +1, du = +1;
  if (rv > 0 || (rv == 0 && ru > 0)) { v0 = H - 1; v1 = -1; u0 = W - 1; u1 = -1; dv = -1; du = -1; }
  const float FMAX = std::numeric_limits<float>::max();
  for (int v = v0; v != v1; v += dv)
    for (int u = u0; u != u1; u += du) {
      const float* C = vol + ((long)v * W + u) * D;
      float* out = Lr + ((long)v * W + u) * D;
      bool inner = !(v + rv > H - 1 || v + rv < 0 || u + ru > W - 1 || u + ru < 0);
      if (!inner) { for (int d = 0; d < D; d++) out[d] = C[d]; continue; }
      int D1 = 0;
      const u8* p = bgr + ((long)v * W + u) * 3;
      const u8* q = bgr + ((long)(v + rv) * W + (u + ru)) * 3;
      for (int c = 0; c < 3; c++) D1 = std::max(D1, std::abs((int)p[c] - (int)q[c]));
      const float* prev = Lr + ((long)(v + rv) * W + (u + ru)) * D;
      float minC = FMAX;
      for (int d = 0; d < D; d++) minC = std::min(prev[d], minC);
      float P1 = 1.0f, P2 = 3.0f;
      if (D1 > corDifThres) { P1 /= reduCoeffi1; P2 /= reduCoeffi1; }
      P1 -= minC;
      for (int d = 0; d < D; d++) {
        float S1 = prev[d] - minC;
        float S2 = d - 1 >= 0 ? prev[d - 1] + P1 : FMAX;
        float S3 = d + 1 < D ? prev[d + 1] + P1 : FMAX;
        float S4 = P2;
        out[d] = C[d] + std::min(std::min(S1, S2), std::min(S3, S4));
      }
    }
}

// sgm + gen_sgm_vm (stereoMatching.cpp:2031-2056): vm = ((0+L0)+L1)+... , no
// averaging.  P = number of paths (reference compiles 4; 8 = same table).
// The paths are independent of each other, so they may run on separate threads.
void orc_sgm(float* vol, const u8* bgr, int H, int W, int D, int P, int corDifThres,
             int reduCoeffi1) {
  const long n = (long)H * W * D;
  std::vector<std::vector<float>> L(P);
  for (int i = 0; i < P; i++) L[i].resize(n);
  ORC_PAR_FOR
  for (int i = 0; i < P; i++)
    orc_sgm_path(vol, bgr, H, W, D, SGM_RV[i], SGM_RU[i], corDifThres, reduCoeffi1, L[i].data());
  ORC_PAR_FOR
  for (long i = 0; i < n; i++) {
    float sum = 0;
    for (int k = 0; k < P; k++) sum += L[k][i];
    vol[i] = sum;
  }
}

// ---------------------------------------------------------------------------
// WTA: gen_dispFromVm (stereoMatching.cpp:3928-3967), ChooseSmall=true:
// strict '>' so the lowest d wins ties; -1 when nothing beats FLT_MAX.
// ---------------------------------------------------------------------------
// vmTop, second half: the disparity of a pixel from its candidate list
// topDisp[v][u][k] = {d, cost}, k < num, topDisp[v][u][num][0] = candidate count (selectTopCostFromVolumn).
//
// version 1 = genDispFromTopCostVm (stereoMatching.h:2466-2545): own candidates + those of the left / right
//   neighbour vote per disparity (count, then summed cost); note the reference's `dNum = dispNum && cost_ < cost`
//   (an assignment, :2531): the tie branch sets the running count to 1.  Reads topDisp only.
// version 2 = genDispFromTopCostVm2 (stereoMatching.cpp:1514-1886), param_.vmTop_method 0 / 1 / 2.  Method 0:
//   border pixels and single-candidate pixels take candidate 0; otherwise the candidates that have another candidate
//   within `ts` disparities (vmTop_hasCir2) are kept, keyed by cost in a std::map (equal cost keys: first insertion
//   wins); none kept -> the candidate nearest to one of the already written neighbours (left, up, up-left, up-right:
//   the raster-order dependency); else the kept candidates + the 8 neighbours' candidates of the same disparity vote
//   (count, then summed cost, then lowest disparity).  Methods 1 / 2: row scans against the previous pixel's result.
// std::map<int, .> iteration = ascending disparity; std::map<float, int> iteration = ascending cost.
// ---------------------------------------------------------------------------
namespace {
struct TopView {
  const float* top; int H, W, num;
  const float* at(int v, int u, int k) const { return top + (((long)v * W + u) * (num + 1) + k) * 2; }
  int count(int v, int u) const { return (int)at(v, u, num)[0]; }
};
struct Vote { int d; int n; float c; };
static void vote_add(std::vector<Vote>& vs, int d, float c, bool create) {
  for (auto& x : vs) if (x.d == d) { x.n++; x.c += c; return; }
  if (create) vs.push_back({d, 1, 0.f + c});   // map's operator[] value-initialises to 0, then += cost
}
}  // namespace

void orc_disp_from_top(const float* top, const u8* bgr, int H, int W, int num, int version, int method, int ts,
                       int hasCir2, int colorLimit, i16* disp) {
  TopView T{top, H, W, num};
  if (version == 1) {
    for (int v = 0; v < H; v++)
      for (int u = 0; u < W; u++) {
        const float cntf = T.at(v, u, num)[0];
        if (cntf == 1) { disp[(long)v * W + u] = (i16)T.at(v, u, 0)[0]; continue; }
        if (!(cntf > 1)) continue;
        std::vector<Vote> vs;
        for (int i = 0; i < cntf; i++) vote_add(vs, (int)T.at(v, u, i)[0], T.at(v, u, i)[1], true);
        for (int du = -1; du <= 1; du += 2) {
          const int u_ = u + du;
          if (u_ < 0 || u_ >= W) continue;
          const int n_ = T.count(v, u_);
          for (int k = 0; k < n_; k++) vote_add(vs, (int)T.at(v, u_, k)[0], T.at(v, u_, k)[1], false);
        }
        std::sort(vs.begin(), vs.end(), [](const Vote& a, const Vote& b) { return a.d < b.d; });
        int best = -1, bestN = -1;
        float bestC = std::numeric_limits<float>::max();
        for (const auto& x : vs) {
          int dNum = x.n;
          bool take = dNum > bestN;
          if (!take) { dNum = (bestN && x.c < bestC) ? 1 : 0; take = dNum != 0; }   // `dNum = dispNum && cost_ < cost`
          if (take) { bestN = dNum; bestC = x.c; best = x.d; }
        }
        disp[(long)v * W + u] = (i16)best;
      }
    return;
  }
  if (method == 0) {
    static const int NV[8] = {0, -1, 0, 1, -1, 1, -1, 1}, NU[8] = {-1, 0, 1, 0, -1, 1, 1, -1};
    for (int v = 0; v < H; v++)
      for (int u = 0; u < W; u++) {
        i16& out = disp[(long)v * W + u];
        if (u == 0 || v == 0) { out = (i16)T.at(v, u, 0)[0]; continue; }
        const int n = T.count(v, u);
        if (n == 1) { out = (i16)T.at(v, u, 0)[0]; continue; }
        if (n < 1) continue;
        // candidates kept, keyed by cost (std::map<float,int>::insert keeps the first entry of an equivalent key)
        std::vector<std::pair<float, int>> kept;
        auto insert = [&](float c, int d) {
          for (auto& e : kept) if (!(e.first < c) && !(c < e.first)) return;
          kept.push_back({c, d});
        };
        for (int i = 0; i < n; i++) {
          const int d0 = (int)T.at(v, u, i)[0];
          const float c0 = T.at(v, u, i)[1];
          if (!hasCir2) { insert(c0, d0); continue; }
          for (int j = i + 1; j < n; j++) {
            const int d1 = (int)T.at(v, u, j)[0];
            if (std::abs(d0 - d1) < ts) { insert(c0, d0); insert(T.at(v, u, j)[1], d1); }
          }
        }
        if (kept.empty()) {
          const int pre1 = disp[(long)v * W + u - 1], pre2 = disp[(long)(v - 1) * W + u], lt = disp[(long)(v - 1) * W + u - 1];
          const int rt = u != W - 1 ? disp[(long)(v - 1) * W + u + 1] : 10000;
          const int ref[4] = {pre1, pre2, rt, lt};
          int small[4], pick[4];
          for (int k = 0; k < 4; k++) { small[k] = std::numeric_limits<int>::max(); pick[k] = -1; }
          for (int i = 0; i < n; i++) {
            const int dd = (int)T.at(v, u, i)[0];
            for (int k = 0; k < 4; k++) {
              const int dif = std::abs(dd - ref[k]);
              if (dif < small[k]) { small[k] = dif; pick[k] = dd; }
            }
          }
          const int m = std::min(std::min(small[2], small[3]), std::min(small[0], small[1]));
          int d = -1;
          if (m == small[3]) d = pick[3];
          else if (m == small[0]) d = pick[0];
          else if (m == small[1]) d = pick[1];
          else if (m == small[2]) d = pick[2];
          out = m < 1000 ? (i16)d : (i16)T.at(v, u, 0)[0];
        } else {
          std::sort(kept.begin(), kept.end(), [](const std::pair<float, int>& a, const std::pair<float, int>& b) { return a.first < b.first; });
          std::vector<Vote> vs;
          for (auto& e : kept) vote_add(vs, e.second, e.first, true);
          const u8* tar = bgr + ((long)v * W + u) * 3;
          for (int k = 0; k < 8; k++) {
            const int v_ = v + NV[k], u_ = u + NU[k];
            if (v_ < 0 || v_ >= H || u_ < 0 || u_ >= W) continue;
            if (colorLimit) {
              const u8* nei = bgr + ((long)v_ * W + u_) * 3;
              bool ok = true;
              for (int c = 0; c < 3; c++) if (std::abs((int)tar[c] - (int)nei[c]) > 10) { ok = false; break; }
              if (!ok) continue;
            }
            const int n_ = T.count(v_, u_);
            for (int x = 0; x < n_; x++) vote_add(vs, (int)T.at(v_, u_, x)[0], T.at(v_, u_, x)[1], false);
          }
          std::sort(vs.begin(), vs.end(), [](const Vote& a, const Vote& b) { return a.d < b.d; });
          int best = -1, bestN = -1;
          float bestC = std::numeric_limits<float>::max();
          for (const auto& x : vs)
            if (x.n > bestN || (x.n == bestN && x.c < bestC)) { bestN = x.n; bestC = x.c; best = x.d; }
          out = (i16)best;
        }
      }
  } else if (method == 1) {
    for (int v = 0; v < H; v++)
      for (int u = 0; u < W; u++) {
        i16& out = disp[(long)v * W + u];
        const int n = T.count(v, u);
        if (u == 0 || n == 1) { out = (i16)T.at(v, u, 0)[0]; continue; }
        int dp = -1, best = 10000;
        const i16 pre = disp[(long)v * W + u - 1];
        for (int k = 0; k < n; k++) {
          const int s_ = (int)std::fabs((float)pre - T.at(v, u, k)[0]);   // abs(short - float) -> float, then int
          if (s_ < 2 && s_ < best) { best = s_; dp = (int)T.at(v, u, k)[0]; }
        }
        out = dp == -1 ? (i16)T.at(v, u, 0)[0] : (i16)dp;
      }
  } else if (method == 2) {
    for (int v = 0; v < H; v++)
      for (int u = 0; u < W; u++) {
        i16& out = disp[(long)v * W + u];
        const int n = T.count(v, u);
        if (u == 0 || n == 1) { out = (i16)T.at(v, u, 0)[0]; continue; }
        int bestPre = 1000000, bestAft = 1000000, d0 = -1, d1 = -1;
        const int pre = disp[(long)v * W + u - 1];
        for (int k = 0; k < n; k++) {
          const int dif = (int)std::fabs(T.at(v, u, k)[0] - (float)pre);
          if (dif < 2 && dif < bestPre) { bestPre = dif; d0 = (int)T.at(v, u, k)[0]; }
        }
        if (u < W - 1) {
          const int aft = (int)T.at(v, u + 1, 0)[0];
          for (int k = 0; k < n; k++) {
            const int dif = (int)std::fabs(T.at(v, u, k)[0] - (float)aft);
            if (dif < 2 && dif < bestAft) { bestAft = dif; d1 = (int)T.at(v, u, k)[0]; }
          }
        }
        if (d0 != -1 && d1 == -1) out = (i16)d0;
        else if (d0 == -1 && d1 != -1) out = (i16)d1;
        else if (d0 == -1 && d1 == -1) out = (i16)T.at(v, u, 0)[0];
        else {
          int cpre = 0, caft = 0;
          const u8* c = bgr + ((long)v * W + u) * 3;
          for (int k = 0; k < 3; k++) { cpre += std::abs((int)c[k] - (int)c[k - 3]); caft += std::abs((int)c[k] - (int)c[k + 3]); }
          out = cpre <= caft ? (i16)d0 : (i16)d1;
        }
      }
  }
}

// ---------------------------------------------------------------------------
void orc_wta(const float* vol, int H, int W, int D, i16* disp) {
  ORC_PAR_FOR
  for (long i = 0; i < (long)H * W; i++) {
    float minC = std::numeric_limits<float>::max();
    int best = -1;
    const float* c = vol + i * D;
    for (int d = 0; d < D; d++)
      if (minC > c[d]) { minC = c[d]; best = d; }
    disp[i] = (i16)best;
  }
}

// subpixelEnhancement (stereoMatching.cpp:6138-6166): parabola offset from the costs at disp-1, disp, disp+1, only for
// 0 < disp < D-1, denom != 0 and -1 < diff < 1.  `disp -= diff` acts on the short: (short)((float)disp - diff), so the
// float map that comes out is integer-valued.
void orc_subpixel(const short* disp, const float* vol, int H, int W, int D, float* out) {
  ORC_PAR_FOR
  for (long i = 0; i < (long)H * W; i++) {
    short d = disp[i];
    if (d > 0 && d < D - 1) {
      const float* c = vol + i * D;
      float cost = c[d], costPlus = c[d + 1], costMinus = c[d - 1];
      float denom = 2 * (costPlus + costMinus - 2 * cost);
      if (denom != 0) {
        float diff = (costPlus - costMinus) / denom;
        if (diff > -1 && diff < 1) d = (short)((float)d - diff);
      }
    }
    out[i] = (float)d;
  }
}

// selectTopCostFromVolumn (stereoMatching.h:2405-2461; the caller hands it a clone of vm,
// stereoMatching.cpp:1118-1119): per pixel, up to num rounds of a first-minimum scan (strict '>'); the winner is taken
// out by overwriting it with FLT_MAX; round 0 is always kept, round k > 0 only while cost < firstCost * thres.
// top = float [H][W][num+1][2], zero where the reference writes nothing; top[..][num][0] = candidate count.
void orc_select_top(const float* vol, int H, int W, int D, int num, float thres, float* top) {
  ORC_PAR_FOR
  for (long i = 0; i < (long)H * W; i++) {
    std::vector<float> c(vol + i * D, vol + i * D + D);
    float* o = top + i * (long)(num + 1) * 2;
    for (int k = 0; k < (num + 1) * 2; k++) o[k] = 0.f;
    float firstV = 0.f;
    for (int k = 0; k < num; k++) {
      float m = c[0];
      int disp = 0;
      for (int d = 1; d < D; d++)
        if (m > c[d]) { m = c[d]; disp = d; }
      if (k == 0) firstV = m;
      else if (!(m < firstV * thres)) break;
      c[disp] = std::numeric_limits<float>::max();
      o[2 * num] += 1.f;
      o[2 * k] = (float)disp;
      o[2 * k + 1] = m;
    }
  }
}

// wta_Co (stereoMatching.cpp:2709-2792) with UniqCk=SubIpl=0: left WTA over
// d<=u, right map from the LEFT volume along the diagonal, both x DISP_SCALE.
void orc_wta_co(const float* vol, int H, int W, int D, int scale, i16* D1, i16* D2) {
  ORC_PAR_FOR
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      float mR = std::numeric_limits<float>::max(), mL = mR;
      int dR = 0, dL = 0;
      for (int d = 0; d < D; d++) {
        if (u + d >= W) break;
        float c = vol[((long)v * W + u + d) * D + d];
        if (c < mR) { mR = c; dR = d; }
      }
      D2[(long)v * W + u] = (i16)(dR * scale);
      const float* c = vol + ((long)v * W + u) * D;
      for (int d = 0; d < D; d++) {
        if (u - d < 0) break;
        if (c[d] < mL) { dL = d; mL = c[d]; }
      }
      D1[(long)v * W + u] = (i16)(dL * scale);
    }
}

// LRConsistencyCheck_normal (stereoMatching.cpp:2262-2282): invalid -> -1.
void orc_lrc_normal(i16* D1, const i16* D2, int H, int W, float maxDiff) {
  ORC_PAR_FOR
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      i16 d = D1[(long)v * W + u];
      if (d < 0 || u - d < 0 || std::abs(d - D2[(long)v * W + u - d]) > maxDiff)
        D1[(long)v * W + u] = -1;
    }
}

// LRConsistencyCheck, LOR=0 branch (stereoMatching.cpp:2284-2335): labels
// occluded (no D2[u-d']==d' exists) vs mismatched, writes an error mask.
void orc_lrc_label(i16* D1, const i16* D2, int H, int W, int D, float maxDiff, int DISP_OCC,
                   int DISP_MIS, u8* errMask) {
  ORC_PAR_FOR
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      long i = (long)v * W + u;
      i16 d = D1[i];
      errMask[i] = 0;
      if (d < 0 || u - d < 0 || std::abs(d - D2[i - d]) > maxDiff) {
        errMask[i] = 255;
        int disp = DISP_OCC;
        for (int dd = 0; dd < D && u - dd >= 0; dd++)
          if (D2[i - dd] == dd) { disp = DISP_MIS; break; }
        D1[i] = (i16)disp;
      }
    }
}

// LRConsistencyCheck, LOR=1 branch (stereoMatching.cpp:2336-2364): the right map checked against the left one and
// labelled in place (D1 is only read).  errMask1 = the flags of the reference's local mask (its errMask stays zero).
void orc_lrc_label_right(const i16* D1, i16* D2, int H, int W, int D, float maxDiff, int DISP_OCC, int DISP_MIS,
                         u8* errMask1) {
  ORC_PAR_FOR
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      long i = (long)v * W + u;
      i16 d = D2[i];
      errMask1[i] = 0;
      if (d < 0 || u + d >= W || std::abs(d - D1[i + d]) > maxDiff) {
        int disp = DISP_OCC;
        for (int dd = 0; dd < D && u + dd <= W - 1; dd++)
          if (D1[i + dd] == dd) { disp = DISP_MIS; break; }
        D2[i] = (i16)disp;
        errMask1[i] = 255;
      }
    }
}

// regionVote_my (stereoMatching.cpp:7219-7277).  arms = HVL[0] (image-space
// arms of the LEFT image, not intersected).  Note hist[mode]/validNum is an
// integer division (line 7270).  Jacobi update (reads Dp, writes a clone).
void orc_region_vote(i16* Dp, const u16* arms, int H, int W, int D, float ratio, int S) {
  std::vector<i16> res(Dp, Dp + (long)H * W);
  ORC_PAR_FOR
  for (int v = 0; v < H; v++) {
    std::vector<int> hist(D);
    for (int u = 0; u < W; u++) {
      if (Dp[(long)v * W + u] >= 0) continue;
      std::fill(hist.begin(), hist.end(), 0);
      int valid = 0;
      const u16* a = arms + ((long)v * W + u) * 5;
      for (int vn = v - a[2]; vn <= v + a[3]; vn++) {
        const u16* b = arms + ((long)vn * W + u) * 5;
        for (int un = u - b[0]; un <= u + b[1]; un++) {
          i16 x = Dp[(long)vn * W + un];
          if (x >= 0) { valid++; hist[x]++; }
        }
      }
      if (valid <= S) continue;
      int most = 0;
      for (int d = 1; d < D; d++)
        if (hist[d] > hist[most]) most = d;
      if (hist[most] / valid >= ratio) res[(long)v * W + u] = (i16)most;
    }
  }
  std::memcpy(Dp, res.data(), sizeof(i16) * H * W);
}

// properIpol (stereoMatching.cpp:7395-7490): 16 directions, <= 20 half/full
// steps, first valid pixel per direction; DISP_OCC pixels take the minimum
// disparity, others the disparity of the smallest max-channel colour distance
// (strictly below 255).  Jacobi.
void orc_proper_ipol(i16* Dp, const u8* bgr, int H, int W, int DISP_OCC) {
  static const int dirW[16] = {0, 2, 2, 2, 0, -2, -2, -2, 1, 2, 2, 1, -1, -2, -2, -1};
  static const int dirH[16] = {2, 2, 0, -2, -2, -2, 0, 2, 2, 1, -1, -2, -2, -1, 1, 2};
  std::vector<i16> res((long)H * W);
  ORC_PAR_FOR
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      long i = (long)v * W + u;
      i16 cur = Dp[i];
      if (cur >= 0) { res[i] = cur; continue; }
      int dDisp[16], dDiff[16];
      for (int k = 0; k < 16; k++) {
        dDisp[k] = -1; dDiff[k] = -1;
        int pw = dirW[k], ph = dirH[k], x = u, y = v;
        for (int dep = 0; dep < 20; dep++) {
          if (dep % 2 == 0) { x += pw / 2; y += ph / 2; }
          else { x += pw - pw / 2; y += ph - ph / 2; }
          if (!(x >= 0 && x < W && y >= 0 && y < H)) break;
          i16 q = Dp[(long)y * W + x];
          if (q >= 0) {
            dDisp[k] = q;
            int cd = 0;
            for (int c = 0; c < 3; c++)
              cd = std::max(cd, std::abs((int)bgr[i * 3 + c] - (int)bgr[((long)y * W + x) * 3 + c]));
            dDiff[k] = cd;
            break;
          }
        }
      }
      if (cur == DISP_OCC) {
        int m = std::numeric_limits<int>::max();
        for (int k = 0; k < 16; k++)
          if (dDisp[k] >= 0 && m > dDisp[k]) m = dDisp[k];
        res[i] = (i16)(m != std::numeric_limits<int>::max() ? m : cur);
      } else {
        int mc = 255, disp = -1;
        for (int k = 0; k < 16; k++)
          if (dDiff[k] >= 0 && mc > dDiff[k]) { mc = dDiff[k]; disp = dDisp[k]; }
        res[i] = (i16)(disp >= 0 ? disp : cur);
      }
    }
  std::memcpy(Dp, res.data(), sizeof(i16) * H * W);
}

// cv::medianBlur(CV_16S, ksize 3) == 3x3 median with replicated border
// (stereoMatching.cpp:1499; pinned against cv2 in tests/golden).
void orc_median3_i16(const i16* src, int H, int W, i16* dst) {
  ORC_PAR_FOR
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      i16 w[9];
      int k = 0;
      for (int dv = -1; dv <= 1; dv++)
        for (int du = -1; du <= 1; du++) {
          int y = std::min(std::max(v + dv, 0), H - 1), x = std::min(std::max(u + du, 0), W - 1);
          w[k++] = src[(long)y * W + x];
        }
      std::nth_element(w, w + 4, w + 9);
      dst[(long)v * W + u] = w[4];
    }
}

// WM (stereoMatching.cpp:7340-7393; default off, Do_WM): on every mask > 0 pixel, a 19x19 (REFLECT_101) bilateral
// weighted median of the labels: w = exp(-|dI|^2 / 25^2 - |dx|^2 / 9^2) evaluated in float from exactly converted
// integers, histogram and running sums in float in window raster order, first d with cum >= sum / 2.  The window
// reads the ORIGINAL padded copy of the map, the result goes to the map.  DEFINED ONLY when every label inside the
// windows is in [0, D): the reference indexes dispHist[q] unchecked (:7371); this restatement then restores the map and
// returns -1.  Otherwise returns the number of pixels rewritten.
static int wm_impl(short* disp, const unsigned char* mask, const unsigned char* bgr, int H, int W, int D, bool lenient, int* numInvalid) {
  const int R = 9;
  auto refl = [](int p, int n) { if (n == 1) return 0; while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p; return p; };
  std::vector<short> src(disp, disp + (long)H * W);
  int num = 0;
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      if (!mask[(long)v * W + u]) continue;
      std::vector<float> hist(D, 0.f);
      float wsum = 0.f;
      const unsigned char* ip = bgr + ((long)v * W + u) * 3;
      for (int dv = -R; dv <= R; dv++)
        for (int du = -R; du <= R; du++) {
          const long j = (long)refl(v + dv, H) * W + refl(u + du, W);
          const short q = src[j];
          const bool bad = q < 0 || q >= D;
          if (bad && !lenient) { std::memcpy(disp, src.data(), sizeof(short) * H * W); return -1; }
          if (bad && numInvalid) ++*numInvalid;
          const unsigned char* iq = bgr + j * 3;
          int c2 = 0;
          for (int c = 0; c < 3; c++) c2 += ((int)ip[c] - (int)iq[c]) * ((int)ip[c] - (int)iq[c]);
          const float colDis = (float)c2, spaDis = (float)(dv * dv + du * du);
          const float wgt = std::exp(-colDis / (25.f * 25.f) - spaDis / (9.f * 9.f));
          if (!bad) hist[q] += wgt;   // lenient: the weight still enters the total, as in the reference, but casts no vote
          wsum += wgt;
        }
      const float half = wsum / 2;
      float cum = 0.f;
      for (int d = 0; d < D; d++) {
        cum += hist[d];
        if (cum >= half) { disp[(long)v * W + u] = (short)d; num++; break; }
      }
    }
  return num;
}
int orc_wm(short* disp, const unsigned char* mask, const unsigned char* bgr, int H, int W, int D) {
  return wm_impl(disp, mask, bgr, H, W, D, false, nullptr);
}
// The defined extension the CUDA entry point sm_wm implements for labels outside [0, D) (undefined in the reference):
// such a neighbour adds to the total weight but to no bin.  *numInvalid = how many were seen.
int orc_wm_lenient(short* disp, const unsigned char* mask, const unsigned char* bgr, int H, int W, int D, int* numInvalid) {
  *numInvalid = 0;
  return wm_impl(disp, mask, bgr, H, W, D, true, numInvalid);
}

// discontinuityAdjust (stereoMatching.cpp:6057-6135; Do_discontinuityAdjust, stereoMatching.h:78, off): the map as an
// 8-bit picture (saturating convertTo) -> equalizeHist -> GaussianBlur(3x3, sigma 4) -> Canny(20, 60, 3); an edge pixel
// whose 3x3 edge neighbourhood names a direction takes the disparity of whichever of itself and the two neighbours ACROSS
// that direction has the smallest cost in vm[0] -- in place and in raster order, so the first neighbour (above or to the
// left) may already carry its adjusted value.  edge (optional) receives the Canny map.  Returns the number of labels
// >= D met on the way (the reference would index vm[0] out of bounds there: undefined; here such a centre pixel is left
// alone and such a neighbour is not a candidate -- the same defined extension sm_discontinuity_adjust implements).
void orc_da_edges(const short* disp, int H, int W, u8* edge) {
  std::vector<u8> a((size_t)H * W), b((size_t)H * W);
  for (long i = 0; i < (long)H * W; i++) a[i] = (u8)(disp[i] < 0 ? 0 : disp[i] > 255 ? 255 : disp[i]);
  orc_cv::equalize_hist(a.data(), b.data(), H, W);
  orc_cv::gauss3_sigma4(b.data(), a.data(), H, W);
  orc_cv::canny3_l1(a.data(), edge, H, W, 20, 60);
}
void orc_equalize_hist(const u8* src, int H, int W, u8* dst) { orc_cv::equalize_hist(src, dst, H, W); }
void orc_gauss3_sigma4(const u8* src, int H, int W, u8* dst) { orc_cv::gauss3_sigma4(src, dst, H, W); }
void orc_canny3_l1(const u8* src, int H, int W, int low, int high, u8* dst) { orc_cv::canny3_l1(src, dst, H, W, low, high); }
int orc_disc_adjust(short* disp, const float* vol, int H, int W, int D, u8* edge_out) {
  std::vector<u8> E((size_t)H * W);
  orc_da_edges(disp, H, W, E.data());
  if (edge_out) memcpy(edge_out, E.data(), E.size());
  static const int dH[8] = {-1, 1, -1, 1, -1, 1, 0, 0}, dW[8] = {-1, 1, 0, 0, 1, -1, -1, 1};
  int bad = 0;
  auto e = [&](int v, int u) { return E[(size_t)v * W + u] != 0; };
  for (int v = 1; v < H - 1; v++)
    for (int u = 1; u < W - 1; u++) {
      if (!e(v, u)) continue;
      int dir = -1;
      if (e(v - 1, u - 1) && e(v + 1, u + 1)) dir = 4;
      else if (e(v - 1, u + 1) && e(v + 1, u - 1)) dir = 0;
      else if (e(v - 1, u) || e(v - 1, u - 1) || e(v - 1, u + 1)) {
        if (e(v + 1, u) || e(v + 1, u - 1) || e(v + 1, u + 1)) dir = 6;
      } else if ((e(v - 1, u - 1) || e(v, u - 1) || e(v + 1, u - 1)) && (e(v - 1, u + 1) || e(v, u + 1) || e(v + 1, u + 1)))
        dir = 2;
      if (dir < 0) continue;
      short dp = disp[(size_t)v * W + u];
      if (dp < 0) continue;
      if (dp >= D) { bad++; continue; }
      float cost = vol[((size_t)v * W + u) * D + dp];
      const int v1 = v + dH[dir], u1 = u + dW[dir], v2 = v + dH[dir + 1], u2 = u + dW[dir + 1];
      const short d1 = disp[(size_t)v1 * W + u1], d2 = disp[(size_t)v2 * W + u2];
      if (d1 >= D) bad++;
      if (d2 >= D) bad++;
      const float cost1 = (d1 >= 0 && d1 < D) ? vol[((size_t)v1 * W + u1) * D + d1] : -1;
      const float cost2 = (d2 >= 0 && d2 < D) ? vol[((size_t)v2 * W + u2) * D + d2] : -1;
      if (cost1 >= 0 && cost1 < cost) { dp = d1; cost = cost1; }
      if (cost2 != -1 && cost2 < cost) dp = d2;
      disp[(size_t)v * W + u] = dp;
    }
  return bad;
}

// cv::medianBlur(CV_32F, ksize 3) on the sub-pixel map (stereoMatching.cpp:1490): same window, replicated border
// (pinned against cv2 in tests/golden/subpixel_ref.npz).
void orc_median3_f32(const float* src, int H, int W, float* dst) {
  ORC_PAR_FOR
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      float w[9];
      int k = 0;
      for (int dv = -1; dv <= 1; dv++)
        for (int du = -1; du <= 1; du++) {
          int y = std::min(std::max(v + dv, 0), H - 1), x = std::min(std::max(u + du, 0), W - 1);
          w[k++] = src[(long)y * W + x];
        }
      std::nth_element(w, w + 4, w + 9);
      dst[(long)v * W + u] = w[4];
    }
}

// The caller's single-level cross-scale step: SolveAll with PY_LVL=1
// (stereoMatching.cpp:2142-2208; main_.cpp:158): regInv = 1/(1+lambda) as a
// float, vm = 0 + invWgt*vm.
void orc_solve_all_1level(float* vol, long n, float lambda) {
  float inv = (float)(1. / (double)(1.0f + lambda));  // 1x1 Mat::inv of (1+lambda): cv::invert rounds the double reciprocal
  ORC_PAR_FOR
  for (long i = 0; i < n; i++) { float sum = 0; sum += inv * vol[i]; vol[i] = sum; }
}

// calErr<short> (stereoMatching.h:1748-1825) for one region mask: float accumulation of pow(dif, 2) in pixel order,
// invalid pixels (DP < 0) count as errors and add 2.  out[0] = PBM, out[1] = rms, counts[0] = sumNum, [1] = errorNumer.
void orc_cal_err(const i16* dp, const float* dt, const u8* mask, int H, int W, int thres, float* out, long* counts) {
  int sumNum = 0, errorNumer = 0;
  float errorValueSum = 0;
  for (int v = 0; v < H; v++)
    for (int u = 0; u < W; u++) {
      const long i = (long)v * W + u;
      if (mask[i] != 255) continue;
      sumNum++;
      if (dp[i] >= 0) {
        float dif = std::abs(dt[i] - dp[i]);
        errorValueSum += std::pow(dif, 2);
        if (dif > thres) errorNumer++;
      } else {
        errorNumer++;
        errorValueSum += 2;
      }
    }
  out[0] = (float)errorNumer / sumNum;
  out[1] = std::sqrt(errorValueSum / sumNum);
  counts[0] = sumNum; counts[1] = errorNumer;
}

// ---------------------------------------------------------------------------
// Cross-scale step of the caller: the pyramid loop of main_.cpp:131-158 and
// SolveAll (stereoMatching.cpp:2142-2208).
//   cv::pyrDown on 8-bit images: separable [1 4 6 4 1]/16, BORDER_REFLECT_101, dst = ((W+1)/2, (H+1)/2),
//     (sum + 128) >> 8  [probe: equal to cv2.pyrDown for odd/even/1-wide shapes, tests/golden/opencv_semantics.npz].
//   regMat (float): tridiagonal {1+l, -l; -l, 1+2l, -l; ...; -l, 1+l}; invWgt = row 0 of regMat.inv().
//     cv::invert(DECOMP_LU) on CV_32F: n = 1, 3 closed forms evaluated in double and rounded once, n = 2 determinant
//     in double then float products with (float)(1/det) (the SIMD128 branch); n > 3 Gaussian
//     elimination with partial pivoting in float (hal::LU32f)  [probe: both reproduce cv2.invert bit for bit].
//   SolveAll: vm0[y][x][d] = sum_s invWgt[s] * vm_s[y>>s][x>>s][d_s], d_0 = d, d_{s+1} = (d_s + 1) / 2, float
//     accumulation in level order starting from 0.
// ---------------------------------------------------------------------------
void orc_pyr_down_u8(const u8* src, int H, int W, int cn, u8* dst) {
  static const int k[5] = {1, 4, 6, 4, 1};
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  for (int y = 0; y < Ho; y++)
    for (int x = 0; x < Wo; x++)
      for (int c = 0; c < cn; c++) {
        int s = 0;
        for (int i = 0; i < 5; i++) {
          const int yy = reflect101(2 * y + i - 2, H);
          for (int j = 0; j < 5; j++) s += k[i] * k[j] * src[((long)yy * W + reflect101(2 * x + j - 2, W)) * cn + c];
        }
        dst[((long)y * Wo + x) * cn + c] = (u8)((s + 128) >> 8);
      }
}

void orc_cross_scale_weights(int n, float lambda, float* invWgt) {
  std::vector<float> M((size_t)n * n, 0.f);
  for (int s = 0; s < n; s++) {
    if (s == 0) { M[0] = 1 + lambda; if (n > 1) M[1] = -lambda; }
    else if (s == n - 1) { M[(size_t)s * n + s] = 1 + lambda; M[(size_t)s * n + s - 1] = -lambda; }
    else { M[(size_t)s * n + s] = 1 + 2 * lambda; M[(size_t)s * n + s - 1] = -lambda; M[(size_t)s * n + s + 1] = -lambda; }
  }
  if (n == 1) { invWgt[0] = (float)(1. / (double)M[0]); return; }
  if (n == 2) {
    double d = (double)M[0] * M[3] - (double)M[1] * M[2];
    d = 1. / d;
    // the CV_SIMD128 branch of cv::invert's 2x2 float case multiplies in float by (float)d  [probe: 300 lambdas]
    invWgt[0] = M[3] * (float)d;
    invWgt[1] = -(M[1] * (float)d);
    return;
  }
  if (n == 3) {
    auto S = [&](int r, int c) { return (double)M[r * 3 + c]; };
    double d = S(0, 0) * (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) - S(0, 1) * (S(1, 0) * S(2, 2) - S(1, 2) * S(2, 0)) +
               S(0, 2) * (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0));
    d = 1. / d;
    invWgt[0] = (float)((S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) * d);
    invWgt[1] = (float)((S(0, 2) * S(2, 1) - S(0, 1) * S(2, 2)) * d);
    invWgt[2] = (float)((S(0, 1) * S(1, 2) - S(0, 2) * S(1, 1)) * d);
    return;
  }
  std::vector<float> b((size_t)n * n, 0.f);
  for (int i = 0; i < n; i++) b[(size_t)i * n + i] = 1.f;
  float* A = M.data();
  for (int i = 0; i < n; i++) {
    int k = i;
    for (int j = i + 1; j < n; j++) if (std::fabs(A[j * n + i]) > std::fabs(A[k * n + i])) k = j;
    if (k != i) for (int c = 0; c < n; c++) { std::swap(A[i * n + c], A[k * n + c]); std::swap(b[i * n + c], b[k * n + c]); }
    const float d = -1 / A[i * n + i];
    for (int j = i + 1; j < n; j++) {
      const float alpha = A[j * n + i] * d;
      for (int c = i + 1; c < n; c++) A[j * n + c] += alpha * A[i * n + c];
      for (int c = 0; c < n; c++) b[j * n + c] += alpha * b[i * n + c];
    }
  }
  for (int i = n - 1; i >= 0; i--)
    for (int j = 0; j < n; j++) {
      float sacc = b[i * n + j];
      for (int c = i + 1; c < n; c++) sacc -= A[i * n + c] * b[c * n + j];
      b[i * n + j] = sacc / A[i * n + i];
    }
  for (int s = 0; s < n; s++) invWgt[s] = b[s];
}

// vols[s]: level-s volume [Hs[s]][Ws[s]][Ds[s]]; vols[0] is updated in place.
void orc_solve_all(float** vols, const int* Hs, const int* Ws, const int* Ds, int levels, float lambda) {
  std::vector<float> w(levels);
  orc_cross_scale_weights(levels, lambda, w.data());
  const int H = Hs[0], W = Ws[0], D = Ds[0];
  ORC_PAR_FOR
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++)
      for (int d = 0; d < D; d++) {
        int cy = y, cx = x, cd = d;
        float sum = 0;
        for (int s = 0; s < levels; s++) {
          sum += w[s] * vols[s][((long)cy * Ws[s] + cx) * Ds[s] + cd];
          cy /= 2; cx /= 2; cd = (cd + 1) / 2;
        }
        vols[0][((long)y * W + x) * D + d] = sum;
      }
}

// ---------------------------------------------------------------------------
// Whole default chain for one stereo pair, as main()/pipeline() drive it with
// costcalculation="ADCensus", aggregation="CBCA", optimization="sgm"
// (stereoMatching.cpp:945-1021, 1046-1136, 1364-1506), Do_refine=1,
// Do_LRConsis=1.  stage_ms (nullable, 8 floats) gets per-stage wall times:
// census, cost, arms, cbca, sgm, wta, refine, total.
// ---------------------------------------------------------------------------
struct orc_params {
  int D, censusFunc, paths, iters, L, L_out, tau, tau_out, minL, corDifThres, reduCoeffi1;
  float adTrunc, lamAD, lamCen, LRmaxDiff, voteRatio;
  int voteS, voteNums, DISP_OCC, do_refine;
  int aggregation;  // 1 = "CBCA" (cbca_aggregate), 2 = "NL" (StereoMatching::NL: left volume only), 0 = none
  int costcalc;     // 0 = "ADCensus", 1 = "censusGrad" (main_.cpp:15)
  float cgLamCen, cgLamG, gradTrunc;  // 13, 1 (main_.cpp:60-61), 500 (stereoMatching.cpp:34)
  int pyrLevels;      // PY_LEV of main_.cpp:132 (1 in the reference's driver)
  float crossLambda;  // REG_LAMBDA (0.3, main_.cpp:157); < 0: skip SolveAll
};
void orc_nl(const u8* bgrL, int H, int W, int D, float* vol, i16* disp);  // nl_oracle.cpp
}  // extern "C"

#include <chrono>
static double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// costCalculate() (stereoMatching.cpp:945-1044) for one pyramid level: cost volumes of both views, arms, aggregation.
// scale = Parameters::disSc of the level (arm lengths L/scale, L_out/scale, stereoMatching.cpp:5368-5371).
static void cost_calculate(const u8* bgrL, const u8* bgrR, const u8* grayL, const u8* grayR, int H, int W, int D,
                           const orc_params* p, int scale, int views, std::vector<std::vector<float>>& vm,
                           std::vector<u16>& aL, std::vector<u16>& aR, double* ms, double& t) {
  auto lap = [&](int k) { double x = now_ms(); ms[k] += x - t; t = x; };
  const long n = (long)H * W * D;
  const int codeLen = orc_census_code_length(p->censusFunc, 3, 4);  // stereoMatching.cpp:815
  const int nw = (codeLen + 63) / 64;
  std::vector<u64> cL((long)H * W * nw), cR((long)H * W * nw);
  orc_census(grayL, H, W, p->censusFunc, 3, 4, cL.data());
  orc_census(grayR, H, W, p->censusFunc, 3, 4, cR.data());
  lap(0);
  vm.assign(2, std::vector<float>());
  aL.assign((long)H * W * 5, 0);
  aR.assign((long)H * W * 5, 0);
  auto arms = [&]() {
    orc_arms(bgrL, H, W, 3, p->L / scale, p->L_out / scale, p->tau, p->tau_out, p->minL, aL.data());
    orc_arms(bgrR, H, W, 3, p->L / scale, p->L_out / scale, p->tau, p->tau_out, p->minL, aR.data());
  };
  if (p->costcalc == 0) {
    std::vector<float> ad(n), cen(n);
    for (int i = 0; i < 2; i++) {  // cost volumes are always built for both views (:898)
      vm[i].resize(n);
      orc_ad_vol(bgrL, bgrR, H, W, D, i, p->adTrunc, ad.data());
      orc_hamming_vol(cL.data(), cR.data(), H, W, D, nw, codeLen, 1.0f, i, cen.data());
      orc_combine_exp(ad.data(), cen.data(), n, p->lamAD, p->lamCen, vm[i].data());
    }
    lap(1);
    arms();
  } else if (p->costcalc == 2) {  // "Census": censusCal(vm, 1) (stereoMatching.cpp:975-976): the Hamming volumes
    for (int i = 0; i < 2; i++) {
      vm[i].resize(n);
      orc_hamming_vol(cL.data(), cR.data(), H, W, D, nw, codeLen, 1.0f, i, vm[i].data());
    }
    lap(1);
    arms();
  } else {  // censusGrad (stereoMatching.cpp:25-48); grad() needs the arms first (:628-631)
    const long np = (long)H * W;
    std::vector<float> g(4 * np), gr(n), cen(n);
    orc_grad_xy(grayL, H, W, g.data(), g.data() + np);
    orc_grad_xy(grayR, H, W, g.data() + 2 * np, g.data() + 3 * np);
    arms();
    for (int i = 0; i < 2; i++) {
      vm[i].resize(n);
      orc_grad_vol(g.data(), g.data() + 2 * np, g.data() + np, g.data() + 3 * np, i == 0 ? aL.data() : aR.data(), H, W, D,
                   i, p->gradTrunc, gr.data());
      orc_hamming_vol(cL.data(), cR.data(), H, W, D, nw, codeLen, 1.0f, i, cen.data());
      orc_combine_exp(cen.data(), gr.data(), n, p->cgLamCen, p->cgLamG, vm[i].data());
    }
    lap(1);
  }
  lap(2);
  if (p->aggregation == 1) {
    for (int i = 0; i < views; i++)
      orc_cbca(vm[i].data(), aL.data(), aR.data(), H, W, D, p->iters, i, nullptr);
  } else if (p->aggregation == 2) {
    orc_nl(bgrL, H, W, D, vm[0].data(), nullptr);  // stereoMatching.cpp:4892-4917: vm[0] only
  }
  lap(3);
}

extern "C" {
// aggL_out (nullable): vm[0] as costCalculate() leaves it (after aggregation / SolveAll, before sgm) -- lets a test
// compare the aggregated volume and the optimised one from a single run of the chain.
void orc_pipeline_ex(const u8* bgrL, const u8* bgrR, const u8* grayL, const u8* grayR, int H, int W,
                     const orc_params* p, i16* dispL, i16* dispR, float* volL_out, float* stage_ms, float* aggL_out) {
  const int D = p->D;
  const long n = (long)H * W * D;
  double t0 = now_ms(), t = t0, ms[8] = {0};
  auto lap = [&](int k) { double x = now_ms(); ms[k] += x - t; t = x; };
  const int views = p->do_refine ? 2 : 1;
  std::vector<std::vector<float>> vm;
  std::vector<u16> aL, aR;
  cost_calculate(bgrL, bgrR, grayL, grayR, H, W, D, p, 1, views, vm, aL, aR, ms, t);
  if (p->crossLambda >= 0.f) {
    // the caller's pyramid loop (main_.cpp:131-158): level s+1 = pyrDown of level s's colour AND gray images,
    // maxDisp/2 + 1, disSc*2; costCalculate() per level; then SolveAll
    const int levels = p->pyrLevels < 1 ? 1 : p->pyrLevels;
    std::vector<std::vector<std::vector<float>>> lv(levels);   // [level][view]
    std::vector<int> Hs(levels), Ws(levels), Ds(levels);
    Hs[0] = H; Ws[0] = W; Ds[0] = D;
    std::vector<u8> cb[2], cg[2], nb[2], ng[2];
    const u8 *pbL = bgrL, *pbR = bgrR, *pgL = grayL, *pgR = grayR;
    int maxDisp = D - 1, scale = 1;
    for (int s = 1; s < levels; s++) {
      const int h = Hs[s - 1], w = Ws[s - 1], ho = (h + 1) / 2, wo = (w + 1) / 2;
      nb[0].resize((long)ho * wo * 3); nb[1].resize((long)ho * wo * 3); ng[0].resize((long)ho * wo); ng[1].resize((long)ho * wo);
      orc_pyr_down_u8(pbL, h, w, 3, nb[0].data()); orc_pyr_down_u8(pbR, h, w, 3, nb[1].data());
      orc_pyr_down_u8(pgL, h, w, 1, ng[0].data()); orc_pyr_down_u8(pgR, h, w, 1, ng[1].data());
      for (int i = 0; i < 2; i++) { cb[i].swap(nb[i]); cg[i].swap(ng[i]); }
      pbL = cb[0].data(); pbR = cb[1].data(); pgL = cg[0].data(); pgR = cg[1].data();
      maxDisp = maxDisp / 2 + 1; scale *= 2;
      Hs[s] = ho; Ws[s] = wo; Ds[s] = maxDisp + 1;
      std::vector<u16> a0, a1;
      double dms[8] = {0}, dt = now_ms();
      cost_calculate(pbL, pbR, pgL, pgR, ho, wo, Ds[s], p, scale, views, lv[s], a0, a1, dms, dt);
    }
    for (int i = 0; i < views; i++) {   // SolveAll: img_n = Do_refine ? 2 : 1 (stereoMatching.cpp:2179)
      std::vector<float*> ptr(levels);
      ptr[0] = vm[i].data();
      for (int s = 1; s < levels; s++) ptr[s] = lv[s][i].data();
      orc_solve_all(ptr.data(), Hs.data(), Ws.data(), Ds.data(), levels, p->crossLambda);
    }
    t = now_ms();
  }
  lap(3);
  if (aggL_out) std::memcpy(aggL_out, vm[0].data(), n * sizeof(float));
  for (int i = 0; i < views; i++)
    orc_sgm(vm[i].data(), i == 0 ? bgrL : bgrR, H, W, D, p->paths, p->corDifThres, p->reduCoeffi1);
  lap(4);
  orc_wta(vm[0].data(), H, W, D, dispL);
  if (views == 2) orc_wta(vm[1].data(), H, W, D, dispR);
  lap(5);
  if (p->do_refine) {
    orc_lrc_normal(dispL, dispR, H, W, p->LRmaxDiff);
    for (int i = 0; i < p->voteNums; i++)
      orc_region_vote(dispL, aL.data(), H, W, D, p->voteRatio, p->voteS);
    for (int i = 0; i < p->voteNums; i++) orc_proper_ipol(dispL, bgrL, H, W, p->DISP_OCC);
    std::vector<i16> tmp(dispL, dispL + (long)H * W);
    orc_median3_i16(tmp.data(), H, W, dispL);
  }
  lap(6);
  if (volL_out) std::memcpy(volL_out, vm[0].data(), n * sizeof(float));
  ms[7] = now_ms() - t0;
  if (stage_ms)
    for (int k = 0; k < 8; k++) stage_ms[k] = (float)ms[k];
}

void orc_pipeline(const u8* bgrL, const u8* bgrR, const u8* grayL, const u8* grayR, int H, int W,
                  const orc_params* p, i16* dispL, i16* dispR, float* volL_out, float* stage_ms) {
  orc_pipeline_ex(bgrL, bgrR, grayL, grayR, H, W, p, dispL, dispR, volL_out, stage_ms, nullptr);
}
}  // extern "C"
