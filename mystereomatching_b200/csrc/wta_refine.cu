// K6: disparity selection and the active refinement chain.
//
//   gen_dispFromVm              stereoMatching.cpp:3928-3967   first minimum, unscaled int16, -1 if none
//   wta_Co                      stereoMatching.cpp:2709-2792   left map + right map from the left volume's diagonal
//   LRConsistencyCheck_normal   stereoMatching.cpp:2262-2282
//   LRConsistencyCheck          stereoMatching.cpp:2284-2335   occlusion / mismatch labelling (LOR = 0)
//   regionVote_my               stereoMatching.cpp:7219-7277
//   properIpol                  stereoMatching.cpp:7395-7490
//   cv::medianBlur(CV_16S, 3)   stereoMatching.cpp:1499
//   SolveAll (1 level)          stereoMatching.cpp:2142-2208
//
// WTA is the only volume-sized stage here (one read of the volume, HBM bound):
// a warp per pixel, lanes stride d so every load instruction is a coalesced 128 B
// line, then a shuffle argmin that keeps the LOWEST d among equal minima.  All
// other kernels are O(H*W).
#include <float.h>
#include <limits.h>

#include <stdlib.h>
#include <limits.h>

#include "common.cuh"

// ------------------------------------------------------------------ WTA
#define WTA_WARPS 8

__global__ void __launch_bounds__(WTA_WARPS * 32)
    k_wta(const float* __restrict__ vol, long long npix, int D, int16_t* __restrict__ disp) {
  const int lane = threadIdx.x & 31;
  long long p = (long long)blockIdx.x * WTA_WARPS + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * WTA_WARPS;
  for (; p < npix; p += stride) {
    const float* c = vol + p * D;
    float best = FLT_MAX;
    int bd = -1;
    for (int d = lane; d < D; d += 32) {
      const float x = c[d];
      if (best > x) { best = x; bd = d; }  // strict: an earlier (lower) d keeps the win
    }
    // lanes hold increasing-d candidates; combine preferring smaller value, then smaller d.
    // bd == -1 (nothing below FLT_MAX) must lose against any real candidate with the same value,
    // which cannot exist (a real candidate is strictly below FLT_MAX), so (value, d) order is safe.
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int od = __shfl_xor_sync(0xffffffffu, bd, o);
      if (ob < best || (ob == best && od >= 0 && (bd < 0 || od < bd))) { best = ob; bd = od; }
    }
    if (lane == 0) disp[p] = (int16_t)bd;
  }
}

extern "C" int sm_wta(sm_ctx* ctx, const float* d_vol, int H, int W, int D, int16_t* d_disp) {
  SM_CHECK_ARG(ctx && d_vol && d_disp && H > 0 && W > 0 && D > 0);
  const long long npix = (long long)H * W;
  int grid = (int)min((long long)ctx->num_sms * 8, (npix + WTA_WARPS - 1) / WTA_WARPS);
  SM_LAUNCH(ctx, k_wta, grid, WTA_WARPS * 32, 0, d_vol, npix, D, d_disp);
  return SM_OK;
}

__global__ void __launch_bounds__(WTA_WARPS * 32)
    k_wta_co(const float* __restrict__ vol, int H, int W, int D, int scale, int16_t* __restrict__ D1,
             int16_t* __restrict__ D2) {
  const int lane = threadIdx.x & 31;
  const long long npix = (long long)H * W;
  long long p = (long long)blockIdx.x * WTA_WARPS + (threadIdx.x >> 5);
  const long long stride = (long long)gridDim.x * WTA_WARPS;
  for (; p < npix; p += stride) {
    const int u = (int)(p % W);
    float bl = FLT_MAX, br = FLT_MAX;
    int dl = 0, dr = 0;
    bool hl = false, hr = false;  // "some d improved on FLT_MAX" (else the reference keeps disp 0)
    for (int d = lane; d < D; d += 32) {
      if (u - d >= 0) {
        const float x = vol[p * D + d];
        if (x < bl) { bl = x; dl = d; hl = true; }
      }
      if (u + d < W) {
        const float x = vol[(p + d) * D + d];
        if (x < br) { br = x; dr = d; hr = true; }
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, bl, o);
      int od = __shfl_xor_sync(0xffffffffu, dl, o);
      int oh = __shfl_xor_sync(0xffffffffu, (int)hl, o);
      if (oh && (!hl || ob < bl || (ob == bl && od < dl))) { bl = ob; dl = od; hl = true; }
      ob = __shfl_xor_sync(0xffffffffu, br, o);
      od = __shfl_xor_sync(0xffffffffu, dr, o);
      oh = __shfl_xor_sync(0xffffffffu, (int)hr, o);
      if (oh && (!hr || ob < br || (ob == br && od < dr))) { br = ob; dr = od; hr = true; }
    }
    if (lane == 0) {
      D1[p] = (int16_t)((hl ? dl : 0) * scale);
      D2[p] = (int16_t)((hr ? dr : 0) * scale);
    }
  }
}

extern "C" int sm_wta_co(sm_ctx* ctx, const float* d_vol, int H, int W, int D, int scale, int16_t* d_D1,
                         int16_t* d_D2) {
  SM_CHECK_ARG(ctx && d_vol && d_D1 && d_D2 && H > 0 && W > 0 && D > 0);
  const long long npix = (long long)H * W;
  int grid = (int)min((long long)ctx->num_sms * 8, (npix + WTA_WARPS - 1) / WTA_WARPS);
  SM_LAUNCH(ctx, k_wta_co, grid, WTA_WARPS * 32, 0, d_vol, H, W, D, scale, d_D1, d_D2);
  return SM_OK;
}

// ------------------------------------------------------------------ candidate disparities (first half)
// selectTopCostFromVolumn (stereoMatching.h:2405-2461): per pixel, up to `num` rounds of "first minimum of what is left"
// (strict '>' scan in increasing d -> the lowest d among equal minima; a taken entry becomes FLT_MAX); candidate 0 is
// always taken, candidate k > 0 only while its cost < firstCost * thres.  topDisp[v][u][k] = {d, cost},
// topDisp[v][u][num][0] = number of candidates.  The reference works on a clone of vm, so d_vol is read only here; the
// entries the reference leaves unwritten (its Mat is uninitialised there) are 0.
// One warp per pixel: the D costs sit in shared memory, lane l scans d = l, l + 32, ...; a round is a lane-local first
// minimum, one redux over the order-preserving integer image of the floats and one over the candidate d.
#define TOP_WARPS 8
__device__ __forceinline__ uint32_t top_f2key(float x) {
  uint32_t b = __float_as_uint(x + 0.0f);   // -0 -> +0: the reference's '>' does not tell them apart
  return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__global__ void __launch_bounds__(TOP_WARPS * 32)
    k_select_top(const float* __restrict__ vol, long long npix, int D, int num, float thres, float* __restrict__ top) {
  extern __shared__ float top_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* c = top_smem + (size_t)warp * D;
  long long p = (long long)blockIdx.x * TOP_WARPS + warp;
  const long long stride = (long long)gridDim.x * TOP_WARPS;
  for (; p < npix; p += stride) {
    for (int d = lane; d < D; d += 32) c[d] = vol[p * D + d];
    float* o = top + p * (long long)(num + 1) * 2;
    for (int i = lane; i < (num + 1) * 2; i += 32) o[i] = 0.f;
    __syncwarp();
    float firstV = 0.f;
    int count = 0;
    for (int k = 0; k < num; k++) {
      uint32_t bk = 0xffffffffu;
      int bd = 0x7fffffff;
      for (int d = lane; d < D; d += 32) {
        const uint32_t key = top_f2key(c[d]);
        if (key < bk) { bk = key; bd = d; }     // strict: the lowest d of this lane's equal minima
      }
      const uint32_t km = __reduce_min_sync(0xffffffffu, bk);
      const int d = (int)__reduce_min_sync(0xffffffffu, bk == km ? (unsigned)bd : 0x7fffffffu);
      const float val = c[d];                    // the value as stored (sign of zero included)
      if (k == 0) firstV = val;
      else if (!(val < firstV * thres)) break;
      __syncwarp();
      if (lane == 0) { c[d] = FLT_MAX; o[2 * k] = (float)d; o[2 * k + 1] = val; }
      count++;
      __syncwarp();
    }
    if (lane == 0) o[2 * num] = (float)count;
    __syncwarp();
  }
}

extern "C" int sm_select_top_cost(sm_ctx* ctx, const float* d_vol, int H, int W, int D, int num, float thres,
                                  float* d_top) {
  SM_CHECK_ARG(ctx && d_vol && d_top && H > 0 && W > 0 && D > 0 && D <= 1024 && num >= 1 && num <= 64);   // 8 warps x D floats of shared memory
  const long long npix = (long long)H * W;
  const int grid = (int)min((long long)ctx->num_sms * 8, (npix + TOP_WARPS - 1) / TOP_WARPS);
  SM_LAUNCH(ctx, k_select_top, grid, TOP_WARPS * 32, TOP_WARPS * D * sizeof(float), d_vol, npix, D, num, thres, d_top);
  return SM_OK;
}

// ------------------------------------------------------------------ sub-pixel enhancement
// subpixelEnhancement (stereoMatching.cpp:6138-6166): parabola through the costs at disp-1, disp, disp+1 of view 0.
// The reference applies the offset to the SHORT (`disp -= diff`, i.e. (short)((float)disp - diff), truncation toward
// zero) and only then converts to float, so the output is integer-valued; kept as it is.  One thread per pixel: 2 B of
// disparity in, three gathered floats (one 32 B sector when they do not straddle), 4 B out.
__global__ void k_subpixel(const int16_t* __restrict__ disp, const float* __restrict__ vol, long long npix, int D,
                           float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix; i += (long long)gridDim.x * blockDim.x) {
    short d = disp[i];
    if (d > 0 && d < D - 1) {
      const float* c = vol + i * D + d;
      const float cost = c[0], costPlus = c[1], costMinus = c[-1];
      const float denom = 2 * (costPlus + costMinus - 2 * cost);
      if (denom != 0) {
        const float diff = (costPlus - costMinus) / denom;
        if (diff > -1 && diff < 1) d = (short)((float)d - diff);
      }
    }
    out[i] = (float)d;
  }
}

extern "C" int sm_subpixel_enhancement(sm_ctx* ctx, const int16_t* d_disp, const float* d_vol, int H, int W, int D,
                                       float* d_floatDisp) {
  SM_CHECK_ARG(ctx && d_disp && d_vol && d_floatDisp && H > 0 && W > 0 && D > 0);
  const long long npix = (long long)H * W;
  const int grid = (int)min((long long)ctx->num_sms * 8, (npix + 255) / 256);
  SM_LAUNCH(ctx, k_subpixel, grid, 256, 0, d_disp, d_vol, npix, D, d_floatDisp);
  return SM_OK;
}

// ------------------------------------------------------------------ LR check
// In-place on D1 is safe: each thread reads D1 only at its own pixel.
__global__ void k_lrc(int16_t* __restrict__ D1, const int16_t* __restrict__ D2, int H, int W, float maxDiff) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const size_t i = (size_t)v * W + u;
  const int d = D1[i];
  if (d < 0 || u - d < 0 || (float)abs(d - (int)D2[i - d]) > maxDiff) D1[i] = -1;
}

extern "C" int sm_lrc(sm_ctx* ctx, int16_t* d_D1, const int16_t* d_D2, int H, int W, float LRmaxDiff) {
  SM_CHECK_ARG(ctx && d_D1 && d_D2 && H > 0 && W > 0);
  dim3 grid(sm_div_up(W, 256), H);
  SM_LAUNCH(ctx, k_lrc, grid, 256, 0, d_D1, d_D2, H, W, LRmaxDiff);
  return SM_OK;
}

__global__ void k_lrc_label(int16_t* __restrict__ D1, const int16_t* __restrict__ D2, int H, int W, int D,
                            float maxDiff, int occ, int mis, uint8_t* __restrict__ mask) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const size_t i = (size_t)v * W + u;
  const int d = D1[i];
  uint8_t m = 0;
  if (d < 0 || u - d < 0 || (float)abs(d - (int)D2[i - d]) > maxDiff) {
    m = 255;
    int disp = occ;
    for (int dd = 0; dd < D && u - dd >= 0; dd++)
      if (D2[i - dd] == dd) { disp = mis; break; }
    D1[i] = (int16_t)disp;
  }
  if (mask) mask[i] = m;
}

extern "C" int sm_lrc_label(sm_ctx* ctx, int16_t* d_D1, const int16_t* d_D2, int H, int W, int D, float LRmaxDiff,
                            int DISP_OCC, int DISP_MIS, uint8_t* d_errMask) {
  SM_CHECK_ARG(ctx && d_D1 && d_D2 && H > 0 && W > 0 && D > 0);
  dim3 grid(sm_div_up(W, 256), H);
  SM_LAUNCH(ctx, k_lrc_label, grid, 256, 0, d_D1, d_D2, H, W, D, LRmaxDiff, DISP_OCC, DISP_MIS, d_errMask);
  return SM_OK;
}

// LOR == 1 branch (stereoMatching.cpp:2336-2364): the right map is checked against the left one and labelled in place;
// D1 is only read, so the sweep is data-parallel like the LOR == 0 one.  The reference's errMask stays all zero on this
// branch (the flags go to a local errMask1 that is only written to LR1.png): mask <- 0, mask1 <- the flags.
__global__ void k_lrc_label_right(const int16_t* __restrict__ D1, int16_t* __restrict__ D2, int H, int W, int D,
                                  float maxDiff, int occ, int mis, uint8_t* __restrict__ mask, uint8_t* __restrict__ mask1) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const size_t i = (size_t)v * W + u;
  const int d = D2[i];
  uint8_t m = 0;
  if (d < 0 || u + d >= W || (float)abs(d - (int)D1[i + d]) > maxDiff) {
    m = 255;
    int disp = occ;
    for (int dd = 0; dd < D && u + dd <= W - 1; dd++)
      if (D1[i + dd] == dd) { disp = mis; break; }
    D2[i] = (int16_t)disp;
  }
  if (mask) mask[i] = 0;
  if (mask1) mask1[i] = m;
}

extern "C" int sm_lrc_label_lor(sm_ctx* ctx, int16_t* d_D1, int16_t* d_D2, int H, int W, int D, float LRmaxDiff,
                                int DISP_OCC, int DISP_MIS, int LOR, uint8_t* d_errMask, uint8_t* d_errMask1) {
  SM_CHECK_ARG(ctx && d_D1 && d_D2 && H > 0 && W > 0 && D > 0 && (LOR == 0 || LOR == 1));
  dim3 grid(sm_div_up(W, 256), H);
  if (LOR == 0) {
    SM_LAUNCH(ctx, k_lrc_label, grid, 256, 0, d_D1, d_D2, H, W, D, LRmaxDiff, DISP_OCC, DISP_MIS, d_errMask);
    if (d_errMask1) SM_CUDA(cudaMemsetAsync(d_errMask1, 0, (size_t)H * W, ctx->stream));   // Mat::zeros, untouched on this branch
  } else {
    SM_LAUNCH(ctx, k_lrc_label_right, grid, 256, 0, d_D1, d_D2, H, W, D, LRmaxDiff, DISP_OCC, DISP_MIS, d_errMask, d_errMask1);
  }
  return SM_OK;
}

// ------------------------------------------------------------------ region voting
// One warp per pixel; valid pixels copy through.  The votes of the cross region
// (vertical arm of the anchor, horizontal arm of each pixel on it; image-space
// arms of the left image) go into a per-warp shared-memory histogram.
#define RV_WARPS 8

// Two steps instead of one persistent sweep.  k_region_vote walks ALL pixels one per warp iteration (a dependent 2-byte
// load each) and, inside a vote, reads one row's arms, then that row's disparities, row after row -- two dependent
// global loads per row, ~20 rows per vote: a chain of latencies (0.26 ms per call at 1080p).  Here k_rv_scan copies the
// valid pixels through (one thread per pixel, coalesced) and appends the invalid ones to a list; k_rv_vote takes one
// warp per list entry, fetches the arms of ALL rows of the region in one round (lane = row), broadcasts them by
// shuffles, and reads the disparities of four rows per round.  Same votes, same integer rule: identical output.
__global__ void k_rv_scan(const int16_t* __restrict__ src, int16_t* __restrict__ dst, long long npix, int* __restrict__ count,
                          int* __restrict__ list) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const int16_t cur = src[p];
  if (cur >= 0) dst[p] = cur;
  else list[atomicAdd(count, 1)] = (int)p;   // warp-aggregated by the compiler
}

__global__ void __launch_bounds__(RV_WARPS * 32)
    k_rv_vote(const int16_t* __restrict__ src, int16_t* __restrict__ dst, const uint16_t* __restrict__ arms, int H, int W,
              int D, float ratio, int S, const int* __restrict__ count, const int* __restrict__ list) {
  extern __shared__ int hist_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* hist = hist_all + warp * D;
  const int n = *count;
  for (int e = blockIdx.x * RV_WARPS + warp; e < n; e += gridDim.x * RV_WARPS) {
    const int p = list[e];
    const int16_t cur = src[p];
    const int v = p / W, u = p - v * W;
    for (int d = lane; d < D; d += 32) hist[d] = 0;
    const uint16_t* a = arms + (size_t)p * 5;
    const int vb = v - a[2], ve = v + a[3];
    const int nrows = ve - vb + 1;           // <= 2 * 255 + 1
    __syncwarp();
    int valid = 0;
    for (int r0 = 0; r0 < nrows; r0 += 32) {
      // lane = row: the horizontal arms of 32 rows of the region in one round
      int ub = 0, ue = -1;
      if (r0 + lane < nrows) {
        const uint16_t* b = arms + ((size_t)(vb + r0 + lane) * W + u) * 5;
        ub = u - b[0]; ue = u + b[1];
      }
      const int nr = min(32, nrows - r0);
      for (int r = 0; r < nr; r += 4) {
        int x[4], un[4], uE[4];
        const int16_t* row[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {        // four rows' first 32 columns in flight together
          const int rr = min(r + k, nr - 1);
          const int b0 = __shfl_sync(0xffffffffu, ub, rr);
          uE[k] = r + k < nr ? __shfl_sync(0xffffffffu, ue, rr) : INT_MIN;
          un[k] = b0 + lane;
          row[k] = src + (size_t)(vb + r0 + rr) * W;
          x[k] = un[k] <= uE[k] ? (int)row[k][un[k]] : -1;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (x[k] >= 0) { valid++; if (x[k] < D) atomicAdd(&hist[x[k]], 1); }
          for (int w = un[k] + 32; w <= uE[k]; w += 32) {   // rows wider than a warp
            const int y = row[k][w];
            if (y >= 0) { valid++; if (y < D) atomicAdd(&hist[y], 1); }
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) valid += __shfl_xor_sync(0xffffffffu, valid, o);
    __syncwarp();
    int16_t res = cur;
    if (valid > S) {
      // mode with the lowest d on ties: maximise (count, -d)
      int bc = -1, bd = 0;
      for (int d = lane; d < D; d += 32) {
        const int c = hist[d];
        if (c > bc) { bc = c; bd = d; }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const int oc = __shfl_xor_sync(0xffffffffu, bc, o), od = __shfl_xor_sync(0xffffffffu, bd, o);
        if (oc > bc || (oc == bc && od < bd)) { bc = oc; bd = od; }
      }
      if ((float)(bc / valid) >= ratio) res = (int16_t)bd;  // integer division, as the reference
    }
    if (lane == 0) dst[p] = res;
    __syncwarp();
  }
}

// 0 < ratio <= 1: the reference's test `(float)(maxCount / validCount) >= ratio` divides two ints, so it holds iff the
// quotient is 1, i.e. iff EVERY valid vote of the region names the same disparity (< D).  No histogram is needed then: a
// running minimum and maximum of the votes decide, and the scan of a region stops at the first group of rows in which two
// different votes (or one >= D) have been seen -- which, around the occlusion borders where the invalid pixels sit, is
// almost always the first.  Same result as k_rv_vote for these ratios (the tests run both); the other ratios (<= 0: always
// the mode; > 1: never) keep the histogram kernel.
__global__ void __launch_bounds__(RV_WARPS * 32)
    k_rv_vote_agree(const int16_t* __restrict__ src, int16_t* __restrict__ dst, const uint16_t* __restrict__ arms, int H, int W,
                    int D, int S, const int* __restrict__ count, const int* __restrict__ list) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = *count;
  for (int e = blockIdx.x * RV_WARPS + warp; e < n; e += gridDim.x * RV_WARPS) {
    const int p = list[e];
    const int16_t cur = src[p];
    const int v = p / W, u = p - v * W;
    const uint16_t* a = arms + (size_t)p * 5;
    const int vb = v - a[2], ve = v + a[3];
    const int nrows = ve - vb + 1;
    int valid = 0;
    unsigned lo = 0xffffffffu, hi = 0u;      // votes + 1 (0 = none seen)
    bool mixed = false;
    for (int r0 = 0; r0 < nrows && !mixed; r0 += 32) {
      int ub = 0, ue = -1;
      if (r0 + lane < nrows) {
        const uint16_t* b = arms + ((size_t)(vb + r0 + lane) * W + u) * 5;
        ub = u - b[0]; ue = u + b[1];
      }
      const int nr = min(32, nrows - r0);
      for (int r = 0; r < nr && !mixed; r += 4) {
        int x[4], un[4], uE[4];
        const int16_t* row[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const int rr = min(r + k, nr - 1);
          const int b0 = __shfl_sync(0xffffffffu, ub, rr);
          uE[k] = r + k < nr ? __shfl_sync(0xffffffffu, ue, rr) : INT_MIN;
          un[k] = b0 + lane;
          row[k] = src + (size_t)(vb + r0 + rr) * W;
          x[k] = un[k] <= uE[k] ? (int)row[k][un[k]] : -1;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
          if (x[k] >= 0) { valid++; lo = min(lo, (unsigned)x[k] + 1u); hi = max(hi, (unsigned)x[k] + 1u); }
          for (int w = un[k] + 32; w <= uE[k]; w += 32) {
            const int y = row[k][w];
            if (y >= 0) { valid++; lo = min(lo, (unsigned)y + 1u); hi = max(hi, (unsigned)y + 1u); }
          }
        }
        const unsigned wlo = __reduce_min_sync(0xffffffffu, lo), whi = __reduce_max_sync(0xffffffffu, hi);
        mixed = whi != 0u && (wlo != whi || whi > (unsigned)D);
      }
    }
    int16_t res = cur;
    if (!mixed) {
      valid = __reduce_add_sync(0xffffffffu, valid);
      const unsigned whi = __reduce_max_sync(0xffffffffu, hi);
      if (valid > S && whi != 0u) res = (int16_t)(whi - 1u);
    }
    if (lane == 0) dst[p] = res;
  }
}

extern "C" int sm_region_vote(sm_ctx* ctx, int16_t* d_disp, int16_t* d_tmp, const uint16_t* d_arms, int H, int W, int D,
                              float ratio, int S) {
  SM_CHECK_ARG(ctx && d_disp && d_tmp && d_arms && H > 0 && W > 0 && D > 0 && D <= 1536);   // RV_WARPS histograms of D ints in 48 KB
  const long long npix = (long long)H * W;
  SM_CHECK_ARG(npix < (1ll << 31));
  void* p;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_RVLIST, 256 + (size_t)npix * 4, &p));
  int* count = (int*)p;
  int* list = (int*)((uint8_t*)p + 256);
  SM_CUDA(cudaMemsetAsync(count, 0, sizeof(int), ctx->stream));
  SM_LAUNCH(ctx, k_rv_scan, (int)((npix + 255) / 256), 256, 0, d_disp, d_tmp, npix, count, list);
  const int grid = (int)min((long long)ctx->num_sms * 8, (npix + RV_WARPS - 1) / RV_WARPS);
  if (ratio > 0.f && ratio <= 1.f)
    SM_LAUNCH(ctx, k_rv_vote_agree, grid, RV_WARPS * 32, 0, d_disp, d_tmp, d_arms, H, W, D, S, count, list);
  else
    SM_LAUNCH(ctx, k_rv_vote, grid, RV_WARPS * 32, RV_WARPS * D * sizeof(int), d_disp, d_tmp, d_arms, H, W, D, ratio, S,
              count, list);
  SM_CUDA(cudaMemcpyAsync(d_disp, d_tmp, npix * sizeof(int16_t), cudaMemcpyDeviceToDevice, ctx->stream));
  return SM_OK;
}

// ------------------------------------------------------------------ proper interpolation
__constant__ int c_dirW[16] = {0, 2, 2, 2, 0, -2, -2, -2, 1, 2, 2, 1, -1, -2, -2, -1};
__constant__ int c_dirH[16] = {2, 2, 0, -2, -2, -2, 0, 2, 2, 1, -1, -2, -2, -1, 1, 2};

__global__ void k_proper_ipol(const int16_t* __restrict__ src, int16_t* __restrict__ dst,
                              const uint8_t* __restrict__ bgr, int H, int W, int occ) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const size_t i = (size_t)v * W + u;
  const int cur = src[i];
  if (cur >= 0) { dst[i] = (int16_t)cur; return; }
  const int b0 = bgr[i * 3], g0 = bgr[i * 3 + 1], r0 = bgr[i * 3 + 2];
  int minDisp = INT_MAX;           // DISP_OCC branch
  int minDif = 255, dispC = -1;    // colour branch
#pragma unroll 1
  for (int k = 0; k < 16; k++) {
    const int pw = c_dirW[k], ph = c_dirH[k];
    int x = u, y = v;
    for (int dep = 0; dep < 20; dep++) {
      if ((dep & 1) == 0) { x += pw / 2; y += ph / 2; }
      else { x += pw - pw / 2; y += ph - ph / 2; }
      if (!(x >= 0 && x < W && y >= 0 && y < H)) break;
      const size_t q = (size_t)y * W + x;
      const int dq = src[q];
      if (dq >= 0) {
        int cd = max(abs(b0 - (int)bgr[q * 3]), max(abs(g0 - (int)bgr[q * 3 + 1]), abs(r0 - (int)bgr[q * 3 + 2])));
        if (minDisp > dq) minDisp = dq;
        if (minDif > cd) { minDif = cd; dispC = dq; }  // strict: first direction wins ties
        break;
      }
    }
  }
  int res;
  if (cur == occ) res = minDisp != INT_MAX ? minDisp : cur;
  else res = dispC >= 0 ? dispC : cur;
  dst[i] = (int16_t)res;
}

extern "C" int sm_proper_ipol(sm_ctx* ctx, int16_t* d_disp, int16_t* d_tmp, const uint8_t* d_bgr, int H, int W,
                              int DISP_OCC) {
  SM_CHECK_ARG(ctx && d_disp && d_tmp && d_bgr && H > 0 && W > 0);
  dim3 grid(sm_div_up(W, 128), H);
  SM_LAUNCH(ctx, k_proper_ipol, grid, 128, 0, d_disp, d_tmp, d_bgr, H, W, DISP_OCC);
  SM_CUDA(cudaMemcpyAsync(d_disp, d_tmp, (size_t)H * W * sizeof(int16_t), cudaMemcpyDeviceToDevice, ctx->stream));
  return SM_OK;
}

// ------------------------------------------------------------------ 3x3 median, int16, replicated border
__device__ __forceinline__ void cswap(int& a, int& b) {
  const int lo = min(a, b), hi = max(a, b);
  a = lo; b = hi;
}

__global__ void k_median3_i16(const int16_t* __restrict__ src, int16_t* __restrict__ dst, int H, int W) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  int w[9];
#pragma unroll
  for (int dv = -1; dv <= 1; dv++)
#pragma unroll
    for (int du = -1; du <= 1; du++) {
      const int y = min(max(v + dv, 0), H - 1), x = min(max(u + du, 0), W - 1);
      w[(dv + 1) * 3 + du + 1] = src[(size_t)y * W + x];
    }
  // 19-exchange median-of-9 network (Paeth)
  cswap(w[1], w[2]); cswap(w[4], w[5]); cswap(w[7], w[8]);
  cswap(w[0], w[1]); cswap(w[3], w[4]); cswap(w[6], w[7]);
  cswap(w[1], w[2]); cswap(w[4], w[5]); cswap(w[7], w[8]);
  cswap(w[0], w[3]); cswap(w[5], w[8]); cswap(w[4], w[7]);
  cswap(w[3], w[6]); cswap(w[1], w[4]); cswap(w[2], w[5]);
  cswap(w[4], w[7]); cswap(w[4], w[2]); cswap(w[6], w[4]);
  cswap(w[4], w[2]);
  dst[(size_t)v * W + u] = (int16_t)w[4];
}

extern "C" int sm_median3_i16(sm_ctx* ctx, const int16_t* d_src, int16_t* d_dst, int H, int W) {
  SM_CHECK_ARG(ctx && d_src && d_dst && d_src != d_dst && H > 0 && W > 0);
  dim3 grid(sm_div_up(W, 128), H);
  SM_LAUNCH(ctx, k_median3_i16, grid, 128, 0, d_src, d_dst, H, W);
  return SM_OK;
}

// 3x3 median, float, replicated border: cv::medianBlur(SE, SE, 3) after subpixelEnhancement (stereoMatching.cpp:1490).
// Same 19-exchange network on min/max; the input is NaN-free (SE holds converted shorts).
__device__ __forceinline__ void cswapf(float& a, float& b) {
  const float lo = fminf(a, b), hi = fmaxf(a, b);
  a = lo; b = hi;
}

__global__ void k_median3_f32(const float* __restrict__ src, float* __restrict__ dst, int H, int W) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  float w[9];
#pragma unroll
  for (int dv = -1; dv <= 1; dv++)
#pragma unroll
    for (int du = -1; du <= 1; du++) {
      const int y = min(max(v + dv, 0), H - 1), x = min(max(u + du, 0), W - 1);
      w[(dv + 1) * 3 + du + 1] = src[(size_t)y * W + x];
    }
  cswapf(w[1], w[2]); cswapf(w[4], w[5]); cswapf(w[7], w[8]);
  cswapf(w[0], w[1]); cswapf(w[3], w[4]); cswapf(w[6], w[7]);
  cswapf(w[1], w[2]); cswapf(w[4], w[5]); cswapf(w[7], w[8]);
  cswapf(w[0], w[3]); cswapf(w[5], w[8]); cswapf(w[4], w[7]);
  cswapf(w[3], w[6]); cswapf(w[1], w[4]); cswapf(w[2], w[5]);
  cswapf(w[4], w[7]); cswapf(w[4], w[2]); cswapf(w[6], w[4]);
  cswapf(w[4], w[2]);
  dst[(size_t)v * W + u] = w[4];
}

extern "C" int sm_median3_f32(sm_ctx* ctx, const float* d_src, float* d_dst, int H, int W) {
  SM_CHECK_ARG(ctx && d_src && d_dst && d_src != d_dst && H > 0 && W > 0);
  dim3 grid(sm_div_up(W, 128), H);
  SM_LAUNCH(ctx, k_median3_f32, grid, 128, 0, d_src, d_dst, H, W);
  return SM_OK;
}

// ------------------------------------------------------------------ 1-level cross-scale step
__global__ void k_scale(float* __restrict__ vol, size_t n, float inv) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) vol[i] = 0.0f + inv * vol[i];  // sum = 0; sum += invWgt * cost
}

extern "C" int sm_cross_scale_1level(sm_ctx* ctx, float* d_vol, size_t n, float lambda) {
  SM_CHECK_ARG(ctx && d_vol);
  const float inv = (float)(1. / (double)(1.0f + lambda));   // cv::invert of the 1x1 float matrix (1 + lambda)
  int grid = (int)min((size_t)ctx->num_sms * 16, (n + 255) / 256);
  SM_LAUNCH(ctx, k_scale, grid, 256, 0, d_vol, n, inv);
  return SM_OK;
}

// ------------------------------------------------------------------ gen_sgm_vm accumulation step
__global__ void k_vol_accumulate(float* __restrict__ acc, const float* __restrict__ x, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) acc[i] = acc[i] + x[i];   // sum += Lr[num]
}

extern "C" int sm_vol_accumulate(sm_ctx* ctx, float* d_acc, const float* d_x, size_t n) {
  SM_CHECK_ARG(ctx && d_acc && d_x);
  int grid = (int)min((size_t)ctx->num_sms * 16, (n + 255) / 256);
  SM_LAUNCH(ctx, k_vol_accumulate, grid > 0 ? grid : 1, 256, 0, d_acc, d_x, n);
  return SM_OK;
}
