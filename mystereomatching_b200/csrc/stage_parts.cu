// Stage-API sub-steps that the pipeline kernels fuse, exposed one by one so that EVERY public method of the
// reference's stage API has a device entry point with the reference's semantics:
//
//   gen1DCumu          stereoMatching.cpp:3896-3926   in-place running sum of vm / areaIS along -u or -v
//   cal1DCost          stereoMatching.h:1643-1715     cum[head] - cum[pre_tail] with materialised intersected arms
//   genfinalVm_cbca    stereoMatching.cpp:3969-3992   vm /= areaIS
//   updateCost<float>  stereoMatching.h:2205-2280     the SGM recurrence at ONE pixel
//   LRConsistencyCheck_new  stereoMatching.cpp:2367-2382   0/1 validity mask
//
// The whole-frame path never calls these (k_cbca_pass does gen1DCumu + cal1DCost + genfinalVm_cbca in one sweep
// without the area volume or the 5.3 GB intersection tensor; k_sgm_* run updateCost for a whole scan line per
// warp); they exist for callers that drive the reference's methods individually, and are bit-exact: same float
// additions in the same order.
#include <float.h>

#include "common.cuh"

// one thread per (scan line, d): lanes along d -> coalesced; the running sum stays in a register
__global__ void k_cumsum_1d(float* __restrict__ vol, int32_t* __restrict__ area, int H, int W, int D, int dir) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int line = blockIdx.y;
  if (d >= D) return;
  const int N = dir == 0 ? W : H;
  const size_t step = dir == 0 ? (size_t)D : (size_t)W * D;
  size_t i = (dir == 0 ? (size_t)line * W * D : (size_t)line * D) + d;
  float cum = 0.f;
  int ca = 0;
  for (int x = 0; x < N; x++, i += step) {
    // x = 0: the predecessor is outside the image, the value stays (no `0 + v`, which would turn -0 into +0)
    cum = x == 0 ? vol[i] : __fadd_rn(vol[i], cum);   // vm[x] += vm[x-1]
    vol[i] = cum;
    if (area) { ca = x == 0 ? area[i] : area[i] + ca; area[i] = ca; }
  }
}

extern "C" int sm_cumsum_1d(sm_ctx* ctx, float* d_vol, int32_t* d_areaIS, int H, int W, int D, int dv, int du) {
  SM_CHECK_ARG(ctx && d_vol && H > 0 && W > 0 && D > 0);
  SM_CHECK_ARG((dv == 0 && du == -1) || (dv == -1 && du == 0));   // the two calls cbca_core makes (:5608-5621)
  const int dir = du == -1 ? 0 : 1;
  dim3 block(64), grid(sm_div_up(D, 64), dir == 0 ? H : W);
  SM_LAUNCH(ctx, k_cumsum_1d, grid, block, 0, d_vol, d_areaIS, H, W, D, dir);
  return SM_OK;
}

__global__ void k_span_1d(const float* __restrict__ vol, const int32_t* __restrict__ area,
                          const uint16_t* __restrict__ hvlis, float* __restrict__ ovol, int32_t* __restrict__ oarea,
                          int H, int W, int D, int dv, int du, int direc) {
  const size_t n = (size_t)H * W * D;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int head_num = direc * 2 + 1, tail_num = direc * 2;
  for (; i < n; i += stride) {
    const int d = (int)(i % D);
    const size_t pxl = i / D;
    const int u = (int)(pxl % W), v = (int)(pxl / W);
    const uint16_t* a = hvlis + i * 5;
    const int tail_u = u + du * a[tail_num], tail_v = v + dv * a[tail_num];
    const int head_u = u - du * a[head_num], head_v = v - dv * a[head_num];
    const int pu = tail_u + du, pv = tail_v + dv;
    const bool inner = pu >= 0 && pu < W && pv >= 0 && pv < H;
    const size_t hi = ((size_t)head_v * W + head_u) * D + d, pi = ((size_t)pv * W + pu) * D + d;
    ovol[i] = inner ? __fsub_rn(vol[hi], vol[pi]) : vol[hi];
    if (area) oarea[i] = inner ? area[hi] - area[pi] : area[hi];
  }
}

extern "C" int sm_span_1d(sm_ctx* ctx, float* d_vol, int32_t* d_areaIS, const uint16_t* d_hvl_is, float* d_tmp_vol,
                          int32_t* d_tmp_area, int H, int W, int D, int dv, int du, int direc) {
  SM_CHECK_ARG(ctx && d_vol && d_hvl_is && d_tmp_vol && H > 0 && W > 0 && D > 0 && (direc == 0 || direc == 1));
  SM_CHECK_ARG((d_areaIS == nullptr) == (d_tmp_area == nullptr));
  SM_CHECK_ARG(dv >= -1 && dv <= 1 && du >= -1 && du <= 1);
  const size_t n = (size_t)H * W * D;
  const int grid = (int)min((size_t)ctx->num_sms * 16, (n + 255) / 256);
  SM_LAUNCH(ctx, k_span_1d, grid, 256, 0, d_vol, d_areaIS, d_hvl_is, d_tmp_vol, d_tmp_area, H, W, D, dv, du, direc);
  // vmTemp.copyTo(vm); areaISTemp.copyTo(areaIS)  (stereoMatching.h:1710-1714)
  SM_CUDA(cudaMemcpyAsync(d_vol, d_tmp_vol, n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
  if (d_areaIS) SM_CUDA(cudaMemcpyAsync(d_areaIS, d_tmp_area, n * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
  return SM_OK;
}

__global__ void k_div_area(float* __restrict__ vol, const int32_t* __restrict__ area, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) vol[i] = __fdiv_rn(vol[i], (float)area[i]);   // vmP[d] /= s (int -> float, IEEE division)
}

extern "C" int sm_div_area(sm_ctx* ctx, float* d_vol, const int32_t* d_areaIS, size_t n) {
  SM_CHECK_ARG(ctx && d_vol && d_areaIS);
  const int grid = (int)min((size_t)ctx->num_sms * 16, (n + 255) / 256);
  SM_LAUNCH(ctx, k_div_area, grid > 0 ? grid : 1, 256, 0, d_vol, d_areaIS, n);
  return SM_OK;
}

// one warp, one pixel: the reference's per-pixel primitive
__global__ void k_update_cost(float* __restrict__ Lr, const float* __restrict__ vm, const uint8_t* __restrict__ bgr, int W,
                              int n, int v, int u, int rv, int ru, int preIsInner, int corDifThres, int reduCoeffi1) {
  const int lane = threadIdx.x;
  const size_t cur = ((size_t)v * W + u) * n;
  if (!preIsInner) {
    for (int d = lane; d < n; d += 32) Lr[cur + d] = vm[cur + d];
    return;
  }
  const size_t pre = ((size_t)(v + rv) * W + (u + ru)) * n;
  int D1 = 0;
  for (int c = 0; c < 3; c++)
    D1 = max(D1, abs((int)bgr[((size_t)v * W + u) * 3 + c] - (int)bgr[((size_t)(v + rv) * W + u + ru) * 3 + c]));
  float minC = FLT_MAX;
  for (int d = lane; d < n; d += 32) minC = fminf(Lr[pre + d], minC);
  for (int o = 16; o; o >>= 1) minC = fminf(minC, __shfl_xor_sync(0xffffffffu, minC, o));
  float P1 = 1.0f, P2 = 3.0f;
  if (D1 > corDifThres) { P1 = __fdiv_rn(P1, (float)reduCoeffi1); P2 = __fdiv_rn(P2, (float)reduCoeffi1); }
  P1 = __fsub_rn(P1, minC);
  for (int d = lane; d < n; d += 32) {
    const float S1 = __fsub_rn(Lr[pre + d], minC);
    const float S2 = d - 1 >= 0 ? __fadd_rn(Lr[pre + d - 1], P1) : FLT_MAX;
    const float S3 = d + 1 < n ? __fadd_rn(Lr[pre + d + 1], P1) : FLT_MAX;
    Lr[cur + d] = __fadd_rn(vm[cur + d], fminf(fminf(S1, S2), fminf(S3, P2)));
  }
}

extern "C" int sm_update_cost(sm_ctx* ctx, float* d_Lr, const float* d_vm, const uint8_t* d_bgr, int H, int W, int n, int v,
                              int u, int rv, int ru, int preIsInner, int corDifThres, int reduCoeffi1) {
  SM_CHECK_ARG(ctx && d_Lr && d_vm && d_bgr && H > 0 && W > 0 && n > 0 && v >= 0 && v < H && u >= 0 && u < W);
  SM_CHECK_ARG(reduCoeffi1 != 0);
  if (preIsInner) SM_CHECK_ARG(v + rv >= 0 && v + rv < H && u + ru >= 0 && u + ru < W);
  SM_LAUNCH(ctx, k_update_cost, 1, 32, 0, d_Lr, d_vm, d_bgr, W, n, v, u, rv, ru, preIsInner, corDifThres, reduCoeffi1);
  return SM_OK;
}

__global__ void k_lrc_mask(const int16_t* __restrict__ D1, const int16_t* __restrict__ D2, int H, int W,
                           uint8_t* __restrict__ mask) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const size_t i = (size_t)v * W + u;
  const int d = D1[i];
  if (d < 0 || u - d < 0 || abs(d - (int)D2[i - d]) > 0) mask[i] = 0;   // Thres = 0 (stereoMatching.cpp:2369)
}

extern "C" int sm_lrc_mask(sm_ctx* ctx, const int16_t* d_D1, const int16_t* d_D2, int H, int W, uint8_t* d_mask) {
  SM_CHECK_ARG(ctx && d_D1 && d_D2 && d_mask && H > 0 && W > 0);
  dim3 grid(sm_div_up(W, 128), H);
  SM_LAUNCH(ctx, k_lrc_mask, grid, 128, 0, d_D1, d_D2, H, W, d_mask);
  return SM_OK;
}

// costScan dispatches on vm.depth() in {CV_8U, CV_16U, CV_32F} (stereoMatching.cpp:2007-2021): updateCost<uchar> /
// <ushort> read `T cost = vmPtr[d]` and add it to a float, i.e. they see the integer volume converted to float --
// exactly what this conversion followed by the float sweep computes.
template <typename T>
__global__ void k_to_f32(const T* __restrict__ src, size_t n, float* __restrict__ dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = (float)src[i];
}

extern "C" int sm_vol_to_f32(sm_ctx* ctx, const void* d_src, int elem_bytes, size_t n, float* d_dst) {
  SM_CHECK_ARG(ctx && d_src && d_dst && (elem_bytes == 1 || elem_bytes == 2));
  const int grid = (int)min((size_t)ctx->num_sms * 16, (n + 255) / 256);
  if (elem_bytes == 1) SM_LAUNCH(ctx, k_to_f32<uint8_t>, grid > 0 ? grid : 1, 256, 0, (const uint8_t*)d_src, n, d_dst);
  else SM_LAUNCH(ctx, k_to_f32<uint16_t>, grid > 0 ? grid : 1, 256, 0, (const uint16_t*)d_src, n, d_dst);
  return SM_OK;
}

// calErr<short> (stereoMatching.h:1748-1825), the evaluation the reference prints after every stage: over the pixels
// of a region mask (== 255): sumNum, errorNumer (|DT - DP| > THRES, or DP < 0) and errorValueSum (dif^2, or 2 for an
// invalid pixel).  The two counts are exact; the reference accumulates the squared error in a float sequentially,
// here it is a double sum over block partials (agrees to ~1e-7 relative).
__global__ void k_cal_err(const int16_t* __restrict__ dp, const float* __restrict__ dt, const uint8_t* __restrict__ mask,
                          long long n, int thres, unsigned long long* __restrict__ counts, double* __restrict__ esum) {
  unsigned long long c0 = 0, c1 = 0;
  double e = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (mask[i] != 255) continue;
    c0++;
    const int d = dp[i];
    if (d >= 0) {
      const float dif = fabsf(__fsub_rn(dt[i], (float)d));
      e += (double)dif * (double)dif;     // pow(dif, 2) is evaluated in double
      if (dif > (float)thres) c1++;
    } else {
      c1++;
      e += 2.0;
    }
  }
  for (int o = 16; o; o >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, o);
    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    e += __shfl_xor_sync(0xffffffffu, e, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (c0) atomicAdd(&counts[0], c0);
    if (c1) atomicAdd(&counts[1], c1);
    if (e != 0.0) atomicAdd(esum, e);
  }
}

extern "C" int sm_cal_err(sm_ctx* ctx, const int16_t* d_disp, const float* d_gt, const uint8_t* d_mask, int H, int W,
                          int thres, long long* h_sumNum, long long* h_errorNum, double* h_errorValueSum) {
  SM_CHECK_ARG(ctx && d_disp && d_gt && d_mask && H > 0 && W > 0 && h_sumNum && h_errorNum && h_errorValueSum);
  void* p;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_MISC0, 64, &p));
  SM_CUDA(cudaMemsetAsync(p, 0, 24, ctx->stream));
  const long long n = (long long)H * W;
  const int grid = (int)min((long long)ctx->num_sms * 4, (n + 255) / 256);
  SM_LAUNCH(ctx, k_cal_err, grid, 256, 0, d_disp, d_gt, d_mask, n, thres, (unsigned long long*)p,
            (double*)((char*)p + 16));
  unsigned long long h[3];
  SM_CUDA(cudaMemcpyAsync(h, p, 24, cudaMemcpyDeviceToHost, ctx->stream));
  SM_CUDA(cudaStreamSynchronize(ctx->stream));
  *h_sumNum = (long long)h[0];
  *h_errorNum = (long long)h[1];
  memcpy(h_errorValueSum, &h[2], 8);
  return SM_OK;
}
