// WM (stereoMatching.cpp:7340-7393): the default-off bilateral weighted median of refine() (Do_WM,
// stereoMatching.h:74): on every pixel with mask > 0 the 19 x 19 window (BORDER_REFLECT_101 on both the map and the
// guidance image) votes with weight exp(-|dI|^2 / 25^2 - |dx|^2 / 9^2) into a histogram over the labels, and the
// pixel takes the first label at which the running histogram sum reaches half the total weight.  The window reads
// the map as it was on entry (the reference copies it into disp_Bor first), so the pixels are independent.
//
// Bit-exactness needs three orders kept: a histogram bin adds its weights in window raster order, the total adds all
// 361 weights in raster order, and the running sum goes over the labels in ascending order -- all float.  One warp
// per flagged pixel: the 361 weights are computed lane-parallel into shared memory (expf as the host's libm computes
// it, smd_expf_host), then every lane walks the 361 entries IN ORDER and adds the ones whose label falls into a bin it
// owns (label mod 32 == lane): each bin has one owner, so its additions happen in raster order, and every lane forms
// the same total.  The final scan over the labels is sequential on lane 0.
//
// Labels outside [0, D) inside a window: the reference indexes its histogram with them unchecked (undefined
// behaviour, :7371).  Here such a neighbour still counts in the total weight (as in the reference) but casts no vote,
// and *d_numInvalid (nullable) counts them so a caller can tell.
#include "common.cuh"

#define WM_R 9
#define WM_N 361
#define WM_WARPS 4

__global__ void k_wm_scan(const uint8_t* __restrict__ mask, long long npix, int* __restrict__ count, int* __restrict__ list) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < npix && mask[p] > 0) list[atomicAdd(count, 1)] = (int)p;
}

__device__ __forceinline__ int wm_reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
  return p;
}

__global__ void __launch_bounds__(WM_WARPS * 32)
    k_wm(const int16_t* __restrict__ src, int16_t* __restrict__ dst, const uint8_t* __restrict__ bgr, int H, int W, int D,
         const int* __restrict__ count, const int* __restrict__ list, int* __restrict__ numInvalid) {
  extern __shared__ float wm_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wv = wm_smem + (size_t)warp * (2 * 384 + D);   // weights [384] | labels [384] (as int) | histogram [D]
  int* qv = reinterpret_cast<int*>(wv + 384);
  float* hist = wv + 768;
  const int n = *count;
  for (int e = blockIdx.x * WM_WARPS + warp; e < n; e += gridDim.x * WM_WARPS) {
    const int p = list[e];
    const int v = p / W, u = p - v * W;
    const uint8_t* ip = bgr + (size_t)p * 3;
    const int i0 = ip[0], i1 = ip[1], i2 = ip[2];
    for (int d = lane; d < D; d += 32) hist[d] = 0.0f;
    int bad = 0;
    for (int j = lane; j < WM_N; j += 32) {
      const int dv = j / 19 - WM_R, du = j - (j / 19) * 19 - WM_R;
      const size_t nb = (size_t)wm_reflect101(v + dv, H) * W + wm_reflect101(u + du, W);
      const int q = src[nb];
      const uint8_t* iq = bgr + nb * 3;
      const int a = i0 - iq[0], b = i1 - iq[1], c = i2 - iq[2];
      const float colDis = (float)(a * a + b * b + c * c), spaDis = (float)(dv * dv + du * du);
      // exp(-colDis / (SIG_CLR * SIG_CLR) - spaDis / (SIG_DIS * SIG_DIS)), float operations in this order
      const float arg = __fsub_rn(__fdiv_rn(-colDis, 625.0f), __fdiv_rn(spaDis, 81.0f));
      wv[j] = smd_expf_host(arg);
      qv[j] = q;
      bad += (q < 0 || q >= D);
    }
    __syncwarp();
    float wsum = 0.0f;
    for (int j = 0; j < WM_N; j++) {
      const float w = wv[j];
      const int q = qv[j];
      wsum = __fadd_rn(wsum, w);
      if ((q & 31) == lane && q >= 0 && q < D) hist[q] = __fadd_rn(hist[q], w);
    }
    __syncwarp();
    if (lane == 0) {
      const float half = __fdiv_rn(wsum, 2.0f);
      float cum = 0.0f;
      for (int d = 0; d < D; d++) {
        cum = __fadd_rn(cum, hist[d]);
        if (cum >= half) { dst[p] = (int16_t)d; break; }
      }
    }
    if (numInvalid) {
      for (int o = 16; o; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
      if (lane == 0 && bad) atomicAdd(numInvalid, bad);
    }
    __syncwarp();
  }
}

extern "C" int sm_wm(sm_ctx* ctx, int16_t* d_disp, int16_t* d_tmp, const uint8_t* d_mask, const uint8_t* d_bgr, int H, int W,
                     int D, int* d_numInvalid) {
  SM_CHECK_ARG(ctx && d_disp && d_tmp && d_mask && d_bgr && H > 0 && W > 0 && D > 0 && D <= 4096 && d_disp != d_tmp);
  const long long npix = (long long)H * W;
  SM_CHECK_ARG(npix < (1ll << 31));
  void* p;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_RVLIST, 256 + (size_t)npix * 4, &p));
  int* count = (int*)p;
  int* list = (int*)((uint8_t*)p + 256);
  SM_CUDA(cudaMemsetAsync(count, 0, sizeof(int), ctx->stream));
  if (d_numInvalid) SM_CUDA(cudaMemsetAsync(d_numInvalid, 0, sizeof(int), ctx->stream));
  // the window reads the map as it was on entry: d_tmp keeps that copy, d_disp receives the new labels
  SM_CUDA(cudaMemcpyAsync(d_tmp, d_disp, npix * sizeof(int16_t), cudaMemcpyDeviceToDevice, ctx->stream));
  SM_LAUNCH(ctx, k_wm_scan, (int)((npix + 255) / 256), 256, 0, d_mask, npix, count, list);
  const size_t smem = (size_t)WM_WARPS * (768 + D) * sizeof(float);
  SM_CUDA(cudaFuncSetAttribute(k_wm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)min((long long)ctx->num_sms * 8, (npix + WM_WARPS - 1) / WM_WARPS);
  SM_LAUNCH(ctx, k_wm, grid, WM_WARPS * 32, smem, d_tmp, d_disp, d_bgr, H, W, D, count, list, d_numInvalid);
  return SM_OK;
}
