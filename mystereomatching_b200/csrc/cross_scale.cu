// The caller's cross-scale step -- SURVEY.md 8(f) rank 1: the pyramid loop of main_.cpp:131-158 and SolveAll
// (stereoMatching.cpp:2142-2208), which sit between aggregation and SGM in the reference's main().
//
//   cv::pyrDown (main_.cpp:145-148) on the 8-bit colour and gray images: separable [1 4 6 4 1]/16,
//     BORDER_REFLECT_101, dst = ((W+1)/2, (H+1)/2), (sum + 128) >> 8.  Integer, bit-exact (pinned against cv2 in
//     tests/golden/opencv_semantics.npz).
//   SolveAll: invWgt = row 0 of the inverse of the tridiagonal float matrix {1+l, -l; -l, 1+2l, -l; ...}; then
//     vm0[y][x][d] = sum_s invWgt[s] * vm_s[y>>s][x>>s][d_s], d_{s+1} = (d_s + 1) / 2, accumulated in float in level
//     order from 0.  The inverse reproduces cv::invert(DECOMP_LU) on CV_32F bit for bit (n = 1, 3: closed forms in
//     double; n = 2: determinant in double, float products with (float)(1/det); n > 3: float Gaussian elimination
//     with partial pivoting) -- 1400 (lambda, n) pairs checked against cv2.invert.
//
// HBM traffic of the gather: level 0 is read and written once (2 V b); the coarser levels add 1/8 + 1/64 + ... of a
// volume and are mostly L2 hits (8 fine elements share a coarse one).
#include <math.h>

#include <vector>

#include "common.cuh"

__global__ void k_pyr_down_u8(const uint8_t* __restrict__ src, int H, int W, int cn, int Ho, int Wo,
                              uint8_t* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= Wo || y >= Ho) return;
  const int k[5] = {1, 4, 6, 4, 1};
  int xs[5], ys[5];
#pragma unroll
  for (int i = 0; i < 5; i++) {
    int p = 2 * x + i - 2, q = 2 * y + i - 2;
    if (W == 1) p = 0; else { while (p < 0 || p >= W) p = p < 0 ? -p : 2 * (W - 1) - p; }
    if (H == 1) q = 0; else { while (q < 0 || q >= H) q = q < 0 ? -q : 2 * (H - 1) - q; }
    xs[i] = p; ys[i] = q;
  }
  for (int c = 0; c < cn; c++) {
    int s = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
      int r = 0;
#pragma unroll
      for (int j = 0; j < 5; j++) r += k[j] * src[((size_t)ys[i] * W + xs[j]) * cn + c];
      s += k[i] * r;
    }
    dst[((size_t)y * Wo + x) * cn + c] = (uint8_t)((s + 128) >> 8);
  }
}

extern "C" int sm_pyr_down_u8(sm_ctx* ctx, const uint8_t* d_src, int H, int W, int cn, uint8_t* d_dst) {
  SM_CHECK_ARG(ctx && d_src && d_dst && H > 0 && W > 0 && (cn == 1 || cn == 3));
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  dim3 block(32, 8), grid(sm_div_up(Wo, 32), sm_div_up(Ho, 8));
  SM_LAUNCH(ctx, k_pyr_down_u8, grid, block, 0, d_src, H, W, cn, Ho, Wo, d_dst);
  return SM_OK;
}

extern "C" int sm_cross_scale_weights(int n, float lambda, float* invWgt) {
  SM_CHECK_ARG(n >= 1 && n <= SM_MAX_PYRAMID && invWgt);
  std::vector<float> M((size_t)n * n, 0.f);
  for (int s = 0; s < n; s++) {
    if (s == 0) { M[0] = 1 + lambda; if (n > 1) M[1] = -lambda; }
    else if (s == n - 1) { M[(size_t)s * n + s] = 1 + lambda; M[(size_t)s * n + s - 1] = -lambda; }
    else { M[(size_t)s * n + s] = 1 + 2 * lambda; M[(size_t)s * n + s - 1] = -lambda; M[(size_t)s * n + s + 1] = -lambda; }
  }
  if (n == 1) { invWgt[0] = (float)(1. / (double)M[0]); return SM_OK; }
  if (n == 2) {
    double d = (double)M[0] * M[3] - (double)M[1] * M[2];
    d = 1. / d;
    invWgt[0] = M[3] * (float)d;
    invWgt[1] = -(M[1] * (float)d);
    return SM_OK;
  }
  if (n == 3) {
    auto S = [&](int r, int c) { return (double)M[r * 3 + c]; };
    double d = S(0, 0) * (S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) - S(0, 1) * (S(1, 0) * S(2, 2) - S(1, 2) * S(2, 0)) +
               S(0, 2) * (S(1, 0) * S(2, 1) - S(1, 1) * S(2, 0));
    d = 1. / d;
    invWgt[0] = (float)((S(1, 1) * S(2, 2) - S(1, 2) * S(2, 1)) * d);
    invWgt[1] = (float)((S(0, 2) * S(2, 1) - S(0, 1) * S(2, 2)) * d);
    invWgt[2] = (float)((S(0, 1) * S(1, 2) - S(0, 2) * S(1, 1)) * d);
    return SM_OK;
  }
  std::vector<float> b((size_t)n * n, 0.f);
  for (int i = 0; i < n; i++) b[(size_t)i * n + i] = 1.f;
  volatile float t;   // keeps every product and sum a rounded float (no host-side contraction)
  for (int i = 0; i < n; i++) {
    int k = i;
    for (int j = i + 1; j < n; j++) if (fabsf(M[j * n + i]) > fabsf(M[k * n + i])) k = j;
    if (k != i) for (int c = 0; c < n; c++) { std::swap(M[i * n + c], M[k * n + c]); std::swap(b[i * n + c], b[k * n + c]); }
    const float d = -1 / M[i * n + i];
    for (int j = i + 1; j < n; j++) {
      const float alpha = M[j * n + i] * d;
      for (int c = i + 1; c < n; c++) { t = alpha * M[i * n + c]; M[j * n + c] += t; }
      for (int c = 0; c < n; c++) { t = alpha * b[i * n + c]; b[j * n + c] += t; }
    }
  }
  for (int i = n - 1; i >= 0; i--)
    for (int j = 0; j < n; j++) {
      float sacc = b[i * n + j];
      for (int c = i + 1; c < n; c++) { t = M[i * n + c] * b[c * n + j]; sacc -= t; }
      b[i * n + j] = sacc / M[i * n + i];
    }
  for (int s = 0; s < n; s++) invWgt[s] = b[s];
  return SM_OK;
}

struct cs_levels {
  const float* vol[SM_MAX_PYRAMID];
  int W[SM_MAX_PYRAMID], D[SM_MAX_PYRAMID];
  float w[SM_MAX_PYRAMID];
};

// One CTA walks pixels (grid-stride), its threads run along d: coalesced level-0 accesses, no per-element division
// (the first version's 64-bit i / D, i % D, % W made it issue-bound at 1.1 TB/s; this one: see DESIGN.md).
template <int LEVELS>
__global__ void __launch_bounds__(256) k_cross_scale(float* __restrict__ vol0, cs_levels L, int H, int W, int D) {
  const int npix = H * W;
  for (int p = blockIdx.x; p < npix; p += gridDim.x) {
    const int y = p / W, x = p - y * W;
    const float* src[LEVELS];
    int cy = y, cx = x;
#pragma unroll
    for (int s = 1; s < LEVELS; s++) {
      cy >>= 1; cx >>= 1;
      src[s] = L.vol[s] + ((size_t)cy * L.W[s] + cx) * L.D[s];
    }
    float* o = vol0 + (size_t)p * D;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      float sum = __fadd_rn(0.f, __fmul_rn(L.w[0], o[d]));
      int cd = d;
#pragma unroll
      for (int s = 1; s < LEVELS; s++) {
        cd = (cd + 1) >> 1;
        sum = __fadd_rn(sum, __fmul_rn(L.w[s], src[s][cd]));
      }
      o[d] = sum;
    }
  }
}

extern "C" int sm_cross_scale(sm_ctx* ctx, float* const* d_vols, const int* Hs, const int* Ws, const int* Ds, int levels,
                              float lambda) {
  SM_CHECK_ARG(ctx && d_vols && Hs && Ws && Ds && levels >= 1 && levels <= SM_MAX_PYRAMID);
  cs_levels L;
  SM_TRY(sm_cross_scale_weights(levels, lambda, L.w));
  for (int s = 0; s < levels; s++) {
    SM_CHECK_ARG(d_vols[s] && Hs[s] > 0 && Ws[s] > 0 && Ds[s] > 0);
    if (s > 0) {   // every index the gather forms must exist in the coarser level
      SM_CHECK_ARG(Hs[s] >= (Hs[s - 1] + 1) / 2 && Ws[s] >= (Ws[s - 1] + 1) / 2 && Ds[s] >= Ds[s - 1] / 2 + 1);
    }
    L.vol[s] = d_vols[s]; L.W[s] = Ws[s]; L.D[s] = Ds[s];
  }
  SM_CHECK_ARG((long long)Hs[0] * Ws[0] < (1ll << 31));
  const int npix = Hs[0] * Ws[0];
  const int block = Ds[0] >= 192 ? 256 : (Ds[0] >= 96 ? 128 : (Ds[0] >= 48 ? 64 : 32));
  const int grid = min(npix, ctx->num_sms * (2048 / block));
  float* v0 = d_vols[0];
  switch (levels) {
    case 1: SM_LAUNCH(ctx, k_cross_scale<1>, grid, block, 0, v0, L, Hs[0], Ws[0], Ds[0]); break;
    case 2: SM_LAUNCH(ctx, k_cross_scale<2>, grid, block, 0, v0, L, Hs[0], Ws[0], Ds[0]); break;
    case 3: SM_LAUNCH(ctx, k_cross_scale<3>, grid, block, 0, v0, L, Hs[0], Ws[0], Ds[0]); break;
    case 4: SM_LAUNCH(ctx, k_cross_scale<4>, grid, block, 0, v0, L, Hs[0], Ws[0], Ds[0]); break;
    case 5: SM_LAUNCH(ctx, k_cross_scale<5>, grid, block, 0, v0, L, Hs[0], Ws[0], Ds[0]); break;
    default: SM_LAUNCH(ctx, k_cross_scale<SM_MAX_PYRAMID>, grid, block, 0, v0, L, Hs[0], Ws[0], Ds[0]); break;
  }
  return SM_OK;
}
