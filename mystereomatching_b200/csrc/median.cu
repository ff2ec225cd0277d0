// K7: (2r+1)^2 median of an 8-bit image with cn interleaved channels and an
// edge-replicated border -- the result ctmf() produces (NL/ctmf.c:378-433; the
// helper initialises the first-row histograms with r+1 copies of row 0, :230, and
// clamps column indices, :244-252/:285/:310-311, i.e. replication, not the zero
// padding its doc comment claims).  The value returned is the (2r^2+2r)-th order
// statistic counted from 0 (:296 "t = 2*r*r + 2*r").
//
// The reference's constant-time histogram walk is a CPU cache trick; on the GPU
// the window (9 / 25 / 49 bytes) fits in registers and the median is found by
// rank counting, O((2r+1)^4) compares per output byte -- still a vanishing share
// of a frame (the image is O(H*W)).
#include "common.cuh"

template <int R>
__global__ void k_median_u8(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int H, int W, int cn) {
  constexpr int K = 2 * R + 1, N = K * K, T = 2 * R * R + 2 * R;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;  // byte column: u*cn + c
  const int v = blockIdx.y;
  if (x >= W * cn) return;
  const int u = x / cn, c = x - u * cn;
  int w[N];
#pragma unroll
  for (int dy = -R; dy <= R; dy++) {
    const int yy = min(max(v + dy, 0), H - 1);
#pragma unroll
    for (int dx = -R; dx <= R; dx++) {
      const int xx = min(max(u + dx, 0), W - 1);
      w[(dy + R) * K + dx + R] = src[((size_t)yy * W + xx) * cn + c];
    }
  }
  int res = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    int less = 0, leq = 0;
#pragma unroll
    for (int j = 0; j < N; j++) { less += w[j] < w[i]; leq += w[j] <= w[i]; }
    if (less <= T && T < leq) res = w[i];
  }
  dst[((size_t)v * W + u) * cn + c] = (uint8_t)res;
}

extern "C" int sm_median_u8(sm_ctx* ctx, const uint8_t* d_src, uint8_t* d_dst, int H, int W, int r, int cn) {
  SM_CHECK_ARG(ctx && d_src && d_dst && d_src != d_dst && H > 0 && W > 0 && cn > 0);
  SM_CHECK_ARG(r >= 1 && r <= 3);
  dim3 grid(sm_div_up((long long)W * cn, 128), H);
  if (r == 1) SM_LAUNCH(ctx, k_median_u8<1>, grid, 128, 0, d_src, d_dst, H, W, cn);
  else if (r == 2) SM_LAUNCH(ctx, k_median_u8<2>, grid, 128, 0, d_src, d_dst, H, W, cn);
  else SM_LAUNCH(ctx, k_median_u8<3>, grid, 128, 0, d_src, d_dst, H, W, cn);
  return SM_OK;
}
