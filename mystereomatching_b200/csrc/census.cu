// K1: census transform, 7 rows x 9 columns (R_V=3, R_U=4), centre included,
// bit = (centre < neighbour), MSB-first, BORDER_REFLECT_101.
//   func 0: genCensusCode<uchar>      (stereoMatching.h:634-688)  -> 63 bits, 1 word
//   func 3: genCensusCode_NC_Sur      (stereoMatching.h:867-934)  -> 63 + 8 ring bits, 2 words:
//           word0 = 63 window bits followed by ring bit 0, word1 = ring bits 1..7.
// The image tile (+halo) is staged in shared memory once per CTA; the kernel is
// O(H*W) and a negligible share of the frame.
#include "common.cuh"

#define CEN_RV 3
#define CEN_RU 4
#define CEN_TX 32
#define CEN_TY 8

__device__ __forceinline__ int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

template <int FUNC>
__global__ void __launch_bounds__(CEN_TX* CEN_TY)
    k_census(const uint8_t* __restrict__ gray, int H, int W, uint64_t* __restrict__ words) {
  constexpr int SW = CEN_TX + 2 * CEN_RU, SH = CEN_TY + 2 * CEN_RV;
  __shared__ uint8_t tile[SH][SW + 1];
  const int u0 = blockIdx.x * CEN_TX, v0 = blockIdx.y * CEN_TY;
  for (int i = threadIdx.y * CEN_TX + threadIdx.x; i < SH * SW; i += CEN_TX * CEN_TY) {
    int ty = i / SW, tx = i - ty * SW;
    int v = reflect101(v0 + ty - CEN_RV, H), u = reflect101(u0 + tx - CEN_RU, W);
    tile[ty][tx] = gray[(size_t)v * W + u];
  }
  __syncthreads();
  const int u = u0 + threadIdx.x, v = v0 + threadIdx.y;
  if (u >= W || v >= H) return;
  const int cy = threadIdx.y + CEN_RV, cx = threadIdx.x + CEN_RU;
  const int centre = tile[cy][cx];
  uint64_t cs = 0;
#pragma unroll
  for (int dv = -CEN_RV; dv <= CEN_RV; dv++)
#pragma unroll
    for (int du = -CEN_RU; du <= CEN_RU; du++) cs = (cs << 1) | (uint64_t)(centre < (int)tile[cy + dv][cx + du]);
  if (FUNC == 0) {
    words[(size_t)v * W + u] = cs;
  } else {
    // ring, clockwise from top-left; bit i = ring[i] < ring[i+1]
    const int rv[9] = {-1, -1, -1, 0, 1, 1, 1, 0, -1};
    const int ru[9] = {-1, 0, 1, 1, 1, 0, -1, -1, -1};
    uint32_t ring = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
      ring = (ring << 1) | (uint32_t)((int)tile[cy + rv[i]][cx + ru[i]] < (int)tile[cy + rv[i + 1]][cx + ru[i + 1]]);
    ulonglong2 o;
    o.x = (cs << 1) | (uint64_t)(ring >> 7);  // 63 window bits then ring bit 0
    o.y = (uint64_t)(ring & 0x7f);            // ring bits 1..7, right aligned
    reinterpret_cast<ulonglong2*>(words)[(size_t)v * W + u] = o;
  }
}

extern "C" int sm_census_words(int func) { return func == 3 ? 2 : 1; }
extern "C" int sm_census_code_length(int func) { return func == 3 ? 71 : 63; }

extern "C" int sm_census(sm_ctx* ctx, const uint8_t* d_gray, int H, int W, int func, uint64_t* d_words) {
  SM_CHECK_ARG(ctx && d_gray && d_words && H > 0 && W > 0);
  SM_CHECK_ARG(func == 0 || func == 3);
  SM_CHECK_ARG(((uintptr_t)d_words & 15) == 0);
  dim3 block(CEN_TX, CEN_TY), grid(sm_div_up(W, CEN_TX), sm_div_up(H, CEN_TY));
  if (func == 0)
    SM_LAUNCH(ctx, k_census<0>, grid, block, 0, d_gray, H, W, d_words);
  else
    SM_LAUNCH(ctx, k_census<3>, grid, block, 0, d_gray, H, W, d_words);
  return SM_OK;
}
