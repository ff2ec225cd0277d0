// discontinuityAdjust (stereoMatching.cpp:6057-6135; Do_discontinuityAdjust, stereoMatching.h:78, off).
//
//   disp.convertTo(8U) -> equalizeHist -> GaussianBlur(3x3, sigma 4) -> Canny(20, 60, 3)          the edge map
//   every interior edge pixel whose 3x3 edge neighbourhood names a direction takes, of its own disparity and the two
//   neighbours' ACROSS that direction, the one with the smallest cost in vm[0] -- in place, in raster order.
//
// The three OpenCV calls are restated as integer kernels (semantics pinned against cv2 4.13, tests/golden/da_ref.npz):
//   k_da_hist / k_da_lut   saturating 16S -> 8U, 256-bin histogram, lut[i] = saturate(rint(cumsum * 255.f / (N - hist[i0])))
//   k_da_blur              lut applied on the fly, Q8 kernel {84, 88, 84} both ways, (.. + 2^15) >> 16, BORDER_REFLECT_101
//   k_da_sobel / k_da_nms  3x3 Sobel (replicated border), |dx| + |dy|, sector test with TG22 in Q15
//   hysteresis             = "a candidate pixel is an edge iff its 8-connected candidate component holds a pixel above the
//                          high threshold": union-find over the candidates (k_da_union, atomicMin hooking), roots of strong
//                          pixels flagged (k_da_mark), k_da_edge writes the map.  No host round trip, no iteration count.
// The raster-order pick looks sequential, but a pixel reads exactly ONE already-visited neighbour (above-left / above /
// above-right / left, by direction) and one not-yet-visited one (still original).  So the directions whose first neighbour
// lies in the row above are resolved 32 pixels at a time once that row has published the columns they read, and only the
// horizontal pairs (first neighbour = left pixel) run as a chain inside the warp: one warp per row, rows handed out by an
// atomic ticket, progress counters with release / acquire -- the scheme of vmtop.cu's k_top2_resolve.
// Labels >= D (the reference would index vm[0] out of bounds: undefined) are never used as indices: such a centre pixel is
// left alone, such a neighbour is no candidate (the tests' CPU checker makes the same choice).
#include "common.cuh"

__global__ void k_da_hist(const int16_t* __restrict__ disp, long long n, uint8_t* __restrict__ img, int* __restrict__ hist) {
  __shared__ int h[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int d = disp[i];
    const int v = d < 0 ? 0 : d > 255 ? 255 : d;
    img[i] = (uint8_t)v;
    atomicAdd(&h[v], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (h[i]) atomicAdd(&hist[i], h[i]);
}

__global__ void k_da_lut(const int* __restrict__ hist, long long total, uint8_t* __restrict__ lut) {
  if (threadIdx.x != 0) return;
  int i0 = 0;
  while (i0 < 255 && hist[i0] == 0) i0++;
  for (int i = 0; i < 256; i++) lut[i] = 0;
  if ((long long)hist[i0] == total) { lut[i0] = (uint8_t)i0; return; }   // constant image: kept
  const float scale = __fdiv_rn(255.f, (float)(total - hist[i0]));
  long long sum = 0;
  for (int i = i0 + 1; i < 256; i++) {
    sum += hist[i];
    int r = __float2int_rn(__fmul_rn((float)sum, scale));
    lut[i] = (uint8_t)(r < 0 ? 0 : r > 255 ? 255 : r);
  }
}

__device__ __forceinline__ int da_reflect(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

__global__ void k_da_blur(const uint8_t* __restrict__ img, const uint8_t* __restrict__ lut, int H, int W, uint8_t* __restrict__ out) {
  __shared__ uint8_t sl[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sl[i] = lut[i];
  __syncthreads();
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const int ul = da_reflect(u - 1, W), ur = da_reflect(u + 1, W);
  int hq[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const uint8_t* r = img + (size_t)da_reflect(v - 1 + k, H) * W;
    hq[k] = 84 * ((int)sl[r[ul]] + (int)sl[r[ur]]) + 88 * (int)sl[r[u]];
  }
  out[(size_t)v * W + u] = (uint8_t)((84 * (hq[0] + hq[2]) + 88 * hq[1] + (1 << 15)) >> 16);
}

__global__ void k_da_sobel(const uint8_t* __restrict__ img, int H, int W, short2* __restrict__ dxy, uint16_t* __restrict__ mag) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const int u0 = max(u - 1, 0), u2 = min(u + 1, W - 1);
  const uint8_t* r0 = img + (size_t)max(v - 1, 0) * W;
  const uint8_t* r1 = img + (size_t)v * W;
  const uint8_t* r2 = img + (size_t)min(v + 1, H - 1) * W;
  const int gx = ((int)r0[u2] + 2 * (int)r1[u2] + (int)r2[u2]) - ((int)r0[u0] + 2 * (int)r1[u0] + (int)r2[u0]);
  const int gy = ((int)r2[u0] + 2 * (int)r2[u] + (int)r2[u2]) - ((int)r0[u0] + 2 * (int)r0[u] + (int)r0[u2]);
  dxy[(size_t)v * W + u] = make_short2((short)gx, (short)gy);
  mag[(size_t)v * W + u] = (uint16_t)(abs(gx) + abs(gy));
}

// map: 1 = no edge, 0 = candidate not above `high`, 2 = above `high`; label = own index for candidates, -1 otherwise
__global__ void k_da_nms(const short2* __restrict__ dxy, const uint16_t* __restrict__ mag, int H, int W, int low, int high,
                         uint8_t* __restrict__ map, int* __restrict__ label, uint8_t* __restrict__ strong) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const size_t p = (size_t)v * W + u;
  auto M = [&](int y, int x) -> int { return (y < 0 || y >= H || x < 0 || x >= W) ? 0 : (int)mag[(size_t)y * W + x]; };
  const int m = mag[p];
  int res = 1;
  if (m > low) {
    const int xs = dxy[p].x, ys = dxy[p].y;
    const long long x = abs(xs), y = (long long)abs(ys) << 15;
    const long long tg22x = x * 13573;   // (int)(0.41421356 * (1 << 15) + 0.5)
    bool ok;
    if (y < tg22x) ok = m > M(v, u - 1) && m >= M(v, u + 1);
    else if (y > tg22x + (x << 16)) ok = m > M(v - 1, u) && m >= M(v + 1, u);
    else {
      const int s = (xs ^ ys) < 0 ? -1 : 1;
      ok = m > M(v - 1, u - s) && m > M(v + 1, u + s);
    }
    if (ok) res = m > high ? 2 : 0;
  }
  map[p] = (uint8_t)res;
  label[p] = res != 1 ? (int)p : -1;
  strong[p] = 0;
}

__device__ __forceinline__ int da_find(int* L, int x) {
  int p;
  while ((p = ((volatile int*)L)[x]) != x) x = p;
  return x;
}
__device__ void da_unite(int* L, int a, int b) {
  for (;;) {
    a = da_find(L, a);
    b = da_find(L, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&L[a], b);   // hook the larger root under the smaller
    if (old == a) return;
    a = old;                               // a had been hooked meanwhile: go on with what it pointed to
  }
}

__global__ void k_da_union(int H, int W, int* label) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const int p = v * W + u;
  if (label[p] < 0) return;
  if (u > 0 && label[p - 1] >= 0) da_unite(label, p, p - 1);
  if (v > 0) {
    const int q = p - W;
    if (u > 0 && label[q - 1] >= 0) da_unite(label, p, q - 1);
    if (label[q] >= 0) da_unite(label, p, q);
    if (u < W - 1 && label[q + 1] >= 0) da_unite(label, p, q + 1);
  }
}

__global__ void k_da_mark(const uint8_t* __restrict__ map, long long n, int* label, uint8_t* strong) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n || map[p] != 2) return;
  strong[da_find(label, (int)p)] = 1;
}

__global__ void k_da_edge(long long n, int* label, const uint8_t* __restrict__ strong, uint8_t* __restrict__ edge) {
  const long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n) return;
  edge[p] = (label[p] >= 0 && strong[da_find(label, (int)p)]) ? 255 : 0;
}

// pend: 0 = nothing to do; 1 + direction (1, 3, 5, 7 for the reference's directions 0, 2, 4, 6)
__global__ void k_da_dir(const uint8_t* __restrict__ E, int H, int W, uint8_t* __restrict__ pend) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const size_t p = (size_t)v * W + u;
  int code = 0;
  if (v >= 1 && v < H - 1 && u >= 1 && u < W - 1 && E[p]) {
    const uint8_t *a = E + p - W, *b = E + p, *c = E + p + W;
    int dir = -1;
    if (a[-1] && c[1]) dir = 4;
    else if (a[1] && c[-1]) dir = 0;
    else if (a[0] || a[-1] || a[1]) { if (c[0] || c[-1] || c[1]) dir = 6; }
    else if ((a[-1] || b[-1] || c[-1]) && (a[1] || b[1] || c[1])) dir = 2;
    code = dir + 1;
  }
  pend[p] = (uint8_t)code;
}

__device__ __forceinline__ int da_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void da_st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// the pick of stereoMatching.cpp:6105-6130 for the centre p (label dp) and the neighbours p1 (d1), p2 (d2)
__device__ __forceinline__ int da_pick(const float* __restrict__ vol, int D, size_t p, int dp, size_t p1, int d1, size_t p2, int d2) {
  if (dp < 0 || dp >= D) return dp;
  float cost = vol[p * D + dp];
  const float cost1 = (d1 >= 0 && d1 < D) ? vol[p1 * D + d1] : -1.f;
  const float cost2 = (d2 >= 0 && d2 < D) ? vol[p2 * D + d2] : -1.f;
  if (cost1 >= 0.f && cost1 < cost) { dp = d1; cost = cost1; }
  if (cost2 != -1.f && cost2 < cost) dp = d2;
  return dp;
}

#define DA_RES_WARPS 4
__global__ void __launch_bounds__(DA_RES_WARPS * 32)
    k_da_resolve(const float* __restrict__ vol, const uint8_t* __restrict__ pend, int H, int W, int D, int16_t* disp, int* __restrict__ ticket,
                 int* __restrict__ progress) {
  const int lane = threadIdx.x & 31;
  for (;;) {
    int v = 0;
    if (lane == 0) v = atomicAdd(ticket, 1);
    v = __shfl_sync(0xffffffffu, v, 0);
    if (v >= H) return;
    int carry = 0;   // final value of the pixel left of the chunk
    for (int c0 = 0; c0 < W; c0 += 32) {
      const int u = c0 + lane;
      const bool in = u < W;
      const size_t p = (size_t)v * W + (in ? u : 0);
      const int orig = in ? (int)disp[p] : 0;
      int cur = orig;
      const int code = in ? (int)pend[p] : 0;
      if (__ballot_sync(0xffffffffu, code != 0)) {   // only interior rows / columns carry a code
        if (lane == 0) {
          const int need = min(W, c0 + 33);
          while (da_ld_acquire(progress + v - 1) < need) __nanosleep(64);
        }
        __syncwarp();
        if (code == 1 || code == 3 || code == 5) {
          // first neighbour in the row above (final there; written by another SM: read past L1), second one below (original)
          const int du = code == 1 ? -1 : code == 3 ? 0 : 1;
          const size_t p1 = p - W + du, p2 = p + W - du;
          cur = da_pick(vol, D, p, orig, p1, (int)__ldcg(disp + p1), p2, (int)disp[p2]);
        }
        __syncwarp();
        // horizontal pairs: the left pixel's FINAL value, the right pixel's ORIGINAL one
        int right = __shfl_down_sync(0xffffffffu, orig, 1);
        if (lane == 31 && code == 7) right = disp[p + 1];
        unsigned mask = __ballot_sync(0xffffffffu, code == 7);
        while (mask) {
          const int l = __ffs(mask) - 1;
          mask &= mask - 1;
          int left = __shfl_sync(0xffffffffu, cur, (l + 31) & 31);
          if (l == 0) left = carry;
          if (lane == l) cur = da_pick(vol, D, p, orig, p - 1, left, p + 1, right);
        }
        if (code != 0) disp[p] = (int16_t)cur;
      }
      carry = __shfl_sync(0xffffffffu, cur, 31);
      __threadfence();
      __syncwarp();
      if (lane == 0) da_st_release(progress + v, min(W, c0 + 32));
    }
  }
}

extern "C" int sm_discontinuity_adjust(sm_ctx* ctx, int16_t* d_disp, const float* d_vol, int H, int W, int D, uint8_t* d_edge) {
  SM_CHECK_ARG(ctx && d_disp && d_vol && H > 0 && W > 0 && D > 0);
  SM_CHECK_ARG((long long)H * W < (1ll << 31));
  const size_t n = (size_t)H * W;
  // scratch: hist | lut | ticket | progress[H] | img a | img b | map | strong | edge | pend | mag u16 | dxy | label
  const size_t head = 256 * 4 + 256 + 256 + (size_t)H * 4;
  const size_t headA = (head + 255) / 256 * 256, nA = (n + 255) / 256 * 256;
  void* s;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_RVLIST, headA + nA * 6 + nA * 2 + nA * 4 + nA * 4, &s));
  uint8_t* base = (uint8_t*)s;
  int* hist = (int*)base;
  uint8_t* lut = base + 1024;
  int* ticket = (int*)(base + 1024 + 256);
  int* progress = (int*)(base + 1024 + 512);
  uint8_t* a = base + headA;
  uint8_t* b = a + nA;
  uint8_t* map = b + nA;
  uint8_t* strong = map + nA;
  uint8_t* edge = strong + nA;
  uint8_t* pend = edge + nA;
  uint16_t* mag = (uint16_t*)(pend + nA);
  short2* dxy = (short2*)((uint8_t*)mag + nA * 2);
  int* label = (int*)((uint8_t*)dxy + nA * 4);
  SM_CUDA(cudaMemsetAsync(base, 0, head, ctx->stream));
  const dim3 grid(sm_div_up(W, 128), H);
  const int lin = sm_div_up((long long)n, 256);
  SM_LAUNCH(ctx, k_da_hist, min(lin, ctx->num_sms * 8), 256, 0, d_disp, (long long)n, a, hist);
  SM_LAUNCH(ctx, k_da_lut, 1, 32, 0, hist, (long long)n, lut);
  SM_LAUNCH(ctx, k_da_blur, grid, 128, 0, a, lut, H, W, b);
  SM_LAUNCH(ctx, k_da_sobel, grid, 128, 0, b, H, W, dxy, mag);
  SM_LAUNCH(ctx, k_da_nms, grid, 128, 0, dxy, mag, H, W, 20, 60, map, label, strong);
  SM_LAUNCH(ctx, k_da_union, grid, 128, 0, H, W, label);
  SM_LAUNCH(ctx, k_da_mark, lin, 256, 0, map, (long long)n, label, strong);
  SM_LAUNCH(ctx, k_da_edge, lin, 256, 0, (long long)n, label, strong, edge);
  SM_LAUNCH(ctx, k_da_dir, grid, 128, 0, edge, H, W, pend);
  const int blocks = min(sm_div_up(H, DA_RES_WARPS), ctx->num_sms * 4);
  SM_LAUNCH(ctx, k_da_resolve, blocks, DA_RES_WARPS * 32, 0, d_vol, pend, H, W, D, d_disp, ticket, progress);
  if (d_edge) SM_CUDA(cudaMemcpyAsync(d_edge, edge, n, cudaMemcpyDeviceToDevice, ctx->stream));
  return SM_OK;
}
