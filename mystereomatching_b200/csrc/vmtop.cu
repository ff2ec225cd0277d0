// vmTop, second half: the disparity of a pixel from its candidate list
//   topDisp[v][u][k] = {d, cost}, k < num;  topDisp[v][u][num][0] = candidate count   (sm_select_top_cost)
//
//   genDispFromTopCostVm    stereoMatching.h:2466-2545     -> sm_disp_from_top   (reads topDisp only: one thread per pixel)
//   genDispFromTopCostVm2   stereoMatching.cpp:1514-1886   -> sm_disp_from_top2  (param_.vmTop_method 0 / 1 / 2)
//
// Method 0 is the reference author's own heuristic and looks strictly sequential (a raster scan over std::maps that
// reads the disparities it has just written), but only ONE of its three cases does: a pixel whose candidates have no
// partner within `ts` disparities takes the candidate nearest to its left / up / up-left / up-right neighbour's final
// value.  Every other pixel (image border, a single candidate, the neighbourhood vote) depends on topDisp alone.  So:
//   k_top2_local    one thread per pixel: all independent cases; the dependent ones are flagged;
//   k_top2_resolve  the flagged pixels, in raster-dependency order: one warp per row, rows handed out by an atomic
//                   ticket (a row only ever waits for a row with a lower ticket, whose warp is already running: no
//                   deadlock whatever the residency), 32 pixels per step; a step waits until the row above has
//                   published column u+1 (progress counter, release / acquire), then resolves its flagged pixels left
//                   to right inside the warp (the left neighbour's value travels by shuffle).
// The std::maps become fixed arrays of at most SM_TOP_MAX entries: map<float,int>::insert = "append unless an
// equivalent cost key exists", iteration = ascending key (insertion sort), map<int,.> iteration = ascending disparity.
// Methods 1 / 2 are scans along a row against the previous pixel's result: one thread per row.
#include "common.cuh"

#define SM_TOP_MAX 16   // candidates per pixel the kernels hold in registers / local memory

struct top_view {
  const float* top;
  int H, W, num;
  __device__ __forceinline__ const float* at(int v, int u, int k) const {
    return top + (((size_t)v * W + u) * (num + 1) + k) * 2;
  }
  __device__ __forceinline__ int count(int v, int u) const { return (int)at(v, u, num)[0]; }
};

struct top_votes {
  int d[SM_TOP_MAX], n[SM_TOP_MAX], cnt;
  float c[SM_TOP_MAX];
  __device__ __forceinline__ void add(int dd, float cc, bool create) {
    for (int i = 0; i < cnt; i++)
      if (d[i] == dd) { n[i]++; c[i] = __fadd_rn(c[i], cc); return; }
    if (create && cnt < SM_TOP_MAX) { d[cnt] = dd; n[cnt] = 1; c[cnt] = __fadd_rn(0.0f, cc); cnt++; }
  }
  __device__ __forceinline__ void sort_by_d() {
    for (int i = 1; i < cnt; i++) {
      const int dd = d[i], nn = n[i];
      const float cc = c[i];
      int j = i - 1;
      for (; j >= 0 && d[j] > dd; j--) { d[j + 1] = d[j]; n[j + 1] = n[j]; c[j + 1] = c[j]; }
      d[j + 1] = dd; n[j + 1] = nn; c[j + 1] = cc;
    }
  }
};

// ------------------------------------------------------------------ genDispFromTopCostVm
__global__ void k_top_v1(top_view T, int16_t* __restrict__ disp) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= T.W) return;
  const float cntf = T.at(v, u, T.num)[0];
  int16_t* out = disp + (size_t)v * T.W + u;
  if (cntf == 1.0f) { *out = (int16_t)T.at(v, u, 0)[0]; return; }
  if (!(cntf > 1.0f)) return;
  top_votes vs;
  vs.cnt = 0;
  for (int i = 0; i < cntf && i < SM_TOP_MAX; i++) vs.add((int)T.at(v, u, i)[0], T.at(v, u, i)[1], true);
  for (int du = -1; du <= 1; du += 2) {
    const int u_ = u + du;
    if (u_ < 0 || u_ >= T.W) continue;
    const int n_ = min(T.count(v, u_), T.num);
    for (int k = 0; k < n_; k++) vs.add((int)T.at(v, u_, k)[0], T.at(v, u_, k)[1], false);
  }
  vs.sort_by_d();
  int best = -1, bestN = -1;
  float bestC = 3.402823466e+38f;
  for (int i = 0; i < vs.cnt; i++) {
    int dNum = vs.n[i];
    bool take = dNum > bestN;
    if (!take) { dNum = (bestN != 0 && vs.c[i] < bestC) ? 1 : 0; take = dNum != 0; }   // `dNum = dispNum && cost_ < cost` (:2531)
    if (take) { bestN = dNum; bestC = vs.c[i]; best = vs.d[i]; }
  }
  *out = (int16_t)best;
}

// ------------------------------------------------------------------ genDispFromTopCostVm2, method 0
// the nearest candidate to each of the four already written neighbours (the reference's "case 2")
__device__ __forceinline__ int top2_nearest(const top_view& T, int v, int u, int n, int pre1, int pre2, int rt, int lt) {
  const int ref[4] = {pre1, pre2, rt, lt};
  int small[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff}, pick[4] = {-1, -1, -1, -1};
  for (int i = 0; i < n; i++) {
    const int dd = (int)T.at(v, u, i)[0];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int dif = abs(dd - ref[k]);
      if (dif < small[k]) { small[k] = dif; pick[k] = dd; }
    }
  }
  const int m = min(min(small[2], small[3]), min(small[0], small[1]));
  int d = -1;
  if (m == small[3]) d = pick[3];        // up-left first, then left, up, up-right (stereoMatching.cpp:1640-1647)
  else if (m == small[0]) d = pick[0];
  else if (m == small[1]) d = pick[1];
  else if (m == small[2]) d = pick[2];
  return m < 1000 ? d : (int)T.at(v, u, 0)[0];
}

__global__ void k_top2_local(top_view T, const uint8_t* __restrict__ bgr, int ts, int hasCir2, int colorLimit,
                             int16_t* __restrict__ disp, uint8_t* __restrict__ pend) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= T.W) return;
  const size_t p = (size_t)v * T.W + u;
  pend[p] = 0;
  if (u == 0 || v == 0) { disp[p] = (int16_t)T.at(v, u, 0)[0]; return; }
  int n = T.count(v, u);
  if (n == 1) { disp[p] = (int16_t)T.at(v, u, 0)[0]; return; }
  if (n < 1) return;
  n = min(n, T.num);
  // candidates kept, keyed by cost: std::map<float,int>::insert keeps the first entry of an equivalent key
  float kc[SM_TOP_MAX];
  int kd[SM_TOP_MAX], nk = 0;
  auto insert = [&](float c, int d) {
    for (int i = 0; i < nk; i++)
      if (!(kc[i] < c) && !(c < kc[i])) return;
    if (nk < SM_TOP_MAX) { kc[nk] = c; kd[nk] = d; nk++; }
  };
  for (int i = 0; i < n; i++) {
    const int d0 = (int)T.at(v, u, i)[0];
    const float c0 = T.at(v, u, i)[1];
    if (!hasCir2) { insert(c0, d0); continue; }
    for (int j = i + 1; j < n; j++) {
      const int d1 = (int)T.at(v, u, j)[0];
      if (abs(d0 - d1) < ts) { insert(c0, d0); insert(T.at(v, u, j)[1], d1); }
    }
  }
  if (nk == 0) { pend[p] = 1; return; }   // depends on the neighbours' final values: k_top2_resolve
  for (int i = 1; i < nk; i++) {          // ascending cost
    const float c = kc[i];
    const int d = kd[i];
    int j = i - 1;
    for (; j >= 0 && kc[j] > c; j--) { kc[j + 1] = kc[j]; kd[j + 1] = kd[j]; }
    kc[j + 1] = c; kd[j + 1] = d;
  }
  top_votes vs;
  vs.cnt = 0;
  for (int i = 0; i < nk; i++) vs.add(kd[i], kc[i], true);
  const int NV[8] = {0, -1, 0, 1, -1, 1, -1, 1}, NU[8] = {-1, 0, 1, 0, -1, 1, 1, -1};   // l, u, r, d, lu, rd, ru, ld
  const uint8_t* tar = bgr + p * 3;
  for (int k = 0; k < 8; k++) {
    const int v_ = v + NV[k], u_ = u + NU[k];
    if (v_ < 0 || v_ >= T.H || u_ < 0 || u_ >= T.W) continue;
    if (colorLimit) {   // judgeColorDif(tarP, neiP, 10, 3)
      const uint8_t* nei = bgr + ((size_t)v_ * T.W + u_) * 3;
      if (abs((int)tar[0] - (int)nei[0]) > 10 || abs((int)tar[1] - (int)nei[1]) > 10 || abs((int)tar[2] - (int)nei[2]) > 10) continue;
    }
    const int n_ = min(T.count(v_, u_), T.num);
    for (int x = 0; x < n_; x++) vs.add((int)T.at(v_, u_, x)[0], T.at(v_, u_, x)[1], false);
  }
  // ascending disparity, `num > num_most || num == num_most && cost < cost_A`: most votes, then lowest summed cost, then
  // lowest disparity -- an order-free form of the same rule
  int best = -1, bestN = -1;
  float bestC = 3.402823466e+38f;
  for (int i = 0; i < vs.cnt; i++) {
    const bool better = vs.n[i] > bestN || (vs.n[i] == bestN && (vs.c[i] < bestC || (!(bestC < vs.c[i]) && vs.d[i] < best)));
    if (better) { bestN = vs.n[i]; bestC = vs.c[i]; best = vs.d[i]; }
  }
  disp[p] = (int16_t)best;
}

__device__ __forceinline__ int top_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void top_st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

#define TOP_RES_WARPS 4
__global__ void __launch_bounds__(TOP_RES_WARPS * 32)
    k_top2_resolve(top_view T, const uint8_t* __restrict__ pend, int16_t* disp, int* __restrict__ ticket, int* __restrict__ progress) {
  const int lane = threadIdx.x & 31;
  const int W = T.W;
  for (;;) {
    int v = 0;
    if (lane == 0) v = atomicAdd(ticket, 1);
    v = __shfl_sync(0xffffffffu, v, 0);
    if (v >= T.H) return;
    int carry = 0;   // final value of the pixel left of the chunk
    for (int c0 = 0; c0 < W; c0 += 32) {
      const int u = c0 + lane;
      const bool in = u < W;
      const size_t p = (size_t)v * W + (in ? u : 0);
      int cur = in ? (int)disp[p] : 0;
      const bool isP = in && pend[p] != 0;
      const unsigned mask0 = __ballot_sync(0xffffffffu, isP);
      if (mask0) {   // (row 0 has no flagged pixel)
        if (lane == 0) {
          const int need = min(W, c0 + 33);
          while (top_ld_acquire(progress + v - 1) < need) __nanosleep(64);
        }
        __syncwarp();
        // the row above is final up to column u + 1; its flagged pixels were written by another SM: read past L1
        const int16_t* up = disp + (size_t)(v - 1) * W;
        int upL = 0, upC = 0, upR = 10000;
        if (isP) {
          upL = __ldcg(up + u - 1);
          upC = __ldcg(up + u);
          if (u != W - 1) upR = __ldcg(up + u + 1);
        }
        unsigned mask = mask0;
        while (mask) {
          const int l = __ffs(mask) - 1;
          mask &= mask - 1;
          int left = __shfl_sync(0xffffffffu, cur, (l + 31) & 31);
          if (l == 0) left = carry;
          if (lane == l) {
            const int n = min(T.count(v, u), T.num);
            cur = top2_nearest(T, v, u, n, left, upC, upR, upL);
          }
        }
        if (isP) disp[p] = (int16_t)cur;
      }
      carry = __shfl_sync(0xffffffffu, cur, 31);
      __threadfence();
      __syncwarp();
      if (lane == 0) top_st_release(progress + v, min(W, c0 + 32));
    }
  }
}

// ------------------------------------------------------------------ genDispFromTopCostVm2, methods 1 and 2
__global__ void k_top2_rows(top_view T, const uint8_t* __restrict__ bgr, int method, int16_t* __restrict__ disp) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= T.H) return;
  int16_t* row = disp + (size_t)v * T.W;
  for (int u = 0; u < T.W; u++) {
    const int n = min(T.count(v, u), T.num);
    const int first = (int)T.at(v, u, 0)[0];
    if (u == 0 || T.count(v, u) == 1) { row[u] = (int16_t)first; continue; }
    const int pre = row[u - 1];
    if (method == 1) {
      int dp = -1, best = 10000;
      for (int k = 0; k < n; k++) {
        const int s = (int)fabsf((float)(int16_t)pre - T.at(v, u, k)[0]);
        if (s < 2 && s < best) { best = s; dp = (int)T.at(v, u, k)[0]; }
      }
      row[u] = (int16_t)(dp == -1 ? first : dp);
    } else {
      int bestPre = 1000000, bestAft = 1000000, d0 = -1, d1 = -1;
      for (int k = 0; k < n; k++) {
        const int dif = (int)fabsf(T.at(v, u, k)[0] - (float)pre);
        if (dif < 2 && dif < bestPre) { bestPre = dif; d0 = (int)T.at(v, u, k)[0]; }
      }
      if (u < T.W - 1) {
        const int aft = (int)T.at(v, u + 1, 0)[0];
        for (int k = 0; k < n; k++) {
          const int dif = (int)fabsf(T.at(v, u, k)[0] - (float)aft);
          if (dif < 2 && dif < bestAft) { bestAft = dif; d1 = (int)T.at(v, u, k)[0]; }
        }
      }
      int r;
      if (d0 != -1 && d1 == -1) r = d0;
      else if (d0 == -1 && d1 != -1) r = d1;
      else if (d0 == -1 && d1 == -1) r = first;
      else {
        const uint8_t* c = bgr + ((size_t)v * T.W + u) * 3;
        int cpre = 0, caft = 0;
        for (int k = 0; k < 3; k++) { cpre += abs((int)c[k] - (int)c[k - 3]); caft += abs((int)c[k] - (int)c[k + 3]); }
        r = cpre <= caft ? d0 : d1;
      }
      row[u] = (int16_t)r;
    }
  }
}

// ------------------------------------------------------------------ entry points
extern "C" int sm_disp_from_top(sm_ctx* ctx, const float* d_top, int H, int W, int num, int16_t* d_disp) {
  SM_CHECK_ARG(ctx && d_top && d_disp && H > 0 && W > 0 && num >= 1 && num <= SM_TOP_MAX);
  top_view T{d_top, H, W, num};
  dim3 grid(sm_div_up(W, 128), H);
  SM_LAUNCH(ctx, k_top_v1, grid, 128, 0, T, d_disp);
  return SM_OK;
}

extern "C" int sm_disp_from_top2(sm_ctx* ctx, const float* d_top, const uint8_t* d_bgr, int H, int W, int num, int method,
                                 int ts, int hasCir2, int colorLimit, int16_t* d_disp) {
  SM_CHECK_ARG(ctx && d_top && d_bgr && d_disp && H > 0 && W > 0 && num >= 1 && num <= SM_TOP_MAX);
  SM_CHECK_ARG(method >= 0 && method <= 2);
  top_view T{d_top, H, W, num};
  if (method != 0) {
    SM_LAUNCH(ctx, k_top2_rows, sm_div_up(H, 32), 32, 0, T, d_bgr, method, d_disp);
    return SM_OK;
  }
  const size_t npix = (size_t)H * W;
  void* p;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_RVLIST, 256 + (size_t)H * 4 + npix, &p));
  int* ticket = (int*)p;
  int* progress = (int*)((uint8_t*)p + 256);
  uint8_t* pend = (uint8_t*)p + 256 + (size_t)H * 4;
  SM_CUDA(cudaMemsetAsync(p, 0, 256 + (size_t)H * 4, ctx->stream));
  dim3 grid(sm_div_up(W, 128), H);
  SM_LAUNCH(ctx, k_top2_local, grid, 128, 0, T, d_bgr, ts, hasCir2, colorLimit, d_disp, pend);
  const int blocks = min(sm_div_up(H, TOP_RES_WARPS), ctx->num_sms * 4);
  SM_LAUNCH(ctx, k_top2_resolve, blocks, TOP_RES_WARPS * 32, 0, T, pend, d_disp, ticket, progress);
  return SM_OK;
}
