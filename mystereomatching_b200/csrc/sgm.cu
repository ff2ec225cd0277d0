// K5: semi-global matching, one kernel launch per path.
//
//   sgm         stereoMatching.cpp:6204-6224   direction table rv/ru = offset TO THE PREDECESSOR
//   costScan    stereoMatching.cpp:1983-2029   raster order, reversed when rv>0 || (rv==0 && ru>0)
//   updateCost  stereoMatching.h:2205-2280     literal P1=1, P2=3 (/= reduCoeffi1 when the current-view
//                                              colour step D1 > corDifThres); P1 -= minC;
//                                              Lr = C + min4(Lr'[d]-minC, Lr'[d-1]+P1, Lr'[d+1]+P1, P2);
//                                              first pixel of a path: Lr = C
//   gen_sgm_vm  stereoMatching.cpp:2031-2056   vm = ((0+L0)+L1)+... (no averaging)
//
// The recurrence along a path is strictly serial, so the unit of parallelism is
// the scan line: ONE WARP PER SCAN LINE (row, column or diagonal).  The D
// disparities are spread over the 32 lanes as contiguous runs of VPL values; the
// previous pixel's Lr row lives in registers, the d+-1 neighbours at run edges
// come from two shuffles, and the D-wide minimum is one redux.sync on an
// order-preserving integer image of the floats.  Every path sweep touches HBM
// once: C is read with 16-byte loads PF pixels ahead of the recurrence (register
// prefetch ring, so >= PF KB per warp are in flight and the ~1 us HBM latency is
// off the critical path), and the path's Lr is either written (mode 0) or added
// to the running sum S (mode 1: read-modify-write, in the reference's path
// order so the float sum is bit-identical).
#include <float.h>

#include "common.cuh"

#define SGM_WARPS 2

__device__ __forceinline__ uint32_t f2key(float x) {
  uint32_t b = __float_as_uint(x);
  return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);  // monotone: a<b <=> key(a)<key(b)
}
__device__ __forceinline__ float key2f(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
  return __uint_as_float(b);
}

template <int VPL, bool VEC>
__device__ __forceinline__ void load_run(const float* __restrict__ p, int d0, int D, float (&r)[VPL]) {
  if (VEC) {
#pragma unroll
    for (int k = 0; k < VPL; k += 4) {  // D % 4 == 0: a float4 is entirely inside or outside [0,D)
      float4 t = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
      if (d0 + k < D) t = *reinterpret_cast<const float4*>(p + d0 + k);
      r[k] = t.x; r[k + 1] = t.y; r[k + 2] = t.z; r[k + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < VPL; k++) r[k] = (d0 + k < D) ? p[d0 + k] : FLT_MAX;
  }
}

template <int VPL, bool VEC>
__device__ __forceinline__ void store_run(float* __restrict__ p, int d0, int D, const float (&r)[VPL]) {
  if (VEC) {
#pragma unroll
    for (int k = 0; k < VPL; k += 4)
      if (d0 + k < D) *reinterpret_cast<float4*>(p + d0 + k) = make_float4(r[k], r[k + 1], r[k + 2], r[k + 3]);
  } else {
#pragma unroll
    for (int k = 0; k < VPL; k++)
      if (d0 + k < D) p[d0 + k] = r[k];
  }
}

// Scan-line geometry.  (mv,mu) = direction of travel = -(rv,ru).
struct sgm_geom {
  int H, W, mv, mu, nLines;
};

__device__ __forceinline__ void line_start(const sgm_geom& g, int k, int& v, int& u, int& len) {
  if (g.mv == 0) {  // horizontal: line = row
    v = k; u = g.mu > 0 ? 0 : g.W - 1; len = g.W;
  } else if (g.mu == 0) {  // vertical: line = column
    u = k; v = g.mv > 0 ? 0 : g.H - 1; len = g.H;
  } else {  // diagonal: W lines start on the first row, H-1 more on the entry column
    if (k < g.W) {
      u = k; v = g.mv > 0 ? 0 : g.H - 1;
    } else {
      int j = k - g.W + 1;
      v = g.mv > 0 ? j : g.H - 1 - j;
      u = g.mu > 0 ? 0 : g.W - 1;
    }
    int lv = g.mv > 0 ? g.H - v : v + 1, lu = g.mu > 0 ? g.W - u : u + 1;
    len = min(lv, lu);
  }
}

template <int VPL, int PF, bool VEC, int MODE>
__global__ void __launch_bounds__(SGM_WARPS * 32)
    k_sgm_path(const float* __restrict__ vol, const uint32_t* __restrict__ pix, float* __restrict__ out, sgm_geom g,
               int D, int corDifThres, float redu) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * SGM_WARPS + (threadIdx.x >> 5);
  if (k >= g.nLines) return;
  int v, u, len;
  line_start(g, k, v, u, len);
  const int d0 = lane * VPL;
  const long long pstep = (long long)g.mv * g.W + g.mu;
  long long p = (long long)v * g.W + u;

  float cpf[PF][VPL], spf[MODE == 1 ? PF : 1][VPL];
  uint32_t xpf[PF];
#pragma unroll
  for (int i = 0; i < PF; i++) {
    if (i < len) {
      const long long q = p + pstep * i;
      load_run<VPL, VEC>(vol + q * D, d0, D, cpf[i]);
      if (MODE == 1) load_run<VPL, VEC>(out + q * D, d0, D, spf[i]);
      xpf[i] = pix[q];
    }
  }

  float prev[VPL];
  float minC = 0.f;
  uint32_t xprev = 0;
  for (int t0 = 0; t0 < len; t0 += PF) {
#pragma unroll
    for (int i = 0; i < PF; i++) {
      const int t = t0 + i;
      if (t < len) {
        float c[VPL], s[VPL], lr[VPL];
#pragma unroll
        for (int j = 0; j < VPL; j++) { c[j] = cpf[i][j]; if (MODE == 1) s[j] = spf[i][j]; }
        const uint32_t x = xpf[i];
        if (t + PF < len) {  // refill this prefetch slot for pixel t+PF
          const long long q = p + pstep * PF;
          load_run<VPL, VEC>(vol + q * D, d0, D, cpf[i]);
          if (MODE == 1) load_run<VPL, VEC>(out + q * D, d0, D, spf[i]);
          xpf[i] = pix[q];
        }
        if (t == 0) {
#pragma unroll
          for (int j = 0; j < VPL; j++) lr[j] = c[j];
        } else {
          const int D1 = smd_absdiff_max3(x, xprev);
          float P1 = 1.0f, P2 = 3.0f;
          if (D1 > corDifThres) { P1 = P1 / redu; P2 = P2 / redu; }
          P1 = P1 - minC;
          float lo = __shfl_up_sync(0xffffffffu, prev[VPL - 1], 1);   // Lr'[d0-1]
          float hi = __shfl_down_sync(0xffffffffu, prev[0], 1);       // Lr'[d0+VPL]
#pragma unroll
          for (int j = 0; j < VPL; j++) {
            const int d = d0 + j;
            const float pm = j == 0 ? lo : prev[j - 1];
            const float pp = j == VPL - 1 ? hi : prev[j + 1];
            const float S1 = prev[j] - minC;
            const float S2 = d - 1 >= 0 ? pm + P1 : FLT_MAX;
            const float S3 = d + 1 < D ? pp + P1 : FLT_MAX;
            lr[j] = c[j] + fminf(fminf(S1, S2), fminf(S3, P2));
          }
        }
        // D-wide minimum of the new row (padding lanes hold FLT_MAX via c[])
        float m = FLT_MAX;
#pragma unroll
        for (int j = 0; j < VPL; j++) m = (d0 + j < D) ? fminf(m, lr[j]) : m;
        minC = key2f(__reduce_min_sync(0xffffffffu, f2key(m)));
#pragma unroll
        for (int j = 0; j < VPL; j++) prev[j] = lr[j];
        xprev = x;
        if (MODE == 1) {
#pragma unroll
          for (int j = 0; j < VPL; j++) lr[j] = s[j] + lr[j];  // gen_sgm_vm: sum += L[num]
        }
        store_run<VPL, VEC>(out + p * D, d0, D, lr);
        p += pstep;
      }
    }
  }
}

static const int SGM_RV[8] = {+1, -1, 0, 0, +1, +1, -1, -1};
static const int SGM_RU[8] = {0, 0, +1, -1, -1, +1, +1, -1};

template <int VPL, int PF, bool VEC>
static int launch_sgm(sm_ctx* ctx, const float* vol, const uint32_t* pix, float* out, const sgm_geom& g, int D,
                      int thr, float redu, int mode) {
  int grid = sm_div_up(g.nLines, SGM_WARPS);
  if (mode == 0)
    SM_LAUNCH(ctx, (k_sgm_path<VPL, PF, VEC, 0>), grid, SGM_WARPS * 32, 0, vol, pix, out, g, D, thr, redu);
  else
    SM_LAUNCH(ctx, (k_sgm_path<VPL, PF, VEC, 1>), grid, SGM_WARPS * 32, 0, vol, pix, out, g, D, thr, redu);
  return SM_OK;
}

int smi_sgm_path_packed(sm_ctx* ctx, const float* d_vol, const uint32_t* d_pix, int H, int W, int D, int path,
                        int corDifThres, int reduCoeffi1, int mode, float* d_out) {
  sgm_geom g;
  g.H = H; g.W = W; g.mv = -SGM_RV[path]; g.mu = -SGM_RU[path];
  g.nLines = g.mv == 0 ? H : (g.mu == 0 ? W : W + H - 1);
  const float redu = (float)reduCoeffi1;
  const bool vec = (D % 4 == 0) && (((uintptr_t)d_vol | (uintptr_t)d_out) % 16 == 0);
  const int vpl = D <= 32 ? 1 : D <= 64 ? 2 : D <= 128 ? 4 : D <= 256 ? 8 : 16;
  // 16-byte loads need D % 4 == 0 (then every float4 of a run is wholly inside or outside [0,D))
  switch (vpl) {
    case 1: return launch_sgm<1, 8, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode);
    case 2: return launch_sgm<2, 8, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode);
    case 4:
      if (vec) return launch_sgm<4, 8, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode);
      return launch_sgm<4, 8, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode);
    case 8:
      if (vec) return launch_sgm<8, 4, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode);
      return launch_sgm<8, 4, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode);
    default:
      if (vec) return launch_sgm<16, 2, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode);
      return launch_sgm<16, 2, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode);
  }
}

extern "C" int sm_sgm_path(sm_ctx* ctx, const float* d_vol, const uint8_t* d_bgr, int H, int W, int D, int path,
                           int corDifThres, int reduCoeffi1, int mode, float* d_out) {
  SM_CHECK_ARG(ctx && d_vol && d_bgr && d_out);
  SM_CHECK_ARG(H > 0 && W > 0 && D > 0 && D <= 512);
  SM_CHECK_ARG(path >= 0 && path < 8 && (mode == 0 || mode == 1) && reduCoeffi1 != 0);
  SM_CHECK_ARG(d_vol != d_out);
  const long long npix = (long long)H * W;
  void* pk;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_IMG0, npix * 4, &pk));
  SM_TRY(smi_pack_bgr(ctx, d_bgr, npix, (uint32_t*)pk));
  return smi_sgm_path_packed(ctx, d_vol, (const uint32_t*)pk, H, W, D, path, corDifThres, reduCoeffi1, mode, d_out);
}

extern "C" int sm_sgm(sm_ctx* ctx, const float* d_vol, const uint8_t* d_bgr, int H, int W, int D, int paths,
                      int corDifThres, int reduCoeffi1, float* d_sum) {
  SM_CHECK_ARG(ctx && d_vol && d_bgr && d_sum);
  SM_CHECK_ARG(H > 0 && W > 0 && D > 0 && D <= 512);
  SM_CHECK_ARG(paths >= 1 && paths <= 8 && reduCoeffi1 != 0 && d_vol != d_sum);
  const long long npix = (long long)H * W;
  void* pk;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_IMG0, npix * 4, &pk));
  SM_TRY(smi_pack_bgr(ctx, d_bgr, npix, (uint32_t*)pk));
  for (int i = 0; i < paths; i++)
    SM_TRY(smi_sgm_path_packed(ctx, d_vol, (const uint32_t*)pk, H, W, D, i, corDifThres, reduCoeffi1, i == 0 ? 0 : 1,
                               d_sum));
  return SM_OK;
}
