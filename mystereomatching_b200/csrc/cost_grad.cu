// Gradient cost family -- SURVEY.md 8(f) rank 3: `costcalculation = "censusGrad"` is the selector the reference's
// main_.cpp:15 compiles in, so it is the cost the reference's own driver feeds into CBCA.
//
//   calGrad / calGrad_y   stereoMatching.cpp:271-368   central differences on the gray image
//   grad -> calgradvm     stereoMatching.cpp:603-656, 388-455   arm-weighted |dgx| , |dgy| volume
//   censusGrad            stereoMatching.cpp:25-48     2 - exp(-census/lamCen) - exp(-grad/lamG)
//
// One pass writes vol[v][u][d] (float32, d fastest).  The gradient volume and the census volume of the reference
// (two H*W*D temporaries each, stereoMatching.cpp:27-33, 609-617) are never materialised.  Same shape as k_cost
// (cost.cu): persistent CTAs walk (row, segment) items, the partner image's census words and two gradient planes
// for the segment's D-1+SEG positions are staged in shared memory, a warp takes an anchor pixel and its lanes run
// along d, so every warp store is one coalesced 128-byte line.
//
// Parity.  sm_cost_grad (the gradient volume alone): bit-exact -- gradients are multiples of 0.5 (exact floats),
// a = sH / (sH + sV) is one IEEE division, a*dx + (1-a)*dy is two multiplications and one addition, none contracted
// (-fmad=false; the intrinsics below say so explicitly).  sm_cost_censusgrad: the census term comes from a host
// table built with the same libm expf the reference calls; exp(-grad/lamG) takes a continuous argument, so it is
// evaluated on the device (__expf = MUFU.EX2 of arg*log2e) -> the combined volume agrees to < 1e-6 absolute on values
// in [0, 2]; the tests assert that and the north star's 1e-4 relative.
#include <math.h>

#include "common.cuh"

__global__ void k_grad_xy(const uint8_t* __restrict__ gray, int H, int W, float* __restrict__ gx,
                          float* __restrict__ gy) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
  if (u >= W || v >= H) return;
  const size_t p = (size_t)v * W + u;
  // interior: 0.5 * (next - prev) (double product of an int, exact in float); borders: next - prev of the border pair
  int dx, dy;
  float sx = 0.5f, sy = 0.5f;
  if (u == 0) { dx = (int)gray[p + 1] - (int)gray[p]; sx = 1.f; }
  else if (u == W - 1) { dx = (int)gray[p] - (int)gray[p - 1]; sx = 1.f; }
  else dx = (int)gray[p + 1] - (int)gray[p - 1];
  if (v == 0) { dy = (int)gray[p + W] - (int)gray[p]; sy = 1.f; }
  else if (v == H - 1) { dy = (int)gray[p] - (int)gray[p - W]; sy = 1.f; }
  else dy = (int)gray[p + W] - (int)gray[p - W];
  gx[p] = sx * (float)dx;
  gy[p] = sy * (float)dy;
}

extern "C" int sm_grad_xy(sm_ctx* ctx, const uint8_t* d_gray, int H, int W, float* d_gx, float* d_gy) {
  SM_CHECK_ARG(ctx && d_gray && d_gx && d_gy && H >= 2 && W >= 2);   // the reference reads pixel 1 / H-2 unconditionally
  dim3 block(32, 8), grid(sm_div_up(W, 32), sm_div_up(H, 8));
  SM_LAUNCH(ctx, k_grad_xy, grid, block, 0, d_gray, H, W, d_gx, d_gy);
  return SM_OK;
}

#define CG_THREADS 1024
#define CG_SEG 512
#define CG_MAX_CODE 71

// One staged entry per position of the "other" image: {gx, gy} (FUSED 0) or {gx, gy, census word 0 lo, hi} + a second
// array with census word 1 (FUSED 1, 71-bit codes).  eb / hb point at the entry of d = lane of the current 32-chunk;
// the view's sign and the census word count are template parameters, so the following chunks sit at compile-time
// offsets (same addressing scheme as k_cost, cost.cu).
template <int FUSED, int NW>
__device__ __forceinline__ float cost_grad_one(const float4 e, uint32_t hi, float ax, float ay, float wa, float wb,
                                               uint32_t ca0lo, uint32_t ca0hi, uint32_t ca1, float Trunc, float negInvLamG,
                                               bool unitLam, float lamG, const char* __restrict__ t1) {
  const float dx = fminf(fabsf(__fsub_rn(ax, e.x)), Trunc);
  const float dy = fminf(fabsf(__fsub_rn(ay, e.y)), Trunc);
  const float g = __fadd_rn(__fmul_rn(wa, dx), __fmul_rn(wb, dy));
  if (!FUSED) return g;
  const uint32_t a0 = ca0lo ^ __float_as_uint(e.z), a1 = ca0hi ^ __float_as_uint(e.w);
  uint32_t off;
  if (NW == 2) {
    const uint32_t a2 = ca1 ^ hi;
    off = (uint32_t)__popc(a0 ^ a1 ^ a2) * 128u + (uint32_t)__popc((a0 & a1) | (a0 & a2) | (a1 & a2)) * 256u;
  } else {
    off = (uint32_t)(__popc(a0) + __popc(a1)) * 128u;
  }
  // exp(-g / lamG): the reference divides, then negates nothing (-vm1/ARU1); x / 1.0f == x exactly
  const float arg = unitLam ? -g : __fdiv_rn(-g, lamG);
  (void)negInvLamG;
  // expf as the host libm evaluates it (smd_expf_host, common.cuh): the combined volume equals the reference's bit for bit
  return __fsub_rn(*reinterpret_cast<const float*>(t1 + off), smd_expf_host(arg));
}

// FUSED 0: gradient volume (calgradvm).  FUSED 1: censusGrad.
template <int FUSED, int NW, int SGN>
__global__ void __launch_bounds__(CG_THREADS, 1)
    k_cost_grad(const float* __restrict__ gxA, const float* __restrict__ gyA, const float* __restrict__ gxO,
                const float* __restrict__ gyO, const uint16_t* __restrict__ armsA, const uint64_t* __restrict__ cenA,
                const uint64_t* __restrict__ cenO, int H, int W, int D, int codeLen, float Trunc,
                float oorGrad, float lamG, const float* __restrict__ tabCen, float* __restrict__ vol) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int maxEntries = CG_SEG + D - 1;
  float* sT1 = reinterpret_cast<float*>(smem_raw);                       // [72][32]: 2 - exp(-c/lamCen)
  float4* sEnt = reinterpret_cast<float4*>(sT1 + (FUSED ? (CG_MAX_CODE + 1) * 32 : 0));
  uint32_t* sHi = reinterpret_cast<uint32_t*>(sEnt + maxEntries);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float oor = oorGrad;
  const bool unitLam = lamG == 1.0f;
  if (FUSED) {
    for (int i = tid; i < (codeLen + 1) * 32; i += CG_THREADS) sT1[i] = 2.0f - tabCen[i >> 5];
    oor = __fsub_rn(2.0f - tabCen[codeLen], smd_expf_host(__fdiv_rn(-oorGrad, lamG)));
  }
  const char* t1 = reinterpret_cast<const char*>(sT1 + lane);
  const int nSeg = (W + CG_SEG - 1) / CG_SEG;
  const int nItems = H * nSeg;
  const int nd = (D + 31) >> 5;
  for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
    const int v = item / nSeg, ua = (item - v * nSeg) * CG_SEG;
    const int nA = min(CG_SEG, W - ua);
    const int elo = SGN > 0 ? ua - (D - 1) : ua;
    const int cnt = nA + D - 1;
    __syncthreads();
    for (int i = tid; i < cnt; i += CG_THREADS) {
      const int e = elo + i;
      float4 ent = make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t hi = 0;
      if (e >= 0 && e < W) {
        const size_t p = (size_t)v * W + e;
        ent.x = gxO[p]; ent.y = gyO[p];
        if (FUSED) {
          const uint64_t c0 = cenO[p * NW];
          ent.z = __uint_as_float((uint32_t)c0); ent.w = __uint_as_float((uint32_t)(c0 >> 32));
          if (NW == 2) hi = (uint32_t)cenO[p * NW + 1];
        }
      }
      sEnt[i] = ent;
      if (FUSED && NW == 2) sHi[i] = hi;
    }
    __syncthreads();
    // the anchor's own words (gradients, arms, census) are fetched one anchor ahead
    float ax_n = 0.f, ay_n = 0.f;
    uint16_t ar_n[4] = {0, 0, 0, 0};
    uint64_t c0_n = 0;
    uint32_t c1_n = 0;
    auto fetch = [&](size_t q) {
      ax_n = gxA[q]; ay_n = gyA[q];
#pragma unroll
      for (int k = 0; k < 4; k++) ar_n[k] = armsA[q * 5 + k];
      if (FUSED) {
        c0_n = cenA[q * NW];
        if (NW == 2) c1_n = (uint32_t)cenA[q * NW + 1];
      }
    };
    if (warp < nA) fetch((size_t)v * W + ua + warp);
    for (int a = warp; a < nA; a += CG_THREADS / 32) {
      const int u = ua + a;
      const size_t p = (size_t)v * W + u;
      const float ax = ax_n, ay = ay_n;
      // a = shortestH / (shortestH + shortestV) from the view's own arms, read as short (stereoMatching.cpp:403-420)
      float sH = (float)min((int)(short)ar_n[0], (int)(short)ar_n[1]);
      float sV = (float)min((int)(short)ar_n[2], (int)(short)ar_n[3]);
      const uint32_t ca0lo = (uint32_t)c0_n, ca0hi = (uint32_t)(c0_n >> 32), ca1 = c1_n;
      if (a + CG_THREADS / 32 < nA) fetch(p + CG_THREADS / 32);
      if (sH == 0.f) sH = 1.f;
      if (sV == 0.f) sV = 1.f;
      const float wa = __fdiv_rn(sH, __fadd_rn(sH, sV)), wb = __fsub_rn(1.f, wa);
      float* out = vol + p * D + lane;
      const int nvalid = min(D, SGN > 0 ? u + 1 : W - u);
      const int base = u - elo - SGN * lane;
      const float4* eb = sEnt + base;
      const uint32_t* hb = sHi + base;
      int j = 0;
      for (; (j + 4) * 32 <= nvalid; j += 4) {
#pragma unroll
        for (int q = 0; q < 4; q++)
          out[q * 32] = cost_grad_one<FUSED, NW>(eb[-SGN * q * 32], (FUSED && NW == 2) ? hb[-SGN * q * 32] : 0u, ax, ay, wa, wb,
                                                 ca0lo, ca0hi, ca1, Trunc, 0.f, unitLam, lamG, t1);
        eb -= SGN * 128; hb -= SGN * 128; out += 128;
      }
      for (; j < nd; j++) {
        const int d = lane + j * 32;
        if (d >= D) break;
        float r = oor;
        if (d < nvalid)
          r = cost_grad_one<FUSED, NW>(eb[0], (FUSED && NW == 2) ? hb[0] : 0u, ax, ay, wa, wb, ca0lo, ca0hi, ca1, Trunc, 0.f,
                                       unitLam, lamG, t1);
        out[0] = r;
        eb -= SGN * 32; hb -= SGN * 32; out += 32;
      }
    }
  }
}

template <int FUSED, int NW, int SGN>
static int launch_cost_grad3(sm_ctx* ctx, const float* gxA, const float* gyA, const float* gxO, const float* gyO,
                             const uint16_t* armsA, const uint64_t* cenA, const uint64_t* cenO, int H, int W, int D,
                             int codeLen, float Trunc, float lamG, const float* tabCen, float* vol) {
  size_t smem = (size_t)(CG_SEG + D - 1) * (sizeof(float4) + sizeof(uint32_t));
  if (FUSED) smem += (CG_MAX_CODE + 1) * 32 * sizeof(float);
  SM_CUDA(cudaFuncSetAttribute(k_cost_grad<FUSED, NW, SGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nItems = H * sm_div_up(W, CG_SEG);
  const int grid = min(nItems, ctx->num_sms);
  // out of range: sqrt(pow(Trunc, 2) * 2) in double (pow(float, int) promotes), stored to float (stereoMatching.cpp:433)
  const float oorGrad = (float)sqrt(pow((double)Trunc, 2) * 2);
  SM_LAUNCH(ctx, (k_cost_grad<FUSED, NW, SGN>), grid, CG_THREADS, smem, gxA, gyA, gxO, gyO, armsA, cenA, cenO, H, W, D,
            codeLen, Trunc, oorGrad, lamG, tabCen, vol);
  return SM_OK;
}

template <int FUSED>
static int launch_cost_grad(sm_ctx* ctx, const float* gxA, const float* gyA, const float* gxO, const float* gyO,
                            const uint16_t* armsA, const uint64_t* cenA, const uint64_t* cenO, int nw, int H, int W, int D,
                            int sgn, int codeLen, float Trunc, float lamG, const float* tabCen, float* vol) {
  if (nw == 2)
    return sgn > 0 ? launch_cost_grad3<FUSED, 2, +1>(ctx, gxA, gyA, gxO, gyO, armsA, cenA, cenO, H, W, D, codeLen, Trunc, lamG, tabCen, vol)
                   : launch_cost_grad3<FUSED, 2, -1>(ctx, gxA, gyA, gxO, gyO, armsA, cenA, cenO, H, W, D, codeLen, Trunc, lamG, tabCen, vol);
  return sgn > 0 ? launch_cost_grad3<FUSED, 1, +1>(ctx, gxA, gyA, gxO, gyO, armsA, cenA, cenO, H, W, D, codeLen, Trunc, lamG, tabCen, vol)
                 : launch_cost_grad3<FUSED, 1, -1>(ctx, gxA, gyA, gxO, gyO, armsA, cenA, cenO, H, W, D, codeLen, Trunc, lamG, tabCen, vol);
}

extern "C" int sm_cost_grad(sm_ctx* ctx, const float* d_gxL, const float* d_gyL, const float* d_gxR, const float* d_gyR,
                            const uint16_t* d_armsView, int H, int W, int D, float trunc, int LOR, float* d_vol) {
  SM_CHECK_ARG(ctx && d_gxL && d_gyL && d_gxR && d_gyR && d_armsView && d_vol);
  SM_CHECK_ARG(H > 0 && W > 0 && D > 0 && D <= 512 && (LOR == 0 || LOR == 1));
  if (LOR == 0)
    return launch_cost_grad<0>(ctx, d_gxL, d_gyL, d_gxR, d_gyR, d_armsView, nullptr, nullptr, 1, H, W, D, +1, 0, trunc, 1.f,
                               nullptr, d_vol);
  return launch_cost_grad<0>(ctx, d_gxR, d_gyR, d_gxL, d_gyL, d_armsView, nullptr, nullptr, 1, H, W, D, -1, 0, trunc, 1.f,
                             nullptr, d_vol);
}

extern "C" int sm_cost_censusgrad(sm_ctx* ctx, const uint64_t* d_cenL, const uint64_t* d_cenR, const float* d_gxL,
                                  const float* d_gyL, const float* d_gxR, const float* d_gyR, const uint16_t* d_armsView,
                                  int H, int W, int D, int func, float lamCen, float lamG, float gradTrunc, int LOR,
                                  float* d_vol) {
  SM_CHECK_ARG(ctx && d_cenL && d_cenR && d_gxL && d_gyL && d_gxR && d_gyR && d_armsView && d_vol);
  SM_CHECK_ARG(H > 0 && W > 0 && D > 0 && D <= 512 && (LOR == 0 || LOR == 1) && (func == 0 || func == 3));
  SM_CHECK_ARG(lamCen > 0.f && lamG > 0.f);
  const int codeLen = sm_census_code_length(func), nw = sm_census_words(func);
  const float *tAD, *tCen;
  // only the census table is used; the AD half of the cached pair keeps whatever the pipeline last asked for
  SM_TRY(smi_exp_tables(ctx, ctx->tab_trunc >= 0.f ? ctx->tab_trunc : 1000.f, ctx->tab_lamAD > 0.f ? ctx->tab_lamAD : 10.f,
                        lamCen, codeLen, &tAD, &tCen));
  if (LOR == 0)
    return launch_cost_grad<1>(ctx, d_gxL, d_gyL, d_gxR, d_gyR, d_armsView, d_cenL, d_cenR, nw, H, W, D, +1, codeLen,
                               gradTrunc, lamG, tCen, d_vol);
  return launch_cost_grad<1>(ctx, d_gxR, d_gyR, d_gxL, d_gyL, d_armsView, d_cenR, d_cenL, nw, H, W, D, -1, codeLen,
                             gradTrunc, lamG, tCen, d_vol);
}
