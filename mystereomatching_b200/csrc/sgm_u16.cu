// K5c: SGM on 16-bit integer cost volumes -- the native form of costScan's integer entry
// (stereoMatching.cpp:2007-2014: vm.depth() CV_8U / CV_16U -> updateCost<uchar | ushort>, stereoMatching.h:2205-2280).
//
// With integer costs every quantity of updateCost is a multiple of 1 / reduCoeffi1 (P1 = 1 or 1 / reduCoeffi1, P2 = 3
// or 3 / reduCoeffi1), so for a power-of-two reduCoeffi1 the reference's float arithmetic is EXACT, and
//       Lr * reduCoeffi1   is an integer  <= (Cmax + 3) * reduCoeffi1,
// and the P-path sum times reduCoeffi1 fits 16 bits whenever P * (Cmax + 3) * reduCoeffi1 <= 65535 (Hamming costs:
// Cmax = 71, 8 paths, reduCoeffi1 = 4 -> 2368).  The kernels below keep Lr and the path sum in that fixed point:
// the volume is read as uint16 (2 B per disparity instead of 4), the sum is written as uint16 = reduCoeffi1 x the
// reference's float sum, exactly, and gen_dispFromVm on it (first minimum) picks the same disparity.  Every volume
// pass moves half the bytes of the float path (b = 2).
//
// Shape as k_sgm_path (sgm.cu): one warp per scan line, D spread over the lanes as runs of VPL values, previous Lr row
// in registers, d +- 1 neighbours by two shuffles, D-wide minimum by one redux on the integers, C (and S) prefetched
// PF pixels ahead with vector loads.
#include "sgm_common.cuh"

#define SGMU_WARPS 2
#define SGMU_BIG 0x3fff0000   // "outside [0, D)": never the minimum, never overflows when P1 is added

template <int VPL, bool VEC>
__device__ __forceinline__ void u16_load_run(const uint16_t* __restrict__ p, int d0, int D, int (&r)[VPL]) {
  if (VEC && VPL >= 2) {
    if (d0 < D) {   // D % VPL == 0: a run is entirely inside or outside [0, D)
      if (VPL == 2) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(p + d0);
        r[0] = w & 0xffff; r[1 % VPL] = w >> 16;
      } else if (VPL == 4) {
        const uint2 w = *reinterpret_cast<const uint2*>(p + d0);
        r[0] = w.x & 0xffff; r[1 % VPL] = w.x >> 16; r[2 % VPL] = w.y & 0xffff; r[3 % VPL] = w.y >> 16;
      } else {
#pragma unroll
        for (int k = 0; k < VPL; k += 8) {
          const uint4 w = *reinterpret_cast<const uint4*>(p + d0 + k);
          r[k] = w.x & 0xffff; r[(k + 1) % VPL] = w.x >> 16; r[(k + 2) % VPL] = w.y & 0xffff; r[(k + 3) % VPL] = w.y >> 16;
          r[(k + 4) % VPL] = w.z & 0xffff; r[(k + 5) % VPL] = w.z >> 16; r[(k + 6) % VPL] = w.w & 0xffff; r[(k + 7) % VPL] = w.w >> 16;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < VPL; k++) r[k] = SGMU_BIG;
    }
  } else {
#pragma unroll
    for (int k = 0; k < VPL; k++) r[k] = (d0 + k < D) ? (int)p[d0 + k] : SGMU_BIG;
  }
}

template <int VPL, bool VEC>
__device__ __forceinline__ void u16_store_run(uint16_t* __restrict__ p, int d0, int D, const int (&r)[VPL]) {
  if (VEC && VPL >= 2) {
    if (d0 < D) {
      if (VPL == 2) {
        *reinterpret_cast<uint32_t*>(p + d0) = (uint32_t)r[0] | ((uint32_t)r[1 % VPL] << 16);
      } else if (VPL == 4) {
        *reinterpret_cast<uint2*>(p + d0) = make_uint2((uint32_t)r[0] | ((uint32_t)r[1 % VPL] << 16), (uint32_t)r[2 % VPL] | ((uint32_t)r[3 % VPL] << 16));
      } else {
#pragma unroll
        for (int k = 0; k < VPL; k += 8)
          *reinterpret_cast<uint4*>(p + d0 + k) =
              make_uint4((uint32_t)r[k] | ((uint32_t)r[(k + 1) % VPL] << 16), (uint32_t)r[(k + 2) % VPL] | ((uint32_t)r[(k + 3) % VPL] << 16),
                         (uint32_t)r[(k + 4) % VPL] | ((uint32_t)r[(k + 5) % VPL] << 16), (uint32_t)r[(k + 6) % VPL] | ((uint32_t)r[(k + 7) % VPL] << 16));
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < VPL; k++)
      if (d0 + k < D) p[d0 + k] = (uint16_t)r[k];
  }
}

// MODE 0: out = Lr.  1: out += Lr.  2: out += Lr and d_disp = gen_dispFromVm of the finished sum.  3: that WTA alone.
// FIRST (MODE 0 only): the volume holds raw costs, which are scaled here; the sum volume is already in fixed point.
template <int VPL, int PF, bool VEC, int MODE>
__global__ void __launch_bounds__(SGMU_WARPS * 32)
    k_sgm_path_u16(const uint16_t* __restrict__ vol, const uint32_t* __restrict__ pix, uint16_t* __restrict__ out, sgm_geom g,
                   int D, int corDifThres, int scale, int16_t* __restrict__ disp) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * SGMU_WARPS + (threadIdx.x >> 5);
  if (k >= g.nLines) return;
  int v, u, len;
  line_start(g, k, v, u, len);
  const int d0 = lane * VPL;
  const long long pstep = (long long)g.mv * g.W + g.mu;
  long long p = (long long)v * g.W + u;

  int cpf[PF][VPL], spf[MODE >= 1 ? PF : 1][VPL];
  uint32_t xpf[PF];
#pragma unroll
  for (int i = 0; i < PF; i++) {
    if (i < len) {
      const long long q = p + pstep * i;
      u16_load_run<VPL, VEC>(vol + q * D, d0, D, cpf[i]);
      if (MODE >= 1) u16_load_run<VPL, VEC>(out + q * D, d0, D, spf[i]);
      xpf[i] = pix[q];
    }
  }
  int prev[VPL];
  int minC = 0;
  uint32_t xprev = 0;
  for (int t0 = 0; t0 < len; t0 += PF) {
#pragma unroll
    for (int i = 0; i < PF; i++) {
      const int t = t0 + i;
      if (t < len) {
        int c[VPL], s[VPL], lr[VPL];
#pragma unroll
        for (int j = 0; j < VPL; j++) {
          c[j] = d0 + j < D ? cpf[i][j] * scale : SGMU_BIG;       // raw cost -> fixed point
          s[j] = MODE >= 1 ? spf[MODE >= 1 ? i : 0][j] : 0;
        }
        const uint32_t x = xpf[i];
        if (t + PF < len) {
          const long long q = p + pstep * PF;
          u16_load_run<VPL, VEC>(vol + q * D, d0, D, cpf[i]);
          if (MODE >= 1) u16_load_run<VPL, VEC>(out + q * D, d0, D, spf[MODE >= 1 ? i : 0]);
          xpf[i] = pix[q];
        }
        if (t == 0) {
#pragma unroll
          for (int j = 0; j < VPL; j++) lr[j] = c[j];
        } else {
          const bool step = (int)smd_absdiff_max3(x, xprev) > corDifThres;
          const int P2 = step ? 3 : 3 * scale;                     // 3 / reduCoeffi1 resp. 3, times reduCoeffi1
          const int P1 = (step ? 1 : scale) - minC;                // P1 -= minC
          int lo = __shfl_up_sync(0xffffffffu, prev[VPL - 1], 1);
          int hi = __shfl_down_sync(0xffffffffu, prev[0], 1);
          if (lane == 0) lo = SGMU_BIG;
          if (lane == 31) hi = SGMU_BIG;
#pragma unroll
          for (int j = 0; j < VPL; j++) {
            const int pm = j == 0 ? lo : prev[j - 1];
            const int pp = j == VPL - 1 ? hi : prev[j + 1];
            lr[j] = c[j] + min(min(prev[j] - minC, pm + P1), min(pp + P1, P2));
          }
        }
        int m = SGMU_BIG;
#pragma unroll
        for (int j = 0; j < VPL; j++) m = min(m, lr[j]);
        minC = __reduce_min_sync(0xffffffffu, m);
#pragma unroll
        for (int j = 0; j < VPL; j++) prev[j] = lr[j];
        xprev = x;
#pragma unroll
        for (int j = 0; j < VPL; j++) s[j] = MODE >= 1 ? s[j] + lr[j] : lr[j];
        if (MODE >= 2) {
          // gen_dispFromVm: strict '>' in increasing d -> the lowest d among the minima
          int bm = 0x7fffffff, bd = 0x7fffffff;
#pragma unroll
          for (int j = 0; j < VPL; j++)
            if (d0 + j < D && bm > s[j]) { bm = s[j]; bd = d0 + j; }
          const int gm = __reduce_min_sync(0xffffffffu, bm);
          const int cand = bm == gm ? bd : 0x7fffffff;
          const int best = __reduce_min_sync(0xffffffffu, cand);
          if (lane == 0) disp[p] = (int16_t)(best == 0x7fffffff ? -1 : best);
        }
        if (MODE != 3) u16_store_run<VPL, VEC>(out + p * D, d0, D, s);
        p += pstep;
      }
    }
  }
}

template <int VPL, int PF, bool VEC>
static int launch_u16(sm_ctx* ctx, const uint16_t* vol, const uint32_t* pix, uint16_t* out, const sgm_geom& g, int D, int thr,
                      int scale, int mode, int16_t* disp) {
  const int grid = sm_div_up(g.nLines, SGMU_WARPS);
  if (mode == 0) SM_LAUNCH(ctx, (k_sgm_path_u16<VPL, PF, VEC, 0>), grid, SGMU_WARPS * 32, 0, vol, pix, out, g, D, thr, scale, disp);
  else if (mode == 1) SM_LAUNCH(ctx, (k_sgm_path_u16<VPL, PF, VEC, 1>), grid, SGMU_WARPS * 32, 0, vol, pix, out, g, D, thr, scale, disp);
  else if (mode == 2) SM_LAUNCH(ctx, (k_sgm_path_u16<VPL, PF, VEC, 2>), grid, SGMU_WARPS * 32, 0, vol, pix, out, g, D, thr, scale, disp);
  else SM_LAUNCH(ctx, (k_sgm_path_u16<VPL, PF, VEC, 3>), grid, SGMU_WARPS * 32, 0, vol, pix, out, g, D, thr, scale, disp);
  return SM_OK;
}

int smi_sgm_path_u16(sm_ctx* ctx, const uint16_t* d_vol, const uint32_t* d_pix, int H, int W, int D, int path, int corDifThres,
                     int scale, int mode, uint16_t* d_out, int16_t* d_disp) {
  sgm_geom g;
  g.H = H; g.W = W; g.mv = -SGM_RV[path]; g.mu = -SGM_RU[path];
  g.nLines = g.mv == 0 ? H : (g.mu == 0 ? W : W + H - 1);
  const int per = sm_div_up(D, 32);
  const int vpl = per <= 1 ? 1 : (per <= 2 ? 2 : (per <= 4 ? 4 : (per <= 8 ? 8 : 16)));
  const bool vec = vpl >= 2 && D % vpl == 0 && ((((uintptr_t)d_vol | (uintptr_t)d_out) & 15) == 0);
  switch (vpl) {
    case 1: return launch_u16<1, 8, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp);
    case 2: return vec ? launch_u16<2, 8, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp)
                       : launch_u16<2, 8, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp);
    case 4: return vec ? launch_u16<4, 8, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp)
                       : launch_u16<4, 8, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp);
    case 8: return vec ? launch_u16<8, 8, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp)
                       : launch_u16<8, 4, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp);
    default: return vec ? launch_u16<16, 4, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp)
                        : launch_u16<16, 2, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp);
  }
}

// The fixed point holds iff reduCoeffi1 is a power of two and paths * (maxCost + 3) * reduCoeffi1 fits 16 bits.
bool smi_sgm_u16_ok(int D, int paths, int reduCoeffi1, int maxCost) {
  if (D < 1 || D > 512 || paths < 1 || paths > 8 || reduCoeffi1 < 1 || reduCoeffi1 > 64) return false;
  if (reduCoeffi1 & (reduCoeffi1 - 1)) return false;
  return (long long)paths * (maxCost + 3) * reduCoeffi1 <= 65535;
}

int smi_sgm_u16(sm_ctx* ctx, const uint16_t* d_vol, const uint32_t* d_pix, int H, int W, int D, int paths, int corDifThres,
                int reduCoeffi1, uint16_t* d_sum, int16_t* d_disp, bool keep_sum) {
  for (int i = 0; i < paths; i++) {
    const int mode = i == 0 ? 0 : (i == paths - 1 && d_disp ? (keep_sum ? 2 : 3) : 1);
    SM_TRY(smi_sgm_path_u16(ctx, d_vol, d_pix, H, W, D, i, corDifThres, reduCoeffi1, mode, d_sum, d_disp));
  }
  if (paths == 1 && d_disp) SM_TRY(sm_wta_u16(ctx, d_sum, H, W, D, d_disp));
  return SM_OK;
}

// gen_dispFromVm on a uint16 volume (first minimum; a uint16 volume has no FLT_MAX entries, so never -1)
__global__ void k_wta_u16(const uint16_t* __restrict__ vol, long long npix, int D, int16_t* __restrict__ disp) {
  const int lane = threadIdx.x & 31;
  const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = wid; p < npix; p += nw) {
    const uint16_t* c = vol + p * D;
    unsigned best = 0xffffffffu;   // (cost << 16) | d : the minimum is the lowest cost, then the lowest d
    for (int d = lane; d < D; d += 32) best = min(best, ((unsigned)c[d] << 16) | (unsigned)d);
    best = __reduce_min_sync(0xffffffffu, best);
    if (lane == 0) disp[p] = (int16_t)(best & 0xffff);
  }
}

extern "C" int sm_wta_u16(sm_ctx* ctx, const uint16_t* d_vol, int H, int W, int D, int16_t* d_disp) {
  SM_CHECK_ARG(ctx && d_vol && d_disp && H > 0 && W > 0 && D > 0 && D <= 65535);
  const long long npix = (long long)H * W;
  const int grid = (int)min((long long)ctx->num_sms * 16, (npix + 7) / 8);
  SM_LAUNCH(ctx, k_wta_u16, grid, 256, 0, d_vol, npix, D, d_disp);
  return SM_OK;
}

extern "C" int sm_sgm_u16(sm_ctx* ctx, const uint16_t* d_vol, const uint8_t* d_bgr, int H, int W, int D, int paths,
                          int corDifThres, int reduCoeffi1, int maxCost, uint16_t* d_sum, int16_t* d_disp) {
  SM_CHECK_ARG(ctx && d_vol && d_bgr && d_sum && H > 0 && W > 0 && (const void*)d_vol != (const void*)d_sum);
  if (!smi_sgm_u16_ok(D, paths, reduCoeffi1, maxCost)) {
    sm_set_error("sm_sgm_u16: needs a power-of-two reduCoeffi1 and paths * (maxCost + 3) * reduCoeffi1 <= 65535 "
                 "(got D %d, paths %d, reduCoeffi1 %d, maxCost %d); use sm_vol_to_f32 + sm_sgm", D, paths, reduCoeffi1, maxCost);
    return SM_ERR_UNSUPPORTED;
  }
  const long long npix = (long long)H * W;
  void* pk;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_IMG0, npix * 4, &pk));
  SM_TRY(smi_pack_bgr(ctx, d_bgr, npix, (uint32_t*)pk));
  return smi_sgm_u16(ctx, d_vol, (const uint32_t*)pk, H, W, D, paths, corDifThres, reduCoeffi1, d_sum, d_disp, true);
}
