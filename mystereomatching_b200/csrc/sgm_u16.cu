// K5c: SGM on 16-bit integer cost volumes -- the native form of costScan's integer entry
// (stereoMatching.cpp:2007-2014: vm.depth() CV_8U / CV_16U -> updateCost<uchar | ushort>, stereoMatching.h:2205-2280).
//
// With integer costs every quantity of updateCost is a multiple of 1 / reduCoeffi1 (P1 = 1 or 1 / reduCoeffi1, P2 = 3
// or 3 / reduCoeffi1), so for a power-of-two reduCoeffi1 the reference's float arithmetic is EXACT, and
//       Lr * reduCoeffi1   is an integer  <= (Cmax + 3) * reduCoeffi1,
// and the P-path sum times reduCoeffi1 fits 16 bits with room for an "outside" sentinel whenever
// P * (Cmax + 3) * reduCoeffi1 <= 16000 (Hamming costs: Cmax = 71, 8 paths, reduCoeffi1 = 4 -> 2368).  The kernels below keep Lr and the path sum in that fixed point:
// the volume is read as uint16 (2 B per disparity instead of 4), the sum is written as uint16 = reduCoeffi1 x the
// reference's float sum, exactly, and gen_dispFromVm on it (first minimum) picks the same disparity.  Every volume
// pass moves half the bytes of the float path (b = 2).
//
// Shape as k_sgm_path (sgm.cu): one warp per scan line, D spread over the lanes as runs of VPL values, previous Lr row
// in registers, d +- 1 neighbours by two shuffles, D-wide minimum by one redux on the integers, C (and S) prefetched
// PF pixels ahead with vector loads.
#include "sgm_common.cuh"

#define SGMU_WARPS 2
#define SGMU_BIG 0x3fffu          // "outside [0, D)": above every real value (<= 16000), survives + P1 without wrapping
#define SGMU_BIG2 0x3fff3fffu
#define SGMU_MAXSUM 16000         // bound on paths * (maxCost + 3) * reduCoeffi1 (see smi_sgm_u16_ok)

// A run of VPL uint16 values of one lane as NW = VPL / 2 packed words (low half = lower disparity).  All arithmetic below
// is the packed 16-bit integer SIMD of sm_100a (VIADD.16x2, VIMNMX.U16x2, VIADDMNMX.U16x2): two disparities per instruction.
template <int NW, bool VEC>
__device__ __forceinline__ void u16_load_run(const uint16_t* __restrict__ p, int d0, int D, uint32_t (&r)[NW]) {
  if (VEC) {
    if (d0 < D) {   // D % VPL == 0: a run is entirely inside or outside [0, D)
      if (NW == 1) r[0] = *reinterpret_cast<const uint32_t*>(p + d0);
      else if (NW == 2) { const uint2 w = *reinterpret_cast<const uint2*>(p + d0); r[0] = w.x; r[1 % NW] = w.y; }
      else {
#pragma unroll
        for (int k = 0; k < NW; k += 4) {
          const uint4 w = *reinterpret_cast<const uint4*>(p + d0 + 2 * k);
          r[k] = w.x; r[(k + 1) % NW] = w.y; r[(k + 2) % NW] = w.z; r[(k + 3) % NW] = w.w;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < NW; k++) r[k] = SGMU_BIG2;
    }
  } else {
#pragma unroll
    for (int k = 0; k < NW; k++) {
      const uint32_t lo = d0 + 2 * k < D ? p[d0 + 2 * k] : SGMU_BIG, hi = d0 + 2 * k + 1 < D ? p[d0 + 2 * k + 1] : SGMU_BIG;
      r[k] = lo | (hi << 16);
    }
  }
}

template <int NW, bool VEC>
__device__ __forceinline__ void u16_store_run(uint16_t* __restrict__ p, int d0, int D, const uint32_t (&r)[NW]) {
  if (VEC) {
    if (d0 < D) {
      if (NW == 1) *reinterpret_cast<uint32_t*>(p + d0) = r[0];
      else if (NW == 2) *reinterpret_cast<uint2*>(p + d0) = make_uint2(r[0], r[1 % NW]);
      else {
#pragma unroll
        for (int k = 0; k < NW; k += 4)
          *reinterpret_cast<uint4*>(p + d0 + 2 * k) = make_uint4(r[k], r[(k + 1) % NW], r[(k + 2) % NW], r[(k + 3) % NW]);
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < NW; k++) {
      if (d0 + 2 * k < D) p[d0 + 2 * k] = (uint16_t)(r[k] & 0xffffu);
      if (d0 + 2 * k + 1 < D) p[d0 + 2 * k + 1] = (uint16_t)(r[k] >> 16);
    }
  }
}

// One pixel of the recurrence, in fixed point: with a = Lr' - minC (>= 0),
//   Lr = C + min(a[d], a[d-1] + P1, a[d+1] + P1, P2)       (updateCost<T>, stereoMatching.h:2205-2280)
// c: scaled costs (padding halves = BIG), s: the path sum so far (MODE >= 1), prev / minC / xprev: state of the path.
template <int NW, int MODE>
__device__ __forceinline__ void u16_step(bool first, const uint32_t (&c)[NW], uint32_t (&s)[NW], uint32_t (&prev)[NW],
                                         uint32_t& minC, uint32_t x, uint32_t& xprev, int d0, int D, int corDifThres,
                                         uint32_t scale, int lane, int16_t* __restrict__ disp, long long p) {
  uint32_t lr[NW];
  if (first) {
#pragma unroll
    for (int w = 0; w < NW; w++) lr[w] = c[w];
  } else {
    const bool step = (int)smd_absdiff_max3(x, xprev) > corDifThres;
    const uint32_t P1 = (step ? 1u : scale) * 0x10001u, P2 = (step ? 3u : 3u * scale) * 0x10001u;
    const uint32_t negm = ((0x10000u - minC) & 0xffffu) * 0x10001u;      // a = prev - minC, modulo 2^16 per half
    uint32_t a[NW];
#pragma unroll
    for (int w = 0; w < NW; w++) a[w] = __vadd2(prev[w], negm);
    uint32_t lo = __shfl_up_sync(0xffffffffu, a[NW - 1], 1);   // a'[d0-1] in its high half
    uint32_t hi = __shfl_down_sync(0xffffffffu, a[0], 1);      // a'[d0+VPL] in its low half
    if (lane == 0) lo = SGMU_BIG2;
    if (lane == 31) hi = SGMU_BIG2;
#pragma unroll
    for (int w = 0; w < NW; w++) {
      const uint32_t pm = __byte_perm(w == 0 ? lo : a[w - 1], a[w], 0x5432);        // {a[d-1] of both halves}
      const uint32_t pp = __byte_perm(a[w], w == NW - 1 ? hi : a[w + 1], 0x5432);   // {a[d+1] of both halves}
      const uint32_t t1 = __viaddmin_u16x2(pm, P1, a[w]);                            // min(a[d-1] + P1, a[d])
      const uint32_t t2 = __viaddmin_u16x2(pp, P1, P2);                              // min(a[d+1] + P1, P2)
      lr[w] = __vadd2(c[w], __vminu2(t1, t2));
    }
  }
  uint32_t m2 = SGMU_BIG2;
#pragma unroll
  for (int w = 0; w < NW; w++) m2 = __vminu2(m2, lr[w]);
  minC = __reduce_min_sync(0xffffffffu, min(m2 & 0xffffu, m2 >> 16));
#pragma unroll
  for (int w = 0; w < NW; w++) { prev[w] = lr[w]; s[w] = MODE >= 1 ? __vadd2(s[w], lr[w]) : lr[w]; }
  xprev = x;
  if (MODE >= 2) {
    // gen_dispFromVm: strict '>' in increasing d -> the lowest d among the minima: key = sum << 16 | d
    uint32_t best = 0xffffffffu;
#pragma unroll
    for (int w = 0; w < NW; w++) {
      if (d0 + 2 * w < D) best = min(best, (s[w] << 16) | (uint32_t)(d0 + 2 * w));
      if (d0 + 2 * w + 1 < D) best = min(best, (s[w] & 0xffff0000u) | (uint32_t)(d0 + 2 * w + 1));
    }
    best = __reduce_min_sync(0xffffffffu, best);
    if (lane == 0) disp[p] = (int16_t)(best & 0xffffu);
  }
}

// MODE 0: out = Lr.  1: out += Lr.  2: out += Lr and d_disp = gen_dispFromVm of the finished sum.  3: that WTA alone.
// The cost volume holds raw costs (scaled here by the shift); the sum volume is in fixed point already.
template <int VPL, int PF, bool VEC, int MODE>
__global__ void __launch_bounds__(SGMU_WARPS * 32)
    k_sgm_path_u16(const uint16_t* __restrict__ vol, const uint32_t* __restrict__ pix, uint16_t* __restrict__ out, sgm_geom g,
                   int D, int corDifThres, int shift, int16_t* __restrict__ disp) {
  constexpr int NW = VPL / 2;
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * SGMU_WARPS + (threadIdx.x >> 5);
  if (k >= g.nLines) return;
  int v, u, len;
  line_start(g, k, v, u, len);
  const int d0 = lane * VPL;
  const long long pstep = (long long)g.mv * g.W + g.mu;
  long long p = (long long)v * g.W + u;
  const uint32_t scale = 1u << shift;
  // padding halves (d >= D) of a run: kept at BIG in the cost so they never win a minimum
  uint32_t padMask[NW];
#pragma unroll
  for (int w = 0; w < NW; w++)
    padMask[w] = (d0 + 2 * w < D ? 0u : 0xffffu) | (d0 + 2 * w + 1 < D ? 0u : 0xffff0000u);

  uint32_t cpf[PF][NW], spf[MODE >= 1 ? PF : 1][NW], xpf[PF];
#pragma unroll
  for (int i = 0; i < PF; i++) {
    if (i < len) {
      const long long q = p + pstep * i;
      u16_load_run<NW, VEC>(vol + q * D, d0, D, cpf[i]);
      if (MODE >= 1) u16_load_run<NW, VEC>(out + q * D, d0, D, spf[i]);
      xpf[i] = pix[q];
    }
  }
  uint32_t prev[NW];
  uint32_t minC = 0, xprev = 0;
  for (int t0 = 0; t0 < len; t0 += PF) {
#pragma unroll
    for (int i = 0; i < PF; i++) {
      const int t = t0 + i;
      if (t < len) {
        uint32_t c[NW], s[NW];
#pragma unroll
        for (int w = 0; w < NW; w++) {
          c[w] = ((cpf[i][w] << shift) & ~padMask[w]) | (SGMU_BIG2 & padMask[w]);   // raw cost -> fixed point (halves < 2^14)
          s[w] = MODE >= 1 ? spf[MODE >= 1 ? i : 0][w] : 0u;
        }
        const uint32_t x = xpf[i];
        if (t + PF < len) {
          const long long q = p + pstep * PF;
          u16_load_run<NW, VEC>(vol + q * D, d0, D, cpf[i]);
          if (MODE >= 1) u16_load_run<NW, VEC>(out + q * D, d0, D, spf[MODE >= 1 ? i : 0]);
          xpf[i] = pix[q];
        }
        u16_step<NW, MODE>(t == 0, c, s, prev, minC, x, xprev, d0, D, corDifThres, scale, lane, disp, p);
        if (MODE != 3) u16_store_run<NW, VEC>(out + p * D, d0, D, s);
        p += pstep;
      }
    }
  }
}

// Horizontal paths: a row is contiguous and there are only H of them (7 warps per SM at 1080p), so a register prefetch
// of 8 pixels cannot keep enough bytes in flight (measured 1.39 ms against 0.85 ms for the other directions at 1080p
// D=256).  Here the C and S runs and the pixel word of the next SGMU_NSTG pixels are staged with cp.async into per-lane
// shared-memory slots (as k_sgm_path_h does for the float volumes).  D % VPL == 0, VPL in {4, 8, 16}, aligned volumes.
#define SGMU_NSTG 12
template <int VPL, int MODE>
__global__ void __launch_bounds__(32)
    k_sgm_path_u16_h(const uint16_t* __restrict__ vol, const uint32_t* __restrict__ pix, uint16_t* __restrict__ out, int H, int W,
                     int mu, int D, int corDifThres, int shift, int16_t* __restrict__ disp) {
  extern __shared__ __align__(16) uint8_t sgmu_smem[];
  constexpr int NW = VPL / 2;
  constexpr int NS = SGMU_NSTG + 1;
  constexpr int RUNB = VPL * 2;                                   // bytes of one lane's run
  constexpr int SLOTB = 32 * RUNB * (MODE >= 1 ? 2 : 1) + 128;    // C runs | S runs | pixel words
  const int lane = threadIdx.x;
  const int v = blockIdx.x;
  const int d0 = lane * VPL;
  const bool act = d0 < D;
  const uint32_t scale = 1u << shift;
  const long long p0 = (long long)v * W + (mu > 0 ? 0 : W - 1);
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sgmu_smem);
  const uint32_t cOff = base + lane * RUNB, sOff = cOff + 32 * RUNB, xOff = base + 32 * RUNB * (MODE >= 1 ? 2 : 1) + lane * 4;
  auto cp_run = [&](uint32_t dst, const uint16_t* src) {
    if (RUNB == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    else {
#pragma unroll
      for (int k = 0; k < RUNB / 16; k++)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + k * 16), "l"(src + k * 8) : "memory");
    }
  };
  auto issue = [&](int t, int slot) {
    if (t < W) {
      const long long q = p0 + (long long)mu * t;
      const uint32_t so = slot * SLOTB;
      if (act) {
        cp_run(cOff + so, vol + q * D + d0);
        if (MODE >= 1) cp_run(sOff + so, out + q * D + d0);
      }
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(xOff + so), "l"(pix + q) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int t = 0; t < SGMU_NSTG; t++) issue(t, t);
  uint32_t prev[NW];
  uint32_t minC = 0, xprev = 0;
  int rd = 0, wr = SGMU_NSTG;
  long long pc = p0;
  for (int t = 0; t < W; t++) {
    asm volatile("cp.async.wait_group %0;" ::"n"(SGMU_NSTG - 1) : "memory");
    uint32_t c[NW], s[NW];
    const uint32_t so = rd * SLOTB;
#pragma unroll
    for (int w = 0; w < NW; w++) { c[w] = SGMU_BIG2; s[w] = 0u; }
    if (act) {
      if (RUNB == 8) {
        uint2 a;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a.x), "=r"(a.y) : "r"(cOff + so) : "memory");
        c[0] = a.x << shift; c[1 % NW] = a.y << shift;
        if (MODE >= 1) { asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a.x), "=r"(a.y) : "r"(sOff + so) : "memory"); s[0] = a.x; s[1 % NW] = a.y; }
      } else {
#pragma unroll
        for (int k = 0; k < RUNB / 16; k++) {
          uint4 a;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(cOff + so + k * 16) : "memory");
          c[(4 * k) % NW] = a.x << shift; c[(4 * k + 1) % NW] = a.y << shift; c[(4 * k + 2) % NW] = a.z << shift; c[(4 * k + 3) % NW] = a.w << shift;
          if (MODE >= 1) {
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(sOff + so + k * 16) : "memory");
            s[(4 * k) % NW] = a.x; s[(4 * k + 1) % NW] = a.y; s[(4 * k + 2) % NW] = a.z; s[(4 * k + 3) % NW] = a.w;
          }
        }
      }
    }
    uint32_t x;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x) : "r"(xOff + so) : "memory");
    issue(t + SGMU_NSTG, wr);
    u16_step<NW, MODE>(t == 0, c, s, prev, minC, x, xprev, d0, D, corDifThres, scale, lane, disp, pc);
    if (MODE != 3 && act) u16_store_run<NW, true>(out + pc * D, d0, D, s);
    pc += mu;
    if (++rd == NS) rd = 0;
    if (++wr == NS) wr = 0;
  }
}

template <int VPL>
static int launch_u16_h(sm_ctx* ctx, const uint16_t* vol, const uint32_t* pix, uint16_t* out, int H, int W, int mu, int D, int thr,
                        int shift, int mode, int16_t* disp) {
#define SGMU_H_LAUNCH(M)                                                                                                  \
  do {                                                                                                                    \
    const size_t smem = (size_t)(SGMU_NSTG + 1) * (32 * VPL * 2 * ((M) >= 1 ? 2 : 1) + 128);                              \
    SM_CUDA(cudaFuncSetAttribute(k_sgm_path_u16_h<VPL, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    SM_LAUNCH(ctx, (k_sgm_path_u16_h<VPL, M>), H, 32, smem, vol, pix, out, H, W, mu, D, thr, shift, disp);                 \
  } while (0)
  if (mode == 0) SGMU_H_LAUNCH(0); else if (mode == 1) SGMU_H_LAUNCH(1); else if (mode == 2) SGMU_H_LAUNCH(2); else SGMU_H_LAUNCH(3);
#undef SGMU_H_LAUNCH
  return SM_OK;
}

template <int VPL, int PF, bool VEC>
static int launch_u16(sm_ctx* ctx, const uint16_t* vol, const uint32_t* pix, uint16_t* out, const sgm_geom& g, int D, int thr,
                      int scale, int mode, int16_t* disp) {
  int sh = 0;
  while ((1 << sh) < scale) sh++;
  scale = sh;   // the kernels take log2(reduCoeffi1)
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
  const int vpl = per <= 2 ? 2 : (per <= 4 ? 4 : (per <= 8 ? 8 : 16));
  const bool vec = D % vpl == 0 && ((((uintptr_t)d_vol | (uintptr_t)d_out) & 15) == 0);
  if (g.mv == 0 && vec && vpl >= 4 && ((uintptr_t)d_pix & 3) == 0) {   // horizontal: cp.async staged rows
    int sh = 0;
    while ((1 << sh) < scale) sh++;
    if (vpl == 4) return launch_u16_h<4>(ctx, d_vol, d_pix, d_out, H, W, g.mu, D, corDifThres, sh, mode, d_disp);
    if (vpl == 8) return launch_u16_h<8>(ctx, d_vol, d_pix, d_out, H, W, g.mu, D, corDifThres, sh, mode, d_disp);
    return launch_u16_h<16>(ctx, d_vol, d_pix, d_out, H, W, g.mu, D, corDifThres, sh, mode, d_disp);
  }
  switch (vpl) {
    case 2: return vec ? launch_u16<2, 8, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp)
                       : launch_u16<2, 8, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp);
    case 4: return vec ? launch_u16<4, 8, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp)
                       : launch_u16<4, 8, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp);
    case 8: return vec ? launch_u16<8, 8, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp)
                       : launch_u16<8, 8, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp);
    default: return vec ? launch_u16<16, 8, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp)
                        : launch_u16<16, 4, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, scale, mode, d_disp);
  }
}

// The fixed point holds iff reduCoeffi1 is a power of two and paths * (maxCost + 3) * reduCoeffi1 stays below the sentinel.
bool smi_sgm_u16_ok(int D, int paths, int reduCoeffi1, int maxCost) {
  if (D < 1 || D > 512 || paths < 1 || paths > 8 || reduCoeffi1 < 1 || reduCoeffi1 > 64) return false;
  if (reduCoeffi1 & (reduCoeffi1 - 1)) return false;
  return (long long)paths * (maxCost + 3) * reduCoeffi1 <= SGMU_MAXSUM;   // headroom below the "outside" sentinel 0x3fff
}

int smi_sgm_u16(sm_ctx* ctx, const uint16_t* d_vol, const uint32_t* d_pix, int H, int W, int D, int paths, int corDifThres,
                int reduCoeffi1, uint16_t* d_sum, int16_t* d_disp, bool keep_sum, bool grouped) {
  if (paths == 8 && grouped) {
    // paths {0,4,5} in one upward row sweep, {1,6,7} in one downward sweep (sgm_group.cu), then the two horizontal paths
    const int rc = smi_sgm_group_u16(ctx, d_vol, d_pix, H, W, D, /*up*/1, /*mode*/0, corDifThres, reduCoeffi1, d_sum);
    if (rc == SM_OK) {
      SM_TRY(smi_sgm_group_u16(ctx, d_vol, d_pix, H, W, D, /*up*/0, /*mode*/1, corDifThres, reduCoeffi1, d_sum));
      SM_TRY(smi_sgm_path_u16(ctx, d_vol, d_pix, H, W, D, 2, corDifThres, reduCoeffi1, 1, d_sum, nullptr));
      SM_TRY(smi_sgm_path_u16(ctx, d_vol, d_pix, H, W, D, 3, corDifThres, reduCoeffi1, d_disp ? (keep_sum ? 2 : 3) : 1, d_sum, d_disp));
      return SM_OK;
    }
    if (rc != SM_ERR_UNSUPPORTED) return rc;
  }
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
    sm_set_error("sm_sgm_u16: needs a power-of-two reduCoeffi1 and paths * (maxCost + 3) * reduCoeffi1 <= 16000 "
                 "(got D %d, paths %d, reduCoeffi1 %d, maxCost %d); use sm_vol_to_f32 + sm_sgm", D, paths, reduCoeffi1, maxCost);
    return SM_ERR_UNSUPPORTED;
  }
  const long long npix = (long long)H * W;
  void* pk;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_IMG0, npix * 4, &pk));
  SM_TRY(smi_pack_bgr(ctx, d_bgr, npix, (uint32_t*)pk));
  return smi_sgm_u16(ctx, d_vol, (const uint32_t*)pk, H, W, D, paths, corDifThres, reduCoeffi1, d_sum, d_disp, true);
}
