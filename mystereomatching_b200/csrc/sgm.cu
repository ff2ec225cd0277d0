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
#include <stdlib.h>

#include "common.cuh"
#include "sgm_common.cuh"

#define SGM_WARPS 2
#define SGM_NSTG 12   // staged variant: pixels in flight per warp (cp.async), one extra slot so a refill never
                      // targets the slot being read

// MODE 3 = MODE 2 without the store of the finished sum (the pipeline's right view: nothing reads vm[1] after the WTA).
// MODE 0: out = Lr.  MODE 1: out += Lr (path sum).  MODE 2: out += Lr and the final WTA of gen_dispFromVm
// (stereoMatching.cpp:3928-3967: first minimum, -1 if nothing is below FLT_MAX) fused into the last path, which
// saves the separate read of the summed volume.
// STAGED (horizontal paths only): a row is contiguous in memory and there are only H of them (7 warps per SM at
// 1080p), so register prefetch cannot keep enough bytes in flight; the C and S runs and the pixel word are instead
// staged SGM_NSTG pixels ahead with cp.async into per-lane shared-memory slots.
__device__ __forceinline__ void sgm_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void sgm_cp8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void sgm_cp4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void sgm_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void sgm_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 sgm_lds16(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float2 sgm_lds8(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t sgm_lds4(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}

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

// One step of the recurrence for one pixel: updateCost<float> (stereoMatching.h:2205-2280) + gen_sgm_vm's sum.
template <int VPL, int MODE>
__device__ __forceinline__ void sgm_step(bool first, const float (&c)[VPL], float (&s)[VPL], float (&prev)[VPL],
                                         float& minC, uint32_t x, uint32_t& xprev, int d0, int D, int corDifThres,
                                         float redu, int lane, int16_t* __restrict__ disp, long long p) {
  float lr[VPL];
  if (first) {
#pragma unroll
    for (int j = 0; j < VPL; j++) lr[j] = c[j];
  } else {
    const int D1 = smd_absdiff_max3(x, xprev);
    float P1 = 1.0f, P2 = 3.0f;
    if (D1 > corDifThres) { P1 = P1 / redu; P2 = P2 / redu; }
    P1 = P1 - minC;
    float lo = __shfl_up_sync(0xffffffffu, prev[VPL - 1], 1);   // Lr'[d0-1]
    float hi = __shfl_down_sync(0xffffffffu, prev[0], 1);       // Lr'[d0+VPL]
    // No per-disparity range test.  The reference's out-of-range neighbour is S = FLT_MAX, which can never be the
    // minimum next to the finite P2; here the neighbour below d = 0 and above lane 31 is FLT_MAX, and a disparity >= D
    // is a padding element whose Lr stays >= FLT_MAX from the first pixel on (c = FLT_MAX + a positive minimum), so
    // S2 / S3 come out as FLT_MAX or +inf: never the minimum either, and no NaN (nothing subtracts inf from inf).
    if (lane == 0) lo = FLT_MAX;
    if (lane == 31) hi = FLT_MAX;
#pragma unroll
    for (int j = 0; j < VPL; j++) {
      const float pm = j == 0 ? lo : prev[j - 1];
      const float pp = j == VPL - 1 ? hi : prev[j + 1];
      const float S1 = prev[j] - minC;
      const float S2 = pm + P1;
      const float S3 = pp + P1;
      lr[j] = c[j] + fminf(fminf(S1, S2), fminf(S3, P2));
    }
  }
  // D-wide minimum of the new row (padding elements hold >= FLT_MAX via c[], every real Lr is below it)
  float m = FLT_MAX;
#pragma unroll
  for (int j = 0; j < VPL; j++) m = fminf(m, lr[j]);
  minC = key2f(__reduce_min_sync(0xffffffffu, f2key(m)));
#pragma unroll
  for (int j = 0; j < VPL; j++) prev[j] = lr[j];
  xprev = x;
  if (MODE >= 1) {
#pragma unroll
    for (int j = 0; j < VPL; j++) s[j] = s[j] + lr[j];  // gen_sgm_vm: sum += L[num]
  } else {
#pragma unroll
    for (int j = 0; j < VPL; j++) s[j] = lr[j];
  }
  if (MODE >= 2) {
    // gen_dispFromVm on the finished sum: strict '>' scan in increasing d -> the lowest d among the minima
    float bm = FLT_MAX;
    int bd = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < VPL; j++)
      if (bm > s[j]) { bm = s[j]; bd = d0 + j; }   // a padding element's sum is >= FLT_MAX: never below bm
    const uint32_t km = __reduce_min_sync(0xffffffffu, f2key(bm));
    const int cand = (f2key(bm) == km && bd != 0x7fffffff) ? bd : 0x7fffffff;
    const int best = (int)__reduce_min_sync(0xffffffffu, (unsigned)cand);
    if (lane == 0) disp[p] = (int16_t)(best == 0x7fffffff ? -1 : best);
  }
}

template <int VPL, int PF, bool VEC, int MODE>
__global__ void __launch_bounds__(SGM_WARPS * 32)
    k_sgm_path(const float* __restrict__ vol, const uint32_t* __restrict__ pix, float* __restrict__ out, sgm_geom g,
               int D, int corDifThres, float redu, int16_t* __restrict__ disp) {
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * SGM_WARPS + (threadIdx.x >> 5);
  if (k >= g.nLines) return;
  int v, u, len;
  line_start(g, k, v, u, len);
  const int d0 = lane * VPL;
  const long long pstep = (long long)g.mv * g.W + g.mu;
  long long p = (long long)v * g.W + u;

  float cpf[PF][VPL], spf[MODE >= 1 ? PF : 1][VPL];
  uint32_t xpf[PF];
#pragma unroll
  for (int i = 0; i < PF; i++) {
    if (i < len) {
      const long long q = p + pstep * i;
      load_run<VPL, VEC>(vol + q * D, d0, D, cpf[i]);
      if (MODE >= 1) load_run<VPL, VEC>(out + q * D, d0, D, spf[i]);
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
        float c[VPL], s[VPL];
#pragma unroll
        for (int j = 0; j < VPL; j++) { c[j] = cpf[i][j]; s[j] = MODE >= 1 ? spf[MODE >= 1 ? i : 0][j] : 0.f; }
        const uint32_t x = xpf[i];
        if (t + PF < len) {  // refill this prefetch slot for pixel t+PF
          const long long q = p + pstep * PF;
          load_run<VPL, VEC>(vol + q * D, d0, D, cpf[i]);
          if (MODE >= 1) load_run<VPL, VEC>(out + q * D, d0, D, spf[MODE >= 1 ? i : 0]);
          xpf[i] = pix[q];
        }
        sgm_step<VPL, MODE>(t == 0, c, s, prev, minC, x, xprev, d0, D, corDifThres, redu, lane, disp, p);
        if (MODE != 3) store_run<VPL, VEC>(out + p * D, d0, D, s);
        p += pstep;
      }
    }
  }
}

// Horizontal paths, cp.async staged (VPL % 4 == 0, D % 4 == 0, 16-byte aligned volumes).
template <int VPL, int MODE>
__global__ void __launch_bounds__(32)
    k_sgm_path_h(const float* __restrict__ vol, const uint32_t* __restrict__ pix, float* __restrict__ out, int H, int W,
                 int mu, int D, int corDifThres, float redu, int16_t* __restrict__ disp) {
  extern __shared__ __align__(16) uint8_t sgm_smem[];
  constexpr int NS = SGM_NSTG + 1;                       // slots
  constexpr int RUNB = VPL * 4;                          // bytes of one lane's run
  constexpr int SLOTB = 32 * RUNB * (MODE >= 1 ? 2 : 1) + 128;   // C run | S run | pixel word, per lane
  const int lane = threadIdx.x;
  const int v = blockIdx.x;
  const int d0 = lane * VPL;
  const bool act = d0 < D;                               // D % 4 == 0 and VPL % 4 == 0: a run is wholly in or out
  const int nq = act ? min(VPL, D - d0) / 4 : 0;         // 16-byte pieces of this lane's run
  const long long pstep = mu;
  long long p = (long long)v * W + (mu > 0 ? 0 : W - 1);
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sgm_smem);
  const uint32_t cOff = base + lane * RUNB, sOff = cOff + 32 * RUNB, xOff = base + 32 * RUNB * (MODE >= 1 ? 2 : 1) + lane * 4;

  auto issue = [&](int t, int slot) {
    if (t < W) {
      const long long q = p + pstep * t;
      const uint32_t so = slot * SLOTB;
      for (int k = 0; k < nq; k++) {
        sgm_cp16(cOff + so + k * 16, vol + q * D + d0 + k * 4);
        if (MODE >= 1) sgm_cp16(sOff + so + k * 16, out + q * D + d0 + k * 4);
      }
      sgm_cp4(xOff + so, pix + q);
    }
    sgm_commit();
  };
  const long long p0 = p;
  (void)p0;
  for (int t = 0; t < SGM_NSTG; t++) issue(t, t);
  float prev[VPL];
  float minC = 0.f;
  uint32_t xprev = 0;
  int rd = 0, wr = SGM_NSTG;
  // note: `issue` indexes pixels from the line start, so keep p fixed and carry the running pixel separately
  long long pc = p;
  for (int t = 0; t < W; t++) {
    sgm_wait<SGM_NSTG - 1>();
    float c[VPL], s[VPL];
    const uint32_t so = rd * SLOTB;
#pragma unroll
    for (int k = 0; k < VPL / 4; k++) {
      float4 a = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX), b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < nq) {
        a = sgm_lds16(cOff + so + k * 16);
        if (MODE >= 1) b = sgm_lds16(sOff + so + k * 16);
      }
      c[4 * k] = a.x; c[4 * k + 1] = a.y; c[4 * k + 2] = a.z; c[4 * k + 3] = a.w;
      s[4 * k] = b.x; s[4 * k + 1] = b.y; s[4 * k + 2] = b.z; s[4 * k + 3] = b.w;
    }
    const uint32_t x = sgm_lds4(xOff + so);
    issue(t + SGM_NSTG, wr);
    sgm_step<VPL, MODE>(t == 0, c, s, prev, minC, x, xprev, d0, D, corDifThres, redu, lane, disp, pc);
#pragma unroll
    for (int k = 0; k < VPL / 4; k++)
      if (MODE != 3 && k < nq) *reinterpret_cast<float4*>(out + pc * D + d0 + k * 4) = make_float4(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3]);
    pc += pstep;
    if (++rd == NS) rd = 0;
    if (++wr == NS) wr = 0;
  }
  sgm_wait<0>();
}

// Any direction, cp.async staged: k_sgm_path_h's pipeline on the scan-line geometry of k_sgm_path.  For SMALL frames: a
// sweep is then a few hundred warps (3 per SM at 450x375), each a chain of H or W dependent steps, and k_sgm_path's
// register prefetch cannot run ahead of the chain -- a warp has six scoreboards for all its loads, shuffles and the
// redux, so consuming slot i+1 waits for the refill of slot i just issued: ncu launch list, 450x375 D=64: 0.6 us per
// step = one memory latency per pixel.  cp.async completes in order against ONE counter (wait_group), so SGM_NSTG
// pixels really are in flight.  Runs of 4 disparities per lane (D <= 128: the lanes past D idle).
template <int VPL, int MODE>
__global__ void __launch_bounds__(32)
    k_sgm_path_s(const float* __restrict__ vol, const uint32_t* __restrict__ pix, float* __restrict__ out, sgm_geom g,
                 int D, int corDifThres, float redu, int16_t* __restrict__ disp) {
  extern __shared__ __align__(16) uint8_t sgm_smem[];
  constexpr int NS = SGM_NSTG + 1;                       // slots
  constexpr int PB = VPL >= 4 ? 16 : 8;                  // bytes of a piece: 16, or the whole run of 2 (D <= 64: all 32 lanes work)
  constexpr int PF = PB / 4;                             // floats of a piece
  constexpr int NQ = VPL * 4 / PB;                       // pieces of one lane's run
  constexpr int RUNB = VPL * 4;                          // bytes of one lane's run
  constexpr uint32_t SLOTB = 32 * RUNB * (MODE >= 1 ? 2 : 1) + 128;   // C run | S run | pixel word, per lane
  const int lane = threadIdx.x;
  int v, u, len;
  line_start(g, blockIdx.x, v, u, len);
  const int d0 = lane * VPL;
  // with a single warp per scan line the time per pixel is the warp's own instruction count (ncu: 170 per pixel at
  // 0.33 IPC, no memory stall left), so everything per pixel is incremental: byte pointers advanced by the line step,
  // slot offsets wrapped by one compare, the pieces of a run unrolled under per-lane predicates
  const long long pstep = (long long)g.mv * g.W + g.mu;
  const long long stepB = pstep * D * 4;
  long long pc = (long long)v * g.W + u;
  const uint8_t* cg = reinterpret_cast<const uint8_t*>(vol + pc * D + d0);   // next pixel to fetch: C run, S run, pixel word
  const uint8_t* sg = reinterpret_cast<const uint8_t*>(out + pc * D + d0);
  const uint8_t* xg = reinterpret_cast<const uint8_t*>(pix + pc);
  uint8_t* og = reinterpret_cast<uint8_t*>(out + pc * D + d0);               // pixel being computed: S run
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sgm_smem);
  const uint32_t cOff = base + lane * RUNB, sOff = cOff + 32 * RUNB, xOff = base + 32 * RUNB * (MODE >= 1 ? 2 : 1) + lane * 4;
  int ti = 0;                 // pixels issued so far
  uint32_t wr = 0, rd = 0;    // slot byte offsets

  auto issue = [&]() {
    if (ti < len) {
#pragma unroll
      for (int k = 0; k < NQ; k++)
        if (d0 + PF * k < D) {  // D % 4 == 0: a piece is wholly in or out
          if (PB == 16) {
            sgm_cp16(cOff + wr + k * 16, cg + k * 16);
            if (MODE >= 1) sgm_cp16(sOff + wr + k * 16, sg + k * 16);
          } else {
            sgm_cp8(cOff + wr, cg);
            if (MODE >= 1) sgm_cp8(sOff + wr, sg);
          }
        }
      sgm_cp4(xOff + wr, xg);
    }
    sgm_commit();
    cg += stepB; sg += stepB; xg += pstep * 4;
    ti++;
    wr += SLOTB;
    if (wr == NS * SLOTB) wr = 0;
  };
  for (int t = 0; t < SGM_NSTG; t++) issue();
  float prev[VPL];
  float minC = 0.f;
  uint32_t xprev = 0;
  for (int t = 0; t < len; t++) {
    sgm_wait<SGM_NSTG - 1>();
    float c[VPL], s[VPL];
#pragma unroll
    if (PB == 16) {
#pragma unroll
      for (int k = 0; k < NQ; k++) {
        float4 a = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX), b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d0 + 4 * k < D) {
          a = sgm_lds16(cOff + rd + k * 16);
          if (MODE >= 1) b = sgm_lds16(sOff + rd + k * 16);
        }
        c[4 * k] = a.x; c[4 * k + 1] = a.y; c[4 * k + 2] = a.z; c[4 * k + 3] = a.w;
        s[4 * k] = b.x; s[4 * k + 1] = b.y; s[4 * k + 2] = b.z; s[4 * k + 3] = b.w;
      }
    } else {
      float2 a = make_float2(FLT_MAX, FLT_MAX), b = make_float2(0.f, 0.f);
      if (d0 < D) {
        a = sgm_lds8(cOff + rd);
        if (MODE >= 1) b = sgm_lds8(sOff + rd);
      }
      c[0] = a.x; c[1] = a.y;
      s[0] = b.x; s[1] = b.y;
    }
    const uint32_t x = sgm_lds4(xOff + rd);
    issue();
    sgm_step<VPL, MODE>(t == 0, c, s, prev, minC, x, xprev, d0, D, corDifThres, redu, lane, disp, pc);
    if (MODE == 3) {
    } else if (PB == 16) {
#pragma unroll
      for (int k = 0; k < NQ; k++)
        if (d0 + 4 * k < D)
          *reinterpret_cast<float4*>(og + k * 16) = make_float4(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3]);
    } else if (d0 < D) {
      *reinterpret_cast<float2*>(og) = make_float2(s[0], s[1]);
    }
    og += stepB;
    pc += pstep;
    rd += SLOTB;
    if (rd == NS * SLOTB) rd = 0;
  }
  sgm_wait<0>();
}

// Horizontal paths through the TMA unit.  A row is one contiguous run of W*D floats, so whole groups of SGM_TK
// pixels (C run, S run, pixel words) are fetched with cp.async.bulk by one lane and land on an mbarrier; the
// LDGSTS pipe, which caps cp.async at ~16 B/clk/SM (~4.7 TB/s chip-wide, see scripts/microbench), is not involved.
// Needs D % 4 == 0 and W % SGM_TK == 0 (16-byte granular, 16-byte aligned copies).
#define SGM_TK 4    // pixels per stage
#define SGM_TNS 3   // stages in flight per warp

__device__ __forceinline__ void sgm_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void sgm_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sgm_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void sgm_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int VPL, int MODE>
__global__ void __launch_bounds__(32)
    k_sgm_path_t(const float* __restrict__ vol, const uint32_t* __restrict__ pix, float* __restrict__ out, int H, int W,
                 int mu, int D, int corDifThres, float redu, int16_t* __restrict__ disp) {
  extern __shared__ __align__(128) uint8_t sgm_smem[];
  const int lane = threadIdx.x;
  const int v = blockIdx.x;
  const int d0 = lane * VPL;
  const int nq = d0 < D ? min(VPL, D - d0) / 4 : 0;      // 16-byte pieces of this lane's run
  const uint32_t runB = (uint32_t)D * 4;                 // bytes of one pixel's D values
  const uint32_t cB = SGM_TK * runB, sB = MODE >= 1 ? cB : 0u, stageB = cB + sB + 16;
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sgm_smem);
  const uint32_t bars = base + SGM_TNS * stageB;
  const int nStages = W / SGM_TK;
  const size_t row = (size_t)v * W;

  // stage k covers pixels [x0, x0 + TK) in memory order, x0 = mu > 0 ? TK*k : W - TK*(k+1)
  auto issue = [&](int k, int slot) {
    const int x0 = mu > 0 ? SGM_TK * k : W - SGM_TK * (k + 1);
    const uint32_t st = base + slot * stageB, bar = bars + slot * 8;
    sgm_mbar_expect(bar, stageB);
    sgm_bulk(st, vol + (row + x0) * D, cB, bar);
    if (MODE >= 1) sgm_bulk(st + cB, out + (row + x0) * D, sB, bar);
    sgm_bulk(st + cB + sB, pix + row + x0, 16, bar);
  };
  if (lane == 0) {
    for (int s = 0; s < SGM_TNS; s++) sgm_mbar_init(bars + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int k = 0; k < SGM_TNS && k < nStages; k++) issue(k, k);
  }
  __syncwarp();

  float prev[VPL];
  float minC = 0.f;
  uint32_t xprev = 0;
  int slot = 0;
  uint32_t parity = 0;
  for (int k = 0; k < nStages; k++) {
    const uint32_t st = base + slot * stageB;
    sgm_mbar_wait(bars + slot * 8, parity);
    float c[SGM_TK][VPL], s[SGM_TK][VPL];
    uint32_t x[SGM_TK];
#pragma unroll
    for (int i = 0; i < SGM_TK; i++) {
      const int pi = mu > 0 ? i : SGM_TK - 1 - i;        // pixel inside the stage, in traversal order
#pragma unroll
      for (int q = 0; q < VPL / 4; q++) {
        float4 a = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX), b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < nq) {
          a = sgm_lds16(st + pi * runB + d0 * 4 + q * 16);
          if (MODE >= 1) b = sgm_lds16(st + cB + pi * runB + d0 * 4 + q * 16);
        }
        c[i][4 * q] = a.x; c[i][4 * q + 1] = a.y; c[i][4 * q + 2] = a.z; c[i][4 * q + 3] = a.w;
        s[i][4 * q] = b.x; s[i][4 * q + 1] = b.y; s[i][4 * q + 2] = b.z; s[i][4 * q + 3] = b.w;
      }
      x[i] = sgm_lds4(st + cB + sB + pi * 4);
    }
#pragma unroll
    for (int i = 0; i < SGM_TK; i++) {
      const int xi = mu > 0 ? SGM_TK * k + i : W - 1 - (SGM_TK * k + i);
      const long long p = (long long)row + xi;
      sgm_step<VPL, MODE>(k == 0 && i == 0, c[i], s[i], prev, minC, x[i], xprev, d0, D, corDifThres, redu, lane, disp, p);
#pragma unroll
      for (int q = 0; q < VPL / 4; q++)
        if (MODE != 3 && q < nq)
          *reinterpret_cast<float4*>(out + p * D + d0 + q * 4) = make_float4(s[i][4 * q], s[i][4 * q + 1], s[i][4 * q + 2], s[i][4 * q + 3]);
    }
    // refill the stage: it was only read (generic proxy) and every lane has consumed what it read, so no proxy
    // fence is needed -- one here would also wait for the warp's outstanding stores of S
    __syncwarp();
    if (lane == 0 && k + SGM_TNS < nStages) issue(k + SGM_TNS, slot);
    if (++slot == SGM_TNS) { slot = 0; parity ^= 1u; }
  }
}


template <int VPL, int PF, bool VEC>
static int launch_sgm(sm_ctx* ctx, const float* vol, const uint32_t* pix, float* out, const sgm_geom& g, int D,
                      int thr, float redu, int mode, int16_t* disp) {
  int grid = sm_div_up(g.nLines, SGM_WARPS);
  if (mode == 0)
    SM_LAUNCH(ctx, (k_sgm_path<VPL, PF, VEC, 0>), grid, SGM_WARPS * 32, 0, vol, pix, out, g, D, thr, redu, disp);
  else if (mode == 1)
    SM_LAUNCH(ctx, (k_sgm_path<VPL, PF, VEC, 1>), grid, SGM_WARPS * 32, 0, vol, pix, out, g, D, thr, redu, disp);
  else if (mode == 2)
    SM_LAUNCH(ctx, (k_sgm_path<VPL, PF, VEC, 2>), grid, SGM_WARPS * 32, 0, vol, pix, out, g, D, thr, redu, disp);
  else
    SM_LAUNCH(ctx, (k_sgm_path<VPL, PF, VEC, 3>), grid, SGM_WARPS * 32, 0, vol, pix, out, g, D, thr, redu, disp);
  return SM_OK;
}

template <int VPL, int MODE>
static int launch_sgm_h1(sm_ctx* ctx, const float* vol, const uint32_t* pix, float* out, int H, int W, int mu, int D,
                         int thr, float redu, int16_t* disp) {
  const size_t smem = (size_t)(SGM_NSTG + 1) * (32 * VPL * 4 * (MODE >= 1 ? 2 : 1) + 128);
  SM_CUDA(cudaFuncSetAttribute(k_sgm_path_h<VPL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SM_LAUNCH(ctx, (k_sgm_path_h<VPL, MODE>), H, 32, smem, vol, pix, out, H, W, mu, D, thr, redu, disp);
  return SM_OK;
}
template <int VPL, int MODE>
static int launch_sgm_t1(sm_ctx* ctx, const float* vol, const uint32_t* pix, float* out, int H, int W, int mu, int D,
                         int thr, float redu, int16_t* disp) {
  const size_t stageB = (size_t)SGM_TK * D * 4 * (MODE >= 1 ? 2 : 1) + 16;
  const size_t smem = SGM_TNS * stageB + SGM_TNS * 8;
  SM_CUDA(cudaFuncSetAttribute(k_sgm_path_t<VPL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SM_LAUNCH(ctx, (k_sgm_path_t<VPL, MODE>), H, 32, smem, vol, pix, out, H, W, mu, D, thr, redu, disp);
  return SM_OK;
}
template <int VPL>
static int launch_sgm_t(sm_ctx* ctx, const float* vol, const uint32_t* pix, float* out, int H, int W, int mu, int D,
                        int thr, float redu, int mode, int16_t* disp) {
  if (mode == 0) return launch_sgm_t1<VPL, 0>(ctx, vol, pix, out, H, W, mu, D, thr, redu, disp);
  if (mode == 1) return launch_sgm_t1<VPL, 1>(ctx, vol, pix, out, H, W, mu, D, thr, redu, disp);
  if (mode == 2) return launch_sgm_t1<VPL, 2>(ctx, vol, pix, out, H, W, mu, D, thr, redu, disp);
  return launch_sgm_t1<VPL, 3>(ctx, vol, pix, out, H, W, mu, D, thr, redu, disp);
}

template <int VPL>
static int launch_sgm_h(sm_ctx* ctx, const float* vol, const uint32_t* pix, float* out, int H, int W, int mu, int D,
                        int thr, float redu, int mode, int16_t* disp) {
  if (mode == 0) return launch_sgm_h1<VPL, 0>(ctx, vol, pix, out, H, W, mu, D, thr, redu, disp);
  if (mode == 1) return launch_sgm_h1<VPL, 1>(ctx, vol, pix, out, H, W, mu, D, thr, redu, disp);
  if (mode == 2) return launch_sgm_h1<VPL, 2>(ctx, vol, pix, out, H, W, mu, D, thr, redu, disp);
  return launch_sgm_h1<VPL, 3>(ctx, vol, pix, out, H, W, mu, D, thr, redu, disp);
}

template <int VPL, int MODE>
static int launch_sgm_s1(sm_ctx* ctx, const float* vol, const uint32_t* pix, float* out, const sgm_geom& g, int D, int thr,
                         float redu, int16_t* disp) {
  const size_t smem = (size_t)(SGM_NSTG + 1) * (32 * VPL * 4 * (MODE >= 1 ? 2 : 1) + 128);
  SM_CUDA(cudaFuncSetAttribute(k_sgm_path_s<VPL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  SM_LAUNCH(ctx, (k_sgm_path_s<VPL, MODE>), g.nLines, 32, smem, vol, pix, out, g, D, thr, redu, disp);
  return SM_OK;
}
template <int VPL>
static int launch_sgm_s(sm_ctx* ctx, const float* vol, const uint32_t* pix, float* out, const sgm_geom& g, int D, int thr,
                        float redu, int mode, int16_t* disp) {
  if (mode == 0) return launch_sgm_s1<VPL, 0>(ctx, vol, pix, out, g, D, thr, redu, disp);
  if (mode == 1) return launch_sgm_s1<VPL, 1>(ctx, vol, pix, out, g, D, thr, redu, disp);
  if (mode == 2) return launch_sgm_s1<VPL, 2>(ctx, vol, pix, out, g, D, thr, redu, disp);
  return launch_sgm_s1<VPL, 3>(ctx, vol, pix, out, g, D, thr, redu, disp);
}

// mode 0: d_out = Lr; 1: d_out += Lr; 2: d_out += Lr and d_disp = WTA of the finished sum; 3: the WTA alone (d_out is read, not written)
int smi_sgm_path_packed2(sm_ctx* ctx, const float* d_vol, const uint32_t* d_pix, int H, int W, int D, int path,
                         int corDifThres, int reduCoeffi1, int mode, float* d_out, int16_t* d_disp) {
  sgm_geom g;
  g.H = H; g.W = W; g.mv = -SGM_RV[path]; g.mu = -SGM_RU[path];
  g.nLines = g.mv == 0 ? H : (g.mu == 0 ? W : W + H - 1);
  const float redu = (float)reduCoeffi1;
  const bool vec = (D % 4 == 0) && (((uintptr_t)d_vol | (uintptr_t)d_out) % 16 == 0);
  const int vpl = D <= 32 ? 1 : D <= 64 ? 2 : D <= 128 ? 4 : D <= 256 ? 8 : 16;
  // small frames (<= 1024 scan lines): the cp.async staged kernel in every direction, runs of >= 4 disparities per lane
  if (vec && ((uintptr_t)d_pix & 3) == 0 && g.nLines <= 1024) {
    // 32 < D <= 64: runs of 2, all 32 lanes at work (runs of 4 leave half the warp idle and cost ~25 more instructions
    // per pixel); the TMA row kernel needs runs of >= 4
    if (vpl == 2) return launch_sgm_s<2>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode, d_disp);
    const int vs = max(vpl, 4);
    const bool tma = g.mv == 0 && W % SGM_TK == 0 && ((uintptr_t)d_pix & 15) == 0;
    if (!tma) {
      if (vs == 4) return launch_sgm_s<4>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode, d_disp);
      if (vs == 8) return launch_sgm_s<8>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode, d_disp);
      return launch_sgm_s<16>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode, d_disp);
    }
    if (vs == 4) return launch_sgm_t<4>(ctx, d_vol, d_pix, d_out, H, W, g.mu, D, corDifThres, redu, mode, d_disp);
  }
  // horizontal: few, contiguous lines: TMA bulk staging when the row tiles into groups of SGM_TK pixels, else cp.async staging
  if (g.mv == 0 && vec && vpl >= 4 && W % SGM_TK == 0 && ((uintptr_t)d_pix & 15) == 0) {
    if (vpl == 4) return launch_sgm_t<4>(ctx, d_vol, d_pix, d_out, H, W, g.mu, D, corDifThres, redu, mode, d_disp);
    if (vpl == 8) return launch_sgm_t<8>(ctx, d_vol, d_pix, d_out, H, W, g.mu, D, corDifThres, redu, mode, d_disp);
    return launch_sgm_t<16>(ctx, d_vol, d_pix, d_out, H, W, g.mu, D, corDifThres, redu, mode, d_disp);
  }
  if (g.mv == 0 && vec && vpl >= 4) {   // cp.async staged variant (W % SGM_TK != 0)
    if (vpl == 4) return launch_sgm_h<4>(ctx, d_vol, d_pix, d_out, H, W, g.mu, D, corDifThres, redu, mode, d_disp);
    if (vpl == 8) return launch_sgm_h<8>(ctx, d_vol, d_pix, d_out, H, W, g.mu, D, corDifThres, redu, mode, d_disp);
    return launch_sgm_h<16>(ctx, d_vol, d_pix, d_out, H, W, g.mu, D, corDifThres, redu, mode, d_disp);
  }
  // 16-byte loads need D % 4 == 0 (then every float4 of a run is wholly inside or outside [0,D))
  switch (vpl) {
    case 1: return launch_sgm<1, 8, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode, d_disp);
    case 2: return launch_sgm<2, 8, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode, d_disp);
    case 4:
      if (vec) return launch_sgm<4, 8, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode, d_disp);
      return launch_sgm<4, 8, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode, d_disp);
    case 8:
      if (vec) return launch_sgm<8, 4, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode, d_disp);
      return launch_sgm<8, 4, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode, d_disp);
    default:
      if (vec) return launch_sgm<16, 2, true>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode, d_disp);
      return launch_sgm<16, 2, false>(ctx, d_vol, d_pix, d_out, g, D, corDifThres, redu, mode, d_disp);
  }
}

int smi_sgm_path_packed(sm_ctx* ctx, const float* d_vol, const uint32_t* d_pix, int H, int W, int D, int path,
                        int corDifThres, int reduCoeffi1, int mode, float* d_out) {
  return smi_sgm_path_packed2(ctx, d_vol, d_pix, H, W, D, path, corDifThres, reduCoeffi1, mode, d_out, nullptr);
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

// 8 paths, row-wise groups first (see sgm_group.cu); d_disp (nullable): WTA of the finished sum fused into the last path
int smi_sgm8_grouped(sm_ctx* ctx, const float* d_vol, const uint32_t* d_pix, int H, int W, int D, int corDifThres,
                     int reduCoeffi1, float* d_sum, int16_t* d_disp, cudaEvent_t ev_after_sweeps, bool* used_sweeps,
                     bool keep_sum) {
  const int last = d_disp ? (keep_sum ? 2 : 3) : 1;
  if (used_sweeps) *used_sweeps = false;
  int rc = smi_sgm_group(ctx, d_vol, d_pix, H, W, D, /*up*/1, /*mode*/0, corDifThres, reduCoeffi1, d_sum);
  if (rc == SM_ERR_UNSUPPORTED) {   // shape outside the grouped kernel: reference order, path by path
    for (int i = 0; i < 8; i++)
      SM_TRY(smi_sgm_path_packed2(ctx, d_vol, d_pix, H, W, D, i, corDifThres, reduCoeffi1,
                                  i == 0 ? 0 : (i == 7 ? last : 1), d_sum, d_disp));
    return SM_OK;
  }
  SM_TRY(rc);
  SM_TRY(smi_sgm_group(ctx, d_vol, d_pix, H, W, D, /*up*/0, /*mode*/1, corDifThres, reduCoeffi1, d_sum));
  if (used_sweeps) *used_sweeps = true;
  if (ev_after_sweeps) SM_CUDA(cudaEventRecord(ev_after_sweeps, ctx->stream));
  SM_TRY(smi_sgm_path_packed2(ctx, d_vol, d_pix, H, W, D, 2, corDifThres, reduCoeffi1, 1, d_sum, nullptr));
  SM_TRY(smi_sgm_path_packed2(ctx, d_vol, d_pix, H, W, D, 3, corDifThres, reduCoeffi1, last, d_sum, d_disp));
  return SM_OK;
}

// Both views of a frame: the two row sweeps run the left and the right volume in one launch each (two CTAs per SM),
// then the horizontal paths per view.  SM_ERR_UNSUPPORTED when the shape does not fit two CTAs per SM.
int smi_sgm8_grouped2(sm_ctx* ctx, const float* const* d_vol, const uint32_t* const* d_pix, int H, int W, int D, int corDifThres,
                      int reduCoeffi1, float* const* d_sum, int16_t* const* d_disp, cudaEvent_t ev_after_sweeps, bool* used_sweeps,
                      const bool* keep_sum) {
  if (used_sweeps) *used_sweeps = false;
  int rc = smi_sgm_group2(ctx, d_vol, d_pix, H, W, D, /*up*/1, /*mode*/0, corDifThres, reduCoeffi1, d_sum);
  if (rc != SM_OK) return rc;
  SM_TRY(smi_sgm_group2(ctx, d_vol, d_pix, H, W, D, /*up*/0, /*mode*/1, corDifThres, reduCoeffi1, d_sum));
  if (used_sweeps) *used_sweeps = true;
  if (ev_after_sweeps) SM_CUDA(cudaEventRecord(ev_after_sweeps, ctx->stream));
  for (int i = 0; i < 2; i++) {
    int16_t* dd = d_disp ? d_disp[i] : nullptr;
    SM_TRY(smi_sgm_path_packed2(ctx, d_vol[i], d_pix[i], H, W, D, 2, corDifThres, reduCoeffi1, 1, d_sum[i], nullptr));
    SM_TRY(smi_sgm_path_packed2(ctx, d_vol[i], d_pix[i], H, W, D, 3, corDifThres, reduCoeffi1,
                                dd ? ((!keep_sum || keep_sum[i]) ? 2 : 3) : 1, d_sum[i], dd));
  }
  return SM_OK;
}

extern "C" int sm_sgm_grouped2(sm_ctx* ctx, const float* d_volL, const float* d_volR, const uint8_t* d_bgrL, const uint8_t* d_bgrR,
                               int H, int W, int D, int corDifThres, int reduCoeffi1, float* d_sumL, float* d_sumR) {
  SM_CHECK_ARG(ctx && d_volL && d_volR && d_bgrL && d_bgrR && d_sumL && d_sumR);
  SM_CHECK_ARG(H > 0 && W > 0 && D > 0 && D <= 512 && reduCoeffi1 != 0);
  SM_CHECK_ARG(d_volL != d_sumL && d_volR != d_sumR && d_sumL != d_sumR && d_volL != d_sumR && d_volR != d_sumL);
  const long long npix = (long long)H * W;
  void *pk0, *pk1;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_IMG0, npix * 4, &pk0));
  SM_TRY(sm_scratch_get(ctx, SM_SCR_IMG1, npix * 4, &pk1));
  SM_TRY(smi_pack_bgr(ctx, d_bgrL, npix, (uint32_t*)pk0));
  SM_TRY(smi_pack_bgr(ctx, d_bgrR, npix, (uint32_t*)pk1));
  const float* vols[2] = {d_volL, d_volR};
  const uint32_t* pixs[2] = {(const uint32_t*)pk0, (const uint32_t*)pk1};
  float* sums[2] = {d_sumL, d_sumR};
  int rc = smi_sgm8_grouped2(ctx, vols, pixs, H, W, D, corDifThres, reduCoeffi1, sums, nullptr);
  if (rc != SM_ERR_UNSUPPORTED) return rc;
  for (int i = 0; i < 2; i++)   // shape outside the two-view launch: one view at a time
    SM_TRY(smi_sgm8_grouped(ctx, vols[i], pixs[i], H, W, D, corDifThres, reduCoeffi1, sums[i], nullptr));
  return SM_OK;
}

extern "C" int sm_sgm_grouped(sm_ctx* ctx, const float* d_vol, const uint8_t* d_bgr, int H, int W, int D, int corDifThres,
                              int reduCoeffi1, float* d_sum) {
  SM_CHECK_ARG(ctx && d_vol && d_bgr && d_sum);
  SM_CHECK_ARG(H > 0 && W > 0 && D > 0 && D <= 512 && reduCoeffi1 != 0 && d_vol != d_sum);
  const long long npix = (long long)H * W;
  void* pk;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_IMG0, npix * 4, &pk));
  SM_TRY(smi_pack_bgr(ctx, d_bgr, npix, (uint32_t*)pk));
  return smi_sgm8_grouped(ctx, d_vol, (const uint32_t*)pk, H, W, D, corDifThres, reduCoeffi1, d_sum, nullptr);
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
