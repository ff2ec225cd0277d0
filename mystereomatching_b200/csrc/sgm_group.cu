// K5b: grouped SGM sweeps -- three path directions per pass over the volume.
//
// The per-path kernels (sgm.cu) read C and read-modify-write the path sum S once PER PATH: (3P-1) V b bytes, which
// is what bounds them (83 % of the HBM roofline).  Three of the eight reference directions advance together row by
// row: the predecessors of paths {1, 6, 7} (stereoMatching.cpp:6207-6208: rv = -1, ru = 0 / +1 / -1) all lie in row
// v-1, those of paths {0, 4, 5} (rv = +1, ru = 0 / -1 / +1) in row v+1.  One sweep over the rows can therefore
// carry all three recurrences on ONE read of C and ONE read-modify-write of S: 3 V b per group instead of 9 V b.
//
// Shape: one cooperative launch, one CTA per SM, a CTA owns CW adjacent columns, a warp owns one column (D spread
// over lanes as runs of VPL), rows in lock step:
//   * the vertical path's previous Lr row stays in registers;
//   * the two diagonal paths need the previous row of the NEIGHBOUR column: warps publish their rows (and row
//     minima) in shared memory, double buffered, one __syncthreads per row;
//   * across CTA boundaries the edge columns' rows go through global memory (each CTA waits only for its two
//     neighbours: a wavefront, not a grid barrier).  The hand-off carries no flag and no fence: the edge buffers
//     are pre-set to a NaN bit pattern no Lr value can have, the writer stores the row, and every reader lane polls
//     ITS OWN words until none is the sentinel (each word changes exactly once, so no ordering between words is
//     needed).  That is one L2 round trip per row instead of store + fence + flag + load (measured 5.9 -> see
//     DESIGN.md us per row);
//   * C and S rows of the next NS (4, or 2 when two views share an SM) rows are staged per warp by the TMA unit (cp.async.bulk + mbarrier).
//
// Arithmetic: every Lr value is computed exactly as updateCost<float> does (stereoMatching.h:2205-2280), so each
// path volume is bit-identical to the reference's.  Only the ORDER in which the eight path volumes are added into
// S differs from gen_sgm_vm's (stereoMatching.cpp:2031-2056: L0+L1+...+L7): here ((L0+L4)+L5), then +L1+L6+L7,
// then +L2, +L3.  For integer-valued costs (the Census path) every partial sum is exact, so S is still bit-exact;
// for float costs S agrees to a few ulp (tests: <= 1e-6 relative, the north star allows 1e-4) and the disparity
// maps agree to >= 99.5 %.  The reference-order variant (sm_sgm) stays available and bit-exact.
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"


__device__ __forceinline__ uint32_t g_f2key(float x) {
  uint32_t b = __float_as_uint(x);
  return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float g_key2f(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k ^ 0x80000000u) : ~k;
  return __uint_as_float(b);
}
__device__ __forceinline__ void g_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void g_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void g_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ float4 g_lds16(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void g_sts16(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ int g_ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void g_st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One pixel of one path: updateCost<float> given the predecessor's Lr row (pr), its minimum and the colour step.
// Disparities outside [0, D) are FLT_MAX in pr and c: FLT_MAX + P1 rounds back to FLT_MAX, which never wins a
// minimum against the finite candidates, so no per-disparity range predicate is needed.
// PACK (one view per launch, 128 registers at hand): the additions are issued as packed pairs (add.rn.f32x2 -> FADD2 on
// sm_100a: two IEEE single-precision additions per instruction, each rounded exactly like the scalar one); pr + P1 is formed
// once per disparity and serves both as the "d - 1" term of d + 1 and the "d + 1" term of d - 1.  Same operands, same
// single rounding per sum: bit-identical Lr.  The two-view launch (72 registers) keeps the scalar form, which does not spill.
template <int VPL, bool PACK>
__device__ __forceinline__ void g_lr(const float (&c)[VPL], const float (&pr)[VPL], float minP, bool step, float P1r, float P2r,
                                     int lane, float (&lr)[VPL], float& minNew) {
  const float P2 = step ? P2r : 3.0f;
  const float P1 = (step ? P1r : 1.0f) - minP;
  float lo = __shfl_up_sync(0xffffffffu, pr[VPL - 1], 1);
  float hi = __shfl_down_sync(0xffffffffu, pr[0], 1);
  lo = lane == 0 ? FLT_MAX : lo;
  hi = lane == 31 ? FLT_MAX : hi;
  float m = FLT_MAX;
  if (PACK) {
    const float2 P1v = make_float2(P1, P1), nMv = make_float2(-minP, -minP);
    float q[VPL + 2];   // q[k + 1] = pr[k] + P1, k = -1 .. VPL
    float a[VPL];       // pr[j] - minP
    q[0] = lo + P1;
    q[VPL + 1] = hi + P1;
#pragma unroll
    for (int j = 0; j < VPL; j += 2) {
      const float2 p2 = make_float2(pr[j], pr[j + 1]);
      const float2 t = __fadd2_rn(p2, P1v), u = __fadd2_rn(p2, nMv);
      q[j + 1] = t.x; q[j + 2] = t.y;
      a[j] = u.x; a[j + 1] = u.y;
    }
#pragma unroll
    for (int j = 0; j < VPL; j += 2) {
      const float m0 = fminf(fminf(a[j], q[j]), fminf(q[j + 2], P2));
      const float m1 = fminf(fminf(a[j + 1], q[j + 1]), fminf(q[j + 3], P2));
      const float2 r = __fadd2_rn(make_float2(c[j], c[j + 1]), make_float2(m0, m1));
      lr[j] = r.x; lr[j + 1] = r.y;
      m = fminf(m, fminf(r.x, r.y));
    }
  } else {
#pragma unroll
    for (int j = 0; j < VPL; j++) {
      const float pm = j == 0 ? lo : pr[j - 1];
      const float pp = j == VPL - 1 ? hi : pr[j + 1];
      lr[j] = c[j] + fminf(fminf(pr[j] - minP, pm + P1), fminf(pp + P1, P2));
      m = fminf(m, lr[j]);
    }
  }
  minNew = g_key2f(__reduce_min_sync(0xffffffffu, g_f2key(m)));
}
template <int VPL>
__device__ __forceinline__ float g_rowmin(const float (&lr)[VPL]) {
  float m = FLT_MAX;
#pragma unroll
  for (int j = 0; j < VPL; j++) m = fminf(m, lr[j]);
  return g_key2f(__reduce_min_sync(0xffffffffu, g_f2key(m)));
}

// edge exchange between neighbouring CTAs: rows of one path of one boundary column
struct sgmg_edges {
  float* rowsP;   // [nb][H][Dp]  first column of CTA b, path whose predecessor is column u+1 (read by CTA b-1)
  float* rowsM;   // [nb][H][Dp]  last column of CTA b, path whose predecessor is column u-1 (read by CTA b+1)
  int Dp;         // row pitch in floats: D values + 1 minimum, padded to a multiple of 4
  unsigned long long* trace;   // diagnostics (SM_SGMG_TRACE=<file>): [cta][first|last][row]{t_publish, t_request, t_got, polls} in ns
};
__device__ __forceinline__ unsigned long long g_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define SGMG_SENTINEL 0xFFFFFFFFu   // a NaN pattern: never a valid Lr value or minimum

__device__ __forceinline__ uint4 g_ld_relaxed16(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t g_ld_relaxed4(const void* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float g_lds4(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void g_sts4(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void g_pair_bar(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// a row of D floats of this lane's run (quads beyond D read as FLT_MAX / are not written)
// QS: byte distance between the lane's quads.  16: the global layout (a run of VPL floats per lane; with VPL = 8 the
// lanes of a quarter-warp sit 32 bytes apart, so every LDS.128 / STS.128 is a 2-way bank conflict -- unavoidable
// for the TMA-staged C / S rows, which arrive in global order).  512: the exchange rows' own layout
// [quad][lane][4]: a quad of all 32 lanes is 512 contiguous bytes, conflict-free.
template <int VPL, int QS = 16>
__device__ __forceinline__ void g_ldrow(uint32_t a, int nq, float (&x)[VPL]) {
#pragma unroll
  for (int q = 0; q < VPL / 4; q++) {
    float4 t = q < nq ? g_lds16(a + q * QS) : make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX);
    x[4 * q] = t.x; x[4 * q + 1] = t.y; x[4 * q + 2] = t.z; x[4 * q + 3] = t.w;
  }
}
template <int VPL, int QS = 16>
__device__ __forceinline__ void g_strow(uint32_t a, int nq, const float (&x)[VPL]) {
#pragma unroll
  for (int q = 0; q < VPL / 4; q++)
    if (q < nq) g_sts16(a + q * QS, make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]));
}
// The consumer of a neighbour CTA's published row also RE-ARMS what it has read: every word goes back to the sentinel, so
// the buffers are all-sentinel again when the launch ends and the next launch needs no fill (each word is written once by
// its producer and read by exactly one consumer lane; the producer's next write to it happens in a later launch).  The
// row minimum is ONE word all lanes read: nobody re-arms it before everybody has it.
// The far row by prefetch: the edge column copies ITS words of the neighbour CTA's row into shared memory with cp.async
// (L2 -> shared, no registers, no scoreboard) at the END of the previous sweep row, so the L2 round trip (0.7 us under load)
// runs under the pair barriers and the near diagonal (which is computed first and published at once) instead of sitting
// between the row top and the far diagonal.  Same flag-free protocol: a word still carrying the sentinel after the copy was fetched too early and is
// polled directly, as before.
__device__ __forceinline__ void g_cpasync16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void g_cpasync_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void g_cpasync_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint4 g_lds16u(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
template <int VPL>
__device__ __forceinline__ void g_edge_prefetch(const float* src, int nq, int Dlane, uint32_t dst, int lane) {
#pragma unroll
  for (int q = 0; q < VPL / 4; q++)
    if (q < nq) g_cpasync16(dst + q * 16, src + q * 4);
  if (lane == 0) g_cpasync16(dst + Dlane * 4, src + Dlane);   // the quad that holds the row minimum (lane 0: Dlane = D)
  g_cpasync_commit();
}
// dst: this lane's words of the prefetched row in shared memory; minAddr: the row minimum's word there.
// (Tried and dropped: a second, EARLY copy issued in the middle of the previous row so that nothing is waited for at
// all -- it is fetched before the neighbour has published too often (2.5 polls per row on average, rows of 2.8 us):
// 5.70 instead of 5.31 ms per view.)
template <int VPL>
__device__ __forceinline__ void g_edge_take(const float* src, int nq, int D, uint32_t dst, uint32_t minAddr, float (&pr)[VPL], float& pm,
                                            int& polls) {
  float* w = const_cast<float*>(src);
  g_cpasync_wait0();
  __syncwarp();
#pragma unroll
  for (int q = 0; q < VPL / 4; q++) {
    uint4 t = make_uint4(0x7f7fffffu, 0x7f7fffffu, 0x7f7fffffu, 0x7f7fffffu);   // FLT_MAX padding
    if (q < nq) {
      t = g_lds16u(dst + q * 16);
      while (t.x == SGMG_SENTINEL || t.y == SGMG_SENTINEL || t.z == SGMG_SENTINEL || t.w == SGMG_SENTINEL) { t = g_ld_relaxed16(src + q * 4); polls++; }
      __stcg(reinterpret_cast<uint4*>(w + q * 4), make_uint4(SGMG_SENTINEL, SGMG_SENTINEL, SGMG_SENTINEL, SGMG_SENTINEL));
    }
    pr[4 * q] = __uint_as_float(t.x); pr[4 * q + 1] = __uint_as_float(t.y);
    pr[4 * q + 2] = __uint_as_float(t.z); pr[4 * q + 3] = __uint_as_float(t.w);
  }
  uint32_t m = __float_as_uint(g_lds4(minAddr));
  while (__any_sync(0xffffffffu, m == SGMG_SENTINEL)) {
    if (m == SGMG_SENTINEL) m = g_ld_relaxed4(src + D);
  }
  pm = __uint_as_float(m);
  if ((threadIdx.x & 31) == 0) __stcg(reinterpret_cast<uint32_t*>(w + D), SGMG_SENTINEL);
}

// UP = 0: rows 0 .. H-1, paths {1, 6, 7} (predecessor columns u, u+1, u-1).
// UP = 1: rows H-1 .. 0, paths {0, 4, 5} (predecessor columns u, u-1, u+1).
// MODE 0: S = LA + LB + LC; MODE 1: S = ((S + LA) + LB) + LC.   FULL: D == 32 * VPL (no padding lanes).
// One view of a frame for the sweep: NV = 2 runs the left and the right volume in ONE cooperative launch, two CTAs
// per SM (one per view where the scheduler places them so): the kernel is latency-bound with 13 warps per SM, a
// second, independent set of warps hides that.
struct sgmg_view {
  const float* vol;
  const uint32_t* pix;
  float* out;
  float* rowsP;
  float* rowsM;
};
template <int VPL, int UP, int MODE, bool FULL, int NS, int NV>
__global__ void __launch_bounds__(448, NV)
    k_sgm_group(const sgmg_view v0, const sgmg_view v1, int H, int W, int D, int corDifThres, float redu, int Dp,
                unsigned long long* trace) {
  extern __shared__ __align__(128) uint8_t gsm[];
  const int lane = threadIdx.x & 31;
  const int nwarp = blockDim.x >> 5;   // widest CTA's column count
  const int nb = gridDim.x / NV;
  const int view = NV == 2 ? (int)(blockIdx.x >= (unsigned)nb) : 0;
  const int b = blockIdx.x - view * nb;
  const float* __restrict__ vol = view ? v1.vol : v0.vol;
  const uint32_t* __restrict__ pix = view ? v1.pix : v0.pix;
  float* __restrict__ out = view ? v1.out : v0.out;
  sgmg_edges E;
  E.rowsP = view ? v1.rowsP : v0.rowsP;
  E.rowsM = view ? v1.rowsM : v0.rowsM;
  E.Dp = Dp;
  E.trace = trace;
  const int u0 = (int)(((long long)b * W) / nb);
  const int nCols = (int)(((long long)(b + 1) * W) / nb) - u0;   // >= 4 (host: nb <= W / 4)
  if ((int)(threadIdx.x >> 5) >= nCols) return;                  // no CTA-wide barrier below
  // Column of this warp inside the CTA.  Rotated by two so that the two edge columns (the slowest: their far
  // diagonal cannot interleave with the rest) do not share a warp scheduler (warp id mod 4) with each other, nor,
  // for the usual 4k+1 columns, sit on the scheduler that carries one warp more than the others.
  const int warp = ((int)(threadIdx.x >> 5) + nCols - 2) % nCols;
  const int u = u0 + warp;
  const int d0 = lane * VPL;
  const int nq = FULL ? VPL / 4 : (d0 < D ? min(VPL, D - d0) / 4 : 0);
  const uint32_t runB = (uint32_t)D * 4;
  // exchange rows (written and read only by this kernel): conflict-free layout when a lane owns two quads
  constexpr bool SWZ = FULL && VPL == 8;
  constexpr int XQ = SWZ ? 512 : 16;
  const uint32_t exOff = SWZ ? (uint32_t)lane * 16u : (uint32_t)d0 * 4u;
  constexpr int ob = UP ? -1 : +1, oc = -ob;   // predecessor column offsets of paths B and C (reference order)

  // shared memory: per warp NS stages {C run, S run}; exchange rows [path][buf][CW][D]; minima [path][buf][CW]
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(gsm);
  const uint32_t stageB = runB * (MODE >= 1 ? 2 : 1);
  const uint32_t stLo = base + (uint32_t)warp * NS * stageB;
  const uint32_t exLo = base + (uint32_t)nwarp * NS * stageB;
  const uint32_t exBufB = (uint32_t)nwarp * runB, exPathB = 2u * exBufB;
  const uint32_t minLo = exLo + 2 * exPathB;
  const uint32_t minBufB = (uint32_t)nwarp * 4, minPathB = 2u * minBufB;
  const uint32_t bars = minLo + 2 * minPathB + (uint32_t)warp * NS * 8;   // (8-byte aligned: see host)
  // prefetched far rows of the two edge columns: [first | last][row parity][Dp floats]
  const uint32_t farLo = (minLo + 2 * minPathB + (uint32_t)nwarp * NS * 8 + 15u) & ~15u;

  const long long rowStep = UP ? -(long long)W : (long long)W;   // pixels from one row of the sweep to the next
  const size_t p0 = (size_t)(UP ? H - 1 : 0) * W + u;            // first pixel of this column in sweep order
  auto issue = [&](int r, int slot) {
    const size_t p = (size_t)((long long)p0 + rowStep * r);
    const uint32_t st = stLo + slot * stageB, bar = bars + slot * 8;
    g_mbar_expect(bar, stageB);
    g_bulk(st, vol + p * D, runB, bar);
    if (MODE >= 1) g_bulk(st + runB, out + p * D, runB, bar);
  };
  if (lane == 0) {
    for (int s = 0; s < NS; s++) g_mbar_init(bars + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int r = 0; r < NS && r < H; r++) issue(r, r);
  }
  __syncwarp();

  // roles.  A path whose previous row sits in the neighbour CTA ("far") is fed through the global edge rows; the
  // other diagonal of an edge column ("near") is what the neighbour CTA consumes and is published there.
  const bool isFirst = warp == 0, isLast = warp == nCols - 1;
  const bool special = isFirst || isLast;
  const int nearOff = isFirst ? +1 : -1;
  const bool nearIsB = nearOff == ob;
  const bool nbrCta = isFirst ? b > 0 : b + 1 < nb;   // the CTA on the far side exists (else: image border)
  float* pubRow = (isFirst ? E.rowsP : E.rowsM) + (size_t)b * H * E.Dp + d0;
  const float* farRow = (isFirst ? E.rowsM : E.rowsP) + (size_t)(nbrCta ? (isFirst ? b - 1 : b + 1) : b) * H * E.Dp + d0;
  const float* farNext = farRow;   // the row the NEXT sweep row consumes (prefetched at the end of this one)
  const uint32_t farBuf = farLo + (isLast ? 2u : 0u) * (uint32_t)E.Dp * 4u;
  const float P1r = 1.0f / redu, P2r = 3.0f / redu;

  // pixel words (BGR packed), 32 rows at a time: lane i holds sweep row 32 * chunk + i of this column and of the two
  // diagonal predecessor columns.  One strided fetch per 32 rows and three shuffles per row keep the LSU free for
  // the edge polls, which otherwise queue behind these loads (L1 returns in order).
  const int ub = min(max(u + ob, 0), W - 1) - u, uc = min(max(u + oc, 0), W - 1) - u;
  uint32_t curO = 0, curB = 0, curC = 0, nxtO = 0, nxtB = 0, nxtC = 0;
  auto loadChunk = [&](int chunk, uint32_t& o_, uint32_t& b_, uint32_t& c_) {
    const int rr = chunk * 32 + lane;
    if (rr < H) {
      const uint32_t* q = pix + (size_t)((long long)p0 + rowStep * rr);
      o_ = q[0]; b_ = q[ub]; c_ = q[uc];
    }
  };
  loadChunk(0, curO, curB, curC);
  loadChunk(1, nxtO, nxtB, nxtC);
  uint32_t xrow = __shfl_sync(0xffffffffu, curO, 0), xpA = 0, xpB = 0, xpC = 0;
  float* o = out + p0 * D + d0;
  const long long oStep = rowStep * D;

  float prevA[VPL], minA = 0.f;
  int slot = 0;
  uint32_t parity = 0;
  const int barL = warp, barR = warp + 1;   // named barrier of the pair (w-1, w) is w: ids 1 .. 15
  const bool hasL = warp > 0, hasR = warp + 1 < nCols;

  for (int r = 0; r < H; r++) {
    const uint32_t bufOff = (r & 1) ? exBufB : 0u, pbufOff = exBufB - bufOff;
    const uint32_t mbufOff = (r & 1) ? minBufB : 0u, mpbufOff = minBufB - mbufOff;
    float c[VPL], s[VPL], lrA[VPL], lrB[VPL], lrC[VPL];
    unsigned long long* tl = (E.trace && b == 70 && r >= 400 && r < 432 && lane == 0)
                                 ? E.trace + (size_t)nb * 2 * H * 4 + ((size_t)warp * 32 + (r - 400)) * 8 : nullptr;
    if (tl) tl[0] = g_now();
    const uint32_t st = stLo + slot * stageB + d0 * 4;
    g_mbar_wait(bars + slot * 8, parity);
    g_ldrow<VPL>(st, nq, c);
    if (MODE >= 1) g_ldrow<VPL>(st + runB, nq, s);
    if (tl) tl[1] = g_now();
    const uint32_t myB = exLo + bufOff + (uint32_t)warp * runB + exOff, myC = myB + exPathB;
    const uint32_t myMinB = minLo + mbufOff + (uint32_t)warp * 4, myMinC = myMinB + minPathB;
    if (r == 0) {   // first row of the sweep: no predecessor inside the image -> Lr = C on all three paths
      const float m = g_rowmin<VPL>(c);
#pragma unroll
      for (int j = 0; j < VPL; j++) lrA[j] = lrB[j] = lrC[j] = c[j];
      minA = m;
      if (!special) {
        g_strow<VPL, XQ>(myB, nq, c); g_strow<VPL, XQ>(myC, nq, c);
        if (lane == 0) { g_sts4(myMinB, m); g_sts4(myMinC, m); }
      } else {
        if (nbrCta && H > 1) {   // (the last row of a sweep is never read: not published, see g_edge_take)
#pragma unroll
          for (int q = 0; q < VPL / 4; q++)
            if (q < nq) __stcg(reinterpret_cast<float4*>(pubRow + q * 4), make_float4(c[4 * q], c[4 * q + 1], c[4 * q + 2], c[4 * q + 3]));
          if (lane == 0) __stcg(pubRow + D, m);
        }
        g_strow<VPL, XQ>(nearIsB ? myC : myB, nq, c);
        if (lane == 0) g_sts4(nearIsB ? myMinC : myMinB, m);
      }
    } else if (!special) {
      // interior column: three independent recurrences, all fed from registers / shared memory
      float prB[VPL], prC[VPL];
      g_ldrow<VPL, XQ>(exLo + pbufOff + (uint32_t)(warp + ob) * runB + exOff, nq, prB);
      g_ldrow<VPL, XQ>(exLo + exPathB + pbufOff + (uint32_t)(warp + oc) * runB + exOff, nq, prC);
      const float pmB = g_lds4(minLo + mpbufOff + (uint32_t)(warp + ob) * 4);
      const float pmC = g_lds4(minLo + minPathB + mpbufOff + (uint32_t)(warp + oc) * 4);
      float mB, mC, mA;
      g_lr<VPL, NV == 1>(c, prevA, minA, (int)smd_absdiff_max3(xrow, xpA) > corDifThres, P1r, P2r, lane, lrA, mA);
      g_lr<VPL, NV == 1>(c, prB, pmB, (int)smd_absdiff_max3(xrow, xpB) > corDifThres, P1r, P2r, lane, lrB, mB);
      g_lr<VPL, NV == 1>(c, prC, pmC, (int)smd_absdiff_max3(xrow, xpC) > corDifThres, P1r, P2r, lane, lrC, mC);
      minA = mA;
      g_strow<VPL, XQ>(myB, nq, lrB); g_strow<VPL, XQ>(myC, nq, lrC);
      if (lane == 0) { g_sts4(myMinB, mB); g_sts4(myMinC, mC); }
    } else {
      // edge column.  The near diagonal comes first and goes to the neighbour CTA as soon as it exists; the far
      // diagonal's previous row was prefetched into shared memory at the end of the previous sweep row.
      float prN[VPL], lrN[VPL], lrF[VPL], mN, mF, mA;
      unsigned long long* tr = E.trace ? E.trace + ((size_t)(b * 2 + (isLast ? 1 : 0)) * H + r) * 4 : nullptr;
      if (tr && lane == 0) tr[1] = g_now();
      if (tl) tl[5] = g_now();
      const uint32_t nearPath = nearIsB ? 0u : exPathB, nearMin = nearIsB ? 0u : minPathB;
      g_ldrow<VPL, XQ>(exLo + nearPath + pbufOff + (uint32_t)(warp + nearOff) * runB + exOff, nq, prN);
      const float pmN = g_lds4(minLo + nearMin + mpbufOff + (uint32_t)(warp + nearOff) * 4);
      const uint32_t xpN = nearIsB ? xpB : xpC, xpF = nearIsB ? xpC : xpB;
      g_lr<VPL, NV == 1>(c, prN, pmN, (int)smd_absdiff_max3(xrow, xpN) > corDifThres, P1r, P2r, lane, lrN, mN);
      if (nbrCta) {
        if (r + 1 < H) {   // the neighbour reads this row at ITS row r + 1
#pragma unroll
          for (int q = 0; q < VPL / 4; q++)
            if (q < nq) __stcg(reinterpret_cast<float4*>(pubRow + q * 4), make_float4(lrN[4 * q], lrN[4 * q + 1], lrN[4 * q + 2], lrN[4 * q + 3]));
          if (lane == 0) __stcg(pubRow + D, mN);
        }
        if (tr && lane == 0) tr[0] = g_now();
        if (tl) tl[6] = g_now();
        float prF[VPL], pmF;
        int polls = 0;
        const uint32_t fb = farBuf + ((r & 1) ? (uint32_t)E.Dp * 4u : 0u);
        g_edge_take<VPL>(farRow, nq, D - d0, fb + (uint32_t)d0 * 4u, fb + (uint32_t)D * 4u, prF, pmF, polls);
        if (tr && lane == 0) { tr[2] = g_now(); tr[3] = (unsigned long long)polls; }
        if (tl) tl[7] = g_now();
        g_lr<VPL, NV == 1>(c, prevA, minA, (int)smd_absdiff_max3(xrow, xpA) > corDifThres, P1r, P2r, lane, lrA, mA);
        g_lr<VPL, NV == 1>(c, prF, pmF, (int)smd_absdiff_max3(xrow, xpF) > corDifThres, P1r, P2r, lane, lrF, mF);
      } else {   // predecessor column outside the image: Lr = C
        g_lr<VPL, NV == 1>(c, prevA, minA, (int)smd_absdiff_max3(xrow, xpA) > corDifThres, P1r, P2r, lane, lrA, mA);
#pragma unroll
        for (int j = 0; j < VPL; j++) lrF[j] = c[j];
        mF = g_rowmin<VPL>(c);
      }
      minA = mA;
      g_strow<VPL, XQ>(nearIsB ? myC : myB, nq, lrF);
      if (lane == 0) g_sts4(nearIsB ? myMinC : myMinB, mF);
#pragma unroll
      for (int j = 0; j < VPL; j++) { lrB[j] = nearIsB ? lrN[j] : lrF[j]; lrC[j] = nearIsB ? lrF[j] : lrN[j]; }
    }
#pragma unroll
    for (int j = 0; j < VPL; j++) prevA[j] = lrA[j];
    if (tl) tl[2] = g_now();
    // the far row of the NEXT sweep row: the neighbour published it early in this row period (0.25 us after the row top);
    // requested here, before the path sum, so that its L2 round trip also runs under the sum and the S store (measured
    // neutral for the float kernel -- the take's own ~0.2 us of instructions remains -- and -1 % for the uint16 one)
    if (special && nbrCta && r + 1 < H)
      g_edge_prefetch<VPL>(farNext, nq, D - d0, farBuf + (((r + 1) & 1) ? (uint32_t)E.Dp * 4u : 0u) + (uint32_t)d0 * 4u, lane);
    // ---- path sum (gen_sgm_vm: sum += L[num], within the group in reference path order)
#pragma unroll
    for (int q = 0; q < VPL / 4; q++)
      if (q < nq) {
        float4 t;
        if (NV == 1) {
          float2 h[2];
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const int j = 4 * q + 2 * e;
            const float2 la = make_float2(lrA[j], lrA[j + 1]), lb = make_float2(lrB[j], lrB[j + 1]), lc = make_float2(lrC[j], lrC[j + 1]);
            h[e] = MODE >= 1 ? __fadd2_rn(__fadd2_rn(__fadd2_rn(make_float2(s[j], s[j + 1]), la), lb), lc) : __fadd2_rn(__fadd2_rn(la, lb), lc);
          }
          t.x = h[0].x; t.y = h[0].y; t.z = h[1].x; t.w = h[1].y;
        } else {
          float* tp = &t.x;
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const int j = 4 * q + e;
            tp[e] = MODE >= 1 ? ((s[j] + lrA[j]) + lrB[j]) + lrC[j] : (lrA[j] + lrB[j]) + lrC[j];
          }
        }
        *reinterpret_cast<float4*>(o + q * 4) = t;
      }
    // advance to the next row of the sweep
    xpA = xrow;
    xpB = __shfl_sync(0xffffffffu, curB, r & 31);
    xpC = __shfl_sync(0xffffffffu, curC, r & 31);
    if ((r & 31) == 31) {
      curO = nxtO; curB = nxtB; curC = nxtC;
      loadChunk((r >> 5) + 2, nxtO, nxtB, nxtC);
    }
    xrow = __shfl_sync(0xffffffffu, curO, (r + 1) & 31);
    o += oStep;
    pubRow += E.Dp;
    if (r > 0) farRow += E.Dp;   // the far row read at row r is the neighbour's row r-1
    farNext += E.Dp;
    // Refill this stage.  No proxy fence: the stage was only READ through the generic proxy, and every lane's reads
    // have been consumed by arithmetic above (a fence.proxy.async here also waits for the warp's outstanding global
    // stores and edge loads: measured 250 ns per row, 700 ns on the edge columns).
    __syncwarp();
    if (lane == 0 && r + NS < H) issue(r + NS, slot);
    if (++slot == NS) { slot = 0; parity ^= 1u; }
    if (tl) tl[3] = g_now();
    // the rows published in shared memory become visible to the two neighbour warps (and theirs to this one);
    // even pairs first, odd pairs second, so the pairwise barriers never ripple across the CTA
    if (warp & 1) { if (hasL) g_pair_bar(barL); if (hasR) g_pair_bar(barR); }
    else          { if (hasR) g_pair_bar(barR); if (hasL) g_pair_bar(barL); }
    if (tl) tl[4] = g_now();
  }
}

// ------------------------------------------------------------------ host
template <int VPL, int UP, int NV>
static int launch_group(sm_ctx* ctx, const float* const* vol, const uint32_t* const* pix, float* const* out, int H, int W, int D,
                        int mode, int corDifThres, float redu) {
  constexpr int NS = NV == 2 ? 2 : 4;
  const int nb = min(ctx->num_sms, W / 4);       // every CTA owns >= 4 columns: first and last column are distinct warps
  const int CW = sm_div_up(W, nb);               // widest CTA (columns b*W/nb .. (b+1)*W/nb - 1)
  SM_CHECK_ARG(nb >= 1 && CW <= 14);             // 448 threads (register budget of three recurrences); 15 named barriers
  const size_t runB = (size_t)D * 4;
  const size_t stageB = runB * (mode >= 1 ? 2 : 1);
  size_t smem = (size_t)CW * NS * stageB + 2 * 2 * CW * runB + 2 * 2 * CW * 4;
  smem = (smem + 7) & ~(size_t)7;
  smem += (size_t)CW * NS * 8;
  int Dp = (D + 1 + 3) & ~3;
  smem = ((smem + 15) & ~(size_t)15) + 2 * 2 * (size_t)Dp * 4;   // prefetched far rows of the two edge columns
  SM_CHECK_ARG(smem <= 227 * 1024);
  const size_t rowsBytes = (size_t)nb * H * Dp * sizeof(float);
  void* p;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_SGMEDGE, 2 * NV * rowsBytes, &p));
  // sentinel fill ("row not published yet") only for memory that has never been armed: the consumers restore the
  // sentinel in every word they read, so a completed launch leaves the buffers armed for the next one
  if (ctx->sgm_edge_ptr != p || ctx->sgm_edge_armed < 2 * NV * rowsBytes) {
    SM_CUDA(cudaMemsetAsync(p, 0xFF, 2 * NV * rowsBytes, ctx->stream));
    ctx->sgm_edge_ptr = p;
    ctx->sgm_edge_armed = 2 * NV * rowsBytes;
  }
  sgmg_view v[2];
  for (int i = 0; i < 2; i++) {
    const int k = i < NV ? i : 0;
    v[i].vol = vol[k]; v[i].pix = pix[k]; v[i].out = out[k];
    v[i].rowsP = (float*)((uint8_t*)p + (size_t)(2 * k) * rowsBytes);
    v[i].rowsM = (float*)((uint8_t*)p + (size_t)(2 * k + 1) * rowsBytes);
  }
  unsigned long long* trace = nullptr;
#ifdef SM_SGMG_TRACE_BUILD   // diagnostics build only (scripts/sgmg_timeline.py): time stamps of the hand-offs of one launch
  const char* traceFile = NV == 1 ? getenv("SM_SGMG_TRACE") : nullptr;
#else
  const char* traceFile = nullptr;
#endif
  const size_t traceBytes = ((size_t)nb * 2 * H * 4 + 16 * 32 * 8) * sizeof(unsigned long long);
  if (traceFile) { SM_CUDA(cudaMalloc((void**)&trace, traceBytes)); SM_CUDA(cudaMemsetAsync(trace, 0, traceBytes, ctx->stream)); }
  void* args[] = {(void*)&v[0], (void*)&v[1], (void*)&H, (void*)&W, (void*)&D, (void*)&corDifThres, (void*)&redu, (void*)&Dp,
                  (void*)&trace};
  const bool full = D == 32 * VPL;
  const void* fn = full ? (mode == 0 ? (const void*)k_sgm_group<VPL, UP, 0, true, NS, NV> : (const void*)k_sgm_group<VPL, UP, 1, true, NS, NV>)
                        : (mode == 0 ? (const void*)k_sgm_group<VPL, UP, 0, false, NS, NV> : (const void*)k_sgm_group<VPL, UP, 1, false, NS, NV>);
  SM_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int perSM = 0;
  SM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, fn, CW * 32, smem));
  if (perSM * ctx->num_sms < NV * nb) return SM_ERR_UNSUPPORTED;   // the CTAs wait on each other: all must be resident
  {
    const cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(NV * nb), dim3(CW * 32), args, smem, ctx->stream);
    if (e != cudaSuccess) {
      ctx->sgm_edge_ptr = nullptr;   // state of the edge buffers unknown: fill again next time
      sm_set_error("%s:%d: cudaLaunchCooperativeKernel -> %s", __FILE__, __LINE__, cudaGetErrorString(e));
      return SM_ERR_CUDA;
    }
  }
  ctx->launches++;
  if (traceFile) {   // diagnostics only: synchronous dump of the hand-off time stamps of this launch
    std::vector<unsigned long long> h(traceBytes / 8);
    SM_CUDA(cudaStreamSynchronize(ctx->stream));
    SM_CUDA(cudaMemcpy(h.data(), trace, traceBytes, cudaMemcpyDeviceToHost));
    SM_CUDA(cudaFree(trace));
    if (FILE* f = fopen(traceFile, "wb")) { fwrite(h.data(), 1, traceBytes, f); fclose(f); }
  }
  return SM_OK;
}

static bool group_shape_ok(sm_ctx* ctx, int H, int W, int D, int mode, int nv) {
  if (!(D % 4 == 0 && D > 64 && D <= 256 && H >= 2 && W >= 8)) return false;
  // Two views per launch need two CTAs per SM, i.e. <= 72 registers per thread: runs of 4 disparities per lane fit
  // (D <= 128: 1.62 instead of 2.20 ms per view at 1280x720 D=128), runs of 8 spill and lose (6.1 vs 5.5 ms per
  // view at 1920x1080 D=256).
  if (nv == 2 && D > 128) return false;
  const int ns = nv == 2 ? 2 : 4;
  const int nb = min(ctx->num_sms, W / 4);
  const int CW = sm_div_up(W, nb);
  const size_t smem = (size_t)CW * ns * D * 4 * (mode >= 1 ? 2 : 1) + 4 * (size_t)CW * D * 4 + 16 * CW + CW * ns * 8 + 8 + 16 + 16 * (size_t)(D + 4);
  return CW <= 14 && smem <= (size_t)(nv == 2 ? 113 : 227) * 1024;
}

// The two row-wise groups of the 8-path table: UP = paths {0,4,5}, DOWN = paths {1,6,7}.  mode 0: d_sum = group sum,
// mode 1: d_sum += group sum.  Returns SM_ERR_UNSUPPORTED when the shape does not fit (caller falls back to paths).
int smi_sgm_group(sm_ctx* ctx, const float* d_vol, const uint32_t* d_pix, int H, int W, int D, int up, int mode,
                  int corDifThres, int reduCoeffi1, float* d_sum) {
  if (!group_shape_ok(ctx, H, W, D, mode, 1) || ((((uintptr_t)d_vol | (uintptr_t)d_sum) & 15) != 0)) return SM_ERR_UNSUPPORTED;
  const float redu = (float)reduCoeffi1;
  if (D <= 128) return up ? launch_group<4, 1, 1>(ctx, &d_vol, &d_pix, &d_sum, H, W, D, mode, corDifThres, redu)
                          : launch_group<4, 0, 1>(ctx, &d_vol, &d_pix, &d_sum, H, W, D, mode, corDifThres, redu);
  return up ? launch_group<8, 1, 1>(ctx, &d_vol, &d_pix, &d_sum, H, W, D, mode, corDifThres, redu)
            : launch_group<8, 0, 1>(ctx, &d_vol, &d_pix, &d_sum, H, W, D, mode, corDifThres, redu);
}

// Both views of a frame in one launch (two CTAs per SM).  SM_ERR_UNSUPPORTED: use smi_sgm_group per view.
int smi_sgm_group2(sm_ctx* ctx, const float* const* d_vol, const uint32_t* const* d_pix, int H, int W, int D, int up, int mode,
                   int corDifThres, int reduCoeffi1, float* const* d_sum) {
  if (!group_shape_ok(ctx, H, W, D, mode, 2)) return SM_ERR_UNSUPPORTED;
  for (int i = 0; i < 2; i++)
    if ((((uintptr_t)d_vol[i] | (uintptr_t)d_sum[i]) & 15) != 0) return SM_ERR_UNSUPPORTED;
  const float redu = (float)reduCoeffi1;
  if (D <= 128) return up ? launch_group<4, 1, 2>(ctx, d_vol, d_pix, d_sum, H, W, D, mode, corDifThres, redu)
                          : launch_group<4, 0, 2>(ctx, d_vol, d_pix, d_sum, H, W, D, mode, corDifThres, redu);
  return up ? launch_group<8, 1, 2>(ctx, d_vol, d_pix, d_sum, H, W, D, mode, corDifThres, redu)
            : launch_group<8, 0, 2>(ctx, d_vol, d_pix, d_sum, H, W, D, mode, corDifThres, redu);
}

// =====================================================================================================================
// Grouped sweeps on 16-bit integer cost volumes (the native form of the "Census" path, sgm_u16.cu): the same shape as
// k_sgm_group -- one cooperative launch, one CTA per SM, a warp owns a column, three recurrences per sweep row on one
// read of C and one read-modify-write of S -- in the fixed point of sgm_u16.cu (Lr * reduCoeffi1, exact), two disparities
// per packed 16x2 instruction.  Rows are D * 2 bytes; a lane's run of VPL values is NW = VPL / 2 words (one 16- or 8-byte
// access, conflict-free as it is).  Integer sums are associative, so the grouped order is bit-identical to gen_sgm_vm's.
// Full shapes only (D == 32 * VPL, VPL = 4 or 8); everything else stays on the single-path kernels.
#define SGMGU_BIG2 0x3fff3fffu   // "outside [0, D)" in both halves (sgm_u16.cu: SGMU_BIG2)

template <int NW>
__device__ __forceinline__ void gu_lds(uint32_t a, uint32_t (&x)[NW]) {
  if (NW == 4) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x[0]), "=r"(x[1]), "=r"(x[2 % NW]), "=r"(x[3 % NW]) : "r"(a) : "memory");
  else asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(x[0]), "=r"(x[1]) : "r"(a) : "memory");
}
template <int NW>
__device__ __forceinline__ void gu_sts(uint32_t a, const uint32_t (&x)[NW]) {
  if (NW == 4) asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x[0]), "r"(x[1]), "r"(x[2 % NW]), "r"(x[3 % NW]) : "memory");
  else asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x[0]), "r"(x[1]) : "memory");
}
template <int NW>
__device__ __forceinline__ void gu_stcg(void* p, const uint32_t (&x)[NW]) {
  if (NW == 4) __stcg(reinterpret_cast<uint4*>(p), make_uint4(x[0], x[1], x[2 % NW], x[3 % NW]));
  else __stcg(reinterpret_cast<uint2*>(p), make_uint2(x[0], x[1]));
}
template <int NW>
__device__ __forceinline__ void gu_ld_relaxed(const void* p, uint32_t (&x)[NW]) {
  if (NW == 4) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x[0]), "=r"(x[1]), "=r"(x[2 % NW]), "=r"(x[3 % NW]) : "l"(p) : "memory");
  else asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(x[0]), "=r"(x[1]) : "l"(p) : "memory");
}
template <int NW>
__device__ __forceinline__ bool gu_any_sentinel(const uint32_t (&x)[NW]) {
  bool s = false;
#pragma unroll
  for (int w = 0; w < NW; w++) s |= x[w] == SGMG_SENTINEL;
  return s;
}

// One pixel of one path in fixed point (u16_step of sgm_u16.cu): a = Lr' - minP, Lr = C + min(a[d], a[d-1] + P1, a[d+1] + P1, P2)
template <int NW>
__device__ __forceinline__ void gu_lr(const uint32_t (&c)[NW], const uint32_t (&pr)[NW], uint32_t minP, bool step, uint32_t scale,
                                      int lane, uint32_t (&lr)[NW], uint32_t& minNew) {
  const uint32_t P1 = (step ? 1u : scale) * 0x10001u, P2 = (step ? 3u : 3u * scale) * 0x10001u;
  const uint32_t negm = ((0x10000u - minP) & 0xffffu) * 0x10001u;
  uint32_t a[NW];
#pragma unroll
  for (int w = 0; w < NW; w++) a[w] = __vadd2(pr[w], negm);
  uint32_t lo = __shfl_up_sync(0xffffffffu, a[NW - 1], 1);
  uint32_t hi = __shfl_down_sync(0xffffffffu, a[0], 1);
  if (lane == 0) lo = SGMGU_BIG2;
  if (lane == 31) hi = SGMGU_BIG2;
  uint32_t m2 = SGMGU_BIG2;
#pragma unroll
  for (int w = 0; w < NW; w++) {
    const uint32_t pm = __byte_perm(w == 0 ? lo : a[w - 1], a[w], 0x5432);
    const uint32_t pp = __byte_perm(a[w], w == NW - 1 ? hi : a[w + 1], 0x5432);
    const uint32_t t1 = __viaddmin_u16x2(pm, P1, a[w]);
    const uint32_t t2 = __viaddmin_u16x2(pp, P1, P2);
    lr[w] = __vadd2(c[w], __vminu2(t1, t2));
    m2 = __vminu2(m2, lr[w]);
  }
  minNew = __reduce_min_sync(0xffffffffu, min(m2 & 0xffffu, m2 >> 16));
}
template <int NW>
__device__ __forceinline__ uint32_t gu_rowmin(const uint32_t (&lr)[NW]) {
  uint32_t m2 = SGMGU_BIG2;
#pragma unroll
  for (int w = 0; w < NW; w++) m2 = __vminu2(m2, lr[w]);
  return __reduce_min_sync(0xffffffffu, min(m2 & 0xffffu, m2 >> 16));
}

struct sgmgu_view {
  const uint16_t* vol;
  const uint32_t* pix;
  uint16_t* out;
  uint8_t* rowsP;   // [nb][H][pitchB]: first column of CTA b, the path whose predecessor is column u+1 (read by CTA b-1)
  uint8_t* rowsM;   // [nb][H][pitchB]: last column of CTA b, the path whose predecessor is column u-1 (read by CTA b+1)
};

template <int VPL, int UP, int MODE, int NS>
__global__ void __launch_bounds__(448, 1)
    k_sgm_group_u16(const sgmgu_view V, int H, int W, int D, int corDifThres, int shift, int pitchB) {
  extern __shared__ __align__(128) uint8_t gsm[];
  constexpr int NW = VPL / 2;
  constexpr uint32_t RUN = VPL * 2;          // bytes of a lane's run
  const int lane = threadIdx.x & 31;
  const int nwarp = blockDim.x >> 5;
  const int nb = gridDim.x, b = blockIdx.x;
  const uint16_t* __restrict__ vol = V.vol;
  const uint32_t* __restrict__ pix = V.pix;
  uint16_t* __restrict__ out = V.out;
  const int u0 = (int)(((long long)b * W) / nb);
  const int nCols = (int)(((long long)(b + 1) * W) / nb) - u0;
  if ((int)(threadIdx.x >> 5) >= nCols) return;
  const int warp = ((int)(threadIdx.x >> 5) + nCols - 2) % nCols;
  const int u = u0 + warp;
  const uint32_t runB = (uint32_t)D * 2;     // bytes of a row
  const uint32_t laneOff = (uint32_t)lane * RUN;
  const uint32_t scale = 1u << shift;
  constexpr int ob = UP ? -1 : +1, oc = -ob;

  const uint32_t base = (uint32_t)__cvta_generic_to_shared(gsm);
  const uint32_t stageB = runB * (MODE >= 1 ? 2 : 1);
  const uint32_t stLo = base + (uint32_t)warp * NS * stageB;
  const uint32_t exLo = base + (uint32_t)nwarp * NS * stageB;
  const uint32_t exBufB = (uint32_t)nwarp * runB, exPathB = 2u * exBufB;
  const uint32_t minLo = exLo + 2 * exPathB;
  const uint32_t minBufB = (uint32_t)nwarp * 4, minPathB = 2u * minBufB;
  const uint32_t bars = minLo + 2 * minPathB + (uint32_t)warp * NS * 8;
  const uint32_t farLo = (minLo + 2 * minPathB + (uint32_t)nwarp * NS * 8 + 15u) & ~15u;

  const long long rowStep = UP ? -(long long)W : (long long)W;
  const size_t p0 = (size_t)(UP ? H - 1 : 0) * W + u;
  auto issue = [&](int r, int slot) {
    const size_t p = (size_t)((long long)p0 + rowStep * r);
    const uint32_t st = stLo + slot * stageB, bar = bars + slot * 8;
    g_mbar_expect(bar, stageB);
    g_bulk(st, vol + p * D, runB, bar);
    if (MODE >= 1) g_bulk(st + runB, out + p * D, runB, bar);
  };
  if (lane == 0) {
    for (int s = 0; s < NS; s++) g_mbar_init(bars + s * 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int r = 0; r < NS && r < H; r++) issue(r, r);
  }
  __syncwarp();

  const bool isFirst = warp == 0, isLast = warp == nCols - 1;
  const bool special = isFirst || isLast;
  const int nearOff = isFirst ? +1 : -1;
  const bool nearIsB = nearOff == ob;
  const bool nbrCta = isFirst ? b > 0 : b + 1 < nb;
  uint8_t* pubRow = (isFirst ? V.rowsP : V.rowsM) + (size_t)b * H * pitchB + laneOff;
  const uint8_t* farRow = (isFirst ? V.rowsM : V.rowsP) + (size_t)(nbrCta ? (isFirst ? b - 1 : b + 1) : b) * H * pitchB + laneOff;
  const uint8_t* farNext = farRow;
  const uint32_t minOff = runB - laneOff;   // from a lane's run to the row minimum's word
  const uint32_t farBuf = farLo + (isLast ? 2u : 0u) * (uint32_t)pitchB;

  const int ub = min(max(u + ob, 0), W - 1) - u, uc = min(max(u + oc, 0), W - 1) - u;
  uint32_t curO = 0, curB = 0, curC = 0, nxtO = 0, nxtB = 0, nxtC = 0;
  auto loadChunk = [&](int chunk, uint32_t& o_, uint32_t& b_, uint32_t& c_) {
    const int rr = chunk * 32 + lane;
    if (rr < H) {
      const uint32_t* q = pix + (size_t)((long long)p0 + rowStep * rr);
      o_ = q[0]; b_ = q[ub]; c_ = q[uc];
    }
  };
  loadChunk(0, curO, curB, curC);
  loadChunk(1, nxtO, nxtB, nxtC);
  uint32_t xrow = __shfl_sync(0xffffffffu, curO, 0), xpA = 0, xpB = 0, xpC = 0;
  uint16_t* o = out + p0 * D + lane * VPL;
  const long long oStep = rowStep * D;

  uint32_t prevA[NW], minA = 0;
  int slot = 0;
  uint32_t parity = 0;
  const int barL = warp, barR = warp + 1;
  const bool hasL = warp > 0, hasR = warp + 1 < nCols;

  for (int r = 0; r < H; r++) {
    const uint32_t bufOff = (r & 1) ? exBufB : 0u, pbufOff = exBufB - bufOff;
    const uint32_t mbufOff = (r & 1) ? minBufB : 0u, mpbufOff = minBufB - mbufOff;
    uint32_t c[NW], s[NW], lrA[NW], lrB[NW], lrC[NW];
    const uint32_t st = stLo + slot * stageB + laneOff;
    g_mbar_wait(bars + slot * 8, parity);
    gu_lds<NW>(st, c);
#pragma unroll
    for (int w = 0; w < NW; w++) c[w] <<= shift;      // raw cost -> fixed point (halves stay below 2^14: smi_sgm_u16_ok)
    if (MODE >= 1) gu_lds<NW>(st + runB, s);
    const uint32_t myB = exLo + bufOff + (uint32_t)warp * runB + laneOff, myC = myB + exPathB;
    const uint32_t myMinB = minLo + mbufOff + (uint32_t)warp * 4, myMinC = myMinB + minPathB;
    if (r == 0) {
      const uint32_t m = gu_rowmin<NW>(c);
#pragma unroll
      for (int w = 0; w < NW; w++) lrA[w] = lrB[w] = lrC[w] = c[w];
      minA = m;
      if (!special) {
        gu_sts<NW>(myB, c); gu_sts<NW>(myC, c);
        if (lane == 0) { g_sts4(myMinB, __uint_as_float(m)); g_sts4(myMinC, __uint_as_float(m)); }
      } else {
        if (nbrCta && H > 1) {
          gu_stcg<NW>(pubRow, c);
          if (lane == 0) __stcg(reinterpret_cast<uint32_t*>(pubRow + minOff), m);
        }
        gu_sts<NW>(nearIsB ? myC : myB, c);
        if (lane == 0) g_sts4(nearIsB ? myMinC : myMinB, __uint_as_float(m));
      }
    } else if (!special) {
      uint32_t prB[NW], prC[NW];
      gu_lds<NW>(exLo + pbufOff + (uint32_t)(warp + ob) * runB + laneOff, prB);
      gu_lds<NW>(exLo + exPathB + pbufOff + (uint32_t)(warp + oc) * runB + laneOff, prC);
      const uint32_t pmB = __float_as_uint(g_lds4(minLo + mpbufOff + (uint32_t)(warp + ob) * 4));
      const uint32_t pmC = __float_as_uint(g_lds4(minLo + minPathB + mpbufOff + (uint32_t)(warp + oc) * 4));
      uint32_t mB, mC, mA;
      gu_lr<NW>(c, prevA, minA, (int)smd_absdiff_max3(xrow, xpA) > corDifThres, scale, lane, lrA, mA);
      gu_lr<NW>(c, prB, pmB, (int)smd_absdiff_max3(xrow, xpB) > corDifThres, scale, lane, lrB, mB);
      gu_lr<NW>(c, prC, pmC, (int)smd_absdiff_max3(xrow, xpC) > corDifThres, scale, lane, lrC, mC);
      minA = mA;
      gu_sts<NW>(myB, lrB); gu_sts<NW>(myC, lrC);
      if (lane == 0) { g_sts4(myMinB, __uint_as_float(mB)); g_sts4(myMinC, __uint_as_float(mC)); }
    } else {
      uint32_t prN[NW], lrN[NW], lrF[NW], mN, mF, mA;
      const uint32_t nearPath = nearIsB ? 0u : exPathB, nearMin = nearIsB ? 0u : minPathB;
      gu_lds<NW>(exLo + nearPath + pbufOff + (uint32_t)(warp + nearOff) * runB + laneOff, prN);
      const uint32_t pmN = __float_as_uint(g_lds4(minLo + nearMin + mpbufOff + (uint32_t)(warp + nearOff) * 4));
      const uint32_t xpN = nearIsB ? xpB : xpC, xpF = nearIsB ? xpC : xpB;
      gu_lr<NW>(c, prN, pmN, (int)smd_absdiff_max3(xrow, xpN) > corDifThres, scale, lane, lrN, mN);
      if (nbrCta) {
        if (r + 1 < H) {
          gu_stcg<NW>(pubRow, lrN);
          if (lane == 0) __stcg(reinterpret_cast<uint32_t*>(pubRow + minOff), mN);
        }
        // the far row: prefetched into shared memory at the end of the previous sweep row; words still carrying the
        // sentinel were fetched too early and are polled; every word read is re-armed
        uint32_t prF[NW];
        const uint32_t fb = farBuf + ((r & 1) ? (uint32_t)pitchB : 0u);
        g_cpasync_wait0();
        __syncwarp();
        gu_lds<NW>(fb + laneOff, prF);
        while (gu_any_sentinel<NW>(prF)) gu_ld_relaxed<NW>(farRow, prF);
        {
          uint32_t arm[NW];
#pragma unroll
          for (int w = 0; w < NW; w++) arm[w] = SGMG_SENTINEL;
          gu_stcg<NW>(const_cast<uint8_t*>(farRow), arm);
        }
        uint32_t pmF = __float_as_uint(g_lds4(fb + runB));
        while (__any_sync(0xffffffffu, pmF == SGMG_SENTINEL)) {
          if (pmF == SGMG_SENTINEL) pmF = g_ld_relaxed4(farRow + minOff);
        }
        if (lane == 0) __stcg(reinterpret_cast<uint32_t*>(const_cast<uint8_t*>(farRow) + minOff), SGMG_SENTINEL);
        gu_lr<NW>(c, prevA, minA, (int)smd_absdiff_max3(xrow, xpA) > corDifThres, scale, lane, lrA, mA);
        gu_lr<NW>(c, prF, pmF, (int)smd_absdiff_max3(xrow, xpF) > corDifThres, scale, lane, lrF, mF);
      } else {
        gu_lr<NW>(c, prevA, minA, (int)smd_absdiff_max3(xrow, xpA) > corDifThres, scale, lane, lrA, mA);
#pragma unroll
        for (int w = 0; w < NW; w++) lrF[w] = c[w];
        mF = gu_rowmin<NW>(c);
      }
      minA = mA;
      gu_sts<NW>(nearIsB ? myC : myB, lrF);
      if (lane == 0) g_sts4(nearIsB ? myMinC : myMinB, __uint_as_float(mF));
#pragma unroll
      for (int w = 0; w < NW; w++) { lrB[w] = nearIsB ? lrN[w] : lrF[w]; lrC[w] = nearIsB ? lrF[w] : lrN[w]; }
    }
#pragma unroll
    for (int w = 0; w < NW; w++) prevA[w] = lrA[w];
    if (special && nbrCta && r + 1 < H) {
      const uint32_t dst = farBuf + (((r + 1) & 1) ? (uint32_t)pitchB : 0u);
      if (NW == 4) g_cpasync16(dst + laneOff, farNext);
      else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + laneOff), "l"(farNext) : "memory");
      if (lane == 0) g_cpasync16(dst + runB, farNext + runB);   // the quad that holds the row minimum
      g_cpasync_commit();
    }
    // ---- path sum (integer: any order is the reference's sum)
    {
      uint32_t t[NW];
#pragma unroll
      for (int w = 0; w < NW; w++) {
        t[w] = __vadd2(__vadd2(lrA[w], lrB[w]), lrC[w]);
        if (MODE >= 1) t[w] = __vadd2(t[w], s[w]);
      }
      if (NW == 4) *reinterpret_cast<uint4*>(o) = make_uint4(t[0], t[1], t[2 % NW], t[3 % NW]);
      else *reinterpret_cast<uint2*>(o) = make_uint2(t[0], t[1]);
    }
    xpA = xrow;
    xpB = __shfl_sync(0xffffffffu, curB, r & 31);
    xpC = __shfl_sync(0xffffffffu, curC, r & 31);
    if ((r & 31) == 31) {
      curO = nxtO; curB = nxtB; curC = nxtC;
      loadChunk((r >> 5) + 2, nxtO, nxtB, nxtC);
    }
    xrow = __shfl_sync(0xffffffffu, curO, (r + 1) & 31);
    o += oStep;
    pubRow += pitchB;
    if (r > 0) farRow += pitchB;
    farNext += pitchB;
    __syncwarp();
    if (lane == 0 && r + NS < H) issue(r + NS, slot);
    if (++slot == NS) { slot = 0; parity ^= 1u; }
    if (warp & 1) { if (hasL) g_pair_bar(barL); if (hasR) g_pair_bar(barR); }
    else          { if (hasR) g_pair_bar(barR); if (hasL) g_pair_bar(barL); }
  }
}

template <int VPL, int UP>
static int launch_group_u16(sm_ctx* ctx, const uint16_t* vol, const uint32_t* pix, uint16_t* out, int H, int W, int D, int mode,
                            int corDifThres, int shift) {
  constexpr int NS = 4;
  const int nb = min(ctx->num_sms, W / 4);
  const int CW = sm_div_up(W, nb);
  SM_CHECK_ARG(nb >= 1 && CW <= 14);
  const size_t runB = (size_t)D * 2;
  const size_t stageB = runB * (mode >= 1 ? 2 : 1);
  size_t smem = (size_t)CW * NS * stageB + 2 * 2 * CW * runB + 2 * 2 * CW * 4;
  smem = (smem + 7) & ~(size_t)7;
  smem += (size_t)CW * NS * 8;
  int pitchB = (int)runB + 16;                                  // D values + the row minimum's quad
  smem = ((smem + 15) & ~(size_t)15) + 2 * 2 * (size_t)pitchB;   // prefetched far rows of the two edge columns
  SM_CHECK_ARG(smem <= 227 * 1024);
  const size_t rowsBytes = (size_t)nb * H * pitchB;
  void* p;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_SGMEDGE, 2 * rowsBytes, &p));
  if (ctx->sgm_edge_ptr != p || ctx->sgm_edge_armed < 2 * rowsBytes) {
    SM_CUDA(cudaMemsetAsync(p, 0xFF, 2 * rowsBytes, ctx->stream));
    ctx->sgm_edge_ptr = p;
    ctx->sgm_edge_armed = 2 * rowsBytes;
  }
  sgmgu_view v{vol, pix, out, (uint8_t*)p, (uint8_t*)p + rowsBytes};
  void* args[] = {(void*)&v, (void*)&H, (void*)&W, (void*)&D, (void*)&corDifThres, (void*)&shift, (void*)&pitchB};
  const void* fn = mode == 0 ? (const void*)k_sgm_group_u16<VPL, UP, 0, NS> : (const void*)k_sgm_group_u16<VPL, UP, 1, NS>;
  SM_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int perSM = 0;
  SM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, fn, CW * 32, smem));
  if (perSM * ctx->num_sms < nb) return SM_ERR_UNSUPPORTED;
  const cudaError_t e = cudaLaunchCooperativeKernel(fn, dim3(nb), dim3(CW * 32), args, smem, ctx->stream);
  if (e != cudaSuccess) {
    ctx->sgm_edge_ptr = nullptr;
    sm_set_error("%s:%d: cudaLaunchCooperativeKernel -> %s", __FILE__, __LINE__, cudaGetErrorString(e));
    return SM_ERR_CUDA;
  }
  ctx->launches++;
  return SM_OK;
}

// up = 1: paths {0, 4, 5}; up = 0: paths {1, 6, 7}.  mode 0: d_sum = group sum, 1: d_sum += group sum (fixed point, as
// smi_sgm_path_u16).  SM_ERR_UNSUPPORTED when the shape does not fit: the caller runs the three single-path sweeps.
int smi_sgm_group_u16(sm_ctx* ctx, const uint16_t* d_vol, const uint32_t* d_pix, int H, int W, int D, int up, int mode,
                      int corDifThres, int reduCoeffi1, uint16_t* d_sum) {
  if (!(D == 256 || D == 128) || H < 2 || W < 8 || ((((uintptr_t)d_vol | (uintptr_t)d_sum) & 15) != 0)) return SM_ERR_UNSUPPORTED;
  const int nb = min(ctx->num_sms, W / 4);
  if (sm_div_up(W, nb) > 14) return SM_ERR_UNSUPPORTED;
  int sh = 0;
  while ((1 << sh) < reduCoeffi1) sh++;
  if (D == 256) return up ? launch_group_u16<8, 1>(ctx, d_vol, d_pix, d_sum, H, W, D, mode, corDifThres, sh)
                          : launch_group_u16<8, 0>(ctx, d_vol, d_pix, d_sum, H, W, D, mode, corDifThres, sh);
  return up ? launch_group_u16<4, 1>(ctx, d_vol, d_pix, d_sum, H, W, D, mode, corDifThres, sh)
            : launch_group_u16<4, 0>(ctx, d_vol, d_pix, d_sum, H, W, D, mode, corDifThres, sh);
}
