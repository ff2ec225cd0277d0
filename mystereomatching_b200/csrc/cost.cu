// K2: cost volume.  One pass writes vol[v][u][d] (float32, d fastest); the AD and
// census volumes of the reference (ADCensusCal's four H*W*D temporaries,
// stereoMatching.cpp:899-904) are never materialised.
//
//   AD      gen_ad_sd_vm, AOS=0      stereoMatching.cpp:2468-2509
//   Hamming gen_cenVM_XOR            stereoMatching.h:936-981
//   combine gen_vm_from2vm_exp       stereoMatching.cpp:3566-3590
//
// Bit-exactness of the exp-robust combination: AD takes only the 767 values
// min((float)k/3, trunc), k = sum_c|l_c-r_c| in 0..765, plus `trunc` for
// out-of-range; the Hamming count takes codeLength+1 values.  The host builds
// the two tables exp(-ad/lamAD), exp(-c/lamCen) with the SAME float libm call the
// reference makes, and the kernel evaluates (2 - tabAD[k]) - tabCen[c] -- the
// reference's left-to-right float subtraction -- so the volume equals the CPU
// result bit for bit, and no transcendental runs on the device.
//
// Kernel shape (HBM-write bound, 4 B per element): persistent CTAs (one per SM, 1024 threads).  Each CTA keeps
// lane-replicated copies of both tables in shared memory ([entry][32 lanes] -> lookups are bank-conflict free
// by construction) and walks (row, 512-pixel segment) work items: the "other" image's census words and packed
// pixel for the D-1+512 positions the segment can match against are staged in shared memory once, then every
// warp takes anchor pixels and its lanes run along d, so each warp store is one fully coalesced 128-byte line.
// Per element the steady state is: one 16-byte staged load (census words + pixel of the partner), the 71-bit Hamming
// count as 2 POPC (3-input majority identity), the 3-channel SAD as one VABSDIFF4.U8.ACC, two table loads, one
// subtraction, one store; the view's sign and the census word count are template parameters, so staged entries and
// table slots are addressed as pointer + immediate.
// Disparities whose partner lies outside the image form a contiguous tail d >= nvalid of every pixel and get
// the constant out-of-range value without touching the tables per element.
#include <math.h>

#include "common.cuh"

#define COST_THREADS 1024
#define COST_SEG 512
#define COST_TAB_AD 767   // k = 0..765, slot 766 = out of range
#define COST_MAX_CODE 71

enum { COST_ADCENSUS = 0, COST_HAMMING_F32 = 1, COST_AD_F32 = 2, COST_HAMMING_U16 = 3 };

// One staged entry per position of the "other" image: {census word 0 lo, hi, census word 1 (low 32 bits), packed pixel}
// -> one LDS.128 per element.  eb points at the entry of d = lane of the current 32-chunk; the entries of the
// following chunks sit at compile-time offsets -SGN*32*q, so the steady state does no index arithmetic at all.
// t1 / t2 are per-lane byte pointers into the lane-replicated tables (base + lane*4): an entry is 128 bytes.
template <int MODE, int NW, typename OutT>
__device__ __forceinline__ OutT cost_one(const uint4 e, uint32_t pa, uint32_t ca0lo, uint32_t ca0hi, uint32_t ca1,
                                         const char* __restrict__ t1, const char* __restrict__ t2) {
  constexpr bool NEED_CEN = MODE != COST_AD_F32, NEED_AD = MODE == COST_ADCENSUS || MODE == COST_AD_F32;
  int c = 0;
  uint32_t off2 = 0;
  if (NEED_CEN) {
    const uint32_t a0 = ca0lo ^ e.x, a1 = ca0hi ^ e.y;
    if (NW == 2) {
      const uint32_t a2 = ca1 ^ e.z;
      // popc(a)+popc(b)+popc(c) = popc(a^b^c) + 2*popc(maj(a,b,c)): 2 POPC for 71 bits
      const int p1 = __popc(a0 ^ a1 ^ a2), p2 = __popc((a0 & a1) | (a0 & a2) | (a1 & a2));
      c = p1 + 2 * p2;
      off2 = (uint32_t)p1 * 128u + (uint32_t)p2 * 256u;
    } else {
      c = __popc(a0) + __popc(a1);
      off2 = (uint32_t)c * 128u;
    }
    // min(count, codeLength) of gen_cenVM_XOR is the identity: a code has codeLength bits (truncRat = 1)
  }
  if (MODE == COST_ADCENSUS || MODE == COST_AD_F32) {
    const uint32_t k = __vsadu4(pa, e.w);   // sum_c |l_c - r_c| (byte 3 is zero in both words)
    const float v1 = *reinterpret_cast<const float*>(t1 + k * 128u);
    if (MODE == COST_AD_F32) return (OutT)v1;
    return (OutT)(v1 - *reinterpret_cast<const float*>(t2 + off2));   // (2 - e^-ad/l) - e^-c/l
  }
  return (OutT)c;
}

template <int MODE, int NW, int SGN, typename OutT>
__global__ void __launch_bounds__(COST_THREADS, 1)
    k_cost(const uint32_t* __restrict__ pixA, const uint32_t* __restrict__ pixO, const uint64_t* __restrict__ cenA,
           const uint64_t* __restrict__ cenO, int H, int W, int D, int codeLen,
           const float* __restrict__ tabAD, const float* __restrict__ tabCen, OutT* __restrict__ vol) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr bool NEED_CEN = MODE != COST_AD_F32, NEED_AD = MODE == COST_ADCENSUS || MODE == COST_AD_F32;
  float* sT1 = reinterpret_cast<float*>(smem_raw);                          // [767][32]: 2 - tabAD (or raw AD)
  float* sTabCen = sT1 + (NEED_AD ? COST_TAB_AD * 32 : 0);                  // [72][32]
  uint4* sEnt = reinterpret_cast<uint4*>(sTabCen + (MODE == COST_ADCENSUS ? (COST_MAX_CODE + 1) * 32 : 0));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (NEED_AD)
    for (int i = tid; i < COST_TAB_AD * 32; i += COST_THREADS)
      sT1[i] = MODE == COST_ADCENSUS ? 2.0f - tabAD[i >> 5] : tabAD[i >> 5];   // first operation of "2 - a - b"
  if (MODE == COST_ADCENSUS)
    for (int i = tid; i < (codeLen + 1) * 32; i += COST_THREADS) sTabCen[i] = tabCen[i >> 5];
  // value of every out-of-range disparity (gen_ad_sd_vm: trunc; gen_cenVM_XOR: codeLength)
  OutT oor;
  if (MODE == COST_ADCENSUS) oor = (OutT)((2.0f - tabAD[766]) - tabCen[codeLen]);
  else if (MODE == COST_AD_F32) oor = (OutT)tabAD[766];
  else oor = (OutT)codeLen;
  const char* t1 = reinterpret_cast<const char*>(sT1 + lane);
  const char* t2 = reinterpret_cast<const char*>(sTabCen + lane);

  const int nSeg = (W + COST_SEG - 1) / COST_SEG;
  const int nItems = H * nSeg;
  const int nd = (D + 31) >> 5;
  for (int item = blockIdx.x; item < nItems; item += gridDim.x) {
    const int v = item / nSeg, ua = (item - v * nSeg) * COST_SEG;
    const int nA = min(COST_SEG, W - ua);
    const int elo = SGN > 0 ? ua - (D - 1) : ua;
    const int cnt = nA + D - 1;
    __syncthreads();  // previous item's readers are done (also orders the table fill)
    for (int i = tid; i < cnt; i += COST_THREADS) {
      const int e = elo + i;
      uint4 ent = make_uint4(0u, 0u, 0u, 0u);
      if (e >= 0 && e < W) {
        const size_t p = (size_t)v * W + e;
        if (NEED_AD) ent.w = pixO[p] & 0x00FFFFFFu;
        if (NEED_CEN) {
          const uint64_t c0 = cenO[p * NW];
          ent.x = (uint32_t)c0; ent.y = (uint32_t)(c0 >> 32);
          if (NW == 2) ent.z = (uint32_t)cenO[p * NW + 1];
        }
      }
      sEnt[i] = ent;
    }
    __syncthreads();
    // the anchor's own words are fetched one anchor ahead, so the warp never waits on global memory between anchors
    uint32_t pa_n = 0, ca1_n = 0;
    uint64_t ca0_n = 0;
    if (warp < nA) {
      const size_t p0 = (size_t)v * W + ua + warp;
      if (NEED_AD) pa_n = pixA[p0];
      if (NEED_CEN) {
        ca0_n = cenA[p0 * NW];
        if (NW == 2) ca1_n = (uint32_t)cenA[p0 * NW + 1];
      }
    }
    for (int a = warp; a < nA; a += COST_THREADS / 32) {
      const int u = ua + a;
      const size_t p = (size_t)v * W + u;
      const uint32_t pa = pa_n & 0x00FFFFFFu;
      const uint32_t ca0lo = (uint32_t)ca0_n, ca0hi = (uint32_t)(ca0_n >> 32), ca1 = ca1_n;
      if (a + COST_THREADS / 32 < nA) {
        const size_t pn = p + COST_THREADS / 32;
        if (NEED_AD) pa_n = pixA[pn];
        if (NEED_CEN) {
          ca0_n = cenA[pn * NW];
          if (NW == 2) ca1_n = (uint32_t)cenA[pn * NW + 1];
        }
      }
      OutT* out = vol + p * D + lane;
      // in-range disparities of this pixel: view 0 (SGN +1): u - d >= 0; view 1: u + d < W
      const int nvalid = min(D, SGN > 0 ? u + 1 : W - u);
      const uint4* eb = sEnt + (u - elo - SGN * lane);   // entry of d = lane; d += 32 moves it by -SGN*32
      int j = 0;
      // chunks of 32 disparities that are entirely in range: no predicate, 4 at a time
      for (; (j + 4) * 32 <= nvalid; j += 4) {
#pragma unroll
        for (int q = 0; q < 4; q++)
          out[q * 32] = cost_one<MODE, NW, OutT>(eb[-SGN * q * 32], pa, ca0lo, ca0hi, ca1, t1, t2);
        eb -= SGN * 128;
        out += 128;
      }
      for (; j < nd; j++) {
        const int d = lane + j * 32;
        if (d >= D) break;
        OutT r = oor;                              // chunk (partly) out of range: constant tail
        if (d < nvalid) r = cost_one<MODE, NW, OutT>(eb[0], pa, ca0lo, ca0hi, ca1, t1, t2);
        out[0] = r;
        eb -= SGN * 32;
        out += 32;
      }
    }
  }
}

// Host tables.  MODE ADCENSUS: exp(-ad/lamAD), exp(-c/lamCen).  They are cached in
// the ctx by parameter value, so a stream of frames uploads them once.

int smi_exp_tables(sm_ctx* ctx, float trunc, float lamAD, float lamCen, int codeLen, const float** d_tabAD,
                   const float** d_tabCen) {
  void* p;
  float* h_tab = ctx->h_tab;
  const size_t tab_bytes = sizeof(float) * (COST_TAB_AD + COST_MAX_CODE + 1);
  SM_TRY(sm_scratch_get(ctx, SM_SCR_TAB, tab_bytes, &p));
  float* d = (float*)p;
  if (ctx->tab_trunc != trunc || ctx->tab_lamAD != lamAD || ctx->tab_lamCen != lamCen || ctx->tab_codeLen != codeLen) {
    SM_CUDA(cudaStreamSynchronize(ctx->stream));  // h_tab may still be the source of a queued copy
    for (int k = 0; k <= 765; k++) {
      float ad = fminf((float)k / 3, trunc);  // sum / channels, then min(., trunc)
      h_tab[k] = lamAD > 0 ? expf(-ad / lamAD) : ad;  // lamAD<=0: raw AD table (sm_cost_ad)
    }
    h_tab[766] = lamAD > 0 ? expf(-trunc / lamAD) : trunc;
    for (int c = 0; c <= COST_MAX_CODE; c++) h_tab[COST_TAB_AD + c] = expf(-(float)c / lamCen);
    SM_CUDA(cudaMemcpyAsync(d, h_tab, tab_bytes, cudaMemcpyHostToDevice, ctx->stream));
    SM_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->tab_trunc = trunc; ctx->tab_lamAD = lamAD; ctx->tab_lamCen = lamCen; ctx->tab_codeLen = codeLen;
  }
  *d_tabAD = d;
  *d_tabCen = d + COST_TAB_AD;
  return SM_OK;
}

template <int MODE, int NW, int SGN, typename OutT>
static int launch_cost3(sm_ctx* ctx, const uint32_t* pixA, const uint32_t* pixO, const uint64_t* cenA, const uint64_t* cenO,
                        int H, int W, int D, int codeLen, const float* tAD, const float* tCen, OutT* vol) {
  size_t smem = 0;
  if (MODE == COST_ADCENSUS || MODE == COST_AD_F32) smem += COST_TAB_AD * 32 * sizeof(float);
  if (MODE == COST_ADCENSUS) smem += (COST_MAX_CODE + 1) * 32 * sizeof(float);
  smem += (size_t)(COST_SEG + D - 1) * sizeof(uint4);
  SM_CUDA(cudaFuncSetAttribute(k_cost<MODE, NW, SGN, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int nItems = H * sm_div_up(W, COST_SEG);
  int grid = min(nItems, ctx->num_sms);
  SM_LAUNCH(ctx, (k_cost<MODE, NW, SGN, OutT>), grid, COST_THREADS, smem, pixA, pixO, cenA, cenO, H, W, D, codeLen, tAD,
            tCen, vol);
  return SM_OK;
}

// the word count (1: 63-bit census, 2: 71-bit) and the view's sign are compile-time in the kernel
template <int MODE, typename OutT>
static int launch_cost(sm_ctx* ctx, const uint32_t* pixA, const uint32_t* pixO, const uint64_t* cenA,
                       const uint64_t* cenO, int nw, int H, int W, int D, int sgn, int codeLen, const float* tAD,
                       const float* tCen, OutT* vol) {
  if (nw == 2)
    return sgn > 0 ? launch_cost3<MODE, 2, +1, OutT>(ctx, pixA, pixO, cenA, cenO, H, W, D, codeLen, tAD, tCen, vol)
                   : launch_cost3<MODE, 2, -1, OutT>(ctx, pixA, pixO, cenA, cenO, H, W, D, codeLen, tAD, tCen, vol);
  return sgn > 0 ? launch_cost3<MODE, 1, +1, OutT>(ctx, pixA, pixO, cenA, cenO, H, W, D, codeLen, tAD, tCen, vol)
                 : launch_cost3<MODE, 1, -1, OutT>(ctx, pixA, pixO, cenA, cenO, H, W, D, codeLen, tAD, tCen, vol);
}

int smi_cost_adcensus_packed(sm_ctx* ctx, const uint32_t* d_pixL, const uint32_t* d_pixR, const uint64_t* d_cenL,
                             const uint64_t* d_cenR, int H, int W, int D, int func, float adTrunc, float lamAD,
                             float lamCen, int LOR, float* d_vol) {
  const int codeLen = sm_census_code_length(func), nw = sm_census_words(func);
  const float *tAD, *tCen;
  SM_TRY(smi_exp_tables(ctx, adTrunc, lamAD, lamCen, codeLen, &tAD, &tCen));
  // LOR 0: anchor = left at u, other = right at u-d.  LOR 1: anchor = right at u, other = left at u+d.
  if (LOR == 0)
    return launch_cost<COST_ADCENSUS, float>(ctx, d_pixL, d_pixR, d_cenL, d_cenR, nw, H, W, D, +1, codeLen, tAD, tCen,
                                             d_vol);
  return launch_cost<COST_ADCENSUS, float>(ctx, d_pixR, d_pixL, d_cenR, d_cenL, nw, H, W, D, -1, codeLen, tAD, tCen,
                                           d_vol);
}

static int check_vol_args(sm_ctx* ctx, int H, int W, int D, int LOR) {
  SM_CHECK_ARG(ctx != nullptr);
  SM_CHECK_ARG(H > 0 && W > 0 && D > 0 && D <= 512);
  SM_CHECK_ARG(LOR == 0 || LOR == 1);
  return SM_OK;
}

extern "C" int sm_cost_adcensus(sm_ctx* ctx, const uint8_t* d_bgrL, const uint8_t* d_bgrR, const uint64_t* d_cenL,
                                const uint64_t* d_cenR, int H, int W, int D, int func, float adTrunc, float lamAD,
                                float lamCen, int LOR, float* d_vol) {
  SM_TRY(check_vol_args(ctx, H, W, D, LOR));
  SM_CHECK_ARG(d_bgrL && d_bgrR && d_cenL && d_cenR && d_vol);
  SM_CHECK_ARG(func == 0 || func == 3);
  SM_CHECK_ARG(lamAD > 0 && lamCen > 0);
  const long long npix = (long long)H * W;
  void *pl, *pr;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_IMG0, npix * 4, &pl));
  SM_TRY(sm_scratch_get(ctx, SM_SCR_IMG1, npix * 4, &pr));
  SM_TRY(smi_pack_bgr(ctx, d_bgrL, npix, (uint32_t*)pl));
  SM_TRY(smi_pack_bgr(ctx, d_bgrR, npix, (uint32_t*)pr));
  return smi_cost_adcensus_packed(ctx, (uint32_t*)pl, (uint32_t*)pr, d_cenL, d_cenR, H, W, D, func, adTrunc, lamAD,
                                  lamCen, LOR, d_vol);
}

extern "C" int sm_cost_hamming(sm_ctx* ctx, const uint64_t* d_cenL, const uint64_t* d_cenR, int H, int W, int D,
                               int func, int LOR, float* d_vol) {
  SM_TRY(check_vol_args(ctx, H, W, D, LOR));
  SM_CHECK_ARG(d_cenL && d_cenR && d_vol && (func == 0 || func == 3));
  const int codeLen = sm_census_code_length(func), nw = sm_census_words(func);
  if (LOR == 0)
    return launch_cost<COST_HAMMING_F32, float>(ctx, nullptr, nullptr, d_cenL, d_cenR, nw, H, W, D, +1, codeLen,
                                                nullptr, nullptr, d_vol);
  return launch_cost<COST_HAMMING_F32, float>(ctx, nullptr, nullptr, d_cenR, d_cenL, nw, H, W, D, -1, codeLen, nullptr,
                                              nullptr, d_vol);
}

extern "C" int sm_cost_hamming_u16(sm_ctx* ctx, const uint64_t* d_cenL, const uint64_t* d_cenR, int H, int W, int D,
                                   int func, int LOR, uint16_t* d_vol) {
  SM_TRY(check_vol_args(ctx, H, W, D, LOR));
  SM_CHECK_ARG(d_cenL && d_cenR && d_vol && (func == 0 || func == 3));
  const int codeLen = sm_census_code_length(func), nw = sm_census_words(func);
  if (LOR == 0)
    return launch_cost<COST_HAMMING_U16, uint16_t>(ctx, nullptr, nullptr, d_cenL, d_cenR, nw, H, W, D, +1, codeLen,
                                                   nullptr, nullptr, d_vol);
  return launch_cost<COST_HAMMING_U16, uint16_t>(ctx, nullptr, nullptr, d_cenR, d_cenL, nw, H, W, D, -1, codeLen,
                                                 nullptr, nullptr, d_vol);
}

extern "C" int sm_cost_ad(sm_ctx* ctx, const uint8_t* d_bgrL, const uint8_t* d_bgrR, int H, int W, int D, int LOR,
                          float trunc, float* d_vol) {
  SM_TRY(check_vol_args(ctx, H, W, D, LOR));
  SM_CHECK_ARG(d_bgrL && d_bgrR && d_vol);
  const long long npix = (long long)H * W;
  void *pl, *pr;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_IMG0, npix * 4, &pl));
  SM_TRY(sm_scratch_get(ctx, SM_SCR_IMG1, npix * 4, &pr));
  SM_TRY(smi_pack_bgr(ctx, d_bgrL, npix, (uint32_t*)pl));
  SM_TRY(smi_pack_bgr(ctx, d_bgrR, npix, (uint32_t*)pr));
  const float *tAD, *tCen;
  SM_TRY(smi_exp_tables(ctx, trunc, -1.f, 1.f, 0, &tAD, &tCen));  // lamAD<0: table holds the raw AD values
  if (LOR == 0)
    return launch_cost<COST_AD_F32, float>(ctx, (uint32_t*)pl, (uint32_t*)pr, nullptr, nullptr, 1, H, W, D, +1, 0, tAD,
                                           tCen, d_vol);
  return launch_cost<COST_AD_F32, float>(ctx, (uint32_t*)pr, (uint32_t*)pl, nullptr, nullptr, 1, H, W, D, -1, 0, tAD,
                                         tCen, d_vol);
}

__global__ void k_combine_exp(const float* __restrict__ a, const float* __restrict__ b, size_t n, float l0, float l1,
                              float* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  // gen_vm_from2vm_exp on two materialised volumes (stereoMatching.cpp:3566-3590): expf as the host libm evaluates it
  // (smd_expf_host) and the same two subtractions -> bit-exact.  The pipeline uses the fused table kernel above.
  for (; i < n; i += stride)
    out[i] = __fsub_rn(__fsub_rn(2.0f, smd_expf_host(__fdiv_rn(-a[i], l0))), smd_expf_host(__fdiv_rn(-b[i], l1)));
}

extern "C" int sm_combine_exp(sm_ctx* ctx, const float* d_vm0, const float* d_vm1, size_t n, float aru0, float aru1,
                              float* d_out) {
  SM_CHECK_ARG(ctx && d_vm0 && d_vm1 && d_out);
  int grid = (int)min((size_t)ctx->num_sms * 16, (n + 255) / 256);
  SM_LAUNCH(ctx, k_combine_exp, grid, 256, 0, d_vm0, d_vm1, n, aru0, aru1, d_out);
  return SM_OK;
}
