// Shared declarations of the sm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/sm_b200.h"

#ifndef __CUDA_ARCH_LIST__
#define __CUDA_ARCH_LIST__ 1000
#endif

#define SM_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// A grow-only device scratch slot: stage calls that need small per-frame
// temporaries (packed images, packed arms, exp tables) take them from here so
// no stage allocates on the hot path after the first frame.
struct sm_scratch {
  void* p = nullptr;
  size_t cap = 0;
};

enum { SM_SCR_IMG0 = 0, SM_SCR_IMG1, SM_SCR_ARM0, SM_SCR_ARM1, SM_SCR_TAB, SM_SCR_MISC0, SM_SCR_MISC1,
       SM_SCR_MISC2, SM_SCR_MISC3, SM_SCR_MISC4, SM_SCR_MISC5, SM_SCR_NLWORK, SM_SCR_SGMEDGE, SM_SCR_NLREC, SM_SCR_NLEULER, SM_SCR_RVLIST, SM_SCR_COUNT };

// A launch-bound chain of small kernels with fixed arguments (the MST rounds and the Euler-tour rooting of nl.cu: ~90
// launches of a few microseconds each) replayed as a CUDA graph: see smi_graphed below.
struct sm_graph_slot {
  cudaGraphExec_t exec = nullptr;
  unsigned long long key = 0, last_key = 0;   // arguments the instantiated graph was captured with / seen last
  long long launches = 0;                     // kernel launches inside (for the ctx's launch counter)
};
enum { SM_GRAPH_BORUVKA = 0, SM_GRAPH_EULER, SM_GRAPH_COUNT };

struct sm_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  long long launches = 0;
  sm_scratch scr[SM_SCR_COUNT];
  // host-side cache of the exp tables currently resident in scr[SM_SCR_TAB]
  float tab_trunc = -1.f, tab_lamAD = -1.f, tab_lamCen = -1.f;
  int tab_codeLen = -1;
  float h_tab[767 + 72];  // host staging of those tables (must outlive the async upload)
  int num_sms = SM_NUM_SMS;
  // edge hand-off buffers of the grouped SGM sweeps: armed (all sentinel) up to this many bytes at this address
  void* sgm_edge_ptr = nullptr;
  size_t sgm_edge_armed = 0;
  sm_graph_slot graphs[SM_GRAPH_COUNT];
};

void sm_set_error(const char* fmt, ...);
int sm_scratch_get(sm_ctx* ctx, int slot, size_t bytes, void** out);

#define SM_CHECK_ARG(cond)                                                        \
  do {                                                                            \
    if (!(cond)) {                                                                \
      sm_set_error("%s:%d: precondition failed: %s", __FILE__, __LINE__, #cond);  \
      return SM_ERR_ARG;                                                          \
    }                                                                             \
  } while (0)

#define SM_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      sm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));  \
      return SM_ERR_CUDA;                                                                   \
    }                                                                                       \
  } while (0)

#define SM_TRY(call)            \
  do {                          \
    int r__ = (call);           \
    if (r__ != SM_OK) return r__; \
  } while (0)

static inline unsigned long long smi_key_mix(unsigned long long h, unsigned long long v) {
  return (h ^ v) * 1099511628211ull;   // FNV-1a over 64-bit words
}
// body(): a sequence of SM_LAUNCH / cudaMemsetAsync calls on ctx->stream without host synchronisation or allocation,
// fully determined by `key` (every pointer and size it uses).  First call with a key: run eagerly.  Second consecutive
// call with the same key: capture into a graph, instantiate, launch.  From then on: one cudaGraphLaunch.  A different key
// (other buffers, other size) falls back to eager and starts over, so callers that pass fresh buffers every time
// (the stage API from a test) never pay for a capture.
template <class F>
static inline int smi_graphed(sm_ctx* ctx, sm_graph_slot& g, unsigned long long key, F body) {
  if (g.exec && g.key == key) {
    SM_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
    ctx->launches += g.launches;
    return SM_OK;
  }
  if (g.last_key != key) { g.last_key = key; return body(); }
  if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
  const long long l0 = ctx->launches;
  if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return body(); }
  const int rc = body();
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
  if (rc != SM_OK || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    ctx->launches = l0;
    return rc != SM_OK ? rc : body();
  }
  e = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { g.exec = nullptr; cudaGetLastError(); ctx->launches = l0; return body(); }
  g.key = key;
  g.launches = ctx->launches - l0;
  SM_CUDA(cudaGraphLaunch(g.exec, ctx->stream));
  return SM_OK;
}

// A ctx's stream and scratch memory live on ctx->device; the caller's current device may be another one (two ctxs in
// one process, or the host application switched devices), so every launch / allocation rebinds when it differs.
static inline cudaError_t smi_bind_device(const sm_ctx* ctx) {
  int cur = -1;
  cudaError_t e = cudaGetDevice(&cur);
  if (e == cudaSuccess && cur != ctx->device) e = cudaSetDevice(ctx->device);
  return e;
}

// Launch bookkeeping: every kernel launch of the library goes through this so
// bench.py can report gpu_launches truthfully.
#define SM_LAUNCH(ctx, kernel, grid, block, smem, ...)                                       \
  do {                                                                                       \
    smi_bind_device(ctx);                                                                    \
    kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);                         \
    (ctx)->launches++;                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) {                                                                \
      sm_set_error("%s:%d: launch %s -> %s", __FILE__, __LINE__, #kernel,                    \
                   cudaGetErrorString(e__));                                                 \
      return SM_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

static inline int sm_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- internal cross-file helpers ------------------------------------------------
// BGR u8x3 interleaved -> one uint32 per pixel (b | g<<8 | r<<16), 4-byte loads.
int smi_pack_bgr(sm_ctx* ctx, const uint8_t* d_bgr, long long npix, uint32_t* d_out);
// arms [H][W][5] u16 -> one uint2 per pixel {armH = left | right<<16, armV = up | down<<16} in rows of W + 2*PAD
// entries: pixel u sits at index PAD + u and the PAD entries on either side are zero (partner outside the image).
// Followed by the armH plane and the armV plane (one u32 per entry each): d_out holds 4 * H * (W + 2*PAD) words.
// (Tried in round 2 and dropped: one uint32 {left, right, up, down} per pixel for the second passes -- VIMNMX.U8x4 +
// IDP.4A + PRMT, half the staged arm bytes: bit-exact but SLOWER at equal occupancy, V second 1.66 -> 1.77 ms, H second
// 1.29 -> 1.33 ms at 1080p D=256, and the freed shared memory let a sixth warp in, which costs the vertical second pass
// its L1 carve-out: 2.11 ms.  profiles/r02_cbca_experiments.md.)
int smi_pack_arms(sm_ctx* ctx, const uint16_t* d_arms, int H, int W, int PAD, uint32_t* d_out);
static inline int smi_arm_pad(int D) { return (D + 31) / 32 * 32 + 4; }   // +4: 16-byte aligned supersets stay in the row
// the two exp lookup tables of the fused AD-Census kernel (see cost.cu)
int smi_exp_tables(sm_ctx* ctx, float trunc, float lamAD, float lamCen, int codeLen, const float** d_tabAD,
                   const float** d_tabCen);

// packed-input forms used by the frame pipeline (no per-call packing)
int smi_cost_adcensus_packed(sm_ctx* ctx, const uint32_t* d_pixL, const uint32_t* d_pixR, const uint64_t* d_cenL,
                             const uint64_t* d_cenR, int H, int W, int D, int func, float adTrunc, float lamAD,
                             float lamCen, int LOR, float* d_vol);
// postScale != 1: the last pass also applies the caller's one-level SolveAll (vm = 0 + postScale * vm)
int smi_cbca_packed(sm_ctx* ctx, float* d_vol, float* d_tmp, const uint32_t* d_armL, const uint32_t* d_armR, int H,
                    int W, int D, int iters, int view, int Lmax, int PAD, float postScale = 1.0f);
int smi_arms_packed(sm_ctx* ctx, const uint32_t* d_pix, int H, int W, int L, int L_out, int tau, int tau_out, int minL,
                    uint16_t* d_arms);
int smi_nl(sm_ctx* ctx, const uint8_t* d_bgrL, float* d_vol, double* d_work, int H, int W, int D);
int smi_nl_tree(sm_ctx* ctx, const uint8_t* d_bgrL, int H, int W);
int smi_nl_filter(sm_ctx* ctx, float* d_vol, double* d_work, int H, int W, int D);
int smi_sgm_path_packed(sm_ctx* ctx, const float* d_vol, const uint32_t* d_pix, int H, int W, int D, int path,
                        int corDifThres, int reduCoeffi1, int mode, float* d_out);
// mode 2: accumulate and write the WTA of the finished sum into d_disp (last path of a view); mode 3: the WTA alone,
// the finished sum is not stored
int smi_sgm_path_packed2(sm_ctx* ctx, const float* d_vol, const uint32_t* d_pix, int H, int W, int D, int path,
                         int corDifThres, int reduCoeffi1, int mode, float* d_out, int16_t* d_disp);

// 16-bit integer-cost SGM (sgm_u16.cu): fixed point = reduCoeffi1 x the reference's float values
bool smi_sgm_u16_ok(int D, int paths, int reduCoeffi1, int maxCost);
// grouped = true: the 8-path table runs as two three-path row sweeps (smi_sgm_group_u16) + the two horizontal sweeps when the
// shape fits; integer sums are associative, so the result is the same volume either way
int smi_sgm_u16(sm_ctx* ctx, const uint16_t* d_vol, const uint32_t* d_pix, int H, int W, int D, int paths, int corDifThres,
                int reduCoeffi1, uint16_t* d_sum, int16_t* d_disp, bool keep_sum, bool grouped = true);
int smi_sgm_path_u16(sm_ctx* ctx, const uint16_t* d_vol, const uint32_t* d_pix, int H, int W, int D, int path, int corDifThres,
                     int scale, int mode, uint16_t* d_out, int16_t* d_disp);
int smi_sgm_group_u16(sm_ctx* ctx, const uint16_t* d_vol, const uint32_t* d_pix, int H, int W, int D, int up, int mode,
                      int corDifThres, int reduCoeffi1, uint16_t* d_sum);

// grouped SGM sweep (sgm_group.cu): up = 1 -> paths {0,4,5}, up = 0 -> paths {1,6,7}; mode 0 writes, 1 accumulates
int smi_sgm_group2(sm_ctx* ctx, const float* const* d_vol, const uint32_t* const* d_pix, int H, int W, int D, int up, int mode,
                   int corDifThres, int reduCoeffi1, float* const* d_sum);
int smi_sgm8_grouped2(sm_ctx* ctx, const float* const* d_vol, const uint32_t* const* d_pix, int H, int W, int D, int corDifThres,
                      int reduCoeffi1, float* const* d_sum, int16_t* const* d_disp, cudaEvent_t ev_after_sweeps = nullptr,
                      bool* used_sweeps = nullptr, const bool* keep_sum = nullptr);
int smi_sgm_group(sm_ctx* ctx, const float* d_vol, const uint32_t* d_pix, int H, int W, int D, int up, int mode,
                  int corDifThres, int reduCoeffi1, float* d_sum);

int smi_sgm8_grouped(sm_ctx* ctx, const float* d_vol, const uint32_t* d_pix, int H, int W, int D, int corDifThres,
                     int reduCoeffi1, float* d_sum, int16_t* d_disp, cudaEvent_t ev_after_sweeps = nullptr,
                     bool* used_sweeps = nullptr, bool keep_sum = true);

__device__ __forceinline__ int smd_absdiff_max3(uint32_t a, uint32_t b) {
  // max over the three low bytes of |a_c - b_c|
  uint32_t d = __vabsdiffu4(a, b);
  int d0 = d & 0xff, d1 = (d >> 8) & 0xff, d2 = (d >> 16) & 0xff;
  return max(d0, max(d1, d2));
}

// expf exactly as the host's glibc evaluates it (>= 2.27: sysdeps/ieee754/flt-32/e_expf.c, the table-driven double-
// precision algorithm of ARM's optimized-routines; x86-64 resolves it to the FMA ifunc variant, whose four
// multiply-adds are contracted): exp(x) = 2^(k/32) * 2^(r/32), k = round(x * 32/ln2) by the 1.5 * 2^52 shift trick,
// a cubic in r, all in double, one rounding to float at the end.  Bit-identical to the libm the oracle and the compiled
// reference call on every float in [-320, 100] (2.25e9 inputs checked exhaustively on the host against libm's expf with the
// same operations, oracle/expf_check.c; tests/test_expf_emulation.py re-checks a sample), 0 below -0x1.9fe368p6 and
// +inf above 0x1.62e42ep6 like libm.  This is
// what lets exp-weighted stages (WM, censusGrad's gradient term, gen_vm_from2vm_exp on materialised volumes) match the
// CPU reference bit for bit instead of to 1e-4.
// T[i] = bits(2^(i/32)) - (i << 47).  In GLOBAL memory (an L1-cached 8-byte gather: the 32 entries span two
// 128-byte lines), not __constant__: the index differs per lane, and divergent constant-bank reads serialise
// (measured: the censusGrad cost stage 3.1 ms with the constant-bank table).
static __device__ const unsigned long long smd_expf_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull,
    0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull,
    0x3feedea64c123422ull, 0x3feece086061892dull, 0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull,
    0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, 0x3feee89f995ad3adull,
    0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
    0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
__device__ __forceinline__ float smd_expf_host(float x, const unsigned long long* __restrict__ T = smd_expf_tab) {
  if (x < -0x1.9fe368p6f) return 0.0f;                      // underflow (also -inf)
  if (!(x <= 0x1.62e42ep6f)) return x > 0.f ? __int_as_float(0x7f800000) : x + x;   // overflow -> +inf; NaN -> NaN
  const double InvLn2N = 0x1.71547652b82fep+0 * 32.0, SHIFT = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0, C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0,
               C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
  const double xd = (double)x;
  const double z = __dmul_rn(InvLn2N, xd);
  double kd = __dadd_rn(z, SHIFT);
  const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
  kd = __dsub_rn(kd, SHIFT);
  const double r = __fma_rn(InvLn2N, xd, -kd);
  const double s = __longlong_as_double((long long)(T[ki & 31] + (ki << 47)));
  const double zz = __fma_rn(C0, r, C1);
  const double r2 = __dmul_rn(r, r);
  double y = __fma_rn(C2, r, 1.0);
  y = __fma_rn(zz, r2, y);
  y = __dmul_rn(y, s);
  return __double2float_rn(y);
}
