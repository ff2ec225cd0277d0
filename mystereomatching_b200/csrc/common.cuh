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
int smi_sgm_path_packed(sm_ctx* ctx, const float* d_vol, const uint32_t* d_pix, int H, int W, int D, int path,
                        int corDifThres, int reduCoeffi1, int mode, float* d_out);
// mode 2: accumulate and write the WTA of the finished sum into d_disp (last path of a view)
int smi_sgm_path_packed2(sm_ctx* ctx, const float* d_vol, const uint32_t* d_pix, int H, int W, int D, int path,
                         int corDifThres, int reduCoeffi1, int mode, float* d_out, int16_t* d_disp);

// grouped SGM sweep (sgm_group.cu): up = 1 -> paths {0,4,5}, up = 0 -> paths {1,6,7}; mode 0 writes, 1 accumulates
int smi_sgm_group2(sm_ctx* ctx, const float* const* d_vol, const uint32_t* const* d_pix, int H, int W, int D, int up, int mode,
                   int corDifThres, int reduCoeffi1, float* const* d_sum);
int smi_sgm8_grouped2(sm_ctx* ctx, const float* const* d_vol, const uint32_t* const* d_pix, int H, int W, int D, int corDifThres,
                      int reduCoeffi1, float* const* d_sum, int16_t* const* d_disp, cudaEvent_t ev_after_sweeps = nullptr,
                      bool* used_sweeps = nullptr);
int smi_sgm_group(sm_ctx* ctx, const float* d_vol, const uint32_t* d_pix, int H, int W, int D, int up, int mode,
                  int corDifThres, int reduCoeffi1, float* d_sum);

int smi_sgm8_grouped(sm_ctx* ctx, const float* d_vol, const uint32_t* d_pix, int H, int W, int D, int corDifThres,
                     int reduCoeffi1, float* d_sum, int16_t* d_disp, cudaEvent_t ev_after_sweeps = nullptr,
                     bool* used_sweeps = nullptr);

__device__ __forceinline__ int smd_absdiff_max3(uint32_t a, uint32_t b) {
  // max over the three low bytes of |a_c - b_c|
  uint32_t d = __vabsdiffu4(a, b);
  int d0 = d & 0xff, d1 = (d >> 8) & 0xff, d2 = (d >> 16) & 0xff;
  return max(d0, max(d1, d2));
}
