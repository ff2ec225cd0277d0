// Yang's own non-local stereo driver, qx_nonlocal_cost_aggregation (API surface named by the north star):
//
//   compute_gradient                         NL/qx_nonlocal_cost_aggregation.cpp:219-236
//   matching_cost_from_color_and_gradient    NL/qx_nonlocal_cost_aggregation.cpp:190-218
//   qx_stereo_flip_corr_vol                  NL/qx_basic.cpp:577-588
//   depth_best_cost / vec_min_pos            NL/qx_basic.cpp:589-602
//   qx_detect_occlusion_left_right           NL/qx_basic.cpp:603-624
//   disparity()'s |d - disp| refinement volume   NL/qx_nonlocal_cost_aggregation.cpp:92-99
//
// All volumes are float64 [H][W][D] as in the reference (the tree filter runs in f64, sm_tree_filter_f64).  Every
// kernel is elementwise; the arithmetic keeps the reference's operation order (-fmad=false: no contraction), so
// the volumes are bit-identical to the CPU result.
#include "common.cuh"

// rgb_2_gray (NL/qx_basic.h:72)
__device__ __forceinline__ float nlca_gray(const uint8_t* in) {
  return (float)(unsigned char)(0.299 * in[0] + 0.587 * in[1] + 0.114 * in[2] + 0.5);
}

__global__ void k_nlca_gradient(const uint8_t* __restrict__ img, int H, int W, float* __restrict__ grad) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const uint8_t* row = img + (size_t)y * W * 3;
  float g;
  if (W == 1) g = 127.5f;   // degenerate: the reference reads column 1, which does not exist; defined here as flat
  else if (x == 0) g = (float)((double)(nlca_gray(row + 3) - nlca_gray(row)) + 127.5);
  else if (x == W - 1) g = (float)((double)(nlca_gray(row + 3 * (W - 1)) - nlca_gray(row + 3 * (W - 2))) + 127.5);
  else g = (float)(0.5 * (double)(nlca_gray(row + 3 * (x + 1)) - nlca_gray(row + 3 * (x - 1))) + 127.5);
  grad[(size_t)y * W + x] = g;
}

__global__ void k_nlca_cost(const uint8_t* __restrict__ left, const uint8_t* __restrict__ right,
                            const float* __restrict__ gl, const float* __restrict__ gr, int H, int W, int D, double maxc,
                            double maxg, double wc, double wci, double* __restrict__ vol) {
  const size_t n = (size_t)H * W * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const size_t p = i / D;
    const int x = (int)(p % W);
    const size_t rowp = p - x;
    const int xs = x >= d ? x - d : 0;   // columns left of the shift replicate column 0
    const uint8_t* l = left + p * 3;
    const uint8_t* r = right + (rowp + xs) * 3;
    double cost = 0;
    for (int c = 0; c < 3; c++) cost += (double)abs((int)l[c] - (int)r[c]);
    cost = fmin(cost / 3, maxc);
    const double cg = fmin((double)fabsf(gl[p] - gr[rowp + xs]), maxg);
    vol[i] = wc * cost + wci * cg;
  }
}

__global__ void k_nlca_flip(const double* __restrict__ vol, int H, int W, int D, double* __restrict__ volR) {
  const size_t n = (size_t)H * W * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const size_t p = i / D;
    const int x = (int)(p % W);
    const int dd = min(d, W - 1 - x);   // beyond the image the previous plane's value repeats
    volR[i] = vol[(p + dd) * D + dd];
  }
}

__global__ void k_argmin_f64(const double* __restrict__ vol, size_t npix, int D, uint8_t* __restrict__ depth) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (size_t)gridDim.x * blockDim.x) {
    const double* in = vol + p * D;
    double mv = in[0];
    int mp = 0;
    for (int i = 1; i < D; i++)
      if (in[i] < mv) { mv = in[i]; mp = i; }
    depth[p] = (uint8_t)mp;
  }
}

__global__ void k_nlca_occlusion(const uint8_t* __restrict__ dl, const uint8_t* __restrict__ dr, int H, int W,
                                 uint8_t* __restrict__ mask) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const size_t p = (size_t)y * W + x;
  const int d = dl[p], xr = x - d;
  uint8_t m = 0;
  if (xr >= 0) { if (d == 0 || abs(d - (int)dr[p - d]) >= 1) m = 255; }
  else m = 255;
  mask[p] = m;
}

__global__ void k_nlca_refine_cost(const uint8_t* __restrict__ disp, const uint8_t* __restrict__ mask, size_t npix, int D,
                                   double* __restrict__ vol) {
  const size_t n = npix * D;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const size_t p = i / D;
    vol[i] = mask[p] ? 0.0 : (double)abs((int)disp[p] - d);
  }
}

static int grid_for(sm_ctx* ctx, size_t n) { return (int)max((size_t)1, min((size_t)ctx->num_sms * 16, (n + 255) / 256)); }

extern "C" int sm_nlca_gradient(sm_ctx* ctx, const uint8_t* d_img, int H, int W, float* d_grad) {
  SM_CHECK_ARG(ctx && d_img && d_grad && H > 0 && W > 0);
  dim3 grid(sm_div_up(W, 128), H);
  SM_LAUNCH(ctx, k_nlca_gradient, grid, 128, 0, d_img, H, W, d_grad);
  return SM_OK;
}

extern "C" int sm_nlca_cost(sm_ctx* ctx, const uint8_t* d_left, const uint8_t* d_right, int H, int W, int D,
                            double max_color_difference, double max_gradient_difference, double weight_on_color,
                            double* d_vol) {
  SM_CHECK_ARG(ctx && d_left && d_right && d_vol && H > 0 && W > 1 && D > 0);
  const size_t npix = (size_t)H * W;
  void *gl, *gr;
  SM_TRY(sm_scratch_get(ctx, SM_SCR_MISC0, npix * 4, &gl));
  SM_TRY(sm_scratch_get(ctx, SM_SCR_MISC1, npix * 4, &gr));
  SM_TRY(sm_nlca_gradient(ctx, d_left, H, W, (float*)gl));
  SM_TRY(sm_nlca_gradient(ctx, d_right, H, W, (float*)gr));
  SM_LAUNCH(ctx, k_nlca_cost, grid_for(ctx, npix * D), 256, 0, d_left, d_right, (const float*)gl, (const float*)gr, H, W, D,
            max_color_difference, max_gradient_difference, weight_on_color, 1 - weight_on_color, d_vol);
  return SM_OK;
}

extern "C" int sm_nlca_flip(sm_ctx* ctx, const double* d_vol, int H, int W, int D, double* d_vol_right) {
  SM_CHECK_ARG(ctx && d_vol && d_vol_right && d_vol != d_vol_right && H > 0 && W > 0 && D > 0);
  SM_LAUNCH(ctx, k_nlca_flip, grid_for(ctx, (size_t)H * W * D), 256, 0, d_vol, H, W, D, d_vol_right);
  return SM_OK;
}

extern "C" int sm_depth_best_cost(sm_ctx* ctx, const double* d_vol, int H, int W, int D, uint8_t* d_depth) {
  SM_CHECK_ARG(ctx && d_vol && d_depth && H > 0 && W > 0 && D > 0 && D <= 256);
  SM_LAUNCH(ctx, k_argmin_f64, grid_for(ctx, (size_t)H * W), 256, 0, d_vol, (size_t)H * W, D, d_depth);
  return SM_OK;
}

extern "C" int sm_nlca_occlusion(sm_ctx* ctx, const uint8_t* d_disp_left, const uint8_t* d_disp_right, int H, int W,
                                 uint8_t* d_mask) {
  SM_CHECK_ARG(ctx && d_disp_left && d_disp_right && d_mask && H > 0 && W > 0);
  dim3 grid(sm_div_up(W, 128), H);
  SM_LAUNCH(ctx, k_nlca_occlusion, grid, 128, 0, d_disp_left, d_disp_right, H, W, d_mask);
  return SM_OK;
}

extern "C" int sm_nlca_refine_cost(sm_ctx* ctx, const uint8_t* d_disp, const uint8_t* d_mask, int H, int W, int D,
                                   double* d_vol) {
  SM_CHECK_ARG(ctx && d_disp && d_mask && d_vol && H > 0 && W > 0 && D > 0);
  SM_LAUNCH(ctx, k_nlca_refine_cost, grid_for(ctx, (size_t)H * W * D), 256, 0, d_disp, d_mask, (size_t)H * W, D, d_vol);
  return SM_OK;
}
