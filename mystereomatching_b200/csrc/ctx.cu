// Context, memory and small image-format helpers of the sm_b200 library.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void sm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* sm_last_error(void) { return g_err; }

extern "C" int sm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" void sm_params_default(sm_params* p, int maxDisp) {
  memset(p, 0, sizeof(*p));
  p->numDisparities = maxDisp + 1;
  p->censusFunc = 3;
  p->adTrunc = 1000.f;
  p->lamAD = 10.f;
  p->lamCen = 30.f;
  p->cbca_crossL = 17;
  p->cbca_crossL_out = 34;
  p->cbca_cTresh = 20;
  p->cbca_cTresh_out = 6;
  p->cbca_minArmL = 1;
  p->cbca_iterationNum = 2;
  p->sgm_paths = 4;
  p->sgm_corDifThres = 15;
  p->sgm_reduCoeffi1 = 4;
  p->LRmaxDiff = 0.f;
  p->region_vote_nums = 2;
  p->regVote_SThres = 20;
  p->regVote_hratioThres = 0.4f;
  p->DISP_OCC = -32;
  p->DISP_MIS = -48;
  p->aggregation = 1;
  p->Do_refine = 1;
  p->Do_LRConsis = 1;
  p->Do_regionVote = 1;
  p->Do_properIpol = 1;
  p->Do_lastMedianBlur = 1;
  p->crossScaleLambda = -1.f;
  p->sgm_grouped = 1;
  p->costcalculation = 0;
  p->cg_lamCen = 13.f;
  p->cg_lamG = 1.f;
  p->gradTrunc = 500.f;
  p->pyramidLevels = 1;
  p->Do_vmTop = 0;
  p->vmTop_method = 0;
  p->vmTop_Num = 2;        // M = 2, lamc = 109, ts = 10 in main_.cpp:62-64
  p->vmTop_thres = 1.09f;
  p->vmTop_ts = 10;
  p->vmTop_hasCir2 = 1;
  p->vmTop_cir3_doColorLimit = 0;
  p->keep_right_volume = 0;
}

extern "C" int sm_ctx_create(sm_ctx** out, int device, void* stream) {
  SM_CHECK_ARG(out != nullptr);
  int n = sm_device_count();
  if (n <= 0) {
    sm_set_error("sm_ctx_create: no CUDA device visible (this library has no CPU fallback)");
    return SM_ERR_CUDA;
  }
  SM_CHECK_ARG(device >= 0 && device < n);
  SM_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  SM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    sm_set_error("sm_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                 prop.minor);
    return SM_ERR_UNSUPPORTED;
  }
  sm_ctx* c = new sm_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  if (stream) {
    c->stream = (cudaStream_t)stream;
    c->own_stream = false;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete c;
      sm_set_error("cudaStreamCreate: %s", cudaGetErrorString(e));
      return SM_ERR_CUDA;
    }
    c->own_stream = true;
  }
  *out = c;
  return SM_OK;
}

extern "C" int sm_ctx_destroy(sm_ctx* ctx) {
  if (!ctx) return SM_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < SM_GRAPH_COUNT; i++)
    if (ctx->graphs[i].exec) cudaGraphExecDestroy(ctx->graphs[i].exec);
  for (int i = 0; i < SM_SCR_COUNT; i++)
    if (ctx->scr[i].p) cudaFree(ctx->scr[i].p);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return SM_OK;
}

extern "C" int sm_ctx_sync(sm_ctx* ctx) {
  SM_CHECK_ARG(ctx);
  SM_CUDA(cudaStreamSynchronize(ctx->stream));
  return SM_OK;
}

extern "C" void* sm_ctx_stream(sm_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" long long sm_ctx_launch_count(sm_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int sm_dev_alloc(sm_ctx* ctx, void** d_ptr, size_t bytes) {
  SM_CHECK_ARG(ctx && d_ptr);
  SM_CUDA(cudaSetDevice(ctx->device));
  cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    cudaGetLastError();
    sm_set_error("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return SM_ERR_NOMEM;
  }
  return SM_OK;
}

extern "C" int sm_dev_free(sm_ctx* ctx, void* d_ptr) {
  SM_CHECK_ARG(ctx);
  if (!d_ptr) return SM_OK;
  SM_CUDA(cudaSetDevice(ctx->device));
  SM_CUDA(cudaStreamSynchronize(ctx->stream));
  SM_CUDA(cudaFree(d_ptr));
  return SM_OK;
}

extern "C" int sm_host_alloc_pinned(void** h_ptr, size_t bytes) {
  SM_CHECK_ARG(h_ptr);
  SM_CUDA(cudaMallocHost(h_ptr, bytes ? bytes : 1));
  return SM_OK;
}

extern "C" int sm_host_free_pinned(void* h_ptr) {
  if (h_ptr) SM_CUDA(cudaFreeHost(h_ptr));
  return SM_OK;
}

extern "C" int sm_memcpy_h2d(sm_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  SM_CHECK_ARG(ctx && d_dst && h_src);
  SM_CUDA(smi_bind_device(ctx));
  SM_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return SM_OK;
}

extern "C" int sm_memcpy_d2h(sm_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  SM_CHECK_ARG(ctx && h_dst && d_src);
  SM_CUDA(smi_bind_device(ctx));
  SM_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return SM_OK;
}

extern "C" int sm_memset(sm_ctx* ctx, void* d_dst, int byte, size_t bytes) {
  SM_CHECK_ARG(ctx && d_dst);
  SM_CUDA(smi_bind_device(ctx));
  SM_CUDA(cudaMemsetAsync(d_dst, byte, bytes, ctx->stream));
  return SM_OK;
}

extern "C" int sm_memcpy_d2d(sm_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
  SM_CHECK_ARG(ctx && d_dst && d_src);
  SM_CUDA(smi_bind_device(ctx));
  SM_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return SM_OK;
}

int sm_scratch_get(sm_ctx* ctx, int slot, size_t bytes, void** out) {
  sm_scratch& s = ctx->scr[slot];
  if (s.cap < bytes) {
    SM_CUDA(smi_bind_device(ctx));
    // Growing is rare (first frame of a new size).  The old block may still be
    // read by queued kernels, so drain the stream before freeing it.
    if (s.p) {
      SM_CUDA(cudaStreamSynchronize(ctx->stream));
      SM_CUDA(cudaFree(s.p));
      s.p = nullptr;
      s.cap = 0;
    }
    size_t cap = (bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&s.p, cap);
    if (e != cudaSuccess) {
      cudaGetLastError();
      sm_set_error("scratch cudaMalloc(%zu): %s", cap, cudaGetErrorString(e));
      return SM_ERR_NOMEM;
    }
    s.cap = cap;
    if (slot == SM_SCR_TAB) ctx->tab_codeLen = -1;
  }
  *out = s.p;
  return SM_OK;
}

// ---------------------------------------------------------------------------------
__global__ void k_pack_bgr(const uint8_t* __restrict__ bgr, long long npix, uint32_t* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < npix; i += stride) {
    uint32_t b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
    out[i] = b | (g << 8) | (r << 16);
  }
}

int smi_pack_bgr(sm_ctx* ctx, const uint8_t* d_bgr, long long npix, uint32_t* d_out) {
  int grid = min(sm_div_up(npix, 256), ctx->num_sms * 8);
  SM_LAUNCH(ctx, k_pack_bgr, grid, 256, 0, d_bgr, npix, d_out);
  return SM_OK;
}

__global__ void k_pack_arms(const uint16_t* __restrict__ arms, int H, int W, int PAD, uint32_t* __restrict__ out) {
  const int Wp = W + 2 * PAD;
  const long long n = (long long)H * Wp;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int v = (int)(i / Wp), u = (int)(i - (long long)v * Wp) - PAD;
    uint32_t wh = 0, wv = 0;
    if (u >= 0 && u < W) {
      const uint16_t* a = arms + 5 * ((size_t)v * W + u);
      wh = (uint32_t)a[0] | ((uint32_t)a[1] << 16);
      wv = (uint32_t)a[2] | ((uint32_t)a[3] << 16);
    }
    reinterpret_cast<uint2*>(out)[i] = make_uint2(wh, wv);   // pair map
    out[2 * n + i] = wh << 7;                                // armH plane, lengths x 128 (ring bytes of the first pass)
    out[3 * n + i] = wv << 7;                                // armV plane, likewise
  }
}

int smi_pack_arms(sm_ctx* ctx, const uint16_t* d_arms, int H, int W, int PAD, uint32_t* d_out) {
  const long long n = (long long)H * (W + 2 * PAD);
  int grid = min(sm_div_up(n, 256), ctx->num_sms * 8);
  SM_LAUNCH(ctx, k_pack_arms, grid, 256, 0, d_arms, H, W, PAD, d_out);
  return SM_OK;
}

// cv::cvtColor(BGR2GRAY) for 8-bit images: 14-bit fixed point with rounding.
__global__ void k_bgr2gray(const uint8_t* __restrict__ bgr, long long npix, uint8_t* __restrict__ gray) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < npix; i += stride) {
    int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
    gray[i] = (uint8_t)((1868 * b + 9617 * g + 4899 * r + 8192) >> 14);
  }
}

extern "C" int sm_bgr2gray(sm_ctx* ctx, const uint8_t* d_bgr, int H, int W, uint8_t* d_gray) {
  SM_CHECK_ARG(ctx && d_bgr && d_gray && H > 0 && W > 0);
  long long npix = (long long)H * W;
  int grid = min(sm_div_up(npix, 256), ctx->num_sms * 8);
  SM_LAUNCH(ctx, k_bgr2gray, grid, 256, 0, d_bgr, npix, d_gray);
  return SM_OK;
}
