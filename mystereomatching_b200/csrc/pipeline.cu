// Whole-frame pipeline: StereoMatching::pipeline() (stereoMatching.cpp:1950-1981)
// = costCalculate (stereoMatching.cpp:945-1021) -> dispOptimize (:1046-1136) ->
// refine (:1364-1506), with costcalculation="ADCensus", aggregation "CBCA" | "NL"
// | none, optimization="sgm".  One sm_pipeline owns every device buffer a frame
// of its size needs, so a stream of frames allocates nothing.
//
// HBM residency at 1080p, D=256 (fp32): three volumes (vm[0], vm[1], scratch)
// = 3 x 2.12 GB; everything else (images, census codes, arms, disparities) is
// < 120 MB.  The reference's HVL_INTERSECTION (2 x 5.3 GB), its four cost
// temporaries and its P path volumes L[i] are never materialised.
#include "common.cuh"


enum { ST_CENSUS = 0, ST_COST, ST_ARMS, ST_AGG, ST_SGM, ST_WTA, ST_REFINE, ST_TOTAL, ST_COUNT };

struct sm_pipeline {
  sm_ctx* ctx = nullptr;
  int H = 0, W = 0, D = 0;
  sm_params p;
  uint8_t *bgr[2] = {nullptr, nullptr}, *gray[2] = {nullptr, nullptr};          // the images the stages read
  uint8_t *own_bgr[2] = {nullptr, nullptr}, *own_gray[2] = {nullptr, nullptr};  // the pipeline's own buffers (sm_pipeline_upload)
  uint32_t *pix[2] = {nullptr, nullptr}, *armpk[2] = {nullptr, nullptr};
  uint64_t* cen[2] = {nullptr, nullptr};
  uint16_t* arms[2] = {nullptr, nullptr};
  float* vol[4] = {nullptr, nullptr, nullptr, nullptr};  // vm[0], vm[1], scratch, second scratch (roles rotate after SGM)
  float* grad[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // censusGrad: gx, gy per image
  double* nlwork = nullptr;
  int16_t *disp[2] = {nullptr, nullptr}, *dtmp = nullptr;
  float* top = nullptr;      // vmTop candidate lists [H][W][vmTop_Num + 1][2] (Do_vmTop)
  uint8_t* h_in = nullptr;   // pinned staging: bgrL | bgrR | grayL | grayR
  int16_t* h_out = nullptr;  // pinned staging: dispL | dispR
  // small frames (a single-path SGM sweep is then a chain of H or W dependent steps on a few hundred warps, far from
  // filling the GPU): the two views' sweeps run concurrently, view 1 on a second stream
  cudaStream_t stream2 = nullptr;
  cudaEvent_t evFork = nullptr, evJoin = nullptr;
  cudaStream_t streamA = nullptr;   // arms of both images (CBCA frames) / the MST build (NL frames) under the cost kernels
  bool nl_tree_pending = false;     // NL frames: the tree is being built on streamA
  cudaEvent_t evA0 = nullptr, evA1 = nullptr;
  bool arms_pending = false;
  sm_pipeline* child = nullptr;   // next pyramid level (cost + aggregation only), pyramidLevels > 1
  bool is_child = false;
  bool have_gray = false, have_arms = false, scale_folded = false;
  bool u16 = false;          // "Census" without aggregation: the volumes are uint16 (vol[] then point at uint16 data)
  bool stage_used = false, stage_drained = true;
  bool timing = false;
  cudaEvent_t ev[ST_COUNT + 1];
  cudaEvent_t evs[2][2];          // sgm stage, per view: start, after the grouped sweeps
  bool sweeps[2] = {false, false};
  bool ev_ok = false;
  float ms[ST_COUNT];
};

static int pl_alloc(sm_ctx* ctx, void** p, size_t bytes) { return sm_dev_alloc(ctx, p, bytes); }

extern "C" int sm_pipeline_destroy(sm_pipeline* pl) {
  if (!pl) return SM_OK;
  if (pl->child) sm_pipeline_destroy(pl->child);
  sm_ctx* c = pl->ctx;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (int i = 0; i < 2; i++) {
    cudaFree(pl->own_bgr[i]); cudaFree(pl->own_gray[i]); cudaFree(pl->pix[i]); cudaFree(pl->armpk[i]);
    cudaFree(pl->cen[i]); cudaFree(pl->arms[i]); cudaFree(pl->disp[i]);
    cudaFree(pl->grad[i][0]); cudaFree(pl->grad[i][1]);
  }
  for (int i = 0; i < 4; i++) cudaFree(pl->vol[i]);
  if (pl->stream2) cudaStreamDestroy(pl->stream2);
  if (pl->evFork) cudaEventDestroy(pl->evFork);
  if (pl->evJoin) cudaEventDestroy(pl->evJoin);
  if (pl->streamA) cudaStreamDestroy(pl->streamA);
  if (pl->evA0) cudaEventDestroy(pl->evA0);
  if (pl->evA1) cudaEventDestroy(pl->evA1);
  cudaFree(pl->nlwork);
  cudaFree(pl->top);
  cudaFree(pl->dtmp);
  if (pl->h_in) cudaFreeHost(pl->h_in);
  if (pl->h_out) cudaFreeHost(pl->h_out);
  if (pl->ev_ok) {
    for (int i = 0; i <= ST_COUNT; i++) cudaEventDestroy(pl->ev[i]);
    for (int i = 0; i < 4; i++) cudaEventDestroy(pl->evs[i / 2][i % 2]);
  }
  delete pl;
  return SM_OK;
}

extern "C" int sm_pipeline_create(sm_ctx* ctx, int H, int W, const sm_params* p, sm_pipeline** out) {
  SM_CHECK_ARG(ctx && p && out && H > 0 && W > 0);
  SM_CHECK_ARG(p->numDisparities > 0 && p->numDisparities <= 512);
  SM_CHECK_ARG(p->censusFunc == 0 || p->censusFunc == 3);
  SM_CHECK_ARG(p->sgm_paths >= 0 && p->sgm_paths <= 8);
  SM_CHECK_ARG(p->aggregation >= 0 && p->aggregation <= 2);
  SM_CHECK_ARG(p->costcalculation >= 0 && p->costcalculation <= 2);
  SM_CHECK_ARG(p->costcalculation != 1 || (H >= 2 && W >= 2 && p->cg_lamCen > 0.f && p->cg_lamG > 0.f));
  SM_CHECK_ARG(p->pyramidLevels >= 1 && p->pyramidLevels <= SM_MAX_PYRAMID);
  SM_CHECK_ARG(p->cbca_crossL_out >= 0 && p->cbca_crossL_out <= 255);
  SM_CHECK_ARG(!p->Do_vmTop || (p->vmTop_Num >= 1 && p->vmTop_Num <= 16 && p->vmTop_method >= 0 && p->vmTop_method <= 2));
  SM_CUDA(cudaSetDevice(ctx->device));
  sm_pipeline* pl = new sm_pipeline();
  pl->ctx = ctx; pl->H = H; pl->W = W; pl->D = p->numDisparities; pl->p = *p;
  const size_t npix = (size_t)H * W, nvol = npix * pl->D;
  const int nw = sm_census_words(p->censusFunc);
  // "Census" + no aggregation + a power-of-two sgm_reduCoeffi1: every value on the path is an exact multiple of
  // 1 / reduCoeffi1, so the frame runs on uint16 volumes (sgm_u16.cu)
  pl->u16 = p->costcalculation == 2 && p->aggregation == 0 && !p->Do_vmTop && p->crossScaleLambda < 0.f &&
            (p->sgm_paths == 0 || smi_sgm_u16_ok(pl->D, p->sgm_paths, p->sgm_reduCoeffi1, sm_census_code_length(p->censusFunc)));
  const size_t volB = nvol * (pl->u16 ? sizeof(uint16_t) : sizeof(float));
  int rc = SM_OK;
  for (int i = 0; i < 2 && rc == SM_OK; i++) {
    if (rc == SM_OK) rc = pl_alloc(ctx, (void**)&pl->own_bgr[i], npix * 3);
    if (rc == SM_OK) rc = pl_alloc(ctx, (void**)&pl->own_gray[i], npix);
    pl->bgr[i] = pl->own_bgr[i]; pl->gray[i] = pl->own_gray[i];
    if (rc == SM_OK) rc = pl_alloc(ctx, (void**)&pl->pix[i], npix * 4);
    if (rc == SM_OK) rc = pl_alloc(ctx, (void**)&pl->armpk[i], (size_t)H * (W + 2 * smi_arm_pad(pl->D)) * 16);
    if (rc == SM_OK) rc = pl_alloc(ctx, (void**)&pl->cen[i], npix * 8 * nw);
    if (rc == SM_OK) rc = pl_alloc(ctx, (void**)&pl->arms[i], npix * 5 * 2);
    if (rc == SM_OK) rc = pl_alloc(ctx, (void**)&pl->disp[i], npix * 2);
    for (int k = 0; k < 2 && p->costcalculation == 1; k++)
      if (rc == SM_OK) rc = pl_alloc(ctx, (void**)&pl->grad[i][k], npix * sizeof(float));
  }
  // a fourth volume only where the two views' SGM sweeps can share a launch (two sums are written at once)
  const bool two_view_sgm = !pl->u16 && p->sgm_paths == 8 && p->sgm_grouped && p->Do_refine && p->Do_LRConsis;
  const bool two_stream_sgm = !two_view_sgm && !pl->u16 && p->sgm_paths > 0 && p->Do_refine && p->Do_LRConsis && nvol <= ((size_t)48 << 20);
  for (int i = 0; i < ((two_view_sgm || two_stream_sgm) ? 4 : 3) && rc == SM_OK; i++)
    rc = pl_alloc(ctx, (void**)&pl->vol[i], volB);
  if (rc == SM_OK && two_stream_sgm) {
    if (cudaStreamCreateWithFlags(&pl->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&pl->evFork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&pl->evJoin, cudaEventDisableTiming) != cudaSuccess)
      rc = SM_ERR_CUDA;
  }
  // side stream: the arm maps (CBCA frames; frames without aggregation whose refinement votes over the arms) or the MST
  // build (NL frames) under the cost kernels.  censusGrad needs the arms BEFORE its cost kernel: nothing to overlap there.
  const bool side_arms = p->costcalculation != 1 && (p->aggregation == 1 || (p->aggregation == 0 && p->Do_refine && p->Do_regionVote));
  const bool side_tree = p->costcalculation == 0 && p->aggregation == 2;
  if (rc == SM_OK && (side_arms || side_tree)) {
    if (cudaStreamCreateWithFlags(&pl->streamA, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&pl->evA0, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&pl->evA1, cudaEventDisableTiming) != cudaSuccess)
      rc = SM_ERR_CUDA;
  }
  if (rc == SM_OK) rc = pl_alloc(ctx, (void**)&pl->dtmp, npix * 2);
  if (rc == SM_OK && p->Do_vmTop) rc = pl_alloc(ctx, (void**)&pl->top, npix * (size_t)(p->vmTop_Num + 1) * 2 * sizeof(float));
  if (rc == SM_OK && p->aggregation == 2) rc = pl_alloc(ctx, (void**)&pl->nlwork, npix * (size_t)(pl->D + 1) * sizeof(double));
  if (rc == SM_OK && cudaMallocHost((void**)&pl->h_in, npix * 8) != cudaSuccess) rc = SM_ERR_NOMEM;
  if (rc == SM_OK && cudaMallocHost((void**)&pl->h_out, npix * 4) != cudaSuccess) rc = SM_ERR_NOMEM;
  if (rc == SM_OK && p->pyramidLevels > 1 && p->crossScaleLambda >= 0.f) {
    // next level of the caller's pyramid (main_.cpp:134-151): half-size images, maxDisp/2 + 1, disSc * 2 (arm lengths
    // L/scale, L_out/scale, stereoMatching.cpp:5368-5371); it runs costCalculate() only
    sm_params cp = *p;
    cp.numDisparities = (p->numDisparities - 1) / 2 + 1 + 1;
    cp.cbca_crossL = p->cbca_crossL / 2;
    cp.cbca_crossL_out = p->cbca_crossL_out / 2;
    cp.pyramidLevels = p->pyramidLevels - 1;
    rc = sm_pipeline_create(ctx, (H + 1) / 2, (W + 1) / 2, &cp, &pl->child);
    if (rc == SM_OK) pl->child->is_child = true;
  }
  if (rc != SM_OK) { sm_pipeline_destroy(pl); return rc; }
  *out = pl;
  return SM_OK;
}

// Page-locked caller buffers are copied from / to directly; pageable ones go through the pipeline's own
// pinned staging area (one extra host memcpy) so the PCIe transfer is always a true async DMA.
static bool host_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

static int upload_one(sm_pipeline* pl, void* d_dst, const uint8_t* h_src, uint8_t* stage, size_t bytes) {
  sm_ctx* c = pl->ctx;
  if (host_is_pinned(h_src)) {
    SM_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c->stream));
  } else {
    if (!pl->stage_drained) {  // the previous frame's async copies out of the staging area must have finished
      SM_CUDA(cudaStreamSynchronize(c->stream));
      pl->stage_drained = true;
    }
    memcpy(stage, h_src, bytes);
    SM_CUDA(cudaMemcpyAsync(d_dst, stage, bytes, cudaMemcpyHostToDevice, c->stream));
    pl->stage_used = true;
  }
  return SM_OK;
}

// The next sm_pipeline_run_device reads the stereo pair from caller-owned DEVICE buffers (a frame stream uploads
// frame i+1 into a second set while frame i computes); all NULL: back to the pipeline's own buffers.
extern "C" int sm_pipeline_bind_inputs(sm_pipeline* pl, const uint8_t* d_bgrL, const uint8_t* d_bgrR, const uint8_t* d_grayL,
                                       const uint8_t* d_grayR) {
  SM_CHECK_ARG(pl && !pl->is_child);
  SM_CHECK_ARG((d_bgrL == nullptr) == (d_bgrR == nullptr) && (d_grayL == nullptr) == (d_grayR == nullptr));
  SM_CHECK_ARG(d_bgrL || !d_grayL);
  if (!d_bgrL) {
    for (int i = 0; i < 2; i++) { pl->bgr[i] = pl->own_bgr[i]; pl->gray[i] = pl->own_gray[i]; }
    pl->have_gray = false;
    return SM_OK;
  }
  pl->bgr[0] = const_cast<uint8_t*>(d_bgrL); pl->bgr[1] = const_cast<uint8_t*>(d_bgrR);
  pl->gray[0] = d_grayL ? const_cast<uint8_t*>(d_grayL) : pl->own_gray[0];
  pl->gray[1] = d_grayR ? const_cast<uint8_t*>(d_grayR) : pl->own_gray[1];
  pl->have_gray = d_grayL != nullptr;
  return SM_OK;
}

extern "C" int sm_pipeline_upload(sm_pipeline* pl, const uint8_t* h_bgrL, const uint8_t* h_bgrR, const uint8_t* h_grayL,
                                  const uint8_t* h_grayR) {
  SM_CHECK_ARG(pl && h_bgrL && h_bgrR);
  SM_CHECK_ARG((h_grayL == nullptr) == (h_grayR == nullptr));
  const size_t npix = (size_t)pl->H * pl->W;
  SM_CUDA(cudaSetDevice(pl->ctx->device));
  for (int i = 0; i < 2; i++) { pl->bgr[i] = pl->own_bgr[i]; pl->gray[i] = pl->own_gray[i]; }
  pl->stage_drained = !pl->stage_used;
  pl->stage_used = false;
  SM_TRY(upload_one(pl, pl->bgr[0], h_bgrL, pl->h_in, npix * 3));
  SM_TRY(upload_one(pl, pl->bgr[1], h_bgrR, pl->h_in + npix * 3, npix * 3));
  pl->have_gray = h_grayL != nullptr;
  if (pl->have_gray) {
    SM_TRY(upload_one(pl, pl->gray[0], h_grayL, pl->h_in + npix * 6, npix));
    SM_TRY(upload_one(pl, pl->gray[1], h_grayR, pl->h_in + npix * 7, npix));
  }
  return SM_OK;
}

#define PL_MARK(k)                                                        \
  do {                                                                    \
    if (pl->timing) SM_CUDA(cudaEventRecord(pl->ev[k], c->stream));        \
  } while (0)

static int pl_ensure_arms(sm_pipeline* pl) {
  if (pl->have_arms) return SM_OK;
  sm_ctx* c = pl->ctx;
  const sm_params& P = pl->p;
  for (int i = 0; i < 2; i++) {
    SM_TRY(smi_arms_packed(c, pl->pix[i], pl->H, pl->W, P.cbca_crossL, P.cbca_crossL_out, P.cbca_cTresh, P.cbca_cTresh_out,
                           P.cbca_minArmL, pl->arms[i]));
    SM_TRY(smi_pack_arms(c, pl->arms[i], pl->H, pl->W, smi_arm_pad(pl->D), pl->armpk[i]));
  }
  pl->have_arms = true;
  return SM_OK;
}

// costCalculate() of one pyramid level (stereoMatching.cpp:945-1044): census, cost volumes, arms, aggregation.
static int pl_cost_calculate(sm_pipeline* pl) {
  sm_ctx* c = pl->ctx;
  const sm_params& P = pl->p;
  const int H = pl->H, W = pl->W, D = pl->D;
  const long long npix = (long long)H * W;
  PL_MARK(0);
  // ---- costCalculate: ADCensusCal (stereoMatching.cpp:894-915)
  for (int i = 0; i < 2; i++) {
    SM_TRY(smi_pack_bgr(c, pl->bgr[i], npix, pl->pix[i]));
    if (!pl->have_gray) SM_TRY(sm_bgr2gray(c, pl->bgr[i], H, W, pl->gray[i]));
    SM_TRY(sm_census(c, pl->gray[i], H, W, P.censusFunc, pl->cen[i]));
  }
  PL_MARK(1);
  const int imgNum = P.Do_LRConsis ? 2 : 1;  // stereoMatching.cpp:898
  const int views = (P.Do_refine && P.Do_LRConsis) ? 2 : 1;  // stereoMatching.cpp:1054, 5592
  pl->have_arms = false;
  auto ensure_arms = [&]() -> int { return pl_ensure_arms(pl); };
  bool marked2 = false;
  if (pl->streamA && P.costcalculation != 1 && (P.aggregation == 1 || (P.aggregation == 0 && P.Do_refine && P.Do_regionVote))) {
    // the arm maps depend on the packed images alone: built on a side stream while the cost kernels run (CBCA needs them
    // right after; without aggregation the refinement's region vote is their first reader)
    SM_CUDA(cudaEventRecord(pl->evA0, c->stream));
    SM_CUDA(cudaStreamWaitEvent(pl->streamA, pl->evA0, 0));
    cudaStream_t main_stream = c->stream;
    c->stream = pl->streamA;
    const int rcA = pl_ensure_arms(pl);
    c->stream = main_stream;
    SM_TRY(rcA);
    SM_CUDA(cudaEventRecord(pl->evA1, pl->streamA));
    pl->arms_pending = true;
  }
  if (P.costcalculation == 0) {
    const bool nl_side = pl->streamA && P.aggregation == 2;   // NL frames: the MST build under the cost kernels
    if (nl_side) {   // fork before the cost kernels are queued: the tree needs the image alone
      SM_CUDA(cudaEventRecord(pl->evA0, c->stream));
      SM_CUDA(cudaStreamWaitEvent(pl->streamA, pl->evA0, 0));
    }
    for (int i = 0; i < imgNum; i++)
      SM_TRY(smi_cost_adcensus_packed(c, pl->pix[0], pl->pix[1], pl->cen[0], pl->cen[1], H, W, D, P.censusFunc,
                                      P.adTrunc, P.lamAD, P.lamCen, i, pl->vol[i]));
    if (nl_side) {
      // (tried on top: view 1's sweeps -- vm[1] is final here, NL aggregates vm[0] only -- queued before the tree so that they
      // run under the MST build as well: the MST's ~90 small dependent launches then take 1.2 instead of 0.78 ms, c4 246
      // instead of 260 fps; with the cost kernels alone under it 266 fps.  Under the tree FILTER they measured slower too.)
      PL_MARK(2);   // the cost kernels end here; what is left of the tree build counts as aggregation
      PL_MARK(3);
      marked2 = true;
      // the tree on the side stream (its rounds synchronise the host with THAT stream only; everything above is queued)
      cudaStream_t main_stream = c->stream;
      c->stream = pl->streamA;
      const int rcT = smi_nl_tree(c, pl->bgr[0], H, W);
      c->stream = main_stream;
      SM_TRY(rcT);
      SM_CUDA(cudaEventRecord(pl->evA1, pl->streamA));
      pl->nl_tree_pending = true;
    }
  } else if (P.costcalculation == 2) {
    // "Census": censusCal(vm, 1) (stereoMatching.cpp:975-976, 807-892) = the Hamming volume of both views
    for (int i = 0; i < imgNum; i++) {
      if (pl->u16) SM_TRY(sm_cost_hamming_u16(c, pl->cen[0], pl->cen[1], H, W, D, P.censusFunc, i, (uint16_t*)pl->vol[i]));
      else SM_TRY(sm_cost_hamming(c, pl->cen[0], pl->cen[1], H, W, D, P.censusFunc, i, pl->vol[i]));
    }
  } else {
    // censusGrad (stereoMatching.cpp:25-48): grad() computes the arms first (stereoMatching.cpp:628-631)
    for (int i = 0; i < 2; i++) SM_TRY(sm_grad_xy(c, pl->gray[i], H, W, pl->grad[i][0], pl->grad[i][1]));
    SM_TRY(ensure_arms());
    for (int i = 0; i < imgNum; i++)
      SM_TRY(sm_cost_censusgrad(c, pl->cen[0], pl->cen[1], pl->grad[0][0], pl->grad[0][1], pl->grad[1][0], pl->grad[1][1],
                                pl->arms[i], H, W, D, P.censusFunc, P.cg_lamCen, P.cg_lamG, P.gradTrunc, i, pl->vol[i]));
  }
  if (!marked2) PL_MARK(2);
  // ---- aggregation
  if (pl->arms_pending) {
    SM_CUDA(cudaStreamWaitEvent(c->stream, pl->evA1, 0));
    pl->arms_pending = false;
  }
  if (P.aggregation == 1) {
    SM_TRY(ensure_arms());
    PL_MARK(3);
    const int Lmax = max(1, max(P.cbca_crossL_out, P.cbca_minArmL));
    // a one-level SolveAll (the reference's main() as compiled: PY_LEV = 1) is a plain scale of vm: folded into the
    // last CBCA pass (same two rounded operations per element, one pass over the volume less)
    pl->scale_folded = false;
    float post = 1.0f;
    if (!pl->is_child && !pl->child && P.crossScaleLambda >= 0.f && P.cbca_iterationNum >= 1 && views == (P.Do_refine ? 2 : 1)) {
      SM_TRY(sm_cross_scale_weights(1, P.crossScaleLambda, &post));
      pl->scale_folded = post != 1.0f;
    }
    if (views == 2 && pl->stream2 && pl->vol[3]) {
      // small frames: a pass is a few hundred one-warp blocks marching along their lines (5 per SM at 450x375), so the
      // two views' passes run concurrently, view 1 on the second stream with its own scratch volume
      SM_CUDA(cudaEventRecord(pl->evFork, c->stream));
      SM_CUDA(cudaStreamWaitEvent(pl->stream2, pl->evFork, 0));
      cudaStream_t main_stream = c->stream;
      int rc2 = smi_cbca_packed(c, pl->vol[0], pl->vol[2], pl->armpk[0], pl->armpk[1], H, W, D, P.cbca_iterationNum, 0,
                                Lmax, smi_arm_pad(D), post);
      c->stream = pl->stream2;
      if (rc2 == SM_OK)
        rc2 = smi_cbca_packed(c, pl->vol[1], pl->vol[3], pl->armpk[0], pl->armpk[1], H, W, D, P.cbca_iterationNum, 1,
                              Lmax, smi_arm_pad(D), post);
      c->stream = main_stream;
      SM_TRY(rc2);
      SM_CUDA(cudaEventRecord(pl->evJoin, pl->stream2));
      SM_CUDA(cudaStreamWaitEvent(c->stream, pl->evJoin, 0));
    } else
    for (int i = 0; i < views; i++)
      SM_TRY(smi_cbca_packed(c, pl->vol[i], pl->vol[2], pl->armpk[0], pl->armpk[1], H, W, D, P.cbca_iterationNum, i,
                             Lmax, smi_arm_pad(D), post));
  } else if (P.aggregation == 2) {
    if (!marked2) PL_MARK(3);
    // (tried: view 1's SGM sweeps on the second stream under the MST build and tree filter of view 0 -- its volume is final
    // before the aggregation in NL mode.  Slower: the filter is a latency chain of one CTA per SM, and the sweeps' warps on
    // the same SMs stretch it, 2.93 -> 3.13 ms against 0.14 ms saved.)
    if (pl->nl_tree_pending) {
      SM_CUDA(cudaStreamWaitEvent(c->stream, pl->evA1, 0));
      pl->nl_tree_pending = false;
      SM_TRY(smi_nl_filter(c, pl->vol[0], pl->nlwork, H, W, D));
    } else {
      SM_TRY(smi_nl(c, pl->bgr[0], pl->vol[0], pl->nlwork, H, W, D));
    }
  } else {
    PL_MARK(3);
  }
  return SM_OK;
}

extern "C" int sm_pipeline_run_device(sm_pipeline* pl) {
  SM_CHECK_ARG(pl && !pl->is_child);
  sm_ctx* c = pl->ctx;
  const sm_params& P = pl->p;
  const int H = pl->H, W = pl->W, D = pl->D;
  const long long npix = (long long)H * W;
  SM_CUDA(cudaSetDevice(c->device));
  const int views = (P.Do_refine && P.Do_LRConsis) ? 2 : 1;  // stereoMatching.cpp:1054, 5592
  auto ensure_arms = [&]() -> int { return pl_ensure_arms(pl); };
  pl->scale_folded = false;
  SM_TRY(pl_cost_calculate(pl));
  if (P.crossScaleLambda >= 0.f && !pl->scale_folded) {
    // the caller's pyramid loop + SolveAll (main_.cpp:131-158, stereoMatching.cpp:2142-2208)
    float* vols[2][SM_MAX_PYRAMID];
    int Hs[SM_MAX_PYRAMID], Ws[SM_MAX_PYRAMID], Ds[SM_MAX_PYRAMID], levels = 0;
    for (sm_pipeline* q = pl; q; q = q->child) {
      if (q != pl) {
        // pyrDown of the level above: colour and gray images separately (main_.cpp:145-148)
        sm_pipeline* up = nullptr;
        for (sm_pipeline* r = pl; r != q; r = r->child) up = r;
        for (int i = 0; i < 2; i++) {
          SM_TRY(sm_pyr_down_u8(c, up->bgr[i], up->H, up->W, 3, q->bgr[i]));
          SM_TRY(sm_pyr_down_u8(c, up->gray[i], up->H, up->W, 1, q->gray[i]));
        }
        q->have_gray = true;
        SM_TRY(pl_cost_calculate(q));
      }
      Hs[levels] = q->H; Ws[levels] = q->W; Ds[levels] = q->D;
      vols[0][levels] = q->vol[0]; vols[1][levels] = q->vol[1];
      levels++;
    }
    for (int i = 0; i < (P.Do_refine ? 2 : 1); i++)   // img_n = Do_refine ? 2 : 1 (stereoMatching.cpp:2179)
      SM_TRY(sm_cross_scale(c, vols[i], Hs, Ws, Ds, levels, P.crossScaleLambda));
  }
  PL_MARK(4);
  // ---- dispOptimize: sgm (stereoMatching.cpp:1051-1089) then WTA (:1108-1128)
  if (P.sgm_paths > 0) {
    bool done = false;
    // vm[1]'s finished path sum has no reader after the WTA (the refinement works on DP[] and on vm[0]); unless the caller
    // asks for it (keep_right_volume, vmTop) the last path of view 1 does the WTA without storing the sum: one volume
    // write less per frame
    const bool keep[2] = {true, P.keep_right_volume != 0 || P.Do_vmTop != 0};
    if (pl->u16) {
      for (int i = 0; i < views; i++) {
        SM_TRY(smi_sgm_u16(c, (const uint16_t*)pl->vol[i], pl->pix[i], H, W, D, P.sgm_paths, P.sgm_corDifThres, P.sgm_reduCoeffi1,
                           (uint16_t*)pl->vol[2], pl->disp[i], keep[i], P.sgm_grouped != 0));
        float* t = pl->vol[i];
        pl->vol[i] = pl->vol[2];
        pl->vol[2] = t;
      }
      done = true;
    }
    pl->sweeps[0] = pl->sweeps[1] = false;
    if (!done && P.sgm_paths == 8 && P.sgm_grouped && views == 2 && pl->vol[3]) {
      // both views per sweep launch (two CTAs per SM); falls through to one view at a time if the shape does not fit
      const float* vols[2] = {pl->vol[0], pl->vol[1]};
      const uint32_t* pixs[2] = {pl->pix[0], pl->pix[1]};
      float* sums[2] = {pl->vol[2], pl->vol[3]};
      int16_t* disps[2] = {pl->disp[0], pl->disp[1]};
      if (pl->timing) SM_CUDA(cudaEventRecord(pl->evs[0][0], c->stream));
      const int rc = smi_sgm8_grouped2(c, vols, pixs, H, W, D, P.sgm_corDifThres, P.sgm_reduCoeffi1, sums, disps,
                                       pl->timing ? pl->evs[0][1] : nullptr, &pl->sweeps[0], keep);
      if (rc == SM_OK) {
        for (int i = 0; i < 2; i++) {   // vm[i] <- path sum; the old cost volumes become the scratch
          float* t = pl->vol[i];
          pl->vol[i] = pl->vol[2 + i];
          pl->vol[2 + i] = t;
        }
        done = true;
      } else if (rc != SM_ERR_UNSUPPORTED) {
        return rc;
      }
    }
    if (!done && views == 2 && pl->stream2 && !(P.sgm_paths == 8 && P.sgm_grouped)) {
      // both views at once: view 0 on the ctx stream into vol[2], view 1 on the second stream into vol[3]
      SM_CUDA(cudaEventRecord(pl->evFork, c->stream));
      SM_CUDA(cudaStreamWaitEvent(pl->stream2, pl->evFork, 0));
      cudaStream_t main_stream = c->stream;
      int rc2 = SM_OK;
      for (int i = 0; i < 2 && rc2 == SM_OK; i++) {
        c->stream = i == 0 ? main_stream : pl->stream2;
        for (int k = 0; k < P.sgm_paths && rc2 == SM_OK; k++) {
          const int mode = k == 0 ? 0 : (k == P.sgm_paths - 1 ? (keep[i] ? 2 : 3) : 1);
          rc2 = smi_sgm_path_packed2(c, pl->vol[i], pl->pix[i], H, W, D, k, P.sgm_corDifThres, P.sgm_reduCoeffi1, mode,
                                     pl->vol[2 + i], pl->disp[i]);
        }
      }
      c->stream = main_stream;
      SM_TRY(rc2);
      SM_CUDA(cudaEventRecord(pl->evJoin, pl->stream2));
      SM_CUDA(cudaStreamWaitEvent(c->stream, pl->evJoin, 0));
      for (int i = 0; i < 2; i++) {   // vm[i] <- path sum; the old cost volumes become the scratch
        float* t = pl->vol[i];
        pl->vol[i] = pl->vol[2 + i];
        pl->vol[2 + i] = t;
      }
      done = true;
    }
    for (int i = 0; i < views && !done; i++) {
      pl->sweeps[i] = false;
      if (P.sgm_paths == 8 && P.sgm_grouped) {
        if (pl->timing) SM_CUDA(cudaEventRecord(pl->evs[i][0], c->stream));
        SM_TRY(smi_sgm8_grouped(c, pl->vol[i], pl->pix[i], H, W, D, P.sgm_corDifThres, P.sgm_reduCoeffi1, pl->vol[2],
                                pl->disp[i], pl->timing ? pl->evs[i][1] : nullptr, &pl->sweeps[i], keep[i]));
      } else
      for (int k = 0; k < P.sgm_paths; k++) {
        // the last path also does gen_dispFromVm on the finished sum (saves one read of the volume)
        const int mode = k == 0 ? 0 : (k == P.sgm_paths - 1 ? (keep[i] ? 2 : 3) : 1);
        SM_TRY(smi_sgm_path_packed2(c, pl->vol[i], pl->pix[i], H, W, D, k, P.sgm_corDifThres, P.sgm_reduCoeffi1, mode,
                                    pl->vol[2], pl->disp[i]));
      }
      float* t = pl->vol[i];  // vm[i] <- path sum; the old cost volume becomes the scratch
      pl->vol[i] = pl->vol[2];
      pl->vol[2] = t;
    }
  }
  PL_MARK(5);
  if (P.Do_vmTop) {
    // stereoMatching.cpp:1111-1121: candidates of a clone of vm[i] (sm_select_top_cost leaves the volume alone), then the
    // author's selection; the guidance image is I_c[0] for both views, as in the reference
    for (int i = 0; i < views; i++) {
      SM_TRY(sm_select_top_cost(c, pl->vol[i], H, W, D, P.vmTop_Num, P.vmTop_thres, pl->top));
      SM_TRY(sm_disp_from_top2(c, pl->top, pl->bgr[0], H, W, P.vmTop_Num, P.vmTop_method, P.vmTop_ts, P.vmTop_hasCir2,
                               P.vmTop_cir3_doColorLimit, pl->disp[i]));
    }
  } else if (pl->u16) {
    if (P.sgm_paths == 0)
      for (int i = 0; i < views; i++) SM_TRY(sm_wta_u16(c, (const uint16_t*)pl->vol[i], H, W, D, pl->disp[i]));
  } else if (P.sgm_paths < 2)   // otherwise the WTA was fused into the last SGM path
    for (int i = 0; i < views; i++) SM_TRY(sm_wta(c, pl->vol[i], H, W, D, pl->disp[i]));
  PL_MARK(6);
  // ---- refine (stereoMatching.cpp:1364-1506)
  if (P.Do_refine) {
    if (P.Do_LRConsis) SM_TRY(sm_lrc(c, pl->disp[0], pl->disp[1], H, W, P.LRmaxDiff));
    if (P.Do_regionVote) {
      SM_TRY(ensure_arms());  // :1393-1396
      for (int i = 0; i < P.region_vote_nums; i++)
        SM_TRY(sm_region_vote(c, pl->disp[0], pl->dtmp, pl->arms[0], H, W, D, P.regVote_hratioThres, P.regVote_SThres));
    }
    if (P.Do_properIpol)
      for (int i = 0; i < P.region_vote_nums; i++) SM_TRY(sm_proper_ipol(c, pl->disp[0], pl->dtmp, pl->bgr[0], H, W, P.DISP_OCC));
    if (P.Do_lastMedianBlur) {
      SM_TRY(sm_median3_i16(c, pl->disp[0], pl->dtmp, H, W));
      SM_CUDA(cudaMemcpyAsync(pl->disp[0], pl->dtmp, npix * sizeof(int16_t), cudaMemcpyDeviceToDevice, c->stream));
    }
  }
  PL_MARK(7);
  return SM_OK;
}

extern "C" int sm_pipeline_sgm_split_ms(sm_pipeline* pl, float* out2) {
  SM_CHECK_ARG(pl && out2 && pl->timing && pl->ev_ok);
  SM_CUDA(cudaStreamSynchronize(pl->ctx->stream));
  float sgm = 0.f, sweeps = 0.f;
  SM_CUDA(cudaEventElapsedTime(&sgm, pl->ev[4], pl->ev[5]));
  for (int i = 0; i < 2; i++)
    if (pl->sweeps[i]) {
      float t = 0.f;
      SM_CUDA(cudaEventElapsedTime(&t, pl->evs[i][0], pl->evs[i][1]));
      sweeps += t;
    }
  out2[0] = sweeps;
  out2[1] = sgm - sweeps;
  return SM_OK;
}

extern "C" int sm_pipeline_download(sm_pipeline* pl, int16_t* h_dispL, int16_t* h_dispR) {
  SM_CHECK_ARG(pl && h_dispL);
  sm_ctx* c = pl->ctx;
  const size_t npix = (size_t)pl->H * pl->W;
  const bool pinL = host_is_pinned(h_dispL), pinR = h_dispR && host_is_pinned(h_dispR);
  SM_CUDA(cudaMemcpyAsync(pinL ? h_dispL : pl->h_out, pl->disp[0], npix * 2, cudaMemcpyDeviceToHost, c->stream));
  if (h_dispR)
    SM_CUDA(cudaMemcpyAsync(pinR ? h_dispR : pl->h_out + npix, pl->disp[1], npix * 2, cudaMemcpyDeviceToHost,
                            c->stream));
  SM_CUDA(cudaStreamSynchronize(c->stream));
  if (!pinL) memcpy(h_dispL, pl->h_out, npix * 2);
  if (h_dispR && !pinR) memcpy(h_dispR, pl->h_out + npix, npix * 2);
  return SM_OK;
}

extern "C" int sm_pipeline_run(sm_pipeline* pl, const uint8_t* h_bgrL, const uint8_t* h_bgrR, const uint8_t* h_grayL,
                               const uint8_t* h_grayR, int16_t* h_dispL, int16_t* h_dispR) {
  SM_TRY(sm_pipeline_upload(pl, h_bgrL, h_bgrR, h_grayL, h_grayR));
  SM_TRY(sm_pipeline_run_device(pl));
  return sm_pipeline_download(pl, h_dispL, h_dispR);
}

extern "C" void* sm_pipeline_buffer(sm_pipeline* pl, int which) {
  if (!pl) return nullptr;
  switch (which) {
    case 0: return pl->vol[0];
    case 1: return pl->vol[1];
    case 2: return pl->disp[0];
    case 3: return pl->disp[1];
    case 4: return pl->arms[0];
    case 5: return pl->arms[1];
    case 6: return pl->cen[0];
    case 7: return pl->cen[1];
    default: return nullptr;
  }
}

extern "C" int sm_pipeline_enable_timing(sm_pipeline* pl, int on) {
  SM_CHECK_ARG(pl);
  if (on && !pl->ev_ok) {
    for (int i = 0; i <= ST_COUNT; i++) SM_CUDA(cudaEventCreate(&pl->ev[i]));
    for (int i = 0; i < 4; i++) SM_CUDA(cudaEventCreate(&pl->evs[i / 2][i % 2]));
    pl->ev_ok = true;
  }
  pl->timing = on != 0;
  return SM_OK;
}

extern "C" int sm_pipeline_stage_ms(sm_pipeline* pl, float* out8) {
  SM_CHECK_ARG(pl && out8 && pl->timing && pl->ev_ok);
  SM_CUDA(cudaStreamSynchronize(pl->ctx->stream));
  for (int k = 0; k < 7; k++) SM_CUDA(cudaEventElapsedTime(&out8[k], pl->ev[k], pl->ev[k + 1]));
  SM_CUDA(cudaEventElapsedTime(&out8[7], pl->ev[0], pl->ev[7]));
  return SM_OK;
}
