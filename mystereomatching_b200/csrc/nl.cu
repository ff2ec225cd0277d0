// K8 + K9 placeholder, replaced by the real implementation in this round.
#include "common.cuh"
extern "C" int sm_mst_build(sm_ctx*, const uint8_t*, int, int, int, int32_t*, uint8_t*, int32_t*, int32_t*) {
  sm_set_error("sm_mst_build: not built yet"); return SM_ERR_UNSUPPORTED; }
extern "C" int sm_tree_filter(sm_ctx*, float*, double*, int, int, int, const int32_t*, const uint8_t*, const int32_t*,
                              const int32_t*, double) { sm_set_error("sm_tree_filter: not built yet"); return SM_ERR_UNSUPPORTED; }
extern "C" int sm_nl(sm_ctx*, const uint8_t*, float*, int, int, int) { sm_set_error("sm_nl: not built yet"); return SM_ERR_UNSUPPORTED; }
int smi_nl(sm_ctx*, const uint8_t*, float*, double*, int, int, int) { sm_set_error("nl: not built yet"); return SM_ERR_UNSUPPORTED; }
