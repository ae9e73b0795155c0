// Placeholder used only when the library is built with AVC_NO_TC (tcgen05 kernel compiled out).
#include "common.cuh"
#ifdef AVC_NO_TC
int avc_tc_available(const avc_ctx*) { return 0; }
int avc_tc_eval_avatar(avc_ctx* ctx, const float*, int64_t, const float*, float*, float*, float*, float*, int, int, cudaStream_t) {
  return avc_fail(ctx, AVC_ESTATE, "library built without the tcgen05 kernel");
}
int avc_tc_eval_recon(avc_ctx* ctx, const float*, int64_t, const float*, float*, cudaStream_t) {
  return avc_fail(ctx, AVC_ESTATE, "library built without the tcgen05 kernel");
}
int avc_tc2_eval_avatar(avc_ctx* ctx, const float*, int64_t, const float*, float*, float*, float*, float*, int, int, cudaStream_t) {
  return avc_fail(ctx, AVC_ESTATE, "library built without the tcgen05 kernel");
}
int avc_tc2_eval_recon(avc_ctx* ctx, const float*, int64_t, const float*, float*, cudaStream_t) {
  return avc_fail(ctx, AVC_ESTATE, "library built without the tcgen05 kernel");
}
#endif
