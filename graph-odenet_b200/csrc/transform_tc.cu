// libgode: tensor-core (tcgen05) path of the dense transform S = [t || GroupNorm(y)] W.
// Placeholder in this commit: reports "unsupported" so odefunc.cu takes the SIMT path.
#include "internal.cuh"

namespace gode {
bool transform_tc_supported(const gode_gcn_odefunc_t* f) {
  (void)f;
  return false;
}
int transform_tc(const gode_gcn_odefunc_t* f, const float* y, float t, float* S, cudaStream_t st) {
  (void)f; (void)y; (void)t; (void)S; (void)st;
  set_error("transform_tc: not built");
  return GODE_EINVAL;
}
}  // namespace gode
