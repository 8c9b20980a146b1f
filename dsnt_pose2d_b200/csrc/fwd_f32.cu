// Forward instantiations for fp32 heatmaps.
#include "launch.cuh"

namespace dsnt {
int launch_head_fwd_f32(const HeadFwdParams& p, int vec, bool logits, int variant, cudaStream_t stream) {
  if (vec == 4) return logits ? launch_fwd_reg<float, 4, true>(p, variant, stream) : launch_fwd_reg<float, 4, false>(p, variant, stream);
  return logits ? launch_fwd_reg<float, 1, true>(p, variant, stream) : launch_fwd_reg<float, 1, false>(p, variant, stream);
}
}  // namespace dsnt
