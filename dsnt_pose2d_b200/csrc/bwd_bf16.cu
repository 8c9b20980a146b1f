// Backward instantiations for bf16 heatmaps (dz is written in bf16, round-to-nearest-even).
#include "launch.cuh"

namespace dsnt {
int launch_head_bwd_bf16(const HeadBwdParams& p, int vec, bool logits, int variant, cudaStream_t stream) {
  using T = __nv_bfloat16;
  if (vec == 8) return logits ? launch_bwd_reg<T, 8, true>(p, variant, stream) : launch_bwd_reg<T, 8, false>(p, variant, stream);
  if (vec == 4) return logits ? launch_bwd_reg<T, 4, true>(p, variant, stream) : launch_bwd_reg<T, 4, false>(p, variant, stream);
  return logits ? launch_bwd_reg<T, 1, true>(p, variant, stream) : launch_bwd_reg<T, 1, false>(p, variant, stream);
}
}  // namespace dsnt
