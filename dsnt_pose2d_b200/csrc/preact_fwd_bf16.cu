// Tuned forward for the other pre-activations, bf16 heatmaps (head_fast.cuh with PA != softmax).
#include "launch.cuh"

namespace dsnt {
int launch_preact_fast_fwd_bf16(const HeadFwdParams& p, const PreactCfg& pc, int vec, cudaStream_t stream) {
  return vec == 8 ? launch_preact_fwd_fast<__nv_bfloat16, 8>(p, pc, stream) : 1;
}
}  // namespace dsnt
