// Tuned forward for the other pre-activations, fp32 heatmaps (head_fast.cuh with PA != softmax).
#include "launch.cuh"

namespace dsnt {
int launch_preact_fast_fwd_f32(const HeadFwdParams& p, const PreactCfg& pc, int vec, cudaStream_t stream) {
  return vec == 4 ? launch_preact_fwd_fast<float, 4>(p, pc, stream) : 1;
}
}  // namespace dsnt
