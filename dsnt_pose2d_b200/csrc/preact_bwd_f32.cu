// Streaming backward for the other pre-activations, fp32 heatmaps (head_stream.cuh with PA != softmax).
#include "launch.cuh"

namespace dsnt {
int launch_preact_fast_bwd_f32(const HeadBwdParams& p, const PreactCfg& pc, int vec, cudaStream_t stream) {
  return vec == 4 ? launch_preact_bwd_fast<float, 4>(p, pc, stream) : 1;
}
}  // namespace dsnt
