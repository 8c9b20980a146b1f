// Streaming backward for the other pre-activations, bf16 heatmaps (head_stream.cuh with PA != softmax).
#include "launch.cuh"

namespace dsnt {
int launch_preact_fast_bwd_bf16(const HeadBwdParams& p, const PreactCfg& pc, int vec, cudaStream_t stream) {
  if (vec == 8) return launch_preact_bwd_fast<__nv_bfloat16, 8>(p, pc, stream);
  if (vec == 4) return launch_preact_bwd_fast<__nv_bfloat16, 4>(p, pc, stream);
  return 1;
}
}  // namespace dsnt
