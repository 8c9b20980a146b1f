// finish_common.cuh -- what the finishing reduction (aux_kernels.cuh: finish_loss_kernel) and the single-launch step
// (head_step2.cuh) share: the workspace layout and the composition of the loss block.
#pragma once

#include "common.cuh"

namespace dsnt {

constexpr int kFinishBlock = 256;
constexpr int kFinishMaxCtas = 128;
constexpr int kFinishSlots = 256;       // partial-sum slots in the workspace: also serves the fused step (one slot per SM)
// workspace (floats): [kFinishSlots x 4 partial sums][ticket, mask barrier, 0, 0][kFinishSlots mask partials of the fused step]
constexpr int kFinishWorkspaceFloats = kFinishSlots * 4 + 4 + kFinishSlots;

// out[0..2] = (sum mask*dist, sum mask*D, sum mask) -> the rest of the block (include/dsnt_b200.h: dsnt_finish_loss):
// masked_average (src/dsnt/nn.py:81-94) and loss = euclid + reg_coeff * reg (src/dsnt/model.py:145)
__device__ __forceinline__ void write_loss_tail(float* out, float reg_coeff) {
  const float cnt = out[2];
  const float den = fmaxf(cnt, 1.0f);
  const float eu = out[0] / den, rg = out[1] / den;
  out[3] = den; out[4] = eu; out[5] = rg; out[6] = fmaf(reg_coeff, rg, eu); out[7] = 0.f;
}

}  // namespace dsnt
