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

// ------------------------------------------------------------------------------------------------
// Exchange of the three partial sums between the ranks of one node, INSIDE the finishing kernel (no NCCL launch):
// every rank owns an exchange buffer that all ranks of the group have mapped (torch symmetric memory: P2P over
// NVLink); the CTA that holds the last ticket stores its sums into slot [rank] of EVERY rank's buffer, waits until
// the slots of all ranks in its OWN buffer carry the current epoch, and adds them in rank order -- the same order on
// every rank, so all ranks get bit-identical totals.  Two parities of slots: a rank can be at most one exchange ahead
// of another (it cannot finish exchange e+1 without the other's contribution, sent only after that one finished e).
constexpr int kMaxRanks = DSNT_MAX_RANKS;
struct PeerXchg {
  float4* peers[kMaxRanks];   // peers[r]: rank r's exchange buffer, 2 * kMaxRanks float4 slots, zero before first use
  unsigned* epoch;            // local device counter of exchanges done so far (zero before first use)
  int* error;                 // local device flag, set when a peer did not show up in time
  int rank, world;
};
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Called by the first warp of one CTA; (a, b, c) are valid on lane 0; returns the totals on every lane.
__device__ __forceinline__ void peer_exchange_sum3(const PeerXchg& x, float& a, float& b, float& c) {
  const int lane = threadIdx.x & 31;
  a = __shfl_sync(kFull, a, 0); b = __shfl_sync(kFull, b, 0); c = __shfl_sync(kFull, c, 0);
  // the counter is read and written through L2: in the single-launch step two different CTAs do the two exchanges of a step
  const unsigned e = __shfl_sync(kFull, lane == 0 ? __ldcg(x.epoch) + 1u : 0u, 0);
  const int par = static_cast<int>(e & 1u) * kMaxRanks;
  if (lane < x.world) {            // lane r delivers to rank r
    float4* dst = x.peers[lane] + par + x.rank;
    float* d = reinterpret_cast<float*>(dst);
    d[0] = a; d[1] = b; d[2] = c;
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned*>(d) + 3, e);
  }
  float va = 0.f, vb = 0.f, vc = 0.f;
  if (lane < x.world) {            // lane r collects from rank r
    const float4* src = x.peers[x.rank] + par + lane;
    const unsigned* flag = reinterpret_cast<const unsigned*>(src) + 3;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    bool ok = true;
    while (ld_acquire_sys(flag) != e) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) { ok = false; break; }     // 20 s: a rank is gone; do not hang the GPU
    }
    const volatile float* sv = reinterpret_cast<const volatile float*>(src);
    if (ok) { va = sv[0]; vb = sv[1]; vc = sv[2]; }
    else { va = vb = vc = __int_as_float(0x7fc00000); *x.error = 1; }
  }
  a = 0.f; b = 0.f; c = 0.f;
  for (int r = 0; r < x.world; ++r) {   // rank order: identical totals everywhere
    a += __shfl_sync(kFull, va, r); b += __shfl_sync(kFull, vb, r); c += __shfl_sync(kFull, vc, r);
  }
  if (lane == 0) { __stcg(x.epoch, e); __threadfence(); }
}



}  // namespace dsnt
