// finish_common.cuh -- what the finishing reduction (aux_kernels.cuh: finish_loss_kernel) and the single-launch step
// (head_step2.cuh) share: the workspace layout and the composition of the loss block.
#pragma once

#include "common.cuh"

namespace dsnt {

constexpr int kFinishBlock = 256;
constexpr int kFinishMaxCtas = 128;
constexpr int kFinishSlots = 256;       // partial-sum slots in the workspace: also serves the fused step (one slot per SM)
// workspace (floats):
//   [kFinishSlots x 4 partial sums]
//   [end ticket, start ticket, "count published" flag, count over all ranks]      (kFinishCtl: 4 words)
//   [kFinishSlots mask partials of the fused step]
//   [count over this rank's shard, 0, 0, 0]                                         (kFinishLocal)
//   [8 x u64 %globaltimer stamps of the last single-launch step]                    (kFinishTrace; diagnostics only)
constexpr int kFinishCtl = kFinishSlots * 4;
constexpr int kFinishMaskPart = kFinishCtl + 4;
constexpr int kFinishLocal = kFinishMaskPart + kFinishSlots;
constexpr int kFinishTrace = kFinishLocal + 4;
constexpr int kFinishTraceStamps = 8;
constexpr int kFinishWorkspaceFloats = kFinishTrace + 2 * kFinishTraceStamps;

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

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
// NVLink); one warp stores its sums into slot [rank] of EVERY rank's buffer, waits until the slots of all ranks in
// its OWN buffer carry the current epoch, and adds them in rank order -- the same order on every rank, so all ranks
// get bit-identical totals.
//
// Every value travels as ONE naturally aligned 64-bit word {float bits, epoch}: a 64-bit store is single-copy atomic,
// so each word validates itself and the sender needs neither a system-scope fence (a round trip over NVLink before the
// flag could go) nor a separate flag store: one relaxed store per value, seen by the receiver one NVLink hop later.
// Two parities of slots: a rank can be at most one exchange ahead of another (it cannot finish exchange e+1 without
// the other's contribution, sent only after that one finished e), so a word with tag e is never overwritten by e+2
// before everybody has read it.
constexpr int kMaxRanks = DSNT_MAX_RANKS;
constexpr int kPeerWordsPerSlot = 4;                                     // 3 used, padded to 32 bytes
constexpr int kPeerExchangeBytes = 2 * kMaxRanks * kPeerWordsPerSlot * 8;   // two parities
struct PeerXchg {
  unsigned long long* peers[kMaxRanks];   // peers[r]: rank r's exchange buffer (kPeerExchangeBytes, zero before first use)
  unsigned* epoch;            // local device counter of exchanges done so far (zero before first use)
  int* error;                 // local device flag, set when a peer did not show up in time
  int rank, world;
};
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// Called by ONE full warp of one CTA; (a, b, c) are valid on lane 0; returns the totals on every lane.
__device__ __forceinline__ void peer_exchange_sum3(const PeerXchg& x, float& a, float& b, float& c) {
  const int lane = threadIdx.x & 31;
  a = __shfl_sync(kFull, a, 0); b = __shfl_sync(kFull, b, 0); c = __shfl_sync(kFull, c, 0);
  // the counter is read and written through L2: in the single-launch step two different CTAs do the two exchanges of a step
  unsigned e = __shfl_sync(kFull, lane == 0 ? __ldcg(x.epoch) + 1u : 0u, 0);
  if (e == 0u) e = 1u;                                   // tag 0 is the zero-initialised buffer
  const int par = static_cast<int>(e & 1u) * kMaxRanks * kPeerWordsPerSlot;
  const int nwords = 3 * x.world;                        // word w: value k = w % 3 of rank r = w / 3
  // deliver: word (r, k) goes into slot [x.rank] of rank r's buffer
  for (int w = lane; w < nwords; w += 32) {
    const int r = w / 3, k = w - 3 * r;
    const float v = k == 0 ? a : (k == 1 ? b : c);
    st_relaxed_sys_u64(x.peers[r] + par + x.rank * kPeerWordsPerSlot + k,
                       (static_cast<unsigned long long>(e) << 32) | __float_as_uint(v));
  }
  // collect: word (r, k) of my own buffer
  float got[2] = {0.f, 0.f};
  const unsigned long long t0 = global_timer_ns();
#pragma unroll
  for (int rnd = 0; rnd < 2; ++rnd) {
    const int w = lane + 32 * rnd;
    if (w < nwords) {
      const int r = w / 3, k = w - 3 * r;
      const unsigned long long* src = x.peers[x.rank] + par + r * kPeerWordsPerSlot + k;
      unsigned long long word = ld_relaxed_sys_u64(src);
      while (static_cast<unsigned>(word >> 32) != e) {
        if (global_timer_ns() - t0 > 20000000000ull) {    // 20 s: a rank is gone; do not hang the GPU
          word = 0x7fc00000ull;
          *x.error = 1;
          break;
        }
        word = ld_relaxed_sys_u64(src);
      }
      got[rnd] = __uint_as_float(static_cast<unsigned>(word));
    }
  }
  a = 0.f; b = 0.f; c = 0.f;
  for (int r = 0; r < x.world; ++r) {   // rank order: identical totals everywhere
    const int w0 = 3 * r, w1 = w0 + 1, w2 = w0 + 2;
    a += __shfl_sync(kFull, (w0 >> 5) ? got[1] : got[0], w0 & 31);
    b += __shfl_sync(kFull, (w1 >> 5) ? got[1] : got[0], w1 & 31);
    c += __shfl_sync(kFull, (w2 >> 5) ? got[1] : got[0], w2 & 31);
  }
  if (lane == 0) { __stcg(x.epoch, e); __threadfence(); }
}

}  // namespace dsnt
