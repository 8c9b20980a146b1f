// head_step2.cuh -- the one-pass training step (head_step.cuh) specialised at COMPILE TIME for one heatmap shape and
// written on packed fp32 pairs.
//
// ncu on the generic step kernel with bf16 heatmaps (profiles/r01_v6_step_bf16_*): 70 % issue-active, XU 64 %, DRAM 39 %,
// 28 instructions per pixel for JS -- 17.4 in the three sweeps, 4.9 per-heatmap scalar work, 3.0 in the per-pixel window
// loop and 2.3 in the "heavy" rows of the backward (rcp + lg2 on every pixel of every row that touches the Gaussian
// window, although only a third of their columns do).  The memory system idles because the SM cannot issue faster.
// This kernel removes instructions instead of adding warps:
//
//  * H, W are template parameters: every sweep is fully unrolled, shared-memory and global addresses are the lane base
//    plus an immediate, row coordinates fold into constants -- no loop counters, compares, branches, address arithmetic;
//  * FADD2 / FMUL2 / FFMA2 (sm_100 packed fp32) process two neighbouring pixels per issue slot;
//  * bf16: the maximum runs on packed bf16 pairs (HMNMX2.BF16_V2) without unpacking; the warp maximum is one
//    CREDUX.MAX.F32;
//  * the softmax sums are per-COLUMN accumulators (a lane always sees the same columns) + one row sum per sweep step, so
//    S, S_x, S_y -- and the variance about the mean, without a second sweep -- come out of the same two adds per pixel;
//  * the Gaussian window is processed as whole 128-bit vectors in a COMPACT lane mapping (vector k of the window -> lane
//    k mod 32), forward and backward: log2 M per window pixel is computed once, kept in registers across the
//    reduction, and reused by the backward (log2(1 + G/P) = log2 2M - log2 P), so the backward needs no rcp / lg2 at
//    all and the main backward sweep has no window branch.
//
//  * e = 2^(z log2e - max) is computed ONCE: the sum sweep writes it back over z in the shared-memory buffer (fp32 heatmaps:
//    exactly; bf16 heatmaps: as fp16 of e * 2^15, 11 significant bits for e >= 2e-9, against the 8 of the bf16 dz it
//    produces) and the backward sweep is LDS, add, mul, store -- the profile of the first version of this kernel showed
//    the XU pipe (MUFU.EX2) at 76 % with DRAM at 67 %, i.e. the second exponential per pixel was what bound it.  The
//    window vectors are left alone (the window pass needs the logits themselves);
//  * bulk loads are PACED (pace_reserve / PendingLoad below): a slot schedule per CTA breaks the convoy in which all warps
//    wait for data, finish together and re-issue loads and stores together (copy through the ring: 0.91 -> 1.01 of the
//    measured HBM peak; fp32 JS step 357 -> 334 us);
//  * single-launch form (p.out8): the mask count (CTA slices + grid barrier) and masked_average + loss composition (last
//    ticket) run inside this kernel; with p.xc.world > 1 both cross the ranks of a sharded batch through peer memory;
//    p.st.count > 1 walks the stacks of a stacked hourglass.
//
// Same mathematics and the same closed forms as head_step.cuh (src/dsnt/nn.py:25-116,168-298,
// src/dsnt/model.py:24-63,145); the ring of shared-memory buffers, the bulk loads and the barrier protocol are shared
// with it.  KL walks its window (28x28 at sigma = 1 px) slot by slot instead of keeping it in registers.
#pragma once

#include "finish_common.cuh"
#include "f32x2.cuh"
#include "head_step.cuh"

namespace dsnt {

template <typename T>
__device__ __forceinline__ void unpack_pairs(const uint4& r, f2 (&v)[8 / sizeof(T)]) {
  if constexpr (sizeof(T) == 4) {
    asm("mov.b64 %0, {%1, %2};" : "=l"(v[0].r) : "r"(r.x), "r"(r.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(v[1].r) : "r"(r.z), "r"(r.w));
  } else {
    v[0] = bf16x2_to_f2(r.x); v[1] = bf16x2_to_f2(r.y); v[2] = bf16x2_to_f2(r.z); v[3] = bf16x2_to_f2(r.w);
  }
}
template <typename T>
__device__ __forceinline__ uint4 pack_pairs(const f2 (&o)[8 / sizeof(T)]) {
  uint4 r;
  if constexpr (sizeof(T) == 4) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(r.x), "=r"(r.y) : "l"(o[0].r));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(r.z), "=r"(r.w) : "l"(o[1].r));
  } else {
    float lo, hi;
    upk(o[0], lo, hi); r.x = pack_bf16(lo, hi);
    upk(o[1], lo, hi); r.y = pack_bf16(lo, hi);
    upk(o[2], lo, hi); r.z = pack_bf16(lo, hi);
    upk(o[3], lo, hi); r.w = pack_bf16(lo, hi);
  }
  return r;
}

// e stash: what the sum sweep leaves in the buffer for the backward sweep (see the header)
template <typename T>
__device__ __forceinline__ uint4 stash_pack(const f2 (&e)[8 / sizeof(T)]) {
  uint4 r;
  if constexpr (sizeof(T) == 4) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(r.x), "=r"(r.y) : "l"(e[0].r));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(r.z), "=r"(r.w) : "l"(e[1].r));
  } else {
    float lo, hi;
    upk(e[0], lo, hi); asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r.x) : "f"(hi), "f"(lo));
    upk(e[1], lo, hi); asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r.y) : "f"(hi), "f"(lo));
    upk(e[2], lo, hi); asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r.z) : "f"(hi), "f"(lo));
    upk(e[3], lo, hi); asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r.w) : "f"(hi), "f"(lo));
  }
  return r;
}
__device__ __forceinline__ f2 f16x2_to_f2(uint32_t w) {
  float lo, hi;
  asm("{\n.reg .b16 l, h;\nmov.b32 {l, h}, %2;\ncvt.f32.f16 %0, l;\ncvt.f32.f16 %1, h;\n}" : "=f"(lo), "=f"(hi) : "r"(w));
  return pk(lo, hi);
}
template <typename T>
__device__ __forceinline__ void stash_unpack(const uint4& r, f2 (&e)[8 / sizeof(T)]) {
  if constexpr (sizeof(T) == 4) {
    asm("mov.b64 %0, {%1, %2};" : "=l"(e[0].r) : "r"(r.x), "r"(r.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(e[1].r) : "r"(r.z), "r"(r.w));
  } else {
    e[0] = f16x2_to_f2(r.x); e[1] = f16x2_to_f2(r.y); e[2] = f16x2_to_f2(r.z); e[3] = f16x2_to_f2(r.w);
  }
}

// The sweeps are fully unrolled (static register indices, immediate address offsets).  Left alone, ptxas hoists ALL of
// a sweep's shared-memory loads to its top and spills (an empty asm fence disappears before ptxas schedules); a
// __syncwarp() every kChunk steps is a real instruction with memory ordering that it will not move loads across, and
// costs one issue slot per 4 * VEC pixels on a converged warp.
// y0 is a per-lane constant, so "y0 + it * dyi" is invariant across heatmaps and would be hoisted out of the tile loop
// for every sweep step (16-32 live registers); opaque() makes a copy the optimiser cannot see through.
__device__ __forceinline__ float opaque(float v) {
  asm volatile("" : "+f"(v));
  return v;
}
constexpr int kChunk = 4;
__device__ __forceinline__ void sweep_fence(int it) {
  if ((it % kChunk) == kChunk - 1) __syncwarp();
}

// axis_window (head_stream.cuh) with the approximate square root: a window one pixel larger or smaller than the exact
// one is equally valid (it is padded by a pixel and any superset of it gives the same sums to the tolerance theta).
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void axis_window_fast(float t, int n, float half_n, float two_over_n, float bias_n, float r2,
                                                 int& lo, int& hi) {
  const float nm1 = static_cast<float>(n - 1);
  const float js = fminf(fmaxf(rintf(fmaf(t + 1.0f, half_n, -0.5f)), 0.f), nm1);
  const float ds = fmaf(js, two_over_n, bias_n) - t;
  const float R = sqrt_approx(fmaf(ds, ds, r2)) * 1.000001f;
  const float flo = fmaxf(ceilf(fmaf(t - R + 1.0f, half_n, -0.5f)) - 1.0f, 0.f);
  const float fhi = fminf(floorf(fmaf(t + R + 1.0f, half_n, -0.5f)) + 1.0f, nm1);
  lo = static_cast<int>(flo);
  hi = static_cast<int>(fhi);
}

// Load pacing: every CTA reserves a time slot for each bulk load it issues, p.pace clocks after the previous one.  When
// the warps are faster than memory they all wait for data, finish together when a burst of heatmaps arrives and
// re-issue their loads (and their stores) together, which comes back as the next burst: measured on B200 a pure
// copy through the ring runs at 0.91 of the HBM peak unpaced and at 1.0 paced (profiles/r01_v7_step2_sweeps.txt).
// A load whose slot lies in the future is kept PENDING by lane 0 of the warp, which goes on with its next heatmap and
// issues the load at one of its checkpoints -- or just before it would block on a barrier.
struct PendingLoad {       // one per warp, in shared memory (only lane 0 touches it): tile < 0 means none
  uint32_t when;           // low 32 bits of clock64 at which it may go (compared as a signed difference)
  int tile;
  uint32_t slot;           // buffer index | round << 8
};
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Arms the barrier and starts the bulk load of one heatmap.  Deliberately NOT inlined: it is reached from several places
// of a fully unrolled kernel whose size matters (the instruction stream is walked once per heatmap).
static __device__ __noinline__ void ring_issue(const char* src, uint32_t hm_bytes, uint32_t bar, uint32_t bufs, volatile int* flag, int value) {
  mbar_expect_tx(bar, hm_bytes);
  bulk_load(bufs, src, hm_bytes, bar);
  __threadfence_block();
  *flag = value;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// reserve the next slot; returns the clock at which the load may be issued (<= now: at once)
__device__ __forceinline__ unsigned long long pace_reserve(unsigned long long* slot, int pace, unsigned long long now) {
  unsigned long long mine = atomicAdd(slot, static_cast<unsigned long long>(pace));
  if (mine < now) {                      // the schedule fell behind (start-up, compute-bound phases): restart it from now
    atomicMax(slot, now + pace);
    mine = now;
  }
  return mine;
}

// window slots per lane: the compact mapping covers 32 * MAXS vectors (host-checked bound, step2_window_fits).  Two builds
// per dtype: the small one serves sigma = 1 px (the reference's default: 16 x 16 pixels = 48 bf16 / 80 fp32 vectors), the
// large one up to ~1.4 px; its window arrays are what sets the register count of the kernel.
template <typename T>
__host__ __device__ constexpr int step2_slots_small() { return sizeof(T) == 2 ? 2 : 3; }
template <typename T>
__host__ __device__ constexpr int step2_slots_large() { return sizeof(T) == 2 ? 3 : 4; }

// PACED = false compiles the pacing out (loads are re-issued the moment a buffer is free): for the instantiations that
// are bound by arithmetic anyway (bf16 with a Gaussian window), where the bookkeeping costs more than it brings.
template <typename T, int REG, int H, int W, int NWMAX, bool PACED, int MAXS>
__global__ void __launch_bounds__(NWMAX * 32, 1) head_step2_kernel(const HeadStepParams p) {
  constexpr bool kJS = REG == DSNT_REG_JS;
  constexpr bool kVar = REG == DSNT_REG_VAR;
  constexpr bool kMSE = REG == DSNT_REG_MSE;
  constexpr bool kKL = REG == DSNT_REG_KL;
  constexpr bool kWin = kJS || kMSE || kKL;      // a Gaussian window exists
  constexpr bool kWinRegs = kJS || kMSE;         // ... and its per-pixel terms are kept in registers between forward and backward
  // KL: G is compared with eps = 1e-24, so its window is +-12 sigma (27 x 27 pixels at sigma = 1 px: 7 fp32 vectors per
  // lane).  Nothing is kept: the window is walked slot by slot in the forward (sum P log2(G + eps)) and again in the
  // backward, logits re-read from the shared-memory buffer -- which therefore carries no e stash (every pixel's gradient
  // needs log2 P = t - log2 S, i.e. the logit itself) -- and ANY sigma is served (the loop has a run-time trip count).
  constexpr int VEC = 16 / sizeof(T);      // pixels per 128-bit vector
  constexpr int NP = VEC / 2;              // packed pairs per vector
  constexpr int WV = W / VEC;              // vectors per row
  static_assert(W % VEC == 0 && 32 % WV == 0, "a warp must cover whole rows");
  constexpr int RPI = 32 / WV;             // rows per sweep step
  static_assert(H % RPI == 0, "whole sweep steps");
  constexpr int ITERS = H / RPI;
  constexpr uint32_t hm_bytes = static_cast<uint32_t>(H) * W * sizeof(T);
  constexpr float tow = 2.0f / W, bw = 1.0f / W - 1.0f, toh = 2.0f / H, bh = 1.0f / H - 1.0f;
  constexpr float dyi = RPI * toh;
  // every e below is 2^(z log2e - m2 + kBias): bf16 heatmaps stash e as fp16, whose range [6e-8, 65504] is centred on
  // the values by the bias; sums, 1/S and the probabilities P = e / S carry the factor consistently
  constexpr float kBias = sizeof(T) == 2 ? 15.0f : 0.0f;
  // MSE feeds e into its own gradient factor (2 rho (P - G) next to a x + b y - c): where the two nearly cancel, the 2^-12
  // of an fp16 e would be amplified, so bf16 + MSE recomputes e from the logits in the backward sweep instead
  constexpr bool kStash = !(kMSE && sizeof(T) == 2) && !kKL;

  extern __shared__ __align__(128) unsigned char step_smem[];
  __shared__ __align__(8) unsigned long long bars[kStepMaxBufs];
  __shared__ volatile int issued[kStepMaxBufs];   // see head_step.cuh: loads issued into each buffer so far
  __shared__ unsigned long long pace_next;        // clock at which the CTA may issue its next bulk load (p.pace)
  __shared__ PendingLoad pend_all[NWMAX];
  __shared__ float fin[NWMAX][2];                  // single-launch form: per-warp sums of mask * (distance, divergence)
  __shared__ float fin_mask[NWMAX];                // ... and of the mask slice at the start.  NOT fin[]: a warp without heatmaps
                                                   // is at the end of the kernel (writing fin) while warp 0 still reads these
                                                   // (found by compute-sanitizer --tool racecheck, profiles/r02_sanitizer_*)
  __shared__ bool fin_last;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= p.nwarps) return;
  const Geom& g = p.g;
  const int NW = p.nwarps, NB = p.nbufs;
  const uint32_t smem0_s = smem_u32(step_smem), bars0_s = smem_u32(&bars[0]);
  const char* zsrc = static_cast<const char*>(p.z);
  char* dzdst = static_cast<char*>(p.dz);
  // tiles of this CTA, interleaved across the CTAs: heatmap = tile * gridDim.x + blockIdx.x (one contiguous range per CTA
  // was measured slower, profiles/r01_v7_step2_sweeps.txt)
  const long hm_mul = static_cast<long>(gridDim.x), hm_add = static_cast<long>(blockIdx.x);
  const long ntiles = p.n > blockIdx.x ? (p.n - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (threadIdx.x == 0) {
    pace_next = 0;
    if (!p.denom && blockIdx.x == 0) reinterpret_cast<unsigned long long*>(p.ws + kFinishTrace)[5] = global_timer_ns();   // kernel entry
  }
  if (lane == 0) {
    for (int t = warp; t < NB; t += NW) {
      const uint32_t bar = bars0_s + 8 * t;
      mbar_init(bar, 1);
      if (t < ntiles) {
        mbar_expect_tx(bar, hm_bytes);
        bulk_load(smem0_s + static_cast<uint32_t>(t) * p.buf_bytes, zsrc + locate(p.st, t * hm_mul + hm_add, hm_bytes).z_bytes,
                  hm_bytes, bar);
      }
      issued[t] = 1;
    }
  }
  // Single-launch form: no dsnt_mask_count before this kernel, and NO grid barrier.  The denominator of masked_average
  // (src/dsnt/nn.py:88-92) scales the gradient only, so the forward of a warp's first heatmap does not need it.  Every CTA
  // adds up ITS slice of the mask, parks the partial in the workspace and draws a ticket -- nobody waits; the CTA with the
  // last ticket adds the partials in index order (sharded batch: exchanges the sum with the other ranks through peer
  // memory) and publishes the count; every warp picks it up just before its first backward (need_count below).  The
  // global round trips -- and, sharded, the NVLink hop and the wait for a rank that started later -- are hidden behind the
  // first bulk loads and the first forward.  Co-residency of the grid (one CTA per SM) is guaranteed by the cooperative
  // launch attribute (step.cu).
  float mask_count = 0.f;
  bool have_count = p.denom != nullptr;
  const bool sharded = p.xc.world > 1;
  unsigned* const ctl = reinterpret_cast<unsigned*>(p.ws + kFinishCtl);   // [end ticket, start ticket, published flag, count]
  unsigned long long* const trace = reinterpret_cast<unsigned long long*>(p.ws + kFinishTrace);
  const long nmask = p.st.count > 1 ? p.st.n_per : p.n;       // the stacks share one mask
  if (!p.denom && blockIdx.x == 0 && threadIdx.x == 0) trace[6] = global_timer_ns();     // barriers initialised, first loads issued
  // Only the first kCountCtas CTAs add up the mask: they are the first to be launched, so the count does not wait for the
  // last of 148 CTAs to get going (measured: the local count was there 10 us after kernel entry with all CTAs taking part).
  constexpr unsigned kCountCtas = 16;
  const unsigned nct = gridDim.x < kCountCtas ? gridDim.x : kCountCtas;
  if (!p.denom && p.mask && blockIdx.x < nct) {
    // 128-bit loads, all of a thread's in flight at once: this is the first touch of the mask by this kernel (cold TLB, HBM),
    // and every dependent round of loads costs another ~2 us before the count can go out
    const long chunk = ((nmask + nct - 1) / nct + 3) & ~3L;
    const long lo = blockIdx.x * chunk, hi = lo + chunk < nmask ? lo + chunk : nmask;
    const long bd = blockDim.x;
    float s0 = 0.f, s1 = 0.f;
    long i = lo + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(p.mask) & 15) == 0 && hi > lo) {
      const float4* m4 = reinterpret_cast<const float4*>(p.mask + lo);
      const long n4 = (hi - lo) >> 2;
      float4 v[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const long idx = threadIdx.x + k * bd;
        v[k] = idx < n4 ? __ldg(m4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) { s0 += v[k].x + v[k].y; s1 += v[k].z + v[k].w; }
      for (long idx = threadIdx.x + 3 * bd; idx < n4; idx += bd) {
        const float4 w = __ldg(m4 + idx);
        s0 += w.x + w.y; s1 += w.z + w.w;
      }
      i = lo + 4 * n4 + threadIdx.x;
    }
    for (; i < hi; i += bd) s0 += __ldg(p.mask + i);
    const float sm = warp_sum(s0 + s1);
    if (lane == 0) fin_mask[warp] = sm;
  }
  __syncthreads();           // barriers initialised, issued[] and pace_next set, mask partials in fin[]
  if (!p.denom) {
    if (blockIdx.x == 0 && threadIdx.x == 0) trace[0] = global_timer_ns();
    if (warp == 0) {
      bool publish = blockIdx.x == 0;                            // no mask: every heatmap counts, CTA 0 says so
      float tot = static_cast<float>(nmask);
      if (p.mask) {
        float* mpart = p.ws + kFinishMaskPart;
        unsigned last = 0u;
        if (lane == 0 && blockIdx.x < nct) {
          float t2 = 0.f;
          for (int w2 = 0; w2 < p.nwarps; ++w2) t2 += fin_mask[w2];
          __stcg(mpart + blockIdx.x, t2);
          __threadfence();
          last = atomicAdd(ctl + 1, 1u) == nct - 1 ? 1u : 0u;
        }
        publish = __shfl_sync(kFull, last, 0) != 0u;
        if (publish) {
          __threadfence();
          static_assert(kCountCtas <= 32, "one partial per lane");
          tot = lane < static_cast<int>(nct) ? __ldcg(mpart + lane) : 0.f;
          tot = warp_sum(tot);
        }
      }
      if (publish) {
        if (lane == 0) { __stcg(p.ws + kFinishLocal, tot); trace[1] = global_timer_ns(); }
        if (sharded) {
          // the count over ALL ranks.  Slot layout as in finish_loss_kernel (sum mask*dist, sum mask*D, sum mask), so that a
          // rank on the three-launch form (an empty shard) meets the others in the same exchanges.
          float z0 = 0.f, z1 = 0.f;
          peer_exchange_sum3(p.xc, z0, z1, tot);
        }
        if (lane == 0) {
          __stcg(reinterpret_cast<float*>(ctl + 3), tot);
          __threadfence();
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ctl + 2), "r"(1u) : "memory");
          trace[2] = global_timer_ns();
        }
      }
    }
  }
  // the count, when a warp first needs it (lane 0 polls the flag in L2; it is long there in the steady case)
  auto need_count = [&]() {
    if (have_count) return;
    float tot = 0.f;
    if (lane == 0) {
      while (ld_acquire_gpu(ctl + 2) == 0u) __nanosleep(100);
      tot = __ldcg(reinterpret_cast<const float*>(ctl + 3));
    }
    mask_count = __shfl_sync(kFull, tot, 0);
    have_count = true;
  };
  if (p.stagger_ns > 0) __nanosleep(static_cast<unsigned>(warp * p.stagger_ns + (blockIdx.x & 3) * (p.stagger_ns >> 2)));
  const float gl = p.g_loss ? __ldg(p.g_loss) : 1.0f;
  float inv_denom = p.denom ? 1.0f / __ldg(p.denom) : 0.f;     // single-launch form: set by the first need_count()
  float acc_d = 0.f, acc_r = 0.f;       // this warp's sums of mask * distance, mask * divergence (single-launch form)
  const float s2 = p.sigma * p.sigma;

  // lane geometry: vector column cv (pixels cv*VEC ...), rows r0, r0 + RPI, ...
  const int cv = lane & (WV - 1), r0 = lane / WV;
  float xs[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) xs[c] = fmaf(static_cast<float>(cv * VEC + c), tow, bw);
  const float y0 = fmaf(static_cast<float>(r0), toh, bh);
  const f2 l2e2 = pk1(kLog2e);

  volatile PendingLoad& pend = pend_all[warp];
  if (lane == 0) pend.tile = -1;
  __syncwarp();
  auto issue_load = [&](long tile, uint32_t b, uint32_t rnd) {
    ring_issue(zsrc + locate(p.st, tile * hm_mul + hm_add, hm_bytes).z_bytes, hm_bytes, bars0_s + 8 * b, smem0_s + b * static_cast<uint32_t>(p.buf_bytes),
               &issued[b], static_cast<int>(rnd) + 2);
  };
  // lane 0: the pending load goes if its time has come (force: wait for it)
  auto flush_pending = [&](bool force) {
    if constexpr (!PACED) return;
    if (lane == 0 && pend.tile >= 0) {
      if (force) { while (static_cast<int>(static_cast<uint32_t>(clock64()) - pend.when) < 0) { } }
      if (force || static_cast<int>(static_cast<uint32_t>(clock64()) - pend.when) >= 0) {
        issue_load(pend.tile, pend.slot & 0xffu, pend.slot >> 8);
        pend.tile = -1;
      }
    }
  };

  // tile t lives in buffer bi = t mod NB on its round-th use (NW <= NB: bi wraps at most once per step)
  uint32_t round = 0, bi = static_cast<uint32_t>(warp);
  const int nt32 = static_cast<int>(ntiles);
  // target and mask of the NEXT heatmap are fetched one heatmap ahead (the global-load latency was 12 % of the stall
  // samples at the top of the loop, profiles/r01_v7_*)
  float2 tgt_next = make_float2(0.f, 0.f);
  float msk_next = 1.0f;
  if (warp < nt32) {
    const long hm0 = locate(p.st, warp * hm_mul + hm_add, hm_bytes).nl;
    if (p.target) tgt_next = __ldg(reinterpret_cast<const float2*>(p.target) + hm0);
    if (p.mask) msk_next = __ldg(p.mask + hm0);
  }
  for (int t = warp; t < nt32; t += NW, bi += NW) {
    if (bi >= static_cast<uint32_t>(NB)) { bi -= NB; ++round; }
    const long hm = t * hm_mul + hm_add;
    const uint32_t phase = round & 1u;
    unsigned char* buf = step_smem + static_cast<size_t>(bi) * p.buf_bytes;
    const uint32_t buf_s = smem0_s + bi * static_cast<uint32_t>(p.buf_bytes), bar_s = bars0_s + 8 * bi;
    const uint4* bufv = reinterpret_cast<const uint4*>(buf);
    const uint4* bv = bufv + lane;           // sweep step it reads bv[32 * it]
    uint4* bst = reinterpret_cast<uint4*>(buf) + lane;   // the stash goes where the vector came from
    uint4* dzv = reinterpret_cast<uint4*>(dzdst + locate(p.st, hm, hm_bytes).dz_bytes);

    const float tx = tgt_next.x, ty = tgt_next.y;
    const float mraw = msk_next;
    if (t + NW < nt32) {
      const long hmn = locate(p.st, (t + NW) * hm_mul + hm_add, hm_bytes).nl;
      if (p.target) tgt_next = __ldg(reinterpret_cast<const float2*>(p.target) + hmn);
      if (p.mask) msk_next = __ldg(p.mask + hmn);
    }
    // the Gaussian window, widened to whole vectors along x
    int i_lo = 0, jv_lo = 0, nvw = 0, nwv = 0, j_lo = 0, j_hi = -1, i_hi = -1;
    if constexpr (kWin) {
      Window win;
      axis_window_fast(tx, W, 0.5f * W, tow, bw, g.r2_win, win.j_lo, win.j_hi);
      axis_window_fast(ty, H, 0.5f * H, toh, bh, g.r2_win, win.i_lo, win.i_hi);
      if (!win.empty()) {
        j_lo = win.j_lo; j_hi = win.j_hi; i_lo = win.i_lo; i_hi = win.i_hi;
        jv_lo = win.j_lo / VEC;
        nvw = win.j_hi / VEC - jv_lo + 1;
        nwv = (win.i_hi - win.i_lo + 1) * nvw;
      }
    }

    // about to block on this heatmap's barrier?  then this warp's pending load must not wait for it
    if constexpr (PACED) {
      if (lane == 0 && pend.tile >= 0 && (issued[bi] < static_cast<int>(round) + 1 || !mbar_test(bar_s, phase))) flush_pending(true);
    }
    if (NB != NW) {
      while (issued[bi] < static_cast<int>(round) + 1) { }
    }
    mbar_wait(bar_s, phase);

    if (p.debug & 1) {   // measurement aid: the data movement of the step without its arithmetic (z copied to dz)
#pragma unroll
      for (int it = 0; it < ITERS; ++it) { dzv[lane + 32 * it] = bv[32 * it]; sweep_fence(it); }
      __syncwarp();
      if (lane == 0) {
        const long nt = static_cast<long>(t) + NB;
        if (nt < ntiles) {
          if (p.pace > 0) {
            const unsigned long long when = pace_reserve(&pace_next, p.pace, static_cast<unsigned long long>(clock64()));
            while (static_cast<unsigned long long>(clock64()) < when) { }
          }
          issue_load(nt, bi, round);
        }
      }
      __syncwarp();
      continue;
    }

    // ---------------------------------------------------------------- forward: max
    float mloc;
    if constexpr (sizeof(T) == 2) {
      uint32_t m0 = 0xff80ff80u, m1 = 0xff80ff80u;   // (-inf, -inf)
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const uint4 r = bv[32 * it];
        m0 = max_bf16x2(m0, r.x); m1 = max_bf16x2(m1, r.y);
        m0 = max_bf16x2(m0, r.z); m1 = max_bf16x2(m1, r.w);
        sweep_fence(it);
      }
      m0 = max_bf16x2(m0, m1);
      mloc = fmaxf(bf16lo(m0), bf16hi(m0));
    } else {
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const uint4 r = bv[32 * it];
        m0 = fmaxf(m0, fmaxf(__uint_as_float(r.x), __uint_as_float(r.y)));
        m1 = fmaxf(m1, fmaxf(__uint_as_float(r.z), __uint_as_float(r.w)));
        sweep_fence(it);
      }
      mloc = fmaxf(m0, m1);
    }
    const float m2 = warp_max_redux(mloc) * kLog2e;
    const f2 nm2 = pk1(kBias - m2);
    // vectors of this lane inside the (vector-aligned) window: rows rlo..rhi relative to the lane's first row
    const bool colin = kWin && nwv > 0 && cv >= jv_lo && cv < jv_lo + nvw;
    const int rlo = i_lo - r0, rhi = i_hi - r0;

    // ---------------------------------------------------------------- forward: column sums + row sums of e = 2^(z log2e - m2)
    float S, Sx, Sy = 0.f, Tt = 0.f;
    f2 colE[NP];
    float rsk[kVar ? ITERS : 1];
    {
      f2 tt2 = pk1(0.f);
      const float ya = opaque(y0);
#pragma unroll
      for (int c = 0; c < NP; ++c) colE[c] = pk1(0.f);
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const uint4 raw = bv[32 * it];
        f2 v[NP], e[NP];
        unpack_pairs<T>(raw, v);
#pragma unroll
        for (int c = 0; c < NP; ++c) {
          const f2 tt = fma2(v[c], l2e2, nm2);
          e[c] = ex2_2(tt);
          colE[c] = add2(colE[c], e[c]);
          if (kMSE) tt2 = fma2(e[c], e[c], tt2);
          if constexpr (kKL) {                       // sum e t; t = -inf (a logit of -inf) contributes 0, not 0 * inf
            float t0, t1;
            upk(tt, t0, t1);
            tt2 = fma2(e[c], pk(fmaxf(t0, -1e30f), fmaxf(t1, -1e30f)), tt2);
          }
        }
        f2 s = add2(e[0], e[1]);
        if constexpr (NP == 4) s = add2(s, add2(e[2], e[3]));
        const float rs = hsum(s);
        Sy = fmaf(rs, ya + static_cast<float>(it) * dyi, Sy);
        if constexpr (kVar) rsk[it] = rs;
        const bool inwin = kWin && colin && (it * RPI >= rlo) && (it * RPI <= rhi);
        if (kStash && !inwin) bst[32 * it] = stash_pack<T>(e);
        sweep_fence(it);
      }
      S = 0.f; Sx = 0.f;
#pragma unroll
      for (int c = 0; c < NP; ++c) {
        float lo, hi;
        upk(colE[c], lo, hi);
        S += lo + hi;
        Sx = fmaf(lo, xs[2 * c], fmaf(hi, xs[2 * c + 1], Sx));
      }
      Tt = hsum(tt2);
    }
    {
      const float k = warp_sum4_transposed(S, Sx, Sy, Tt, lane);
      S = __shfl_sync(kFull, k, 0); Sx = __shfl_sync(kFull, k, 8); Sy = __shfl_sync(kFull, k, 16); Tt = __shfl_sync(kFull, k, 24);
    }
    flush_pending(false);
    const float invS = rcp(S);             // S >= 1: one MUFU, 1 ulp
    const float mux = Sx * invS, muy = Sy * invS;

    float D = 0.f, creg = 0.f, ginv = 0.f, vx = 0.f, vy = 0.f;

    // ---------------------------------------------------------------- forward: variance about the mean, from the sums
    if constexpr (kVar) {
      float ax = 0.f, ay = 0.f;
      const float yv = opaque(y0) - muy;
#pragma unroll
      for (int c = 0; c < NP; ++c) {
        float lo, hi;
        upk(colE[c], lo, hi);
        const float d0 = xs[2 * c] - mux, d1 = xs[2 * c + 1] - mux;
        ax = fmaf(lo * d0, d0, fmaf(hi * d1, d1, ax));
      }
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const float d = yv + static_cast<float>(it) * dyi;
        ay = fmaf(rsk[it] * d, d, ay);
      }
      {
        const float k = warp_sum2_transposed(ax, ay, lane);
        ax = __shfl_sync(kFull, k, 0); ay = __shfl_sync(kFull, k, 16);
      }
      vx = ax * invS;
      vy = ay * invS;
      const float ex = vx - s2, ey = vy - s2;
      D = ex * ex + ey * ey;
      creg = 2.f * (ex * vx + ey * vy);
    }

    // ---------------------------------------------------------------- forward: divergence on the window (compact mapping)
    f2 wd[kWinRegs ? MAXS : 1][NP], wP[kWinRegs ? MAXS : 1][NP];   // per window pixel: log2 P - 1 - log2 M (JS) or G (MSE); P
    float inv_nvw = 0.f, l2ginv = 0.f;
    if constexpr (kWin) {
      f2 qa = pk1(0.f), qb = pk1(0.f), qc = pk1(0.f);
      if (nwv > 0) {
        // Normalisation of the target Gaussian (src/dsnt/nn.py:168-180 divides by its sum over the image).  A window the image
        // border does not clip holds the whole sum to theta, and unit-spaced samples of a Gaussian sum to sigma sqrt(2 pi)
        // within 5.4e-9 relative from sigma = 1 px on (Poisson summation); clipped windows and narrower Gaussians are summed.
        // (fp32 builds only: in the bf16 JS build, at 96 registers, the extra branch costs spills -- 189 -> 217 us measured.)
        const float spx = p.sigma * (0.5f * W), spy = p.sigma * (0.5f * H);
        float sx = spx * 2.5066282746310002f, sy = spy * 2.5066282746310002f;
        if (sizeof(T) != 4 || !(spx >= 1.0f && spy >= 1.0f && j_lo > 0 && j_hi < W - 1 && i_lo > 0 && i_hi < H - 1)) {
          sx = 0.f; sy = 0.f;
          for (int j = j_lo + lane; j <= j_hi; j += 32) {
            const float d = fmaf(static_cast<float>(j), tow, bw) - tx;
            sx += ex2(g.k2 * d * d);
          }
          for (int i = i_lo + lane; i <= i_hi; i += 32) {
            const float d = fmaf(static_cast<float>(i), toh, bh) - ty;
            sy += ex2(g.k2 * d * d);
          }
          const float k = warp_sum2_transposed(sx, sy, lane);
          sx = __shfl_sync(kFull, k, 0);
          sy = __shfl_sync(kFull, k, 16);
        }
        ginv = rcp(sx * sy + kEps);
        l2ginv = lg2(ginv);
        const f2 tlm1 = pk1(-lg2(S) - 1.0f);         // log2 P - 1 = t + tlm1
        const f2 hinvS = pk1(0.5f * invS), invS2 = pk1(invS), k2p = pk1(g.k2), half2 = pk1(0.5f), eps2 = pk1(kEps);
        inv_nvw = rcp(static_cast<float>(nvw));
        if constexpr (kKL) {
          // sum over the window of e log2(G + eps) and of e, slot by slot (run-time trip count: any sigma); nothing is kept
#pragma unroll 1
          for (int k = lane; k < nwv; k += 32) {
            const int r = static_cast<int>((static_cast<float>(k) + 0.5f) * inv_nvw);
            const int i = i_lo + r, jv = jv_lo + (k - r * nvw);
            const uint4 raw = bufv[i * WV + jv];
            f2 v[NP];
            unpack_pairs<T>(raw, v);
            const float dy = fmaf(static_cast<float>(i), toh, bh) - ty;
            const f2 rowt = pk1(fmaf(g.k2 * dy, dy, l2ginv));
            const f2 dx0 = pk1(fmaf(static_cast<float>(jv * VEC), tow, bw) - tx);
#pragma unroll
            for (int c = 0; c < NP; ++c) {
              const f2 dx = add2(dx0, pk((2 * c) * tow, (2 * c + 1) * tow));
              const f2 L = lg2_2(add2(ex2_2(fma2(mul2(dx, k2p), dx, rowt)), eps2));      // log2(G + eps)
              const f2 e = ex2_2(fma2(v[c], l2e2, nm2));
              qa = fma2(e, L, qa);
              qb = add2(qb, e);
            }
          }
        }
#pragma unroll
        for (int s = 0; s < (kWinRegs ? MAXS : 0); ++s) {
          const int k = lane + 32 * s;
          if (k < nwv) {
            const int r = static_cast<int>((static_cast<float>(k) + 0.5f) * inv_nvw);
            const int i = i_lo + r, jv = jv_lo + (k - r * nvw);
            const uint4 raw = bufv[i * WV + jv];
            f2 v[NP];
            unpack_pairs<T>(raw, v);
            const float dy = fmaf(static_cast<float>(i), toh, bh) - ty;
            const f2 rowt = pk1(fmaf(g.k2 * dy, dy, l2ginv));
            const f2 dx0 = pk1(fmaf(static_cast<float>(jv * VEC), tow, bw) - tx);
#pragma unroll
            for (int c = 0; c < NP; ++c) {
              const f2 dx = add2(dx0, pk((2 * c) * tow, (2 * c + 1) * tow));
              const f2 lgG = fma2(mul2(dx, k2p), dx, rowt);
              const f2 G = ex2_2(lgG);
              const f2 tt = fma2(v[c], l2e2, nm2);
              const f2 e = ex2_2(tt);
              const f2 P = mul2(e, invS2);
              wP[s][c] = P;
              if constexpr (kJS) {
                const f2 L = lg2_2(fma2(e, hinvS, fma2(half2, G, eps2)));
                const f2 d = sub2(add2(tt, tlm1), L);
                qa = fma2(P, d, qa);
                qb = fma2(G, sub2(lgG, L), qb);
                wd[s][c] = d;
              } else {
                const f2 df = sub2(P, G);
                qa = fma2(df, df, qa);
                qb = fma2(P, P, qb);
                qc = fma2(P, df, qc);
                wd[s][c] = G;
              }
            }
          }
        }
      }
      float a0 = hsum(qa), a1 = hsum(qb), a2 = hsum(qc), a3 = 0.f;
      {
        const float k = warp_sum4_transposed(a0, a1, a2, a3, lane);
        a0 = __shfl_sync(kFull, k, 0); a1 = __shfl_sync(kFull, k, 8); a2 = __shfl_sync(kFull, k, 16);
      }
      if (kMSE) {
        const float outside = fmaxf(fmaf(Tt * invS, invS, -a1), 0.f);
        D = outside + a0;
        creg = 2.f * (outside + a2);
      } else if (kKL) {
        // D = sum P ln P - sum P ln(G + eps); outside the window G + eps = eps exactly (src/dsnt/nn.py:208-211,233)
        const float plnp = kLn2 * fmaf(Tt, invS, -lg2(S));
        D = plnp - kLn2 * fmaf(kLog2Eps, 1.0f - a1 * invS, a0 * invS);
        creg = D + 1.0f;
      } else {
        creg = 0.5f * kLn2 * (1.0f + a0);
        D = fmaf(0.5f * kLn2, a1, creg);
      }
    }

    // ---------------------------------------------------------------- outputs + the scalars of the backward
    if (!have_count) { need_count(); inv_denom = 1.0f / fmaxf(mask_count, 1.0f); }
    const float wgt = mraw * inv_denom;
    float dist = 0.f, a = 0.f, b = 0.f;
    if (p.target) {
      const float dx = mux - tx, dy = muy - ty;
      const float d2 = dx * dx + dy * dy;
      const float rs = rsqrtf(d2);
      dist = d2 > 0.f ? d2 * rs : 0.f;
      const float invd = d2 > 0.f ? rs : ((p.flags & DSNT_FLAG_STRICT_NAN) ? INFINITY : 0.f);
      a = gl * wgt * (dx * invd);
      b = gl * wgt * (dy * invd);
    }
    const float rho = gl * wgt * p.reg_coeff;
    if (lane == 0) {
      reinterpret_cast<float2*>(p.coords)[hm] = make_float2(mux, muy);
      if (p.stats) {
        float4* st = reinterpret_cast<float4*>(p.stats + hm * kStatsK);
        st[0] = make_float4(m2, invS * exp2f(kBias), mux, muy);
        st[1] = make_float4(vx, vy, creg, ginv);
      }
      if (p.terms) reinterpret_cast<float2*>(p.terms)[hm] = make_float2(dist, D);
    }
    acc_d = fmaf(mraw, dist, acc_d);
    acc_r = fmaf(mraw, D, acc_r);
    const float cc = fmaf(a, mux, fmaf(b, muy, rho * creg));
    float cbase = -cc;
    if (kJS) cbase = fmaf(0.5f * kLn2, rho, cbase);
    if (kKL) cbase = fmaf(rho, 1.0f - kLnEps - kLn2 * lg2(S), cbase);      // r = ln2 t + this - [ln(G + eps) - ln eps]
    const float rho_t = kKL ? rho * kLn2 : 0.f;

    // ---------------------------------------------------------------- KL: backward on the window vectors, walked again
    if constexpr (kKL) {
      if (nwv > 0) {
        const f2 a2p = pk1(a), rt2 = pk1(rho_t), kl2 = pk1(-kLn2 * rho), l2eps2 = pk1(kLog2Eps), invS2 = pk1(invS);
        const f2 k2p = pk1(g.k2), eps2 = pk1(kEps);
#pragma unroll 1
        for (int k = lane; k < nwv; k += 32) {
          const int r = static_cast<int>((static_cast<float>(k) + 0.5f) * inv_nvw);
          const int i = i_lo + r, jv = jv_lo + (k - r * nvw);
          const uint4 raw = bufv[i * WV + jv];
          f2 v[NP], o[NP];
          unpack_pairs<T>(raw, v);
          const float y = fmaf(static_cast<float>(i), toh, bh);
          const float dy = y - ty;
          const f2 rowt = pk1(fmaf(g.k2 * dy, dy, l2ginv));
          const f2 rowc = pk1(fmaf(b, y, cbase));
          const f2 x0 = pk1(fmaf(static_cast<float>(jv * VEC), tow, bw));
          const f2 tx2 = pk1(tx);
#pragma unroll
          for (int c = 0; c < NP; ++c) {
            const f2 x = add2(x0, pk((2 * c) * tow, (2 * c + 1) * tow));
            const f2 dx = sub2(x, tx2);
            const f2 L = lg2_2(add2(ex2_2(fma2(mul2(dx, k2p), dx, rowt)), eps2));
            const f2 tt = fma2(v[c], l2e2, nm2);
            const f2 e = ex2_2(tt);
            float t0, t1;
            upk(tt, t0, t1);
            f2 gm = fma2(a2p, x, rowc);
            gm = fma2(rt2, pk(fmaxf(t0, -1e30f), fmaxf(t1, -1e30f)), gm);
            gm = fma2(kl2, sub2(L, l2eps2), gm);
            o[c] = mul2(mul2(e, invS2), gm);
          }
          dzv[i * WV + jv] = pack_pairs<T>(o);
        }
      }
    }

    // ---------------------------------------------------------------- backward on the window vectors, from registers
    if constexpr (kWinRegs) {
      if (nwv > 0) {
        const f2 a2p = pk1(a), kw = pk1(kJS ? 0.5f * kLn2 * rho : 2.f * rho);
#pragma unroll
        for (int s = 0; s < MAXS; ++s) {
          const int k = lane + 32 * s;
          if (k < nwv) {
            const int r = static_cast<int>((static_cast<float>(k) + 0.5f) * inv_nvw);
            const int i = i_lo + r, jv = jv_lo + (k - r * nvw);
            const f2 rowc = pk1(fmaf(b, fmaf(static_cast<float>(i), toh, bh), cbase));
            const f2 x0 = pk1(fmaf(static_cast<float>(jv * VEC), tow, bw));
            f2 o[NP];
#pragma unroll
            for (int c = 0; c < NP; ++c) {
              const f2 x = add2(x0, pk((2 * c) * tow, (2 * c + 1) * tow));
              f2 gm = fma2(a2p, x, rowc);
              if constexpr (kJS) gm = fma2(kw, wd[s][c], gm);                  // -(ln2/2) rho log2(1 + (G + 2 eps)/P)
              else gm = fma2(kw, sub2(wP[s][c], wd[s][c]), gm);                // 2 rho (P - G)
              o[c] = mul2(wP[s][c], gm);
            }
            dzv[i * WV + jv] = pack_pairs<T>(o);
          }
        }
      }
    }

    // ---------------------------------------------------------------- backward: every other vector, dz = e * (A_col + R_row)
    {
      f2 acol[NP];
#pragma unroll
      for (int c = 0; c < NP; ++c) {
        float v0 = a * xs[2 * c], v1 = a * xs[2 * c + 1];
        if (kVar) {
          const float kx = rho * 2.f * (vx - s2);
          const float d0 = xs[2 * c] - mux, d1 = xs[2 * c + 1] - mux;
          v0 = fmaf(kx * d0, d0, v0);
          v1 = fmaf(kx * d1, d1, v1);
        }
        acol[c] = pk(v0 * invS, v1 * invS);
      }
      const float bS = b * invS, cbS = cbase * invS;
      const float kyS = kVar ? rho * 2.f * (vy - s2) * invS : 0.f;
      const f2 rpS = pk1(kMSE ? 2.f * rho * invS * invS : 0.f);
      const f2 rtS = pk1(rho_t * invS);
      const float yb = opaque(y0);
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const uint4 raw = bv[32 * it];
        f2 e[NP], o[NP];
        f2 tk[kKL ? NP : 1];                       // KL: t itself enters the gradient (rho ln2 t)
        if constexpr (kStash) {
          stash_unpack<T>(raw, e);
        } else {
          f2 v[NP];
          unpack_pairs<T>(raw, v);
#pragma unroll
          for (int c = 0; c < NP; ++c) {
            const f2 tt = fma2(v[c], l2e2, nm2);
            e[c] = ex2_2(tt);
            if constexpr (kKL) {
              float t0, t1;
              upk(tt, t0, t1);
              tk[c] = pk(fmaxf(t0, -1e30f), fmaxf(t1, -1e30f));
            }
          }
        }
        const float y = yb + static_cast<float>(it) * dyi;
        float rc = fmaf(bS, y, cbS);
        if (kVar) { const float d = y - muy; rc = fmaf(kyS * d, d, rc); }
        const f2 rc2 = pk1(rc);
#pragma unroll
        for (int c = 0; c < NP; ++c) {
          f2 gm = add2(acol[c], rc2);
          if (kMSE) gm = fma2(rpS, e[c], gm);
          if constexpr (kKL) gm = fma2(rtS, tk[c], gm);
          o[c] = mul2(e[c], gm);
        }
        const bool inwin = kWin && colin && (it * RPI >= rlo) && (it * RPI <= rhi);
        if (!inwin) dzv[lane + 32 * it] = pack_pairs<T>(o);
        sweep_fence(it);
      }
    }

    // ---------------------------------------------------------------- hand the buffer to the next tile
    if (kStash) fence_async_smem();   // this lane's writes of e (generic proxy) are ordered before the bulk load (async proxy)
    __syncwarp();
    flush_pending(true);     // at most one load is pending per warp
    if (lane == 0) {
      const long nt = static_cast<long>(t) + NB;
      if (nt < ntiles) {
        if constexpr (PACED) {
          const unsigned long long now = static_cast<unsigned long long>(clock64());
          const unsigned long long when = p.pace > 0 ? pace_reserve(&pace_next, p.pace, now) : now;
          pend.when = static_cast<uint32_t>(when); pend.tile = static_cast<int>(nt); pend.slot = bi | (round << 8);
        } else {
          mbar_expect_tx(bar_s, hm_bytes);
          bulk_load(buf_s, zsrc + locate(p.st, nt * hm_mul + hm_add, hm_bytes).z_bytes, hm_bytes, bar_s);
          __threadfence_block();
          issued[bi] = static_cast<int>(round) + 2;
        }
      }
    }
    flush_pending(false);    // goes at once unless its slot lies in the future
    __syncwarp();
  }
  flush_pending(true);

  // ---------------------------------------------------------------- single-launch form: masked_average + loss composition
  // (src/dsnt/nn.py:81-94, src/dsnt/model.py:145) as in finish_loss_kernel: warp order inside the CTA, the CTA with the last
  // ticket adds the CTAs' partials in index order -- no float atomics, bit-reproducible for a given grid.
  if (p.out8) {
    if (lane == 0) { fin[warp][0] = acc_d; fin[warp][1] = acc_r; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float sd = 0.f, sr = 0.f;
      for (int w2 = 0; w2 < p.nwarps; ++w2) { sd += fin[w2][0]; sr += fin[w2][1]; }
      reinterpret_cast<float4*>(p.ws)[blockIdx.x] = make_float4(sd, sr, 0.f, 0.f);
      __threadfence();
      fin_last = atomicAdd(ctl, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (fin_last && warp == 0) {
      __threadfence();
      if (lane == 0) trace[3] = global_timer_ns();
      need_count();                 // a warp without heatmaps has not looked yet
      const float4* part = reinterpret_cast<const float4*>(p.ws);
      float4 v[kFinishSlots / 32];
#pragma unroll
      for (int k = 0; k < kFinishSlots / 32; ++k) {
        const int i = lane + 32 * k;
        v[k] = i < static_cast<int>(gridDim.x) ? __ldcg(part + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int k = 0; k < kFinishSlots / 32; ++k) { sa += v[k].x; sb += v[k].y; }
      sa = warp_sum(sa); sb = warp_sum(sb);
      if (sharded) {                // totals over the ranks; the count goes round again so that the layout of the exchange
        float cnt = __ldcg(p.ws + kFinishLocal);      // is that of finish_loss_kernel
        peer_exchange_sum3(p.xc, sa, sb, cnt);
      }
      if (lane == 0) {
        p.out8[0] = sa; p.out8[1] = sb;
        p.out8[2] = p.denom ? __ldg(p.denom) : mask_count;   // with an external denominator out8[2..3] repeat it
        write_loss_tail(p.out8, p.reg_coeff);
        ctl[0] = 0u;         // end ticket
        ctl[1] = 0u;         // start ticket: every CTA drew it long ago
        ctl[2] = 0u;         // and the "count published" flag: every warp with a heatmap has read it
        trace[4] = global_timer_ns();
      }
    }
  }
}

}  // namespace dsnt
