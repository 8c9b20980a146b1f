// common.cuh -- device helpers shared by the DSNT head kernels (sm_100a).
//
// Everything here is fp32 math on data that is streamed once from HBM; there is no contraction on
// this path, so no tensor-core code.  What matters is 128-bit coalesced access, few instructions
// per pixel (the forward has a budget of ~17 issue slots and 2.7 MUFU ops per pixel at the HBM
// roofline of a 64x64 fp32 heatmap), and cheap reductions.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dsnt_b200.h"

namespace dsnt {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kEps = 1e-24f;  // src/dsnt/nn.py:202,208,214

constexpr int kStatsK = DSNT_STATS_K;

// ---------------------------------------------------------------------------------- stacked launches
// A stacked-hourglass model emits one heatmap tensor per stack (src/dsnt/hourglass.py:166-177), all with the same
// shape, target and mask (src/dsnt/model.py:238-246).  One launch covers every stack: heatmap index hm runs over
// count*n_per, stack s = hm / n_per lives at byte offset z_off[s] from stack 0.  Per-heatmap inputs shared by the
// stacks (target, mask) are indexed by the position inside the stack, outputs (coords, stats, terms) by hm.
constexpr int kMaxStacks = DSNT_MAX_STACKS;
struct Stacks {
  long n_per;                 // heatmaps per stack
  long z_off[kMaxStacks];     // byte offsets of the logits of stack s relative to stack 0
  long dz_off[kMaxStacks];    // same for the gradient output (backward only)
  int count;
};
struct HmRef {
  long nl;                    // index inside the stack (target / mask)
  long z_bytes, dz_bytes;     // byte offsets of this heatmap from the stack-0 base pointers
};
__device__ __forceinline__ HmRef locate(const Stacks& st, long hm, long hm_bytes) {
  HmRef r;
  if (st.count <= 1) {
    r.nl = hm; r.z_bytes = hm * hm_bytes; r.dz_bytes = r.z_bytes;
  } else {
    const unsigned h = static_cast<unsigned>(hm), per = static_cast<unsigned>(st.n_per);
    const unsigned s = h / per;
    r.nl = h - s * per;
    r.z_bytes = st.z_off[s] + r.nl * hm_bytes;
    r.dz_bytes = st.dz_off[s] + r.nl * hm_bytes;
  }
  return r;
}

// ---------------------------------------------------------------------------------- MUFU wrappers
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------- vector access
// Loads VEC consecutive elements starting at element index `idx` (a multiple of VEC, so the access
// is VEC*sizeof(T) aligned) and widens to fp32.  Heatmaps are streamed exactly once per kernel, so
// loads bypass L1 allocation.
template <typename T, int VEC>
struct VecIO;

template <>
struct VecIO<float, 4> {
  static __device__ __forceinline__ void load(const float* base, long idx, float (&v)[4]) {
    const float4* p = reinterpret_cast<const float4*>(base + idx);
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3])
                 : "l"(p));
  }
  static __device__ __forceinline__ void store(float* base, long idx, const float (&v)[4]) {
    *reinterpret_cast<float4*>(base + idx) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <>
struct VecIO<float, 1> {
  static __device__ __forceinline__ void load(const float* base, long idx, float (&v)[1]) { v[0] = __ldg(base + idx); }
  static __device__ __forceinline__ void store(float* base, long idx, const float (&v)[1]) { base[idx] = v[0]; }
};

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);  // .x = lo (low 16 bits), round-to-nearest-even
  return *reinterpret_cast<uint32_t*>(&h);
}

template <>
struct VecIO<__nv_bfloat16, 8> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* base, long idx, float (&v)[8]) {
    uint32_t a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                 : "l"(base + idx));
    v[0] = bf16lo(a); v[1] = bf16hi(a); v[2] = bf16lo(b); v[3] = bf16hi(b);
    v[4] = bf16lo(c); v[5] = bf16hi(c); v[6] = bf16lo(d); v[7] = bf16hi(d);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* base, long idx, const float (&v)[8]) {
    uint4 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
    o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(base + idx) = o;
  }
};

template <>
struct VecIO<__nv_bfloat16, 4> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* base, long idx, float (&v)[4]) {
    uint32_t a, b;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(base + idx));
    v[0] = bf16lo(a); v[1] = bf16hi(a); v[2] = bf16lo(b); v[3] = bf16hi(b);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* base, long idx, const float (&v)[4]) {
    uint2 o;
    o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]);
    *reinterpret_cast<uint2*>(base + idx) = o;
  }
};

template <>
struct VecIO<__nv_bfloat16, 1> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* base, long idx, float (&v)[1]) {
    v[0] = __bfloat162float(base[idx]);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* base, long idx, const float (&v)[1]) {
    base[idx] = __float2bfloat16_rn(v[0]);
  }
};

// ---------------------------------------------------------------------------------- warp reductions
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// Transposed butterfly: reduces FOUR values over the warp with 6 shuffles instead of 20.
// On return lane L holds the complete sum of value number ((L >> 3) & 3)  (lanes 0-7: a, 8-15: b, ...).
__device__ __forceinline__ float warp_sum4_transposed(float a, float b, float c, float d, int lane) {
  const bool hi16 = lane & 16;
  // lanes with bit4 == 0 keep (a,b) and receive the partner's (a,b); lanes with bit4 == 1 keep (c,d)
  float k0 = hi16 ? c : a, k1 = hi16 ? d : b;
  float s0 = hi16 ? a : c, s1 = hi16 ? b : d;
  k0 += __shfl_xor_sync(kFull, s0, 16);
  k1 += __shfl_xor_sync(kFull, s1, 16);
  const bool hi8 = lane & 8;
  float k = hi8 ? k1 : k0, s = hi8 ? k0 : k1;
  k += __shfl_xor_sync(kFull, s, 8);
  k += __shfl_xor_sync(kFull, k, 4);
  k += __shfl_xor_sync(kFull, k, 2);
  k += __shfl_xor_sync(kFull, k, 1);
  return k;
}

// Same idea for TWO values (5 shuffles instead of 10): lane L ends with value number ((L >> 4) & 1).
__device__ __forceinline__ float warp_sum2_transposed(float a, float b, int lane) {
  const bool hi16 = lane & 16;
  float k = hi16 ? b : a, s = hi16 ? a : b;
  k += __shfl_xor_sync(kFull, s, 16);
  k += __shfl_xor_sync(kFull, k, 8);
  k += __shfl_xor_sync(kFull, k, 4);
  k += __shfl_xor_sync(kFull, k, 2);
  k += __shfl_xor_sync(kFull, k, 1);
  return k;
}

// ---------------------------------------------------------------------------------- group reductions
// A "group" is the set of GROUP threads that cooperate on one heatmap: one warp (GROUP == 32, several
// heatmaps per CTA, no block barrier at all) or the whole CTA (GROUP == blockDim.x).
// `red` points to a per-call-site shared scratch of at least 4 * (GROUP/32) floats; distinct call sites
// use distinct scratch so no extra barrier is needed between consecutive reductions.

template <int GROUP>
__device__ __forceinline__ void group_barrier() {
  if constexpr (GROUP == 32) __syncwarp(); else __syncthreads();
}

template <int GROUP>
__device__ __forceinline__ float group_max(float v, float* red, int warp_g, int lane) {
  v = warp_max(v);
  if constexpr (GROUP > 32) {
    constexpr int NW = GROUP / 32;
    if (lane == 0) red[warp_g] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) r = fmaxf(r, red[w]);
    v = r;
  }
  return v;
}

template <int GROUP>
__device__ __forceinline__ void group_sum4(float& a, float& b, float& c, float& d, float* red, int warp_g, int lane) {
  float k = warp_sum4_transposed(a, b, c, d, lane);
  if constexpr (GROUP == 32) {
    a = __shfl_sync(kFull, k, 0);
    b = __shfl_sync(kFull, k, 8);
    c = __shfl_sync(kFull, k, 16);
    d = __shfl_sync(kFull, k, 24);
  } else {
    constexpr int NW = GROUP / 32;
    if ((lane & 7) == 0) red[(lane >> 3) * NW + warp_g] = k;   // layout [value][warp]
    __syncthreads();
    float r[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) acc += red[v * NW + w];
      r[v] = acc;
    }
    a = r[0]; b = r[1]; c = r[2]; d = r[3];
  }
}

template <int GROUP>
__device__ __forceinline__ void group_sum2(float& a, float& b, float* red, int warp_g, int lane) {
  float k = warp_sum2_transposed(a, b, lane);
  if constexpr (GROUP == 32) {
    a = __shfl_sync(kFull, k, 0);
    b = __shfl_sync(kFull, k, 16);
  } else {
    constexpr int NW = GROUP / 32;
    if ((lane & 15) == 0) red[(lane >> 4) * NW + warp_g] = k;
    __syncthreads();
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) { r0 += red[w]; r1 += red[NW + w]; }
    a = r0; b = r1;
  }
}

// ---------------------------------------------------------------------------------- pixel geometry
// Pixel-centre coordinate of index j on an axis of n pixels: (2j+1)/n - 1  (src/dsnt/nn.py:30-37).
__device__ __forceinline__ float axis_coord(int j, float two_over_n, float bias) {
  return fmaf(static_cast<float>(j), two_over_n, bias);  // bias = 1/n - 1
}

// Walks the vectors f = lane, lane+GROUP, lane+2*GROUP ... of one heatmap and tracks (row, first column)
// without a division per vector.  wv = W / VEC vectors per row.
struct VecWalker {
  int row, cv, step_r, step_c, wv;
  __device__ __forceinline__ VecWalker(int first_vec, int stride, int wv_) : wv(wv_) {
    row = first_vec / wv_;
    cv = first_vec - row * wv_;
    step_r = stride / wv_;
    step_c = stride - step_r * wv_;
  }
  __device__ __forceinline__ void next() {
    cv += step_c;
    row += step_r;
    if (cv >= wv) { cv -= wv; ++row; }
  }
};

// ---------------------------------------------------------------------------------- regulariser traits
__host__ __device__ constexpr bool reg_needs_gauss(int reg) {
  return reg == DSNT_REG_KL || reg == DSNT_REG_JS || reg == DSNT_REG_MSE;
}

// Unnormalised separable Gaussian tables for one heatmap, built by ONE warp per axis:
//   tab[j] = exp(k (c_j - t)^2),  returns (sum tab, sum tab*log2(tab)) on every lane of the calling warp.
__device__ __forceinline__ void gauss_axis_table(float* tab, int n, float t, float k2 /* -0.5/sigma^2 * log2e */,
                                                 int lane, float& sum, float& ent2) {
  const float two_over_n = 2.0f / static_cast<float>(n), bias = 1.0f / static_cast<float>(n) - 1.0f;
  float s = 0.f, h = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float dx = axis_coord(j, two_over_n, bias) - t;
    const float a = k2 * dx * dx;
    const float g = ex2(a);
    tab[j] = g;
    s += g;
    h = fmaf(g, fmaxf(a, -1e30f), h);  // g * log2(g); clamp avoids 0 * -inf
  }
  sum = warp_sum(s);
  ent2 = warp_sum(h);
}

}  // namespace dsnt
