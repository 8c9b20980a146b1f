// f32x2.cuh -- packed two-wide fp32 arithmetic of sm_100 (PTX add/sub/mul/fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2),
// packed bf16 maximum (HMNMX2.BF16_V2) and the one-instruction warp maximum (redux.sync.max.f32 -> CREDUX.MAX.F32).
//
// A packed instruction does two fp32 operations per lane in ONE issue slot.  The DSNT head is issue-bound on bf16
// heatmaps (ncu: 70 % issue-active, 28 instructions per pixel, profiles/r01_v6_step_bf16_js_*), so the per-pixel FMA /
// ADD / MUL work of the sweeps is written on pairs of neighbouring pixels.
#pragma once

#include "common.cuh"

namespace dsnt {

struct f2 {
  unsigned long long r;   // .x = low 32 bits, .y = high 32 bits
};

__device__ __forceinline__ f2 pk(float lo, float hi) {
  f2 o;
  asm("mov.b64 %0, {%1, %2};" : "=l"(o.r) : "f"(lo), "f"(hi));
  return o;
}
__device__ __forceinline__ f2 pk1(float v) { return pk(v, v); }
__device__ __forceinline__ void upk(f2 a, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.r));
}
__device__ __forceinline__ float hsum(f2 a) {
  float lo, hi;
  upk(a, lo, hi);
  return lo + hi;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 o;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(o.r) : "l"(a.r), "l"(b.r), "l"(c.r));
  return o;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 o;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(o.r) : "l"(a.r), "l"(b.r));
  return o;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
  f2 o;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(o.r) : "l"(a.r), "l"(b.r));
  return o;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 o;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(o.r) : "l"(a.r), "l"(b.r));
  return o;
}
// element-wise 2^x / log2 x on a pair (two MUFU ops; there is no packed MUFU)
__device__ __forceinline__ f2 ex2_2(f2 a) {
  float lo, hi;
  upk(a, lo, hi);
  return pk(ex2(lo), ex2(hi));
}
__device__ __forceinline__ f2 lg2_2(f2 a) {
  float lo, hi;
  upk(a, lo, hi);
  return pk(lg2(lo), lg2(hi));
}

// one 32-bit word holding two bf16 (element 2k in the low half) -> the pair of fp32 values
__device__ __forceinline__ f2 bf16x2_to_f2(uint32_t w) { return pk(bf16lo(w), bf16hi(w)); }

__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t o;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(b));
  return o;
}

// maximum over the 32 lanes of a warp in one instruction (all lanes must take part)
__device__ __forceinline__ float warp_max_redux(float v) {
  float o;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(o) : "f"(v));
  return o;
}

}  // namespace dsnt
