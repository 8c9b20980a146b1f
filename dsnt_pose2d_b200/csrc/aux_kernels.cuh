// aux_kernels.cuh -- the small kernels around the fused head: the deterministic finishing reduction
// (masked_average + loss composition), the standalone Euclidean distance, (thresholded) softmax with
// materialised output, and make_gauss.  None of these is on the bandwidth-critical path; they exist so
// that every function of src/dsnt/nn.py has a CUDA implementation behind the C ABI (no CPU/torch fallback).
#pragma once

#include "common.cuh"
#include "finish_common.cuh"

namespace dsnt {

// ------------------------------------------------------------------------------------------------
// Block-wide sum of up to 4 values with a fixed reduction tree (deterministic for a fixed block size).
template <int BLOCK>
__device__ __forceinline__ void block_sum4(float& a, float& b, float& c, float& d, float* red /*4*BLOCK/32*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  group_sum4<BLOCK>(a, b, c, d, red, warp, lane);
}

// ------------------------------------------------------------------------------------------------
// masked_average + loss composition (src/dsnt/nn.py:81-94, src/dsnt/model.py:145).
// Two-level reduction in ONE launch: every CTA sums a contiguous slice of the per-heatmap terms in a fixed
// order and parks its partial in the workspace; the CTA that draws the last ticket adds the partials in
// index order.  No float atomics, fixed traversal => bit-reproducible for a given (N, grid).
// workspace: kFinishSlots*4 floats of partials followed by one unsigned ticket counter that must be zero
// before the first launch and is reset by the kernel itself (stream-ordered use only).

// Stacked form: terms holds n = count*n_per rows (stack-major); the mask (one stack long) is shared, the mask
// count -- the denominator of every per-stack average -- is taken over the first stack only, so
// out[6] = sum_s (euclid_s + reg_coeff*reg_s) as in src/dsnt/model.py:238-246.
__global__ void __launch_bounds__(kFinishBlock) finish_loss_kernel(const float* __restrict__ terms,
                                                                  const float* __restrict__ mask, long n, long n_per,
                                                                  float reg_coeff, float* __restrict__ out,
                                                                  float* __restrict__ workspace, const PeerXchg xc) {
  __shared__ float red[4 * kFinishBlock / 32];
  __shared__ bool is_last;
  const long chunk = (n + gridDim.x - 1) / gridDim.x;
  const long lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
  float sd = 0.f, sr = 0.f, sm = 0.f, unused = 0.f;
  const bool one_stack = n_per >= n;     // no 64-bit modulo per element in the common single-tensor case
  for (long i = lo + threadIdx.x; i < hi; i += kFinishBlock) {
    const float2 t = terms ? __ldg(reinterpret_cast<const float2*>(terms) + i) : make_float2(0.f, 0.f);   // null: mask count only
    const float w = mask ? __ldg(mask + (one_stack ? i : i % n_per)) : 1.0f;
    sd = fmaf(w, t.x, sd);
    sr = fmaf(w, t.y, sr);
    if (i < n_per) sm += w;
  }
  block_sum4<kFinishBlock>(sd, sr, sm, unused, red);
  unsigned* ticket = reinterpret_cast<unsigned*>(workspace + kFinishCtl);
  if (threadIdx.x == 0) {
    float4* part = reinterpret_cast<float4*>(workspace);
    part[blockIdx.x] = make_float4(sd, sr, sm, 0.f);
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    // L2 loads (the partials were written by other SMs; ld.cg never looks at this SM's L1), all issued before the first
    // is consumed: a volatile float4 is four dependent round trips per partial
    const float4* part = reinterpret_cast<const float4*>(workspace);
    float4 v[kFinishMaxCtas / 32];
#pragma unroll
    for (int k = 0; k < kFinishMaxCtas / 32; ++k) {
      const int i = threadIdx.x + 32 * k;
      v[k] = i < static_cast<int>(gridDim.x) ? __ldcg(part + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
    for (int k = 0; k < kFinishMaxCtas / 32; ++k) {  // fixed order per lane
      a += v[k].x; b += v[k].y; c += v[k].z;
    }
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (xc.world > 1) peer_exchange_sum3(xc, a, b, c);   // sharded batch: totals over the ranks
    if (threadIdx.x == 0) {
      out[0] = a; out[1] = b; out[2] = c;
      write_loss_tail(out, reg_coeff);
      *ticket = 0u;
    }
  }
}

// x *= *g unless *g == 1 (then nothing is read or written): lets the one-pass step hand out the gradient it already
// computed for d(loss) = 1 and stay exact -- and CUDA-graph capturable -- for any other upstream gradient.
constexpr int kScaleBlock = 1024;
template <typename T>
__global__ void __launch_bounds__(kScaleBlock) scale_unless_one_kernel(T* __restrict__ x, long nvec16, long numel,
                                                                       const float* __restrict__ g) {
  const float s = __ldg(g);
  if (s == 1.0f) return;
  constexpr int PER = 16 / sizeof(T);
  constexpr int UNROLL = 4;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < nvec16; i += UNROLL * stride) {
    float v[UNROLL][PER];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) VecIO<T, PER>::load(x, (i + u * stride) * PER, v[u]);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
      for (int c = 0; c < PER; ++c) v[u][c] *= s;
      VecIO<T, PER>::store(x, (i + u * stride) * PER, v[u]);
    }
  }
  for (; i < nvec16; i += stride) {
    float v[PER];
    VecIO<T, PER>::load(x, i * PER, v);
#pragma unroll
    for (int c = 0; c < PER; ++c) v[c] *= s;
    VecIO<T, PER>::store(x, i * PER, v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long k = nvec16 * PER; k < numel; ++k) {
      float v[1];
      VecIO<T, 1>::load(x, k, v);
      v[0] *= s;
      VecIO<T, 1>::store(x, k, v);
    }
  }
}

__global__ void combine_loss_kernel(float* out, float reg_coeff) {
  if (threadIdx.x == 0 && blockIdx.x == 0) write_loss_tail(out, reg_coeff);
}

// ------------------------------------------------------------------------------------------------
// Standalone euclidean_loss pieces (src/dsnt/nn.py:112-114): one thread per point.
// The distance is written in the [n,2] "terms" layout (dist, 0) so dsnt_finish_loss can average it.
__global__ void euclid_fwd_kernel(const float* __restrict__ actual, const float* __restrict__ target, long n, int d,
                                  float* __restrict__ terms) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int k = 0; k < d; ++k) {
    const float df = actual[i * d + k] - target[i * d + k];
    acc = fmaf(df, df, acc);
  }
  reinterpret_cast<float2*>(terms)[i] = make_float2(sqrtf(acc), 0.f);
}

// g_actual[i,k] = g_loss * w_i * (actual - target)/dist,  w_i = (mask ? mask[i] : 1)/denom
__global__ void euclid_bwd_kernel(const float* __restrict__ actual, const float* __restrict__ target,
                                  const float* __restrict__ terms, const float* __restrict__ mask,
                                  const float* __restrict__ g_loss, const float* __restrict__ denom, long n, int d,
                                  int flags, float* __restrict__ g_actual) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float ds = terms[2 * i];
  const float invd = ds > 0.f ? 1.0f / ds : ((flags & DSNT_FLAG_STRICT_NAN) ? INFINITY : 0.f);
  const float g = __ldg(g_loss) * (mask ? mask[i] : 1.0f) / __ldg(denom);
  for (int k = 0; k < d; ++k) g_actual[i * d + k] = g * ((actual[i * d + k] - target[i * d + k]) * invd);
}

// ------------------------------------------------------------------------------------------------
// (Thresholded) softmax over rows of a [rows, len] matrix with the output materialised
// (src/dsnt/nn.py:119-139,160-165).  One CTA per row; the row is re-read from L1/L2 between the three
// passes (max, sum, write).  Used by flat_softmax/softmax_2d/thresholded_softmax and the lazy `.heatmaps`.
constexpr int kRowBlock = 256;

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p, long i);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, long i) { return p[i]; }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p, long i) {
  return __bfloat162float(p[i]);
}
template <typename T>
__device__ __forceinline__ void store_from_float(T* p, long i, float v);
template <>
__device__ __forceinline__ void store_from_float<float>(float* p, long i, float v) { p[i] = v; }
template <>
__device__ __forceinline__ void store_from_float<__nv_bfloat16>(__nv_bfloat16* p, long i, float v) {
  p[i] = __float2bfloat16_rn(v);
}

template <typename T>
__global__ void __launch_bounds__(kRowBlock) tsoftmax_fwd_kernel(const T* __restrict__ x, long len, float threshold,
                                                                float eps, T* __restrict__ out) {
  __shared__ float red_m[kRowBlock / 32];
  __shared__ float red_s[4 * kRowBlock / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const T* xr = x + static_cast<long>(blockIdx.x) * len;
  T* orow = out + static_cast<long>(blockIdx.x) * len;
  float m = -INFINITY;
  for (long i = threadIdx.x; i < len; i += kRowBlock) m = fmaxf(m, load_as_float(xr, i));
  m = group_max<kRowBlock>(m, red_m, warp, lane);  // max over ALL entries, kept or not (nn.py:124)
  const float m2 = m * kLog2e;
  float s = 0.f, u0 = 0.f, u1 = 0.f, u2 = 0.f;
  for (long i = threadIdx.x; i < len; i += kRowBlock) {
    const float v = load_as_float(xr, i);
    if (v >= threshold) s += ex2(fmaf(v, kLog2e, -m2));
  }
  group_sum4<kRowBlock>(s, u0, u1, u2, red_s, warp, lane);
  const float inv = 1.0f / (s + eps);
  for (long i = threadIdx.x; i < len; i += kRowBlock) {
    const float v = load_as_float(xr, i);
    store_from_float(orow, i, v >= threshold ? ex2(fmaf(v, kLog2e, -m2)) * inv : 0.f);
  }
}

template <typename T>
__global__ void __launch_bounds__(kRowBlock) tsoftmax_bwd_kernel(const T* __restrict__ out, const T* __restrict__ g,
                                                                long len, T* __restrict__ dx) {
  __shared__ float red_s[4 * kRowBlock / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long off = static_cast<long>(blockIdx.x) * len;
  float s = 0.f, u0 = 0.f, u1 = 0.f, u2 = 0.f;
  for (long i = threadIdx.x; i < len; i += kRowBlock) s = fmaf(load_as_float(g + off, i), load_as_float(out + off, i), s);
  group_sum4<kRowBlock>(s, u0, u1, u2, red_s, warp, lane);
  for (long i = threadIdx.x; i < len; i += kRowBlock) {
    const float o = load_as_float(out + off, i);
    store_from_float(dx + off, i, o * (load_as_float(g + off, i) - s));
  }
}

// ------------------------------------------------------------------------------------------------
// make_gauss (src/dsnt/nn.py:168-205): out[n,i,j] = gx_j gy_i / (sum + 1e-24).  One CTA per Gaussian.
constexpr int kGaussBlock = 256;

__global__ void __launch_bounds__(kGaussBlock) make_gauss_fwd_kernel(const float* __restrict__ mu, int W, int H,
                                                                    float sigma, float* __restrict__ out) {
  extern __shared__ __align__(16) float dyn_smem[];
  float* tabx = dyn_smem;
  float* taby = tabx + ((W + 3) & ~3);
  float* scal = taby + ((H + 3) & ~3);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long n = blockIdx.x;
  const float tx = mu[2 * n], ty = mu[2 * n + 1];
  const float k2 = -0.5f / (sigma * sigma) * kLog2e;
  if (warp == 0) {
    float s, h;
    gauss_axis_table(tabx, W, tx, k2, lane, s, h);
    if (lane == 0) scal[0] = s;
  } else if (warp == 1) {
    float s, h;
    gauss_axis_table(taby, H, ty, k2, lane, s, h);
    if (lane == 0) scal[2] = s;
  }
  __syncthreads();
  const float ginv = 1.0f / (scal[0] * scal[2] + kEps);
  float* o = out + n * static_cast<long>(H) * W;
  for (int idx = threadIdx.x; idx < H * W; idx += kGaussBlock) {
    const int i = idx / W, j = idx - i * W;
    o[idx] = tabx[j] * (taby[i] * ginv);
  }
}

// d out / d mu (make_gauss is differentiable wrt its centres, src/dsnt/nn.py:170):
//   dmu_x = sum g G u_x - (sum g G)(sum G u_x),  u_x = (x_j - mu_x)/sigma^2   (likewise y)
__global__ void __launch_bounds__(kGaussBlock) make_gauss_bwd_kernel(const float* __restrict__ mu,
                                                                    const float* __restrict__ g, int W, int H,
                                                                    float sigma, float* __restrict__ dmu) {
  extern __shared__ __align__(16) float dyn_smem[];
  __shared__ float red_a[4 * kGaussBlock / 32];
  __shared__ float red_b[2 * kGaussBlock / 32];
  float* tabx = dyn_smem;
  float* taby = tabx + ((W + 3) & ~3);
  float* scal = taby + ((H + 3) & ~3);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long n = blockIdx.x;
  const float tx = mu[2 * n], ty = mu[2 * n + 1];
  const float k2 = -0.5f / (sigma * sigma) * kLog2e;
  if (warp == 0) {
    float s, h;
    gauss_axis_table(tabx, W, tx, k2, lane, s, h);
    if (lane == 0) scal[0] = s;
  } else if (warp == 1) {
    float s, h;
    gauss_axis_table(taby, H, ty, k2, lane, s, h);
    if (lane == 0) scal[2] = s;
  }
  __syncthreads();
  const float ginv = 1.0f / (scal[0] * scal[2] + kEps);
  const float is2 = 1.0f / (sigma * sigma);
  const float two_over_w = 2.0f / W, bias_w = 1.0f / W - 1.0f;
  const float two_over_h = 2.0f / H, bias_h = 1.0f / H - 1.0f;
  const float* gr = g + n * static_cast<long>(H) * W;
  float a1x = 0.f, a1y = 0.f, a0 = 0.f, unused = 0.f, ux = 0.f, uy = 0.f;
  for (int idx = threadIdx.x; idx < H * W; idx += kGaussBlock) {
    const int i = idx / W, j = idx - i * W;
    const float G = tabx[j] * (taby[i] * ginv);
    const float vx = (axis_coord(j, two_over_w, bias_w) - tx) * is2;
    const float vy = (axis_coord(i, two_over_h, bias_h) - ty) * is2;
    const float gg = gr[idx] * G;
    a1x = fmaf(gg, vx, a1x);
    a1y = fmaf(gg, vy, a1y);
    a0 += gg;
    ux = fmaf(G, vx, ux);
    uy = fmaf(G, vy, uy);
  }
  group_sum4<kGaussBlock>(a1x, a1y, a0, unused, red_a, warp, lane);
  group_sum2<kGaussBlock>(ux, uy, red_b, warp, lane);
  if (threadIdx.x == 0) {
    dmu[2 * n] = a1x - a0 * ux;
    dmu[2 * n + 1] = a1y - a0 * uy;
  }
}

// ------------------------------------------------------------------------------------------------
// Gradient of the Gaussian-target regularisers w.r.t. the target centres mu_t.  In the reference kl / js / mse_reg_loss
// build their target with make_gauss (src/dsnt/nn.py:232,250,268), which is differentiable w.r.t. mu_t (:170), so a
// caller whose targets require grad gets d(loss)/d(mu_t) from autograd.  Here: one CTA per heatmap evaluates
//   dD/dG_ij   KL: -P/(G+eps)    JS: 1/2 [ln(G+eps) - ln(M+eps) + G/(G+eps) - M/(M+eps)]    MSE: -2 (P - G)
// on the fly and pushes it through the Jacobian of the normalised Gaussian exactly as make_gauss_bwd_kernel does:
//   dmu_x = sum g G u_x - (sum g G)(sum G u_x),  u_x = (x_j - mu_x)/sigma^2,
// scaled by d(loss) * reg_coeff * mask / denominator.  P comes from normalised heatmaps, or from logits and the saved
// softmax statistics (m log2e, 1/S).  Not on the bandwidth path: the targets of a training run carry no gradient.
template <typename T, int REG>
__global__ void __launch_bounds__(kGaussBlock) reg_dmu_kernel(const T* __restrict__ z, int input_is_logits,
                                                             const float* __restrict__ stats, const float* __restrict__ mu,
                                                             const float* __restrict__ mask, const float* __restrict__ g_loss,
                                                             const float* __restrict__ denom, float reg_coeff, int W, int H,
                                                             float sigma, float* __restrict__ dmu) {
  extern __shared__ __align__(16) float dyn_smem[];
  __shared__ float red_a[4 * kGaussBlock / 32];
  __shared__ float red_b[2 * kGaussBlock / 32];
  float* tabx = dyn_smem;
  float* taby = tabx + ((W + 3) & ~3);
  float* scal = taby + ((H + 3) & ~3);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long n = blockIdx.x;
  const float tx = mu[2 * n], ty = mu[2 * n + 1];
  const float k2 = -0.5f / (sigma * sigma) * kLog2e;
  if (warp == 0) {
    float s, h;
    gauss_axis_table(tabx, W, tx, k2, lane, s, h);
    if (lane == 0) scal[0] = s;
  } else if (warp == 1) {
    float s, h;
    gauss_axis_table(taby, H, ty, k2, lane, s, h);
    if (lane == 0) scal[2] = s;
  }
  __syncthreads();
  const float ginv = 1.0f / (scal[0] * scal[2] + kEps);
  const float is2 = 1.0f / (sigma * sigma);
  const float two_over_w = 2.0f / W, bias_w = 1.0f / W - 1.0f;
  const float two_over_h = 2.0f / H, bias_h = 1.0f / H - 1.0f;
  const T* zr = z + n * static_cast<long>(H) * W;
  float m2 = 0.f, invS = 1.f;
  if (input_is_logits) { m2 = stats[n * kStatsK]; invS = stats[n * kStatsK + 1]; }
  float a1x = 0.f, a1y = 0.f, a0 = 0.f, unused = 0.f, ux = 0.f, uy = 0.f;
  for (int idx = threadIdx.x; idx < H * W; idx += kGaussBlock) {
    const int i = idx / W, j = idx - i * W;
    const float G = tabx[j] * (taby[i] * ginv);
    const float v = load_as_float(zr, idx);
    const float P = input_is_logits ? ex2(fmaf(v, kLog2e, -m2)) * invS : v;
    float g;
    if constexpr (REG == DSNT_REG_KL) {
      g = -P / (G + kEps);
    } else if constexpr (REG == DSNT_REG_JS) {
      const float M = 0.5f * (P + G);
      g = 0.5f * (kLn2 * (lg2(G + kEps) - lg2(M + kEps)) + G / (G + kEps) - M / (M + kEps));
    } else {
      g = -2.0f * (P - G);
    }
    const float vx = (axis_coord(j, two_over_w, bias_w) - tx) * is2;
    const float vy = (axis_coord(i, two_over_h, bias_h) - ty) * is2;
    const float gg = g * G;
    a1x = fmaf(gg, vx, a1x);
    a1y = fmaf(gg, vy, a1y);
    a0 += gg;
    ux = fmaf(G, vx, ux);
    uy = fmaf(G, vy, uy);
  }
  group_sum4<kGaussBlock>(a1x, a1y, a0, unused, red_a, warp, lane);
  group_sum2<kGaussBlock>(ux, uy, red_b, warp, lane);
  if (threadIdx.x == 0) {
    const float sc = __ldg(g_loss) * reg_coeff * (mask ? mask[n] : 1.0f) / __ldg(denom);
    dmu[2 * n] = sc * (a1x - a0 * ux);
    dmu[2 * n + 1] = sc * (a1y - a0 * uy);
  }
}

}  // namespace dsnt
