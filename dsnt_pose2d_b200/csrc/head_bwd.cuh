// head_bwd.cuh -- fused, reduction-free backward of the DSNT head.
//
// Replaces the autograd replay of src/dsnt/nn.py:25-298 + F.softmax backward.  The forward saved O(1)
// statistics per heatmap (DSNT_STATS_K floats), so this kernel is pure streaming: read z, recompute
// P = 2^(z*log2e - m2) / S and the separable Gaussian, write
//     dz = P (a x_j + b y_i + rho r_ij - c),   c = a mu_x + b mu_y + rho c_reg     
// (or dP = a x + b y + rho r for heatmap input).  8 bytes/pixel for fp32 and no reduction or barrier
// beyond the one that publishes the W+H Gaussian table entries.
//
// Work decomposition: blockIdx.x = heatmap (or GPB heatmaps on the warp path), blockIdx.y = chunk of
// GROUP*NV vectors, so any heatmap size is covered by the same kernel.
#pragma once

#include "common.cuh"
#include "head_fwd.cuh"  // table_floats, kWarpPathBlock

namespace dsnt {

struct HeadBwdParams {
  const void* z;
  const float* target;    // [N,2] or null
  const float* mask;      // [N] or null
  const float* stats;     // [N,8]
  const float* g_coords;  // [N,2] or null
  const float* g_reg;     // [N] or null
  const float* g_loss;    // device scalar or null
  const float* denom;     // device scalar (with g_loss)
  void* dz;
  long n;
  int H, W;
  int reg;
  int flags;
  float sigma, reg_coeff;
  Stacks st;
};

// Everything a thread needs to know about its heatmap (uniform across the group).
struct BwdScalars {
  float m2, invS, mux, muy;
  float a, b, rho, c;
  float kx, ky;        // var: rho*2*(vx - s^2), rho*2*(vy - s^2)
  float cx, cy;        // heatmap-input var: mu_x (1 - s0), mu_y (1 - s0)
  float l2is;          // log2(1/S)
  float ginv;
  float tx, ty;
};

template <bool LOGITS>
__device__ __forceinline__ BwdScalars load_bwd_scalars(const HeadBwdParams& p, long hm, long nl, int reg) {
  BwdScalars s;
  const float4* st = reinterpret_cast<const float4*>(p.stats + hm * kStatsK);
  // L2 loads: a caller may have written the statistics a moment ago in the same stream; the read-only path (ld.global.nc)
  // is for data that is constant for the lifetime of the kernel only, L2 loads cost the same here
  const float4 s0 = __ldcg(st), s1 = __ldcg(st + 1);
  s.m2 = s0.x; s.invS = LOGITS ? s0.y : 1.0f; s.mux = s0.z; s.muy = s0.w;
  const float vx = s1.x, vy = s1.y, creg = s1.z;
  s.ginv = s1.w;
  s.tx = 0.f; s.ty = 0.f;
  if (p.target) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p.target) + nl);
    s.tx = t.x; s.ty = t.y;
  }
  float a = 0.f, b = 0.f, rho = 0.f;
  if (p.g_loss) {
    const float gl = __ldg(p.g_loss);
    const float w = __fdividef(p.mask ? __ldg(p.mask + nl) : 1.0f, __ldg(p.denom));
    if (p.target && !(p.flags & DSNT_FLAG_NO_EUCLID)) {
      const float dx = s.mux - s.tx, dy = s.muy - s.ty;
      const float d2 = dx * dx + dy * dy;
      // sqrt'(0) = inf: the reference back-propagates NaN there; default is 0.
      const float invd = d2 > 0.f ? rsqrtf(d2) : ((p.flags & DSNT_FLAG_STRICT_NAN) ? INFINITY : 0.f);
      a = gl * w * (dx * invd);
      b = gl * w * (dy * invd);
    }
    rho = gl * w * p.reg_coeff;
  }
  if (p.g_coords) {
    const float2 gc = __ldg(reinterpret_cast<const float2*>(p.g_coords) + hm);
    a += gc.x; b += gc.y;
  }
  if (p.g_reg) rho += __ldg(p.g_reg + hm);
  s.a = a; s.b = b; s.rho = rho;
  s.c = LOGITS ? fmaf(a, s.mux, fmaf(b, s.muy, rho * creg)) : 0.f;
  const float s2 = p.sigma * p.sigma;
  s.kx = reg == DSNT_REG_VAR ? rho * 2.f * (vx - s2) : 0.f;
  s.ky = reg == DSNT_REG_VAR ? rho * 2.f * (vy - s2) : 0.f;
  // heatmap input: sum P (x - mu_x) = mu_x (1 - sum P); stats[0] holds sum P there
  s.cx = LOGITS ? 0.f : s.mux * (1.0f - s0.x);
  s.cy = LOGITS ? 0.f : s.muy * (1.0f - s0.x);
  s.l2is = LOGITS ? __log2f(s.invS) : 0.f;
  return s;
}

template <typename T, int VEC, int GROUP, int NV, int REG, bool LOGITS>
__global__ void __launch_bounds__(fwd_block_threads<GROUP>()) head_bwd_kernel(const HeadBwdParams p) {
  constexpr int BLOCK = fwd_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  constexpr int NW = GROUP / 32;
  extern __shared__ __align__(16) float dyn_smem[];

  const int reg = REG >= 0 ? REG : p.reg;
  const int tid = threadIdx.x;
  const int gid = tid / GROUP, lane_g = tid % GROUP, warp_g = lane_g >> 5, lane = tid & 31;
  const long hm = static_cast<long>(blockIdx.x) * GPB + gid;
  if (hm >= p.n) return;  // warp path only

  const int H = p.H, W = p.W;
  const int wv = W / VEC, nvec = H * wv;
  const int f0 = blockIdx.y * (GROUP * NV) + lane_g;
  const HmRef ref = locate(p.st, hm, static_cast<long>(H) * W * sizeof(T));
  const T* zb = reinterpret_cast<const T*>(static_cast<const char*>(p.z) + ref.z_bytes);
  T* dzb = reinterpret_cast<T*>(static_cast<char*>(p.dz) + ref.dz_bytes);

  const bool gauss = reg_needs_gauss(reg);
  const bool need_z = LOGITS || gauss;

  // ---- issue the loads first
  float v[NV][VEC];
  if (need_z) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int f = f0 + k * GROUP;
      if (f < nvec) VecIO<T, VEC>::load(zb, static_cast<long>(f) * VEC, v[k]);
    }
  }

  const BwdScalars s = load_bwd_scalars<LOGITS>(p, hm, ref.nl, reg);

  // ---- Gaussian tables: tabx[j] = gx_j, taby[i] = gy_i * ginv (the forward saved the normaliser)
  float* tabx = dyn_smem + gid * table_floats(H, W);
  float* taby = tabx + ((W + 3) & ~3);
  if (gauss) {
    const float k2 = -0.5f / (p.sigma * p.sigma) * kLog2e;
    const float two_over_w = 2.0f / W, bias_w = 1.0f / W - 1.0f;
    const float two_over_h = 2.0f / H, bias_h = 1.0f / H - 1.0f;
    for (int j = lane_g; j < W + H; j += GROUP) {
      if (j < W) {
        const float dx = axis_coord(j, two_over_w, bias_w) - s.tx;
        tabx[j] = ex2(k2 * dx * dx);
      } else {
        const float dy = axis_coord(j - W, two_over_h, bias_h) - s.ty;
        taby[j - W] = ex2(k2 * dy * dy) * s.ginv;
      }
    }
    group_barrier<GROUP>();
  }

  const float two_over_w = 2.0f / W, bias_w = 1.0f / W - 1.0f;
  const float two_over_h = 2.0f / H, bias_h = 1.0f / H - 1.0f;

  VecWalker wk(f0, GROUP, wv);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int f = f0 + k * GROUP;
    if (f < nvec) {
      const int col0 = wk.cv * VEC;
      const float y = axis_coord(wk.row, two_over_h, bias_h);
      // row-constant part of (g - c)
      float rowc = fmaf(s.b, y, -s.c);
      if (reg == DSNT_REG_VAR) {
        const float dy = y - s.muy;
        rowc = fmaf(s.ky, LOGITS ? dy * dy : fmaf(-2.f * y, s.cy, dy * dy), rowc);
      }
      if (LOGITS && reg == DSNT_REG_KL) rowc += s.rho;  // the "+1" of r = ln P - ln G' + 1
      const float gyn = gauss ? taby[wk.row] : 0.f;
      float out[VEC];
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float x = axis_coord(col0 + c, two_over_w, bias_w);
        float gmc = fmaf(s.a, x, rowc);  // a x + b y - c (+ row terms)
        if (reg == DSNT_REG_VAR) {
          const float dx = x - s.mux;
          gmc = fmaf(s.kx, LOGITS ? dx * dx : fmaf(-2.f * x, s.cx, dx * dx), gmc);
        }
        if constexpr (LOGITS) {
          float t = fmaf(v[k][c], kLog2e, -s.m2);
          const float P = ex2(t) * s.invS;
          if (reg == DSNT_REG_KL) {
            t = fmaxf(t, -1e30f);  // z = -inf: P = 0, keep 0 * t finite
            const float G = tabx[col0 + c] * gyn;
            gmc = fmaf(s.rho * kLn2, (t + s.l2is) - lg2(G + kEps), gmc);
          } else if (reg == DSNT_REG_JS) {
            // r = 1/2 [ln P - ln(M+eps)] = -1/2 ln2 (lg2(1+q) - 1),  q = (G + 2 eps)/P.  The ratio form keeps the
            // absolute error of lg2.approx (2^-22 near 1) instead of differencing two logs of magnitude ~10-20,
            // which matters exactly where P ~ G (a trained heatmap's peak).
            const float q = fmaf(tabx[col0 + c], gyn, 2.f * kEps) * rcp(fmaxf(P, 1e-37f));
            gmc = fmaf(-0.5f * s.rho * kLn2, lg2(1.0f + q) - 1.0f, gmc);
          } else if (reg == DSNT_REG_MSE) {
            gmc = fmaf(2.f * s.rho, fmaf(-tabx[col0 + c], gyn, P), gmc);
          }
          out[c] = P * gmc;
        } else {
          if (gauss) {
            const float P = v[k][c];
            const float G = tabx[col0 + c] * gyn;
            float r;
            if (reg == DSNT_REG_KL) {
              r = kLn2 * (lg2(P + kEps) - lg2(G + kEps)) + P / (P + kEps);
            } else if (reg == DSNT_REG_JS) {
              const float M = 0.5f * (P + G);
              r = 0.5f * (kLn2 * (lg2(P + kEps) - lg2(M + kEps)) + P / (P + kEps) - M / (M + kEps));
            } else {
              r = 2.f * (P - G);
            }
            gmc = fmaf(s.rho, r, gmc);
          }
          out[c] = gmc;
        }
      }
      VecIO<T, VEC>::store(dzb, static_cast<long>(f) * VEC, out);
    }
    wk.next();
  }
}

}  // namespace dsnt
