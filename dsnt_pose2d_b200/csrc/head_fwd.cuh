// head_fwd.cuh -- fused forward of the DSNT head.
//
// One pass over each heatmap: flat softmax statistics (max, sum), coordinate expectations, and the
// regulariser against an on-the-fly separable Gaussian.  Replaces src/dsnt/model.py:24-30 (softmax),
// src/dsnt/nn.py:25-78 (grid, expectation, dsnt), :112-114 (distance), :168-216 + :219-298 (regularisers).
//
// Two kernels:
//   head_fwd_kernel        the heatmap lives in registers (GROUP threads x NV vectors x VEC elements);
//                          GROUP == 32: one warp per heatmap, no block barrier; GROUP == blockDim: one CTA.
//   head_fwd_large_kernel  heatmaps too big for registers (e.g. 256x256): two streaming passes, the
//                          second one served by L2 (148 CTAs x 256 KiB in flight << 126 MB).
//
// Arithmetic (the closed forms are pinned in oracle/closed_form.py):
//   t_ij = z_ij*log2e - m2,  e_ij = 2^t_ij,  S = sum e,  P = e/S,  mu = sum P (x_j, y_i)
//   T    = sum e t           (sum P ln P = ln2*T/S - ln S, no per-pixel log)
//   KL   D = sum P ln P - ln2 * sum P lg2(G+eps)                              c_reg = D + 1
//   JS   D = 1/2 [sum P ln P + sum G ln G - 2 ln2 sum M lg2(M+eps)]           c_reg = 1/2 [sum P ln P - ln2 sum P lg2(M+eps)]
//        sum G ln G is closed-form from the two axis tables (separable Gaussian)
//   MSE  D = sum (P-G)^2                                                       c_reg = 2 sum P (P-G)
//   var  D = (vx-s^2)^2 + (vy-s^2)^2,  vx = sum P (x-mu_x)^2 (two-pass, no cancellation)   c_reg = 2(vx-s^2)vx + 2(vy-s^2)vy
// Epsilon handling: ln(P+1e-24) -> ln P and P/(P+eps) -> 1 in logits mode; every such term is weighted by
// P, so the deviation is < 1e-15 absolute (Appendix B.4).  ln(G+eps) and ln(M+eps) keep their epsilon
// wherever it can matter (KL against a vanishing Gaussian).  Heatmap-input mode keeps every epsilon.
#pragma once

#include "common.cuh"

namespace dsnt {

struct HeadFwdParams {
  const void* z;
  const float* target;  // [N,2] or null
  float* coords;        // [N,2]
  float* stats;         // [N,8] or null
  float* terms;         // [N,2] or null
  long n;               // heatmaps in this launch (all stacks)
  int H, W;
  int reg;              // used only by REG < 0 (dynamic) instantiations
  float sigma;
  Stacks st;
};

// Flip test-time augmentation fused into the load (src/dsnt/inference.py:36-46): the raw heatmaps of the
// original images are heatmaps [0, batch*C), those of the mirrored images [batch*C, 2*batch*C); what the head sees is
//     z'[b,c,i,j] = (z[b,c,i,j] + z[batch+b, perm[c], i, W-1-j]) / 2
struct FlipCfg {
  const int* perm;   // [C] joint permutation under a horizontal flip (device), or null = identity
  void* avg_out;     // optional [batch*C,H,W]: the averaged raw heatmaps, same dtype as z
  int C;
};

// How raw heatmaps become a distribution (src/dsnt/model.py:24-45): P = f(z) / (sum f(z) + eps).
struct PreactCfg {
  int preact;        // DSNT_PREACT_*
  float threshold;   // thresholded softmax: keep z >= threshold
  float eps;         // added to the normaliser sum (1e-12 in the reference, 0 for plain softmax)
};

__host__ __device__ constexpr bool preact_is_softmax(int pa) {
  return pa == DSNT_PREACT_SOFTMAX || pa == DSNT_PREACT_TSOFTMAX;
}

constexpr int kWarpPathBlock = 128;  // 4 heatmaps per CTA on the warp-per-heatmap path

template <int GROUP>
__host__ __device__ constexpr int fwd_block_threads() { return GROUP >= 64 ? GROUP : kWarpPathBlock; }

// Shared-memory floats needed per group for the Gaussian tables (16-byte aligned sections).
__host__ __device__ inline int table_floats(int H, int W) { return ((W + 3) & ~3) + ((H + 3) & ~3) + 8; }

// Per-heatmap Gaussian context shared by phase B of both forward kernels.
struct GaussCtx {
  const float* tx;  // gx_j (unnormalised)
  const float* ty;  // gy_i (unnormalised)
  float ginv;       // 1 / (sum_x * sum_y + 1e-24)
  float sumGlnG;    // sum G ln G (natural log), closed form
};

// Builds the tables with warps `wx` and `wy` of the group, then (after the caller's barrier) finish() derives
// the normaliser.  scal[0..3] = sum_x, ent_x, sum_y, ent_y.
template <int GROUP>
__device__ __forceinline__ void gauss_tables_build(float* tabx, float* taby, float* scal, int H, int W, float tx,
                                                   float ty, float sigma, int warp_g, int lane) {
  const float k2 = -0.5f / (sigma * sigma) * kLog2e;
  constexpr int NW = GROUP / 32;
  if (warp_g == 0) {
    float s, h;
    gauss_axis_table(tabx, W, tx, k2, lane, s, h);
    if (lane == 0) { scal[0] = s; scal[1] = h; }
  }
  if (warp_g == (NW > 1 ? 1 : 0)) {
    float s, h;
    gauss_axis_table(taby, H, ty, k2, lane, s, h);
    if (lane == 0) { scal[2] = s; scal[3] = h; }
  }
}

__device__ __forceinline__ GaussCtx gauss_tables_finish(const float* tabx, const float* taby, const float* scal) {
  GaussCtx g;
  g.tx = tabx;
  g.ty = taby;
  const float sx = scal[0], ex = scal[1], sy = scal[2], ey = scal[3];
  const float tot = sx * sy;
  g.ginv = 1.0f / (tot + kEps);
  // sum G ln G = ginv * [ sy*Ent_x + sx*Ent_y + tot*ln ginv ],  Ent = ln2 * sum g log2 g
  g.sumGlnG = g.ginv * (kLn2 * (sy * ex + sx * ey) + tot * __logf(g.ginv));
  return g;
}

// ------------------------------------------------------------------------------------------------
// Phase-B accumulation for one vector of VEC pixels in row `row`, columns col0..col0+VEC-1.
//   e[]   : 2^t (logits mode) or P itself (heatmap mode)
//   q0,q1 : the two running sums whose meaning depends on `reg` (see finalize_terms)
template <int VEC, bool LOGITS>
__device__ __forceinline__ void phase_b_vec(int reg, const float (&e)[VEC], int row, int col0, float invS, float mux,
                                            float muy, float two_over_w, float bias_w, float yv, const GaussCtx& g,
                                            float& q0, float& q1) {
  if (reg == DSNT_REG_VAR) {
    float rs = 0.f;
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      const float dx = axis_coord(col0 + c, two_over_w, bias_w) - mux;
      q0 = fmaf(e[c] * dx, dx, q0);
      rs += e[c];
    }
    const float dy = yv - muy;
    q1 = fmaf(rs * dy, dy, q1);
    return;
  }
  if (!reg_needs_gauss(reg)) return;
  float gx[VEC];
  if constexpr (VEC == 4) {
    const float4 t = *reinterpret_cast<const float4*>(g.tx + col0);
    gx[0] = t.x; gx[1] = t.y; gx[2] = t.z; gx[3] = t.w;
  } else if constexpr (VEC == 8) {
    const float4 t0 = *reinterpret_cast<const float4*>(g.tx + col0);
    const float4 t1 = *reinterpret_cast<const float4*>(g.tx + col0 + 4);
    gx[0] = t0.x; gx[1] = t0.y; gx[2] = t0.z; gx[3] = t0.w;
    gx[4] = t1.x; gx[5] = t1.y; gx[6] = t1.z; gx[7] = t1.w;
  } else {
#pragma unroll
    for (int c = 0; c < VEC; ++c) gx[c] = g.tx[col0 + c];
  }
  const float gyn = g.ty[row] * g.ginv;  // normalised row factor: G_ij = gx_j * gyn
  if (reg == DSNT_REG_KL) {
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      const float G = gx[c] * gyn;
      if constexpr (LOGITS) q0 = fmaf(e[c], lg2(G + kEps), q0);
      else q0 = fmaf(e[c], lg2(e[c] + kEps) - lg2(G + kEps), q0);
    }
  } else if (reg == DSNT_REG_JS) {
    if constexpr (LOGITS) {
      const float gyh = 0.5f * gyn, hs = 0.5f * invS;
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float M = fmaf(e[c], hs, gx[c] * gyh);
        const float L = lg2(M + kEps);
        q0 = fmaf(M, L, q0);
        q1 = fmaf(e[c], L, q1);
      }
    } else {
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float G = gx[c] * gyn, P = e[c];
        const float L = lg2(0.5f * (P + G) + kEps);
        q0 += P * (lg2(P + kEps) - L) + G * (lg2(G + kEps) - L);
      }
    }
  } else {  // MSE
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      const float diff = fmaf(e[c], invS, -gx[c] * gyn);
      q0 = fmaf(diff, diff, q0);
      q1 = fmaf(e[c], diff, q1);
    }
  }
}

// Turns the reduced sums into (D, c_reg).   plnp = sum P ln P (logits mode only).
template <bool LOGITS>
__device__ __forceinline__ void finalize_terms(int reg, float q0, float q1, float invS, float plnp, float sigma,
                                               const GaussCtx& g, float& D, float& creg, float& vx, float& vy) {
  D = 0.f; creg = 0.f; vx = 0.f; vy = 0.f;
  if (reg == DSNT_REG_VAR) {
    vx = q0 * invS; vy = q1 * invS;
    const float s2 = sigma * sigma, ex = vx - s2, ey = vy - s2;
    D = ex * ex + ey * ey;
    creg = 2.f * (ex * vx + ey * vy);
  } else if (reg == DSNT_REG_KL) {
    if constexpr (LOGITS) { D = plnp - kLn2 * invS * q0; creg = D + 1.f; }
    else D = kLn2 * q0;
  } else if (reg == DSNT_REG_JS) {
    if constexpr (LOGITS) {
      D = 0.5f * (plnp + g.sumGlnG - 2.f * kLn2 * q0);
      creg = 0.5f * (plnp - kLn2 * invS * q1);
    } else D = 0.5f * kLn2 * q0;
  } else if (reg == DSNT_REG_MSE) {
    D = q0;
    if constexpr (LOGITS) creg = 2.f * invS * q1;
  }
}

__device__ __forceinline__ void write_outputs(const HeadFwdParams& p, long hm, float s0, float s1, float mux, float muy,
                                              float vx, float vy, float creg, float ginv, float tx, float ty, float D) {
  reinterpret_cast<float2*>(p.coords)[hm] = make_float2(mux, muy);
  if (p.stats) {
    float4* st = reinterpret_cast<float4*>(p.stats + hm * kStatsK);
    st[0] = make_float4(s0, s1, mux, muy);
    st[1] = make_float4(vx, vy, creg, ginv);
  }
  if (p.terms) {
    float dist = 0.f;
    if (p.target) {
      const float dx = mux - tx, dy = muy - ty;
      dist = sqrtf(dx * dx + dy * dy);
    }
    reinterpret_cast<float2*>(p.terms)[hm] = make_float2(dist, D);
  }
}

// ================================================================================================
// Register-resident kernel.  REG < 0 selects the regulariser at run time (scalar / odd-size path).
template <typename T, int VEC, int GROUP, int NV, int REG, bool LOGITS>
__global__ void __launch_bounds__(fwd_block_threads<GROUP>()) head_fwd_kernel(const HeadFwdParams p) {
  constexpr int BLOCK = fwd_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  constexpr int NW = GROUP / 32;
  extern __shared__ __align__(16) float dyn_smem[];
  __shared__ float red_m[GPB * NW];
  __shared__ float red_a[GPB * NW * 4];
  __shared__ float red_b[GPB * NW * 2];

  const int reg = REG >= 0 ? REG : p.reg;
  const int tid = threadIdx.x;
  const int gid = tid / GROUP, lane_g = tid % GROUP, warp_g = lane_g >> 5, lane = tid & 31;
  const long hm = static_cast<long>(blockIdx.x) * GPB + gid;
  if (hm >= p.n) return;  // GROUP == 32 only (grid is exact otherwise): the whole warp leaves together

  const int H = p.H, W = p.W;
  const int wv = W / VEC, nvec = H * wv;
  const HmRef ref = locate(p.st, hm, static_cast<long>(H) * W * sizeof(T));
  const T* zb = reinterpret_cast<const T*>(static_cast<const char*>(p.z) + ref.z_bytes);

  // ---- 1. issue every load of this thread up front (NV x 128-bit in flight per thread)
  float v[NV][VEC];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int f = lane_g + k * GROUP;
    if (f < nvec) {
      VecIO<T, VEC>::load(zb, static_cast<long>(f) * VEC, v[k]);
    } else {
#pragma unroll
      for (int c = 0; c < VEC; ++c) v[k][c] = LOGITS ? -INFINITY : 0.f;
    }
  }

  float tx = 0.f, ty = 0.f;
  if (p.target) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p.target) + ref.nl);
    tx = t.x; ty = t.y;
  }

  // ---- Gaussian tables (built by warps 0/1 while the loads are in flight)
  float* tabx = dyn_smem + gid * table_floats(H, W);
  float* taby = tabx + ((W + 3) & ~3);
  float* scal = taby + ((H + 3) & ~3);
  const bool gauss = reg_needs_gauss(reg);
  if (gauss) gauss_tables_build<GROUP>(tabx, taby, scal, H, W, tx, ty, p.sigma, warp_g, lane);

  // ---- 2. max
  float m2 = 0.f;
  if constexpr (LOGITS) {
    float mloc = -INFINITY;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int c = 0; c < VEC; ++c) mloc = fmaxf(mloc, v[k][c]);
    const float m = group_max<GROUP>(mloc, red_m + gid * NW, warp_g, lane);
    m2 = m * kLog2e;
  }

  // ---- 3. phase A: e = 2^t, S, Sx, Sy, T
  const float two_over_w = 2.0f / W, bias_w = 1.0f / W - 1.0f;
  const float two_over_h = 2.0f / H, bias_h = 1.0f / H - 1.0f;
  float S = 0.f, Sx = 0.f, Sy = 0.f, Tt = 0.f;
  {
    VecWalker wk(lane_g, GROUP, wv);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if (lane_g + k * GROUP < nvec) {
        const int col0 = wk.cv * VEC;
        float rs = 0.f;
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          float e;
          if constexpr (LOGITS) {
            const float t = fmaf(v[k][c], kLog2e, -m2);
            e = ex2(t);
            if (reg == DSNT_REG_KL || reg == DSNT_REG_JS) Tt = fmaf(e, fmaxf(t, -1e30f), Tt);
            v[k][c] = e;
          } else {
            e = v[k][c];
          }
          rs += e;
          Sx = fmaf(e, axis_coord(col0 + c, two_over_w, bias_w), Sx);
        }
        S += rs;
        Sy = fmaf(rs, axis_coord(wk.row, two_over_h, bias_h), Sy);
      }
      wk.next();
    }
  }
  group_sum4<GROUP>(S, Sx, Sy, Tt, red_a + gid * NW * 4, warp_g, lane);  // barrier: tables are visible too
  if constexpr (GROUP == 32) __syncwarp();

  const float invS = LOGITS ? 1.0f / S : 1.0f;
  const float mux = Sx * invS, muy = Sy * invS;
  const float plnp = LOGITS ? kLn2 * invS * Tt - __logf(S) : 0.f;

  // ---- 4. phase B
  GaussCtx g{tabx, taby, 0.f, 0.f};
  if (gauss) g = gauss_tables_finish(tabx, taby, scal);
  float q0 = 0.f, q1 = 0.f;
  if (reg != DSNT_REG_NONE) {
    VecWalker wk(lane_g, GROUP, wv);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if (lane_g + k * GROUP < nvec)
        phase_b_vec<VEC, LOGITS>(reg, v[k], wk.row, wk.cv * VEC, invS, mux, muy, two_over_w, bias_w,
                                 axis_coord(wk.row, two_over_h, bias_h), g, q0, q1);
      wk.next();
    }
    group_sum2<GROUP>(q0, q1, red_b + gid * NW * 2, warp_g, lane);
  }

  // ---- 5. outputs
  if (lane_g == 0) {
    float D, creg, vx, vy;
    finalize_terms<LOGITS>(reg, q0, q1, invS, plnp, p.sigma, g, D, creg, vx, vy);
    write_outputs(p, hm, LOGITS ? m2 : S, LOGITS ? invS : 0.f, mux, muy, vx, vy, creg, g.ginv, tx, ty, D);
  }
}

// ================================================================================================
// Streaming two-pass kernel for heatmaps that do not fit in registers.  One CTA per heatmap.
constexpr int kLargeBlock = 512;

template <typename T, int VEC, int REG, bool LOGITS>
__global__ void __launch_bounds__(kLargeBlock) head_fwd_large_kernel(const HeadFwdParams p) {
  constexpr int GROUP = kLargeBlock, NW = GROUP / 32;
  extern __shared__ __align__(16) float dyn_smem[];
  __shared__ float red_m[NW];
  __shared__ float red_a[NW * 4];
  __shared__ float red_b[NW * 4];

  const int reg = REG >= 0 ? REG : p.reg;
  const int tid = threadIdx.x, warp_g = tid >> 5, lane = tid & 31;
  const long hm = blockIdx.x;
  const int H = p.H, W = p.W;
  const int wv = W / VEC, nvec = H * wv;
  const HmRef ref = locate(p.st, hm, static_cast<long>(H) * W * sizeof(T));
  const T* zb = reinterpret_cast<const T*>(static_cast<const char*>(p.z) + ref.z_bytes);

  float tx = 0.f, ty = 0.f;
  if (p.target) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p.target) + ref.nl);
    tx = t.x; ty = t.y;
  }
  float* tabx = dyn_smem;
  float* taby = tabx + ((W + 3) & ~3);
  float* scal = taby + ((H + 3) & ~3);
  const bool gauss = reg_needs_gauss(reg);
  if (gauss) gauss_tables_build<GROUP>(tabx, taby, scal, H, W, tx, ty, p.sigma, warp_g, lane);

  const float two_over_w = 2.0f / W, bias_w = 1.0f / W - 1.0f;
  const float two_over_h = 2.0f / H, bias_h = 1.0f / H - 1.0f;

  // ---- pass 1: online max / sum / first moments (per-thread running max, rescale only when it grows)
  float mt = -INFINITY, S = 0.f, Sx = 0.f, Sy = 0.f;
  {
    VecWalker wk(tid, GROUP, wv);
    for (int f = tid; f < nvec; f += GROUP) {
      float v[VEC];
      VecIO<T, VEC>::load(zb, static_cast<long>(f) * VEC, v);
      const int col0 = wk.cv * VEC;
      if constexpr (LOGITS) {
        float vm = v[0];
#pragma unroll
        for (int c = 1; c < VEC; ++c) vm = fmaxf(vm, v[c]);
        if (vm > mt) {
          const float sc = ex2((mt - vm) * kLog2e);  // 2^-inf = 0 on the first vector
          S *= sc; Sx *= sc; Sy *= sc;
          mt = vm;
        }
      }
      const float mt2 = mt * kLog2e;
      float rs = 0.f;
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float e = LOGITS ? ex2(fmaf(v[c], kLog2e, -mt2)) : v[c];
        rs += e;
        Sx = fmaf(e, axis_coord(col0 + c, two_over_w, bias_w), Sx);
      }
      S += rs;
      Sy = fmaf(rs, axis_coord(wk.row, two_over_h, bias_h), Sy);
      wk.next();
    }
  }
  float m2 = 0.f;
  if constexpr (LOGITS) {
    const float m = group_max<GROUP>(mt, red_m, warp_g, lane);
    const float sc = (mt == -INFINITY) ? 0.f : ex2((mt - m) * kLog2e);  // threads that saw no vector
    S *= sc; Sx *= sc; Sy *= sc;
    m2 = m * kLog2e;
  }
  float dummy = 0.f;
  group_sum4<GROUP>(S, Sx, Sy, dummy, red_a, warp_g, lane);
  const float invS = LOGITS ? 1.0f / S : 1.0f;
  const float mux = Sx * invS, muy = Sy * invS;

  // ---- pass 2 (L2-resident re-read): sum P t and the regulariser sums
  GaussCtx g{tabx, taby, 0.f, 0.f};
  if (gauss) g = gauss_tables_finish(tabx, taby, scal);
  float q0 = 0.f, q1 = 0.f, Tt = 0.f;
  if (reg != DSNT_REG_NONE) {
    VecWalker wk(tid, GROUP, wv);
    for (int f = tid; f < nvec; f += GROUP) {
      float v[VEC];
      VecIO<T, VEC>::load(zb, static_cast<long>(f) * VEC, v);
      if constexpr (LOGITS) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const float t = fmaf(v[c], kLog2e, -m2);
          const float e = ex2(t);
          if (reg == DSNT_REG_KL || reg == DSNT_REG_JS) Tt = fmaf(e, fmaxf(t, -1e30f), Tt);
          v[c] = e;
        }
      }
      phase_b_vec<VEC, LOGITS>(reg, v, wk.row, wk.cv * VEC, invS, mux, muy, two_over_w, bias_w,
                               axis_coord(wk.row, two_over_h, bias_h), g, q0, q1);
      wk.next();
    }
    float dummy2 = 0.f;
    group_sum4<GROUP>(q0, q1, Tt, dummy2, red_b, warp_g, lane);
  }
  if (tid == 0) {
    const float plnp = LOGITS ? kLn2 * invS * Tt - __logf(S) : 0.f;
    float D, creg, vx, vy;
    finalize_terms<LOGITS>(reg, q0, q1, invS, plnp, p.sigma, g, D, creg, vx, vy);
    write_outputs(p, hm, LOGITS ? m2 : S, LOGITS ? invS : 0.f, mux, muy, vx, vy, creg, g.ginv, tx, ty, D);
  }
}

}  // namespace dsnt
