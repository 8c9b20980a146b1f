// head_preact.cuh -- the fused head for the reference's other heatmap pre-activations.
//
// src/dsnt/model.py:24-45 offers five ways to turn raw heatmaps z into a distribution P:
//     softmax | thresholded_softmax (threshold -0.5, src/dsnt/nn.py:119-157) | abs | relu | sigmoid
// the last three as  P = f(z) / (sum f(z) + 1e-12).  The tuned logits kernels (head_fast/head_stream) cover plain
// softmax, where P > 0 everywhere and every epsilon of the regularisers is droppable.  Here P can be exactly 0
// (thresholded softmax, relu) or far from normalised (an all-negative relu map sums to 0), so these kernels keep
// the reference's arithmetic term for term -- every 1e-24 inside the logs, the 1e-12 in the normaliser, sum P != 1
// in the variance gradient -- while still never materialising P:
//
//   forward   f_ij = act(z_ij),  S = sum f,  inv = 1/(S + eps_n),  P = f inv
//             mu = sum P (x_j, y_i);  D and c_reg = sum P r  (r = dD/dP) in the same launch
//   backward  dz_ij = f'_ij inv (a x_j + b y_i + rho r_ij - c),   c = a mu_x + b mu_y + rho c_reg
//             (thresholded softmax: f' = f, which is exactly the reference's custom backward out*(g - sum g out),
//              src/dsnt/nn.py:131-139)
//
// The forward keeps the heatmap in registers when it fits (one warp or one CTA per heatmap) and otherwise makes
// three passes over global memory, the later ones served by L2.  The backward is reduction-free streaming.
// These pre-activations are ablation settings (experiments/preact.json), so the regulariser and the activation are
// run-time switches of one kernel per shape rather than one tuned instantiation each.
#pragma once

#include "head_bwd.cuh"
#include "head_fwd.cuh"

namespace dsnt {

struct HeadPreactFwdParams {
  HeadFwdParams base;
  PreactCfg pc;
  FlipCfg fl;
};
struct HeadPreactBwdParams {
  HeadBwdParams base;
  PreactCfg pc;
};

// f(z).  m2 = max(z) * log2(e) for the softmax family.
__device__ __forceinline__ float preact_f(const PreactCfg& pc, float z, float m2) {
  switch (pc.preact) {
    case DSNT_PREACT_SOFTMAX: return ex2(fmaf(z, kLog2e, -m2));
    case DSNT_PREACT_TSOFTMAX: return z >= pc.threshold ? ex2(fmaf(z, kLog2e, -m2)) : 0.f;
    case DSNT_PREACT_ABS: return fabsf(z);
    case DSNT_PREACT_RELU: return fmaxf(z, 0.f);
    default: return __fdividef(1.0f, 1.0f + ex2(-z * kLog2e));   // sigmoid
  }
}

// f'(z) given f = f(z); abs'(0) = relu'(0) = 0 as in torch.
__device__ __forceinline__ float preact_df(const PreactCfg& pc, float z, float f) {
  switch (pc.preact) {
    case DSNT_PREACT_SOFTMAX:
    case DSNT_PREACT_TSOFTMAX: return f;
    case DSNT_PREACT_ABS: return z > 0.f ? 1.0f : (z < 0.f ? -1.0f : 0.f);
    case DSNT_PREACT_RELU: return z > 0.f ? 1.0f : 0.f;
    default: return f * (1.0f - f);
  }
}

// log2((a + eps)/(b + eps)) as ONE logarithm of a ratio: keeps the absolute error of lg2.approx instead of
// differencing two logs of magnitude up to 80.  a, b >= 0; the ratio spans [1e-24, 1e24], inside fp32 range.
__device__ __forceinline__ float lg2_ratio(float a, float b) { return lg2((a + kEps) * rcp(b + kEps)); }

// dD/dP at one pixel (every epsilon kept), natural-log units.
__device__ __forceinline__ float preact_r(int reg, float P, float G) {
  if (reg == DSNT_REG_KL) return fmaf(kLn2, lg2_ratio(P, G), P * rcp(P + kEps));
  if (reg == DSNT_REG_JS) {
    const float M = 0.5f * (P + G);
    return 0.5f * (fmaf(kLn2, lg2_ratio(P, M), P * rcp(P + kEps)) - M * rcp(M + kEps));
  }
  return 2.f * (P - G);   // MSE
}

// Visits the vectors of one heatmap owned by this thread: from registers (NV > 0) or from global memory (NV == 0).
// FLIP: every vector is averaged with the mirrored vector of the partner heatmap zf as it is loaded.
template <typename T, int VEC, int GROUP, int NV, bool FLIP = false>
struct HmSweep {
  float v[NV > 0 ? NV : 1][VEC];
  const T* zb;
  const T* zf;
  int nvec, wv, lane_g;

  __device__ __forceinline__ void load_vec(int f, int row, int cv, float (&out)[VEC]) const {
    VecIO<T, VEC>::load(zb, static_cast<long>(f) * VEC, out);
    if constexpr (FLIP) {
      float o[VEC];
      VecIO<T, VEC>::load(zf, static_cast<long>(row * wv + (wv - 1 - cv)) * VEC, o);
#pragma unroll
      for (int c = 0; c < VEC; ++c) out[c] = 0.5f * (out[c] + o[VEC - 1 - c]);
    }
  }

  __device__ __forceinline__ void load_all() {
    if constexpr (NV > 0) {
      VecWalker wk(lane_g, GROUP, wv);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int f = lane_g + k * GROUP;
        if (f < nvec) {
          load_vec(f, wk.row, wk.cv, v[k]);
          wk.next();
        } else {
#pragma unroll
          for (int c = 0; c < VEC; ++c) v[k][c] = 0.f;
        }
      }
    }
  }

  // fn(vals, row, col0): vals may be modified in place; the change persists only when the heatmap is resident.
  template <typename F>
  __device__ __forceinline__ void sweep(F&& fn) {
    VecWalker wk(lane_g, GROUP, wv);
    if constexpr (NV > 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        if (lane_g + k * GROUP < nvec) fn(v[k], wk.row, wk.cv * VEC);
        wk.next();
      }
    } else {
      for (int f = lane_g; f < nvec; f += GROUP) {
        float t[VEC];
        load_vec(f, wk.row, wk.cv, t);
        fn(t, wk.row, wk.cv * VEC);
        wk.next();
      }
    }
  }
};

// ================================================================================================ forward
template <typename T, int VEC, int GROUP, int NV, bool FLIP = false>
__global__ void __launch_bounds__(fwd_block_threads<GROUP>()) head_preact_fwd_kernel(const HeadPreactFwdParams ps) {
  constexpr int BLOCK = fwd_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  constexpr int NW = GROUP / 32;
  constexpr bool RESIDENT = NV > 0;
  extern __shared__ __align__(16) float dyn_smem[];
  __shared__ float red_m[GPB * NW];
  __shared__ float red_a[GPB * NW * 4];
  __shared__ float red_b[GPB * NW * 4];

  const HeadFwdParams& p = ps.base;
  const PreactCfg pc = ps.pc;
  const int reg = p.reg;
  const int tid = threadIdx.x;
  const int gid = tid / GROUP, lane_g = tid % GROUP, warp_g = lane_g >> 5, lane = tid & 31;
  const long hm = static_cast<long>(blockIdx.x) * GPB + gid;
  if (hm >= p.n) return;  // GROUP == 32 only: the whole warp leaves together

  const int H = p.H, W = p.W;
  const HmRef ref = locate(p.st, hm, static_cast<long>(H) * W * sizeof(T));
  HmSweep<T, VEC, GROUP, NV, FLIP> hs;
  hs.zb = reinterpret_cast<const T*>(static_cast<const char*>(p.z) + ref.z_bytes);
  hs.zf = nullptr;
  T* avg_out = nullptr;
  if constexpr (FLIP) {
    const int C = ps.fl.C;
    const long b = hm / C;
    const int c = static_cast<int>(hm - b * C);
    const long partner = p.n + b * C + (ps.fl.perm ? __ldg(ps.fl.perm + c) : c);
    hs.zf = reinterpret_cast<const T*>(p.z) + partner * H * W;
    if (ps.fl.avg_out) avg_out = reinterpret_cast<T*>(ps.fl.avg_out) + hm * H * W;
  }
  hs.wv = W / VEC;
  hs.nvec = H * hs.wv;
  hs.lane_g = lane_g;
  hs.load_all();

  float tx = 0.f, ty = 0.f;
  if (p.target) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p.target) + ref.nl);
    tx = t.x; ty = t.y;
  }
  float* tabx = dyn_smem + gid * table_floats(H, W);
  float* taby = tabx + ((W + 3) & ~3);
  float* scal = taby + ((H + 3) & ~3);
  const bool gauss = reg_needs_gauss(reg);
  if (gauss) gauss_tables_build<GROUP>(tabx, taby, scal, H, W, tx, ty, p.sigma, warp_g, lane);

  const float two_over_w = 2.0f / W, bias_w = 1.0f / W - 1.0f;
  const float two_over_h = 2.0f / H, bias_h = 1.0f / H - 1.0f;

  // ---- 1. max over ALL entries, also the ones below the threshold (src/dsnt/nn.py:124)
  float m2 = 0.f;
  if (preact_is_softmax(pc.preact)) {
    float mloc = -INFINITY;
    hs.sweep([&](float (&z)[VEC], int, int) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) mloc = fmaxf(mloc, z[c]);
    });
    m2 = group_max<GROUP>(mloc, red_m + gid * NW, warp_g, lane) * kLog2e;
  }

  // ---- 2. f = act(z) (kept in the registers when resident), S, sum f x, sum f y
  float S = 0.f, Sx = 0.f, Sy = 0.f, dummy = 0.f;
  hs.sweep([&](float (&z)[VEC], int row, int col0) {
    if constexpr (FLIP) {
      if (avg_out) VecIO<T, VEC>::store(avg_out, static_cast<long>(row) * W + col0, z);
    }
    float rs = 0.f;
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      const float f = preact_f(pc, z[c], m2);
      z[c] = f;
      rs += f;
      Sx = fmaf(f, axis_coord(col0 + c, two_over_w, bias_w), Sx);
    }
    S += rs;
    Sy = fmaf(rs, axis_coord(row, two_over_h, bias_h), Sy);
  });
  group_sum4<GROUP>(S, Sx, Sy, dummy, red_a + gid * NW * 4, warp_g, lane);  // barrier: tables are visible too
  if constexpr (GROUP == 32) __syncwarp();
  const float inv = 1.0f / (S + pc.eps);
  const float om = pc.eps * inv;   // 1 - sum P = eps/(S + eps), exact (S*inv would round it away)
  const float mux = Sx * inv, muy = Sy * inv;

  // ---- 3. regulariser value D and c_reg = sum P r
  GaussCtx g{tabx, taby, 0.f, 0.f};
  if (gauss) g = gauss_tables_finish(tabx, taby, scal);
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
  if (reg != DSNT_REG_NONE) {
    hs.sweep([&](float (&val)[VEC], int row, int col0) {
      float P[VEC];
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        if constexpr (RESIDENT) P[c] = val[c] * inv;
        else P[c] = preact_f(pc, val[c], m2) * inv;
      }
      if (reg == DSNT_REG_VAR) {
        float rs = 0.f;
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const float dx = axis_coord(col0 + c, two_over_w, bias_w) - mux;
          q0 = fmaf(P[c] * dx, dx, q0);
          rs += P[c];
        }
        const float dy = axis_coord(row, two_over_h, bias_h) - muy;
        q1 = fmaf(rs * dy, dy, q1);
        return;
      }
      const float gyn = g.ty[row] * g.ginv;
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float G = g.tx[col0 + c] * gyn, Pc = P[c];
        if (reg == DSNT_REG_KL) {
          q0 = fmaf(Pc, lg2_ratio(Pc, G), q0);                    // D / ln2
          q1 = fmaf(Pc, Pc * rcp(Pc + kEps), q1);                 // sum P^2/(P+eps)
        } else if (reg == DSNT_REG_JS) {
          const float M = 0.5f * (Pc + G);
          const float lp = lg2_ratio(Pc, M);
          q0 = fmaf(Pc, lp, q0);                                   // sum P (log2(P+e) - log2(M+e))
          q1 = fmaf(G, lg2_ratio(G, M), q1);                       // sum G (log2(G+e) - log2(M+e))
          q2 = fmaf(Pc, Pc * rcp(Pc + kEps) - M * rcp(M + kEps), q2);
        } else {
          const float df = Pc - G;
          q0 = fmaf(df, df, q0);
          q1 = fmaf(Pc, df, q1);
        }
      }
    });
    group_sum4<GROUP>(q0, q1, q2, q3, red_b + gid * NW * 4, warp_g, lane);
  }

  if (lane_g == 0) {
    float D = 0.f, creg = 0.f, vx = 0.f, vy = 0.f;
    if (reg == DSNT_REG_VAR) {
      vx = q0; vy = q1;
      const float s2 = p.sigma * p.sigma, ex = vx - s2, ey = vy - s2;
      D = ex * ex + ey * ey;
      // sum P r with r = 2(vx-s2)[(x-mu)^2 - 2 x mu (1 - sum P)] + ... : sum P x = mu_x exactly
      creg = 2.f * (ex * (vx - 2.f * mux * mux * om) + ey * (vy - 2.f * muy * muy * om));
    } else if (reg == DSNT_REG_KL) {
      D = kLn2 * q0;
      creg = D + q1;
    } else if (reg == DSNT_REG_JS) {
      D = 0.5f * kLn2 * (q0 + q1);
      creg = 0.5f * fmaf(kLn2, q0, q2);
    } else if (reg == DSNT_REG_MSE) {
      D = q0;
      creg = 2.f * q1;
    }
    // stats: [0] m2  [1] inv  [2] mu_x [3] mu_y [4] v_x [5] v_y [6] c_reg  [7] Gaussian normaliser, or 1 - sum P for var
    write_outputs(p, hm, m2, inv, mux, muy, vx, vy, creg, reg == DSNT_REG_VAR ? om : g.ginv, tx, ty, D);
  }
}

// ================================================================================================ backward
template <typename T, int VEC, int GROUP, int NV>
__global__ void __launch_bounds__(fwd_block_threads<GROUP>()) head_preact_bwd_kernel(const HeadPreactBwdParams ps) {
  constexpr int BLOCK = fwd_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  extern __shared__ __align__(16) float dyn_smem[];

  const HeadBwdParams& p = ps.base;
  const PreactCfg pc = ps.pc;
  const int reg = p.reg;
  const int tid = threadIdx.x;
  const int gid = tid / GROUP, lane_g = tid % GROUP;
  const long hm = static_cast<long>(blockIdx.x) * GPB + gid;
  if (hm >= p.n) return;  // warp path only

  const int H = p.H, W = p.W;
  const int wv = W / VEC, nvec = H * wv;
  const int f0 = blockIdx.y * (GROUP * NV) + lane_g;
  const HmRef ref = locate(p.st, hm, static_cast<long>(H) * W * sizeof(T));
  const T* zb = reinterpret_cast<const T*>(static_cast<const char*>(p.z) + ref.z_bytes);
  T* dzb = reinterpret_cast<T*>(static_cast<char*>(p.dz) + ref.dz_bytes);

  float v[NV][VEC];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int f = f0 + k * GROUP;
    if (f < nvec) VecIO<T, VEC>::load(zb, static_cast<long>(f) * VEC, v[k]);
  }

  // load_bwd_scalars<true>: m2, inv (= "invS"), mu, a, b, rho and c = a mu_x + b mu_y + rho c_reg
  const BwdScalars s = load_bwd_scalars<true>(p, hm, ref.nl, reg);
  const bool gauss = reg_needs_gauss(reg);
  const float two_over_w = 2.0f / W, bias_w = 1.0f / W - 1.0f;
  const float two_over_h = 2.0f / H, bias_h = 1.0f / H - 1.0f;

  float* tabx = dyn_smem + gid * table_floats(H, W);
  float* taby = tabx + ((W + 3) & ~3);
  if (gauss) {
    const float k2 = -0.5f / (p.sigma * p.sigma) * kLog2e;
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
  // variance: stats[7] carries 1 - sum P (there is no Gaussian); d v_x/dP = (x-mu)^2 - 2 x mu_x (1 - sum P)
  const float om = reg == DSNT_REG_VAR ? s.ginv : 0.f;
  const float cx = s.mux * om, cy = s.muy * om;

  VecWalker wk(f0, GROUP, wv);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int f = f0 + k * GROUP;
    if (f < nvec) {
      const int col0 = wk.cv * VEC;
      const float y = axis_coord(wk.row, two_over_h, bias_h);
      float rowc = fmaf(s.b, y, -s.c);
      if (reg == DSNT_REG_VAR) {
        const float dy = y - s.muy;
        rowc = fmaf(s.ky, fmaf(-2.f * y, cy, dy * dy), rowc);
      }
      const float gyn = gauss ? taby[wk.row] : 0.f;
      float out[VEC];
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float x = axis_coord(col0 + c, two_over_w, bias_w);
        const float z = v[k][c];
        const float fz = preact_f(pc, z, s.m2);
        float gmc = fmaf(s.a, x, rowc);
        if (reg == DSNT_REG_VAR) {
          const float dx = x - s.mux;
          gmc = fmaf(s.kx, fmaf(-2.f * x, cx, dx * dx), gmc);
        } else if (gauss) {
          gmc = fmaf(s.rho, preact_r(reg, fz * s.invS, tabx[col0 + c] * gyn), gmc);
        }
        out[c] = preact_df(pc, z, fz) * s.invS * gmc;
      }
      VecIO<T, VEC>::store(dzb, static_cast<long>(f) * VEC, out);
    }
    wk.next();
  }
}

}  // namespace dsnt
