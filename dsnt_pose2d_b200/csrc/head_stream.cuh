// head_stream.cuh -- the bandwidth-path kernels of the DSNT head (logits input, vectorised layouts).
//
// v0 of this library kept the heatmap in registers and evaluated the divergence on every pixel; ncu showed
// both kernels ISSUE-bound (76-84 % issue-active, 26-46 instructions per pixel; profiles/r01_v0_*).  These
// kernels cut the work per pixel to ~7 instructions:
//
//  * forward = ONE streaming pass with a per-thread online softmax (running max, rescale only when it
//    grows), accumulating S, S_x, S_y (and T = sum e*t for KL) -- no second look at the heatmap, nothing
//    kept in registers, any heatmap size;
//  * the Gaussian target is separable and, for the sigma the models use (1 px), negligible outside a
//    window of ~17x17 (JS) / 25x25 (KL) pixels around the target.  Outside that window every G-dependent
//    term has a closed form (G+eps == eps exactly for KL; ln P - ln M = ln 2 for JS), so the divergence is
//    evaluated ONLY on the window pixels, re-read from L2 right after the streaming pass.  The window is
//    derived per heatmap from the requested tolerance (geom.r2_win), so a large sigma simply makes the
//    window the whole image: the result never depends on this optimisation (error bound: W*H*theta);
//  * the window sums are accumulated in DIFFERENCE form (P (log2 P - log2 M' - 1), G (log2 G - log2 M')), which
//    also removes the cancellation between sum P ln P and sum P ln M' that limited c_reg in v0;
//  * geometry (coordinates of a thread's columns, row stride) is hoisted out of the pixel loop whenever the
//    group size is a multiple of the vectors per row (FIXC), which holds for 32/64/128/256-wide maps;
//  * backward = the same streaming loop: out-of-window pixels cost FFMA, EX2, FMUL, FADD, FMUL.
//
// A "group" is one warp (GROUP == 32: several heatmaps per CTA, no block barrier) or one CTA (GROUP == 256).
#pragma once

#include <cmath>

#include "common.cuh"
#include "head_bwd.cuh"
#include "head_fwd.cuh"

namespace dsnt {

// JS / MSE: pixels whose Gaussian weight is below theta * max along EITHER axis take the closed forms (ln P - ln M = ln 2,
// (P - G)^2 = P^2).  theta = 1e-10 is a window of +-6.8 sigma (16 pixels wide at sigma = 1 px, one pixel of padding included); measured
// in fp64 over random / peaked / trained-like heatmaps the truncation changes D by < 1e-11 and dL/dz by < 3e-10 relative
// (1e-14, the first choice, gave a 19-pixel window: a third more window work for nothing).
constexpr float kThetaJS = 1e-10f;
constexpr float kThetaKL = 3e-32f;   // G + 1e-24 == 1e-24 exactly in fp32 below this
constexpr float kLog2Eps = -79.726274277296700f;   // log2(1e-24)
constexpr float kLnEps = -55.262042231857095f;     // ln(1e-24)

// f(z) for the other pre-activations of the reference (src/dsnt/model.py:31-41) on the tuned kernels.  PA is a
// template parameter, so the softmax instantiations are unchanged.  Softmax family: t = z log2e - max log2e is given
// and e = 2^t (masked below the threshold); the others return f and leave t alone (see act_log2).
template <int PA>
__device__ __forceinline__ float act_fast(float z, float t, float thr) {
  if constexpr (PA == DSNT_PREACT_SOFTMAX) return ex2(t);
  else if constexpr (PA == DSNT_PREACT_TSOFTMAX) return z >= thr ? ex2(t) : 0.f;
  else if constexpr (PA == DSNT_PREACT_ABS) return fabsf(z);
  else if constexpr (PA == DSNT_PREACT_RELU) return fmaxf(z, 0.f);
  else return rcp(1.0f + ex2(-z * kLog2e));
}
// log2 f(z): free for the softmax family (it is t), one MUFU otherwise; 0 where f = 0 (every use is weighted by f)
template <int PA>
__device__ __forceinline__ float act_log2(float e, float t) {
  if constexpr (preact_is_softmax(PA)) return t;
  else return e > 0.f ? lg2(e) : 0.f;
}
// f'(z) given f(z) (thresholded softmax: the reference's custom backward uses out itself, src/dsnt/nn.py:131-139)
template <int PA>
__device__ __forceinline__ float act_grad(float z, float e) {
  if constexpr (preact_is_softmax(PA)) return e;
  else if constexpr (PA == DSNT_PREACT_ABS) return z > 0.f ? 1.0f : (z < 0.f ? -1.0f : 0.f);
  else if constexpr (PA == DSNT_PREACT_RELU) return z > 0.f ? 1.0f : 0.f;
  else return e * (1.0f - e);
}

// Launch-uniform geometry, computed on the host (no divisions in the kernels).
struct Geom {
  float two_over_w, bias_w, two_over_h, bias_h;  // x_j = j*two_over_w + bias_w  (src/dsnt/nn.py:30-37)
  float half_w, half_h;                          // W/2, H/2: pixel index of a coordinate = (x+1)*W/2 - 1/2
  float k2;                                      // -0.5/sigma^2 * log2(e)
  float r2_win;                                  // window: dx^2 - dx*^2 <= r2_win  (= log2(theta)/k2)
  float dy_step;                                 // FIXC: y advance per group stride = rstep * two_over_h
  int wv, nvec, rstep;                           // vectors per row, per heatmap; FIXC: rows per group stride
};

inline Geom make_geom(int H, int W, int vec, int group, float sigma, int reg) {
  Geom g;
  g.two_over_w = 2.0f / W; g.bias_w = 1.0f / W - 1.0f;
  g.two_over_h = 2.0f / H; g.bias_h = 1.0f / H - 1.0f;
  g.half_w = 0.5f * W; g.half_h = 0.5f * H;
  const double k2 = -0.5 / (static_cast<double>(sigma) * sigma) * 1.4426950408889634;
  g.k2 = static_cast<float>(k2);
  const double theta = reg == DSNT_REG_KL ? kThetaKL : kThetaJS;
  g.r2_win = static_cast<float>(std::log2(theta) / k2);
  g.wv = W / vec; g.nvec = H * g.wv;
  g.rstep = group / g.wv;
  g.dy_step = g.rstep * g.two_over_h;
  return g;
}

template <int GROUP>
__host__ __device__ constexpr int stream_block_threads() { return GROUP >= 64 ? GROUP : 128; }

// ---------------------------------------------------------------------------------- window of the Gaussian
struct Window {
  int j_lo, j_hi, i_lo, i_hi;
  __device__ __forceinline__ bool empty() const { return j_lo > j_hi || i_lo > i_hi; }
};

// Indices on an n-pixel axis where g(c_j)/g_max >= theta, g(c) = exp(k (c - t)^2), padded by one pixel.
__device__ __forceinline__ void axis_window(float t, int n, float half_n, float two_over_n, float bias_n, float r2,
                                            int& lo, int& hi) {
  const float nm1 = static_cast<float>(n - 1);
  const float js = fminf(fmaxf(rintf(fmaf(t + 1.0f, half_n, -0.5f)), 0.f), nm1);  // nearest in-image centre
  const float ds = fmaf(js, two_over_n, bias_n) - t;
  const float R = sqrtf(fmaf(ds, ds, r2));
  const float flo = fmaxf(ceilf(fmaf(t - R + 1.0f, half_n, -0.5f)) - 1.0f, 0.f);
  const float fhi = fminf(floorf(fmaf(t + R + 1.0f, half_n, -0.5f)) + 1.0f, nm1);
  lo = static_cast<int>(flo);
  hi = static_cast<int>(fhi);
}

__device__ __forceinline__ Window make_window(const Geom& g, int H, int W, float tx, float ty) {
  Window w;
  axis_window(tx, W, g.half_w, g.two_over_w, g.bias_w, g.r2_win, w.j_lo, w.j_hi);
  axis_window(ty, H, g.half_h, g.two_over_h, g.bias_h, g.r2_win, w.i_lo, w.i_hi);
  return w;
}

// sum over the window of the unnormalised axis factors (every warp computes it redundantly)
__device__ __forceinline__ float axis_window_sum(int lo, int hi, float t, float two_over_n, float bias_n, float k2,
                                                 int lane) {
  float s = 0.f;
  for (int j = lo + lane; j <= hi; j += 32) {
    const float d = fmaf(static_cast<float>(j), two_over_n, bias_n) - t;
    s += ex2(k2 * d * d);
  }
  return s;
}

template <int N>
__device__ __forceinline__ float max_of(const float (&v)[N]) {
  float m = v[0];
#pragma unroll
  for (int i = 1; i + 1 < N; i += 2) m = fmaxf(m, fmaxf(v[i], v[i + 1]));  // FMNMX3
  if (N % 2 == 0) m = fmaxf(m, v[N - 1]);
  return m;
}

// ================================================================================================ forward
struct HeadFwdStreamParams {
  HeadFwdParams base;
  Geom g;
};

template <typename T, int VEC, int GROUP, int REG, bool FIXC>
__global__ void __launch_bounds__(stream_block_threads<GROUP>()) head_fwd_stream_kernel(const HeadFwdStreamParams ps) {
  constexpr int BLOCK = stream_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  constexpr int NW = GROUP / 32;
  constexpr int U = VEC == 8 ? 4 : 8;  // vectors in flight per thread: 128 B (fp32) / 64 B (bf16) x U
  constexpr bool kKL = REG == DSNT_REG_KL;
  constexpr bool kJS = REG == DSNT_REG_JS;
  constexpr bool kVar = REG == DSNT_REG_VAR;
  constexpr bool kMSE = REG == DSNT_REG_MSE;
  static_assert(!kVar || FIXC, "the single-pass variance needs thread-fixed columns");
  __shared__ float red_m[GPB * NW];
  __shared__ float red_a[GPB * NW * 4];
  __shared__ float red_b[GPB * NW * 4];

  const HeadFwdParams& p = ps.base;
  const Geom& g = ps.g;
  const int tid = threadIdx.x;
  const int gid = tid / GROUP, lane_g = tid % GROUP, warp_g = lane_g >> 5, lane = tid & 31;
  const long hm = static_cast<long>(blockIdx.x) * GPB + gid;
  if (hm >= p.n) return;  // GROUP == 32 only; the grid is exact otherwise

  const int H = p.H, W = p.W;
  const int nvec = g.nvec;
  const HmRef ref = locate(p.st, hm, static_cast<long>(H) * W * sizeof(T));
  const T* zb = reinterpret_cast<const T*>(static_cast<const char*>(p.z) + ref.z_bytes);

  // ---- per-thread geometry
  float xs[VEC];
  float ybase = 0.f;
  VecWalker wk(lane_g, GROUP, g.wv);
  if constexpr (FIXC) {
#pragma unroll
    for (int c = 0; c < VEC; ++c) xs[c] = axis_coord(wk.cv * VEC + c, g.two_over_w, g.bias_w);
    ybase = axis_coord(wk.row, g.two_over_h, g.bias_h);
  }

  // ---- streaming pass: online softmax statistics.
  // FIXC: a thread always sees the same VEC columns, so it keeps one accumulator per column (E[c] = sum over its
  // rows of e) and S, S_x follow exactly at the end.  The variance regulariser additionally runs a weighted
  // Welford recurrence over the thread's rows (my, M2y), which is cancellation-free for any sigma.
  float mt2 = -INFINITY, S = 0.f, Sx = 0.f, Sy = 0.f, Tt = 0.f, Q = 0.f;
  float E[VEC];
  float my = 0.f, M2y = 0.f;
#pragma unroll
  for (int c = 0; c < VEC; ++c) E[c] = 0.f;
  for (int f0 = lane_g; f0 < nvec; f0 += GROUP * U) {
    float v[U][VEC];
    const bool full = f0 + (U - 1) * GROUP < nvec;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (full || f0 + u * GROUP < nvec) {
        VecIO<T, VEC>::load(zb, static_cast<long>(f0 + u * GROUP) * VEC, v[u]);
      } else {
#pragma unroll
        for (int c = 0; c < VEC; ++c) v[u][c] = -INFINITY;
      }
    }
    float bm = max_of(v[0]);
#pragma unroll
    for (int u = 1; u < U; ++u) bm = fmaxf(bm, max_of(v[u]));
    const float bm2 = bm * kLog2e;
    if (bm2 > mt2) {  // rare after the first batches: rescale the running sums to the new maximum
      const float sc = ex2(mt2 - bm2);
      if (kKL) Tt = S > 0.f ? sc * fmaf(mt2 - bm2, S, Tt) : 0.f;  // sum e'(t - d) = sc (T - d S)
      S *= sc; Sy *= sc;
      if constexpr (FIXC) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) E[c] *= sc;
      } else {
        Sx *= sc;
      }
      if (kVar) M2y *= sc;
      if (kMSE) Q *= sc * sc;
      mt2 = bm2;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (full || f0 + u * GROUP < nvec) {
        float y;
        if constexpr (FIXC) {
          y = fmaf(static_cast<float>(u), g.dy_step, ybase);
        } else {
#pragma unroll
          for (int c = 0; c < VEC; ++c) xs[c] = axis_coord(wk.cv * VEC + c, g.two_over_w, g.bias_w);
          y = axis_coord(wk.row, g.two_over_h, g.bias_h);
        }
        float ev[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const float t = fmaf(v[u][c], kLog2e, -mt2);
          ev[c] = ex2(t);
          if (kKL) Tt = fmaf(ev[c], t, Tt);
          if (kMSE) Q = fmaf(ev[c], ev[c], Q);
          if constexpr (FIXC) E[c] += ev[c]; else Sx = fmaf(ev[c], xs[c], Sx);
        }
        float rs;
        if constexpr (VEC == 8) rs = ((ev[0] + ev[1]) + (ev[2] + ev[3])) + ((ev[4] + ev[5]) + (ev[6] + ev[7]));
        else if constexpr (VEC == 4) rs = (ev[0] + ev[1]) + (ev[2] + ev[3]);
        else rs = ev[0];
        if constexpr (kVar) {
          // weighted Welford: S += rs; my += (y - my) rs / S; M2y += rs (y - my_old)(y - my_new)
          S += rs;
          const float dl = y - my;
          const float wgt = rs * rcp(fmaxf(S, 1e-30f));
          my = fmaf(dl, wgt, my);
          M2y = fmaf(rs * dl, y - my, M2y);
        } else {
          if (!FIXC || kKL) S += rs;   // FIXC: S is the sum of the column accumulators (KL needs it on the fly)
          Sy = fmaf(rs, y, Sy);
        }
      }
      if constexpr (!FIXC) wk.next();
    }
    if constexpr (FIXC) ybase = fmaf(static_cast<float>(U), g.dy_step, ybase);
  }

  // ---- thread-local wrap-up
  float mx = 0.f, M2x = 0.f;
  if constexpr (FIXC) {
    float s = 0.f, sx = 0.f;
#pragma unroll
    for (int c = 0; c < VEC; ++c) { s += E[c]; sx = fmaf(E[c], xs[c], sx); }
    if (!kVar && !kKL) S = s;   // (Welford / KL already own S; it is the same sum)
    Sx = sx;
    if constexpr (kVar) {
      mx = s > 0.f ? sx / s : 0.f;
#pragma unroll
      for (int c = 0; c < VEC; ++c) { const float d = xs[c] - mx; M2x = fmaf(E[c] * d, d, M2x); }
      Sy = S * my;
    }
  }

  // ---- merge the per-thread statistics
  const float m2 = group_max<GROUP>(mt2, red_m + gid * NW, warp_g, lane);
  {
    const float sc = mt2 > -INFINITY ? ex2(mt2 - m2) : 0.f;
    if (kKL) Tt = S > 0.f ? sc * fmaf(mt2 - m2, S, Tt) : 0.f;
    S *= sc; Sx *= sc; Sy *= sc;
    if (kVar) { M2x *= sc; M2y *= sc; }
    if (kMSE) Tt = Q * sc * sc;     // MSE rides in the 4th slot of the reduction
  }
  const float S_loc = S;
  group_sum4<GROUP>(S, Sx, Sy, Tt, red_a + gid * NW * 4, warp_g, lane);
  const float invS = 1.0f / S;
  const float mux = Sx * invS, muy = Sy * invS;

  float tx = 0.f, ty = 0.f;
  if (p.target) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p.target) + ref.nl);
    tx = t.x; ty = t.y;
  }

  float D = 0.f, creg = 0.f, ginv = 0.f, vx = 0.f, vy = 0.f;

  // ---- variance regulariser: second moments about the GLOBAL mean (Chan's update, all terms positive)
  if constexpr (kVar) {
    const float ddx = mx - mux, ddy = my - muy;
    float ax = fmaf(S_loc * ddx, ddx, M2x), ay = fmaf(S_loc * ddy, ddy, M2y);
    group_sum2<GROUP>(ax, ay, red_b + gid * NW * 4, warp_g, lane);
    vx = ax * invS; vy = ay * invS;
    const float s2 = p.sigma * p.sigma, ex = vx - s2, ey = vy - s2;
    D = ex * ex + ey * ey;
    creg = 2.f * (ex * vx + ey * vy);
  }

  // ---- divergence on the window of the Gaussian only
  if constexpr (kKL || kJS || kMSE) {
    const Window win = make_window(g, H, W, tx, ty);
    float qa = 0.f, qb = 0.f, qc = 0.f, qd = 0.f;
    if (!win.empty()) {
      float sx = axis_window_sum(win.j_lo, win.j_hi, tx, g.two_over_w, g.bias_w, g.k2, lane);
      float sy = axis_window_sum(win.i_lo, win.i_hi, ty, g.two_over_h, g.bias_h, g.k2, lane);
      {
        const float k = warp_sum2_transposed(sx, sy, lane);
        sx = __shfl_sync(kFull, k, 0);
        sy = __shfl_sync(kFull, k, 16);
      }
      ginv = 1.0f / (sx * sy + kEps);
      const float l2ginv = log2f(ginv);
      const float l2is = -log2f(S);         // log2 P = t + l2is
      const float tlm1 = l2is - 1.0f;
      const float hinvS = 0.5f * invS;
      constexpr int RU = 4;
      const int nwr = win.i_hi - win.i_lo + 1;
      for (int j0 = win.j_lo; j0 <= win.j_hi; j0 += 32) {
        const int j = j0 + lane;
        const bool cact = j <= win.j_hi;
        const float dx = fmaf(static_cast<float>(j), g.two_over_w, g.bias_w) - tx;
        const float ax = g.k2 * dx * dx;
        for (int r = warp_g * RU; r < nwr; r += NW * RU) {
          float zv[RU];
#pragma unroll
          for (int rr = 0; rr < RU; ++rr) {
            zv[rr] = 0.f;
            if (cact && r + rr < nwr) {
              float one[1];
              VecIO<T, 1>::load(zb, static_cast<long>(win.i_lo + r + rr) * W + j, one);
              zv[rr] = one[0];
            }
          }
#pragma unroll
          for (int rr = 0; rr < RU; ++rr) {
            if (cact && r + rr < nwr) {
              const float dy = fmaf(static_cast<float>(win.i_lo + r + rr), g.two_over_h, g.bias_h) - ty;
              const float lgG = fmaf(g.k2 * dy, dy, ax) + l2ginv;  // log2 G, closed form
              const float G = ex2(lgG);
              const float t = fmaf(zv[rr], kLog2e, -m2);
              const float e = ex2(t);
              if (kJS) {
                const float Mp = fmaf(e, hinvS, fmaf(0.5f, G, kEps));  // M + eps
                const float L = lg2(Mp);
                qa = fmaf(e * invS, (t + tlm1) - L, qa);                // P (log2 P - log2 M' - 1)
                qb = fmaf(G, lgG - L, qb);                             // G (log2 G - log2 M')
              } else if (kKL) {
                qa = fmaf(e * invS, lg2(G + kEps), qa);                // see head_fast.cuh: D stays free of the -79.7 offset
                qb += e * invS;
              } else {
                const float P = e * invS, df = P - G;
                qa = fmaf(df, df, qa);                                 // (P - G)^2
                qb = fmaf(P, P, qb);                                   // P^2 (to take the window out of sum P^2)
                qc = fmaf(P, df, qc);                                  // P (P - G)
              }
            }
          }
        }
      }
    }
    if constexpr (kMSE) {
      group_sum4<GROUP>(qa, qb, qc, qd, red_b + gid * NW * 4, warp_g, lane);
      const float outside = fmaxf(fmaf(Tt * invS, invS, -qb), 0.f);   // sum of P^2 where G is negligible
      D = outside + qa;
      creg = 2.f * (outside + qc);
    } else {
      group_sum2<GROUP>(qa, qb, red_b + gid * NW * 4, warp_g, lane);
      if (kJS) {
        creg = 0.5f * kLn2 * (1.0f + qa);        // 1/2 sum P (ln P - ln M')
        D = fmaf(0.5f * kLn2, qb, creg);         // + 1/2 sum G (ln G - ln M')
      } else {
        const float plnp = fmaf(kLn2 * invS, Tt, -logf(S));  // sum P ln P
        D = plnp - kLn2 * fmaf(kLog2Eps, 1.0f - qb, qa);   // outside the window G + eps = eps exactly
        creg = D + 1.0f;
      }
    }
  }

  if (lane_g == 0) write_outputs(p, hm, m2, invS, mux, muy, vx, vy, creg, ginv, tx, ty, D);
}

// ================================================================================================ backward
struct HeadBwdStreamParams {
  HeadBwdParams base;
  Geom g;
  PreactCfg pc;   // PA != DSNT_PREACT_SOFTMAX instantiations only (dsnt_head_preact_bwd)
};

// PA != SOFTMAX (act_fast / act_log2 / act_grad above): dz = f'(z)/(S+eps) (g - c) for the reference's other
// pre-activations; where f' = 0 (masked, clipped or z = 0) the gradient is exactly 0 whatever the log terms say.
template <typename T, int VEC, int GROUP, int REG, bool FIXC, int PA = DSNT_PREACT_SOFTMAX>
__global__ void __launch_bounds__(stream_block_threads<GROUP>()) head_bwd_stream_kernel(const HeadBwdStreamParams ps) {
  constexpr bool kPlain = PA == DSNT_PREACT_SOFTMAX;
  constexpr int BLOCK = stream_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  constexpr int U = VEC == 8 ? 4 : 8;
  constexpr bool kGauss = reg_needs_gauss(REG);

  const HeadBwdParams& p = ps.base;
  const Geom& g = ps.g;
  const int tid = threadIdx.x;
  const int gid = tid / GROUP, lane_g = tid % GROUP;
  const long hm = static_cast<long>(blockIdx.x) * GPB + gid;
  if (hm >= p.n) return;

  const int H = p.H, W = p.W;
  const int nvec = g.nvec;
  const HmRef ref = locate(p.st, hm, static_cast<long>(H) * W * sizeof(T));
  const T* zb = reinterpret_cast<const T*>(static_cast<const char*>(p.z) + ref.z_bytes);
  T* dzb = reinterpret_cast<T*>(static_cast<char*>(p.dz) + ref.dz_bytes);

  const BwdScalars s = load_bwd_scalars<true>(p, hm, ref.nl, REG);
  Window win{1, 0, 1, 0};
  if constexpr (kGauss) win = make_window(g, H, W, s.tx, s.ty);
  // var with sum P != 1 (eps in the normaliser; stats[7] = 1 - sum P): d v/dP gains -2 x mu (1 - sum P)
  const float cx2 = (!kPlain && REG == DSNT_REG_VAR) ? -2.f * s.mux * s.ginv : 0.f;
  const float cy2 = (!kPlain && REG == DSNT_REG_VAR) ? -2.f * s.muy * s.ginv : 0.f;
  const float thr = ps.pc.threshold;

  // constant part of (g - c):  -c, plus the out-of-window value of rho*r
  float cbase = -s.c;
  if (REG == DSNT_REG_JS) cbase = fmaf(0.5f * kLn2, s.rho, cbase);                          // r -> 1/2 ln 2
  if (REG == DSNT_REG_KL) cbase = fmaf(s.rho, 1.0f - kLnEps + kLn2 * s.l2is, cbase);         // r = ln2 t + this
  const float rho_t = REG == DSNT_REG_KL ? s.rho * kLn2 : 0.f;
  const float rho_p = REG == DSNT_REG_MSE ? 2.f * s.rho : 0.f;

  // ---- per-column context: x, a x (+ var term), unnormalised Gaussian factor (0 outside the window)
  float acol[VEC], gxs[VEC];
  bool anycol = false;
  auto init_cols = [&](int col0) {
    anycol = false;
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      const float x = axis_coord(col0 + c, g.two_over_w, g.bias_w);
      float a = s.a * x;
      if (REG == DSNT_REG_VAR) {
        const float dx = x - s.mux;
        a = fmaf(s.kx * dx, dx, a);
        if constexpr (!kPlain) a = fmaf(s.kx * x, cx2, a);
      }
      acol[c] = a;
      gxs[c] = 0.f;
      if (kGauss) {
        const bool in = col0 + c >= win.j_lo && col0 + c <= win.j_hi;
        const float dx = x - s.tx;
        gxs[c] = in ? ex2(g.k2 * dx * dx) : 0.f;
        anycol |= in;
      }
    }
  };
  VecWalker wk(lane_g, GROUP, g.wv);
  float ybase = 0.f;
  int rowbase = 0;
  if constexpr (FIXC) {
    init_cols(wk.cv * VEC);
    ybase = axis_coord(wk.row, g.two_over_h, g.bias_h);
    rowbase = wk.row;
  }

  for (int f0 = lane_g; f0 < nvec; f0 += GROUP * U) {
    float v[U][VEC];
    const bool full = f0 + (U - 1) * GROUP < nvec;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (full || f0 + u * GROUP < nvec) VecIO<T, VEC>::load(zb, static_cast<long>(f0 + u * GROUP) * VEC, v[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (full || f0 + u * GROUP < nvec) {
        float y;
        int row;
        if constexpr (FIXC) {
          y = fmaf(static_cast<float>(u), g.dy_step, ybase);
          row = rowbase + u * g.rstep;
        } else {
          init_cols(wk.cv * VEC);
          y = axis_coord(wk.row, g.two_over_h, g.bias_h);
          row = wk.row;
        }
        float rowc = fmaf(s.b, y, cbase);
        if (REG == DSNT_REG_VAR) {
          const float dy = y - s.muy;
          rowc = fmaf(s.ky * dy, dy, rowc);
          if constexpr (!kPlain) rowc = fmaf(s.ky * y, cy2, rowc);
        }
        float gyn = 0.f;
        bool heavy = false;
        if (kGauss) {
          heavy = anycol && row >= win.i_lo && row <= win.i_hi;
          if (heavy) {
            const float dy = y - s.ty;
            gyn = ex2(g.k2 * dy * dy) * s.ginv;
          }
        }
        float out[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const float t0 = fmaf(v[u][c], kLog2e, -s.m2);
          const float e = act_fast<PA>(v[u][c], t0, thr);
          const float t = REG == DSNT_REG_KL ? act_log2<PA>(e, t0) : t0;   // log2 f
          const float P = e * s.invS;
          float gmc = acol[c] + rowc;
          if (REG == DSNT_REG_KL) gmc = fmaf(rho_t, t, gmc);
          if (REG == DSNT_REG_MSE) gmc = fmaf(rho_p, P, gmc);
          if (kGauss && heavy) {
            const float G = gxs[c] * gyn;
            if (REG == DSNT_REG_JS) {
              // rho r = rho/2 ln2 (1 - lg2(1+q)),  q = (G + 2 eps)/P;  the "1" is already in cbase
              const float q = (G + 2.f * kEps) * rcp(fmaxf(P, 1e-37f));
              gmc = fmaf(-0.5f * kLn2 * s.rho, lg2(1.0f + q), gmc);
            } else if (REG == DSNT_REG_KL) {
              gmc = fmaf(-kLn2 * s.rho, lg2(G + kEps) - kLog2Eps, gmc);
            } else {
              gmc = fmaf(-rho_p, G, gmc);
            }
          }
          if constexpr (kPlain) {
            out[c] = P * gmc;
          } else {
            const float fp = act_grad<PA>(v[u][c], e);
            out[c] = fp != 0.f ? fp * s.invS * gmc : 0.f;
          }
        }
        VecIO<T, VEC>::store(dzb, static_cast<long>(f0 + u * GROUP) * VEC, out);
      }
      if constexpr (!FIXC) wk.next();
    }
    if constexpr (FIXC) {
      ybase = fmaf(static_cast<float>(U), g.dy_step, ybase);
      rowbase += U * g.rstep;
    }
  }
}

}  // namespace dsnt
