// head_fast.cuh -- the tuned bandwidth-path kernels (v2) for the common layouts: logits input, 16-byte
// vectors, power-of-two vectors per row dividing the group size (thread-fixed columns) and a vector count that
// is a multiple of GROUP*U (no tail).  64x64, 128x128, 256x256 (fp32 and bf16) and 32x32 fp32 all qualify;
// everything else takes the generic kernels of head_stream.cuh / head_fwd.cuh / head_bwd.cuh.
//
// What v2 changes against head_stream.cuh (profiles/r01_v1_*: forward 70 % issue-active, 19 instructions per
// pixel, window re-read missing L2 -> 13 % extra DRAM traffic):
//   * no bounds predicates, no integer division, raw 128-bit loads unpacked at the point of use (bf16 keeps
//     eight vectors = 128 B in flight per thread like fp32);
//   * the pixels of the Gaussian window are STASHED in shared memory while they stream past (the window depends
//     only on the target, so it is known up front) and the divergence is evaluated from the stash with every
//     lane busy (flattened window index) -- no second look at global memory;
//   * backward: vectors that touch the window are stashed instead of taking a divergent slow path inside the
//     streaming loop; a dense epilogue computes and stores them.
#pragma once

#include "head_stream.cuh"

namespace dsnt {

constexpr int kFastU = 8;            // 128-bit vectors in flight per thread
constexpr int kStashFloats = 896;    // window pixels kept on chip per heatmap (e.g. 28 rows x 32 columns)
constexpr int kTabN = 32;            // max window rows / columns for the stash path

// ---------------------------------------------------------------------------------- raw 16-byte access
__device__ __forceinline__ uint4 ld_raw16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

template <typename T, int VEC>
__device__ __forceinline__ void unpack16(const uint4& r, float (&v)[VEC]) {
  if constexpr (sizeof(T) == 4) {
    static_assert(VEC == 4, "fp32 vectors are 4 wide");
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  } else {
    static_assert(VEC == 8, "bf16 vectors are 8 wide");
    v[0] = bf16lo(r.x); v[1] = bf16hi(r.x); v[2] = bf16lo(r.y); v[3] = bf16hi(r.y);
    v[4] = bf16lo(r.z); v[5] = bf16hi(r.z); v[6] = bf16lo(r.w); v[7] = bf16hi(r.w);
  }
}

template <typename T, int VEC>
__device__ __forceinline__ uint4 pack16(const float (&o)[VEC]) {
  uint4 r;
  if constexpr (sizeof(T) == 4) {
    r.x = __float_as_uint(o[0]); r.y = __float_as_uint(o[1]); r.z = __float_as_uint(o[2]); r.w = __float_as_uint(o[3]);
  } else {
    r.x = pack_bf16(o[0], o[1]); r.y = pack_bf16(o[2], o[3]); r.z = pack_bf16(o[4], o[5]); r.w = pack_bf16(o[6], o[7]);
  }
  return r;
}

// max over the elements of U raw vectors (bf16: packed HMNMX2 on pairs, widened once at the end)
template <typename T, int U>
__device__ __forceinline__ float raw_max(const uint4 (&raw)[U]) {
  if constexpr (sizeof(T) == 4) {
    float m = fmaxf(fmaxf(__uint_as_float(raw[0].x), __uint_as_float(raw[0].y)),
                    fmaxf(__uint_as_float(raw[0].z), __uint_as_float(raw[0].w)));
#pragma unroll
    for (int u = 1; u < U; ++u) {
      m = fmaxf(m, fmaxf(__uint_as_float(raw[u].x), __uint_as_float(raw[u].y)));
      m = fmaxf(m, fmaxf(__uint_as_float(raw[u].z), __uint_as_float(raw[u].w)));
    }
    return m;
  } else {
    auto as2 = [](uint32_t w) { return *reinterpret_cast<const __nv_bfloat162*>(&w); };
    __nv_bfloat162 m = __hmax2(__hmax2(as2(raw[0].x), as2(raw[0].y)), __hmax2(as2(raw[0].z), as2(raw[0].w)));
#pragma unroll
    for (int u = 1; u < U; ++u) {
      m = __hmax2(m, __hmax2(as2(raw[u].x), as2(raw[u].y)));
      m = __hmax2(m, __hmax2(as2(raw[u].z), as2(raw[u].w)));
    }
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&m);
    return fmaxf(bf16lo(w), bf16hi(w));
  }
}

// ptxas sinks independent loads below the stores of earlier vectors to save registers, which leaves one or two
// 128-bit loads in flight per thread (seen in SASS: 32 registers, LDG/compute/STG interleaved, 4 % slower than
// the generic kernel).  A real use of every loaded vector ahead of the first store keeps all U loads in flight:
// xor one word of each and trap on an impossible value (three LOP3 + one compare per batch).
template <int U>
__device__ __forceinline__ void require_all_loaded(const uint4 (&raw)[U]) {
  uint32_t acc = raw[0].x;
#pragma unroll
  for (int u = 1; u < U; ++u) acc ^= raw[u].x;
  uint32_t other = raw[0].w;
#pragma unroll
  for (int u = 1; u < U; ++u) other += raw[u].w;
  if (acc == 0x7fc5a5a5u && other == 0x7fc3c3c3u) __trap();
}

// Launch-uniform extras of the fast path (host-computed).
struct FastGeom {
  int wv_shift;       // log2(vectors per row)
  int nbatch;         // nvec / (GROUP * U)
  int rstep;          // rows advanced per vector slot u: GROUP / wv
  float dy_step;      // rstep * 2/H
  float dy_batch;     // U * dy_step
};

struct HeadFwdFastParams {
  HeadFwdParams base;
  Geom g;
  FastGeom f;
  FlipCfg fl;   // FLIP instantiations only (dsnt_flip_tta_fwd)
  PreactCfg pc; // PA != DSNT_PREACT_SOFTMAX instantiations only (dsnt_head_preact_fwd)
};

constexpr int kFlipU = 4;   // FLIP loads two vectors per slot: 4 slots = the same 128 B (fp32) in flight per thread

// Per-heatmap window bookkeeping for the stash.
struct StashWin {
  int i_lo, nrw;        // first window row, number of rows
  int cv_lo, nvc;       // first window vector column, vectors per row
  int coff, wcols;      // first window column relative to the stash row start; exact window width
  bool ok;              // window non-empty and small enough for the stash
};

template <int VEC>
__device__ __forceinline__ StashWin make_stash_window(const Window& w) {
  StashWin s;
  s.i_lo = w.i_lo; s.nrw = w.i_hi - w.i_lo + 1;
  s.cv_lo = w.j_lo / VEC;
  s.nvc = w.j_hi / VEC - s.cv_lo + 1;
  s.coff = w.j_lo - s.cv_lo * VEC;
  s.wcols = w.j_hi - w.j_lo + 1;
  s.ok = !w.empty() && s.nrw <= kTabN && s.nvc * VEC <= kTabN && s.nrw * s.nvc * VEC <= kStashFloats;
  return s;
}

// ================================================================================================ forward
// FLIP (inference, src/dsnt/inference.py:36-46): heatmap hm = (b, c) is averaged on the fly with the mirrored heatmap
// (batch + b, perm[c]) of the same tensor; REG must be NONE then (no target at inference).
// PA != SOFTMAX: P = f(z)/(sum f + eps) for the reference's other pre-activations.  P may be exactly 0 and sum P may
// differ from 1 (eps), which the closed forms below carry as sumP / om; stats[7] holds om = 1 - sum P for `var`.
// The body takes the index of the "block" of heatmaps it works on (a persistent caller could loop over it);
// head_fwd_fast_kernel passes blockIdx.x.
template <typename T, int VEC, int GROUP, int REG, bool FLIP = false, int PA = DSNT_PREACT_SOFTMAX>
__device__ __forceinline__ void head_fwd_fast_body(const HeadFwdFastParams& ps, long block_index) {
  static_assert(sizeof(T) * VEC == 16, "fast path = 16-byte vectors");
  static_assert(!FLIP || REG == DSNT_REG_NONE, "flip test-time augmentation is forward-only, no regulariser");
  constexpr bool kSM = preact_is_softmax(PA);       // needs the running maximum
  constexpr bool kPlain = PA == DSNT_PREACT_SOFTMAX;
  const float thr = ps.pc.threshold;
  constexpr int BLOCK = stream_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  constexpr int NW = GROUP / 32;
  constexpr int U = FLIP ? kFlipU : kFastU;
  constexpr bool kKL = REG == DSNT_REG_KL;
  constexpr bool kJS = REG == DSNT_REG_JS;
  constexpr bool kVar = REG == DSNT_REG_VAR;
  constexpr bool kMSE = REG == DSNT_REG_MSE;
  constexpr bool kWin = kKL || kJS || kMSE;
  __shared__ float red_m[GPB * NW];
  __shared__ float red_a[GPB * NW * 4];
  __shared__ float red_b[GPB * NW * 4];
  __shared__ __align__(16) float stash_all[kWin ? GPB * (kStashFloats + 2 * kTabN) : 4];

  const HeadFwdParams& p = ps.base;
  const Geom& g = ps.g;
  const FastGeom& fg = ps.f;
  const int tid = threadIdx.x;
  const int gid = tid / GROUP, lane_g = tid % GROUP, warp_g = lane_g >> 5, lane = tid & 31;
  const long hm = block_index * GPB + gid;
  if (hm >= p.n) return;  // GROUP == 32 only; the grid is exact otherwise

  const int H = p.H, W = p.W;
  const HmRef ref = locate(p.st, hm, static_cast<long>(H) * W * sizeof(T));
  const char* zb = static_cast<const char*>(p.z) + ref.z_bytes;

  float tx = 0.f, ty = 0.f;
  if (p.target) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p.target) + ref.nl);
    tx = t.x; ty = t.y;
  }

  // ---- per-thread geometry: fixed columns, rows row0 + k*rstep
  const int cv = lane_g & ((1 << fg.wv_shift) - 1);
  const int row0 = lane_g >> fg.wv_shift;
  float xs[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) xs[c] = axis_coord(cv * VEC + c, g.two_over_w, g.bias_w);
  float ybase = axis_coord(row0, g.two_over_h, g.bias_h);

  // ---- window of the Gaussian (depends on the target only) and this thread's slot in the stash
  float* stash = stash_all + gid * (kStashFloats + 2 * kTabN);
  float* tabx = stash + kStashFloats;
  float* taby = tabx + kTabN;
  Window win{1, 0, 1, 0};
  StashWin sw{0, 0, 0, 0, 0, 0, false};
  int srow = 0;             // (row - i_lo) of the thread's current row, advanced with the loop
  int scol = -1;            // float offset of the thread's vector inside a stash row, -1 = not in the window
  int sstride = 0;
  if constexpr (kWin) {
    win = make_window(g, H, W, tx, ty);
    sw = make_stash_window<VEC>(win);
    if (sw.ok) {
      const int dc = cv - sw.cv_lo;
      scol = (dc >= 0 && dc < sw.nvc) ? dc * VEC : -1;
      sstride = sw.nvc * VEC;
      srow = row0 - sw.i_lo;
    }
  }

  // ---- streaming pass: online softmax statistics, one accumulator per column
  float mt2 = kSM ? -INFINITY : 0.f, S = 0.f, Sy = 0.f, Tt = 0.f, Q = 0.f;
  float E[VEC];
  float my = 0.f, M2y = 0.f;
#pragma unroll
  for (int c = 0; c < VEC; ++c) E[c] = 0.f;
  const char* src = zb + static_cast<size_t>(lane_g) * 16;
  const char* srcf = nullptr;   // FLIP: the mirrored vector of the partner heatmap, same row
  char* avg = nullptr;          // FLIP: optional averaged raw heatmaps
  if constexpr (FLIP) {
    const int C = ps.fl.C;
    const long bi = hm / C;
    const int ci = static_cast<int>(hm - bi * C);
    const long partner = p.n + bi * C + (ps.fl.perm ? __ldg(ps.fl.perm + ci) : ci);
    const int wv = 1 << fg.wv_shift;
    srcf = static_cast<const char*>(p.z) + partner * (static_cast<long>(H) * W * sizeof(T)) +
           static_cast<size_t>((row0 << fg.wv_shift) + (wv - 1 - cv)) * 16;
    if (ps.fl.avg_out)
      avg = static_cast<char*>(ps.fl.avg_out) + hm * (static_cast<long>(H) * W * sizeof(T)) + static_cast<size_t>(lane_g) * 16;
  }
  for (int b = 0; b < fg.nbatch; ++b) {
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) raw[u] = ld_raw16(src + static_cast<size_t>(u) * (GROUP * 16));
    src += static_cast<size_t>(U) * (GROUP * 16);
    float fv[FLIP ? U : 1][VEC];
    float bm2 = 0.f;
    if constexpr (FLIP) {
      uint4 rawf[U];
#pragma unroll
      for (int u = 0; u < U; ++u) rawf[u] = ld_raw16(srcf + static_cast<size_t>(u) * (GROUP * 16));
      srcf += static_cast<size_t>(U) * (GROUP * 16);
      float bm = -INFINITY;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float a[VEC], m[VEC];
        unpack16<T, VEC>(raw[u], a);
        unpack16<T, VEC>(rawf[u], m);
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          fv[u][c] = 0.5f * (a[c] + m[VEC - 1 - c]);      // (hm1 + reversed hm2) / 2, inference.py:44-46
          bm = fmaxf(bm, fv[u][c]);
        }
        if (avg) *reinterpret_cast<uint4*>(avg + static_cast<size_t>(u) * (GROUP * 16)) = pack16<T, VEC>(fv[u]);
      }
      if (avg) avg += static_cast<size_t>(U) * (GROUP * 16);
      bm2 = bm * kLog2e;
    } else if constexpr (kSM) {
      bm2 = raw_max<T, U>(raw) * kLog2e;
    }
    if (kSM && bm2 > mt2) {  // rare after the first batches: rescale the running sums to the new maximum
      const float sc = ex2(mt2 - bm2);
      if (kKL) Tt = S > 0.f ? sc * fmaf(mt2 - bm2, S, Tt) : 0.f;  // sum e'(t - d) = sc (T - d S)
      S *= sc; Sy *= sc;
#pragma unroll
      for (int c = 0; c < VEC; ++c) E[c] *= sc;
      if (kVar) M2y *= sc;
      if (kMSE) Q *= sc * sc;
      mt2 = bm2;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float v[VEC];
      if constexpr (FLIP) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) v[c] = fv[u][c];
      } else {
        unpack16<T, VEC>(raw[u], v);
      }
      if constexpr (kWin) {
        // stash the raw logits of window vectors (unsigned compare = 0 <= r < nrw)
        const int r = srow + u * fg.rstep;
        if (scol >= 0 && static_cast<unsigned>(r) < static_cast<unsigned>(sw.nrw)) {
          float4* dst = reinterpret_cast<float4*>(stash + r * sstride + scol);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          if constexpr (VEC == 8) dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      const float y = fmaf(static_cast<float>(u), fg.dy_step, ybase);
      float ev[VEC];
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float t = fmaf(v[c], kLog2e, -mt2);
        ev[c] = act_fast<PA>(v[c], t, thr);
        if (kKL) Tt = fmaf(ev[c], act_log2<PA>(ev[c], t), Tt);
        if (kMSE) Q = fmaf(ev[c], ev[c], Q);
        E[c] += ev[c];
      }
      float rs;
      if constexpr (VEC == 8) rs = ((ev[0] + ev[1]) + (ev[2] + ev[3])) + ((ev[4] + ev[5]) + (ev[6] + ev[7]));
      else rs = (ev[0] + ev[1]) + (ev[2] + ev[3]);
      if constexpr (kVar) {
        // weighted Welford over the thread's rows: cancellation-free second moment for any sigma
        S += rs;
        const float dl = y - my;
        const float wgt = rs * rcp(fmaxf(S, 1e-30f));
        my = fmaf(dl, wgt, my);
        M2y = fmaf(rs * dl, y - my, M2y);
      } else {
        if (kKL) S += rs;
        Sy = fmaf(rs, y, Sy);
      }
    }
    ybase += fg.dy_batch;
    if constexpr (kWin) srow += U * fg.rstep;
  }

  // ---- thread-local wrap-up: S and S_x from the column accumulators
  float Sx = 0.f, mx = 0.f, M2x = 0.f;
  {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < VEC; ++c) { s += E[c]; Sx = fmaf(E[c], xs[c], Sx); }
    if (!kVar && !kKL) S = s;
    if constexpr (kVar) {
      mx = s > 0.f ? Sx / s : 0.f;
#pragma unroll
      for (int c = 0; c < VEC; ++c) { const float d = xs[c] - mx; M2x = fmaf(E[c] * d, d, M2x); }
      Sy = S * my;
    }
  }

  // ---- merge the per-thread statistics
  const float m2 = group_max<GROUP>(mt2, red_m + gid * NW, warp_g, lane);
  {
    const float sc = ex2(mt2 - m2);
    if (kKL) Tt = S > 0.f ? sc * fmaf(mt2 - m2, S, Tt) : 0.f;
    S *= sc; Sx *= sc; Sy *= sc;
    if (kVar) { M2x *= sc; M2y *= sc; }
    if (kMSE) Tt = Q * sc * sc;     // sum e^2 rides in the 4th slot of the reduction
  }
  const float S_loc = S;
  group_sum4<GROUP>(S, Sx, Sy, Tt, red_a + gid * NW * 4, warp_g, lane);
  if constexpr (GROUP == 32) __syncwarp();   // stash writes of the other lanes are visible from here on
  const float invS = kPlain ? 1.0f / S : 1.0f / (S + ps.pc.eps);
  const float mux = Sx * invS, muy = Sy * invS;
  // sum P and 1 - sum P = eps/(S + eps): 1 and 0 for plain softmax; an all-masked / all-negative map has sum P = 0
  const float sumP = kPlain ? 1.0f : S * invS;
  const float om = kPlain ? 0.f : ps.pc.eps * invS;

  float D = 0.f, creg = 0.f, ginv = 0.f, vx = 0.f, vy = 0.f;

  if constexpr (kVar) {  // second moments about the GLOBAL mean (Chan's update, all terms positive)
    const float ddx = mx - mux, ddy = my - muy;
    float ax = fmaf(S_loc * ddx, ddx, M2x), ay = fmaf(S_loc * ddy, ddy, M2y);
    group_sum2<GROUP>(ax, ay, red_b + gid * NW * 4, warp_g, lane);
    vx = ax * invS; vy = ay * invS;
    const float s2 = p.sigma * p.sigma, ex = vx - s2, ey = vy - s2;
    D = ex * ex + ey * ey;
    creg = 2.f * (ex * vx + ey * vy);
    if constexpr (!kPlain) {
      // d v_x/dP = (x - mu_x)^2 - 2 x mu_x (1 - sum P): the second term survives when sum P != 1
      creg = 2.f * (ex * (vx - 2.f * mux * mux * om) + ey * (vy - 2.f * muy * muy * om));
      ginv = om;   // stats[7] (no Gaussian for `var`): the backward needs 1 - sum P
    }
  }

  if constexpr (kWin) {
    float qa = 0.f, qb = 0.f, qc = 0.f, qd = 0.f;
    if (!win.empty()) {
      // axis factors of the window: every warp needs the sums; warp 0 also publishes the exponents
      float sx = 0.f, sy = 0.f;
      for (int j = win.j_lo + lane; j <= win.j_hi; j += 32) {
        const float d = fmaf(static_cast<float>(j), g.two_over_w, g.bias_w) - tx;
        const float a = g.k2 * d * d;
        sx += ex2(a);
        if (sw.ok && warp_g == 0) tabx[j - win.j_lo] = a;
      }
      for (int i = win.i_lo + lane; i <= win.i_hi; i += 32) {
        const float d = fmaf(static_cast<float>(i), g.two_over_h, g.bias_h) - ty;
        const float a = g.k2 * d * d;
        sy += ex2(a);
        if (sw.ok && warp_g == 0) taby[i - win.i_lo] = a;
      }
      {
        const float k = warp_sum2_transposed(sx, sy, lane);
        sx = __shfl_sync(kFull, k, 0);
        sy = __shfl_sync(kFull, k, 16);
      }
      ginv = 1.0f / (sx * sy + kEps);
      const float l2ginv = log2f(ginv);
      const float l2is = kPlain ? -log2f(S) : log2f(invS);   // log2 P = log2 f + l2is
      const float tlm1 = l2is - 1.0f;
      const float hinvS = 0.5f * invS;

      auto pixel = [&](float z, float lgG) {
        const float G = ex2(lgG);
        const float t0 = fmaf(z, kLog2e, -m2);
        const float e = act_fast<PA>(z, t0, thr);
        const float t = act_log2<PA>(e, t0);                     // log2 f (0 where f = 0: the term is weighted by f)
        if (kJS) {
          const float Mp = fmaf(e, hinvS, fmaf(0.5f, G, kEps));  // M + eps
          const float L = lg2(Mp);
          qa = fmaf(e * invS, (t + tlm1) - L, qa);                // P (log2 P - log2 M' - 1)
          qb = fmaf(G, lgG - L, qb);                             // G (log2 G - log2 M')
        } else if (kKL) {
          // sum_win P log2(G+eps) and sum_win P separately: folding the out-of-window constant log2(eps) = -79.7 into
          // every term (P (log2(G+eps) - log2 eps)) made D the difference of two ~55-sized numbers
          qa = fmaf(e * invS, lg2(G + kEps), qa);
          qb += e * invS;
        } else {
          const float P = e * invS, df = P - G;
          qa = fmaf(df, df, qa);                                 // (P - G)^2
          qb = fmaf(P, P, qb);                                   // P^2 (takes the window out of sum P^2)
          qc = fmaf(P, df, qc);                                  // P (P - G)
        }
      };

      if (sw.ok) {
        if constexpr (GROUP == 32) __syncwarp(); else __syncthreads();   // tables (and stash) visible
        const int npx = sw.nrw * sw.wcols;
        const float inv_wc = 1.0f / static_cast<float>(sw.wcols);
        for (int idx = lane_g; idx < npx; idx += GROUP) {
          const int r = static_cast<int>((static_cast<float>(idx) + 0.5f) * inv_wc);
          const int c = idx - r * sw.wcols;
          pixel(stash[r * sstride + sw.coff + c], (tabx[c] + taby[r]) + l2ginv);
        }
      } else {
        // window too large for the stash (large sigma): re-read it from global memory, lane <-> column
        const T* zt = reinterpret_cast<const T*>(zb);
        const int nwr = win.i_hi - win.i_lo + 1;
        for (int j0 = win.j_lo; j0 <= win.j_hi; j0 += 32) {
          const int j = j0 + lane;
          const bool cact = j <= win.j_hi;
          const float dx = fmaf(static_cast<float>(j), g.two_over_w, g.bias_w) - tx;
          const float ax = g.k2 * dx * dx;
          for (int r = warp_g; r < nwr; r += NW) {
            if (cact) {
              float one[1];
              VecIO<T, 1>::load(zt, static_cast<long>(win.i_lo + r) * W + j, one);
              const float dy = fmaf(static_cast<float>(win.i_lo + r), g.two_over_h, g.bias_h) - ty;
              pixel(one[0], fmaf(g.k2 * dy, dy, ax) + l2ginv);
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
        creg = 0.5f * kLn2 * (sumP + qa);        // 1/2 sum P (ln P - ln M'); outside the window every term is ln 2
        D = fmaf(0.5f * kLn2, qb, creg);         // + 1/2 sum G (ln G - ln M')
      } else {
        // sum P ln P = (ln2 sum f log2 f)/(S + eps) + sum P ln(1/(S + eps))
        const float plnp = kPlain ? fmaf(kLn2 * invS, Tt, -logf(S)) : fmaf(kLn2 * invS, Tt, sumP * logf(invS));
        D = plnp - kLn2 * fmaf(kLog2Eps, sumP - qb, qa);   // outside the window G + eps = eps exactly
        creg = D + sumP;                         // + sum P^2/(P + eps) = sum P (P is 0 or >> eps)
      }
    }
  }

  if (lane_g == 0) write_outputs(p, hm, m2, invS, mux, muy, vx, vy, creg, ginv, tx, ty, D);
}

template <typename T, int VEC, int GROUP, int REG, bool FLIP = false, int PA = DSNT_PREACT_SOFTMAX>
__global__ void __launch_bounds__(stream_block_threads<GROUP>()) head_fwd_fast_kernel(const HeadFwdFastParams ps) {
  head_fwd_fast_body<T, VEC, GROUP, REG, FLIP, PA>(ps, blockIdx.x);
}

// ================================================================================================ backward
struct HeadBwdFastParams {
  HeadBwdParams base;
  Geom g;
  FastGeom f;
};

__device__ __forceinline__ void st_vec16(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

template <typename T>
__device__ __forceinline__ void st_scalar(T* p, float v);
template <>
__device__ __forceinline__ void st_scalar<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_scalar<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

template <typename T, int VEC, int GROUP, int REG>
__device__ __forceinline__ void head_bwd_fast_body(const HeadBwdFastParams& ps, long block_index) {
  static_assert(sizeof(T) * VEC == 16, "fast path = 16-byte vectors");
  constexpr int BLOCK = stream_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  constexpr int U = kFastU;
  constexpr bool kKL = REG == DSNT_REG_KL;
  constexpr bool kJS = REG == DSNT_REG_JS;
  constexpr bool kVar = REG == DSNT_REG_VAR;
  constexpr bool kMSE = REG == DSNT_REG_MSE;
  constexpr bool kWin = kKL || kJS || kMSE;
  __shared__ __align__(16) float stash_all[kWin ? GPB * (kStashFloats + 2 * kTabN) : 4];

  const HeadBwdParams& p = ps.base;
  const Geom& g = ps.g;
  const FastGeom& fg = ps.f;
  const int tid = threadIdx.x;
  const int gid = tid / GROUP, lane_g = tid % GROUP, warp_g = lane_g >> 5, lane = tid & 31;
  const long hm = block_index * GPB + gid;
  if (hm >= p.n) return;  // GROUP == 32 only

  const int H = p.H, W = p.W;
  const HmRef ref = locate(p.st, hm, static_cast<long>(H) * W * sizeof(T));
  const char* src = static_cast<const char*>(p.z) + ref.z_bytes + static_cast<size_t>(lane_g) * 16;
  char* dst = static_cast<char*>(p.dz) + ref.dz_bytes + static_cast<size_t>(lane_g) * 16;

  const BwdScalars s = load_bwd_scalars<true>(p, hm, ref.nl, REG);

  // constant part of (g - c):  -c, plus the out-of-window value of rho*r
  float cbase = -s.c;
  if (kJS) cbase = fmaf(0.5f * kLn2, s.rho, cbase);                          // r -> 1/2 ln 2
  if (kKL) cbase = fmaf(s.rho, 1.0f - kLnEps + kLn2 * s.l2is, cbase);         // r = ln2 t + this
  const float rho_t = kKL ? s.rho * kLn2 : 0.f;
  const float rho_p = kMSE ? 2.f * s.rho : 0.f;

  // ---- thread geometry (fixed columns) and the column part of (g - c)
  const int cv = lane_g & ((1 << fg.wv_shift) - 1);
  const int row0 = lane_g >> fg.wv_shift;
  float acol[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) {
    const float x = axis_coord(cv * VEC + c, g.two_over_w, g.bias_w);
    float a = s.a * x;
    if (kVar) { const float dx = x - s.mux; a = fmaf(s.kx * dx, dx, a); }
    acol[c] = a;
  }
  float ybase = axis_coord(row0, g.two_over_h, g.bias_h);

  // ---- window and stash slot
  float* stash = stash_all + gid * (kStashFloats + 2 * kTabN);
  float* tabx = stash + kStashFloats;   // gx_j (unnormalised) for the window columns
  float* taby = tabx + kTabN;           // gy_i * ginv for the window rows
  Window win{1, 0, 1, 0};
  StashWin sw{0, 0, 0, 0, 0, 0, false};
  int srow = 0, scol = -1, sstride = 0;
  if constexpr (kWin) {
    win = make_window(g, H, W, s.tx, s.ty);
    sw = make_stash_window<VEC>(win);
    if (sw.ok) {
      const int dc = cv - sw.cv_lo;
      scol = (dc >= 0 && dc < sw.nvc) ? dc * VEC : -1;
      sstride = sw.nvc * VEC;
      srow = row0 - sw.i_lo;
    }
  }

  // ---- streaming loop: every pixel gets the out-of-window gradient; window vectors go to the stash instead
  for (int b = 0; b < fg.nbatch; ++b) {
    uint4 raw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) raw[u] = ld_raw16(src + static_cast<size_t>(u) * (GROUP * 16));
    require_all_loaded<U>(raw);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float v[VEC];
      unpack16<T, VEC>(raw[u], v);
      bool stashed = false;
      if constexpr (kWin) {
        const int r = srow + u * fg.rstep;
        stashed = scol >= 0 && static_cast<unsigned>(r) < static_cast<unsigned>(sw.nrw);
        if (stashed) {
          float4* sd = reinterpret_cast<float4*>(stash + r * sstride + scol);
          sd[0] = make_float4(v[0], v[1], v[2], v[3]);
          if constexpr (VEC == 8) sd[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      const float y = fmaf(static_cast<float>(u), fg.dy_step, ybase);
      float rowc = fmaf(s.b, y, cbase);
      if (kVar) { const float dy = y - s.muy; rowc = fmaf(s.ky * dy, dy, rowc); }
      float out[VEC];
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float t = fmaf(v[c], kLog2e, -s.m2);
        const float P = ex2(t) * s.invS;
        float gmc = acol[c] + rowc;
        if (kKL) gmc = fmaf(rho_t, t, gmc);
        if (kMSE) gmc = fmaf(rho_p, P, gmc);
        out[c] = P * gmc;
      }
      if (!stashed) st_vec16(dst + static_cast<size_t>(u) * (GROUP * 16), pack16<T, VEC>(out));
    }
    src += static_cast<size_t>(U) * (GROUP * 16);
    dst += static_cast<size_t>(U) * (GROUP * 16);
    ybase += fg.dy_batch;
    if constexpr (kWin) srow += U * fg.rstep;
  }

  // ---- epilogue: the pixels where the Gaussian matters
  if constexpr (kWin) {
    if (win.empty()) return;     // uniform over the group
    T* dzb = reinterpret_cast<T*>(static_cast<char*>(p.dz) + ref.dz_bytes);

    auto pixel = [&](float z, float G, float x, float y) -> float {
      const float t = fmaf(z, kLog2e, -s.m2);
      const float P = ex2(t) * s.invS;
      float gmc = fmaf(s.a, x, fmaf(s.b, y, cbase));
      if (kJS) {
        // rho r = rho/2 ln2 (1 - lg2(1+q)),  q = (G + 2 eps)/P;  the "1" is already in cbase
        const float q = (G + 2.f * kEps) * rcp(fmaxf(P, 1e-37f));
        gmc = fmaf(-0.5f * kLn2 * s.rho, lg2(1.0f + q), gmc);
      } else if (kKL) {
        gmc = fmaf(rho_t, t, gmc);
        gmc = fmaf(-kLn2 * s.rho, lg2(G + kEps) - kLog2Eps, gmc);
      } else {
        gmc = fmaf(rho_p, P - G, gmc);
      }
      return P * gmc;
    };

    if (sw.ok) {
      // tables: unnormalised gx per window column (of the vector-aligned stash row), normalised gy per row
      if (warp_g == 0) {
        for (int k = lane; k < sstride; k += 32) {
          const float d = axis_coord(sw.cv_lo * VEC + k, g.two_over_w, g.bias_w) - s.tx;
          tabx[k] = ex2(g.k2 * d * d);
        }
        for (int k = lane; k < sw.nrw; k += 32) {
          const float d = axis_coord(sw.i_lo + k, g.two_over_h, g.bias_h) - s.ty;
          taby[k] = ex2(g.k2 * d * d) * s.ginv;
        }
      }
      if constexpr (GROUP == 32) __syncwarp(); else __syncthreads();
      const int nv = sw.nrw * sw.nvc;
      const float inv_nvc = 1.0f / static_cast<float>(sw.nvc);
      for (int idx = lane_g; idx < nv; idx += GROUP) {
        const int r = static_cast<int>((static_cast<float>(idx) + 0.5f) * inv_nvc);
        const int k = (idx - r * sw.nvc) * VEC;
        const float y = axis_coord(sw.i_lo + r, g.two_over_h, g.bias_h);
        const float gyn = taby[r];
        float z[VEC], out[VEC];
        const float4* sp = reinterpret_cast<const float4*>(stash + r * sstride + k);
        const float4 z0 = sp[0];
        z[0] = z0.x; z[1] = z0.y; z[2] = z0.z; z[3] = z0.w;
        if constexpr (VEC == 8) { const float4 z1 = sp[1]; z[4] = z1.x; z[5] = z1.y; z[6] = z1.z; z[7] = z1.w; }
        const int col0 = sw.cv_lo * VEC + k;
#pragma unroll
        for (int c = 0; c < VEC; ++c)
          out[c] = pixel(z[c], tabx[k + c] * gyn, axis_coord(col0 + c, g.two_over_w, g.bias_w), y);
        st_vec16(dzb + static_cast<long>(sw.i_lo + r) * W + col0, pack16<T, VEC>(out));
      }
    } else {
      // window too large for the stash: the loop stored the out-of-window value everywhere; overwrite the
      // window pixels (ordered after the loop's stores by the barrier), lane <-> column
      if constexpr (GROUP == 32) __syncwarp(); else __syncthreads();
      const T* zt = reinterpret_cast<const T*>(static_cast<const char*>(p.z) + ref.z_bytes);
      constexpr int NW = GROUP / 32;
      for (int j0 = win.j_lo; j0 <= win.j_hi; j0 += 32) {
        const int j = j0 + lane;
        if (j > win.j_hi) continue;
        const float x = axis_coord(j, g.two_over_w, g.bias_w);
        const float dx = x - s.tx;
        const float gx = ex2(g.k2 * dx * dx);
        for (int i = win.i_lo + warp_g; i <= win.i_hi; i += NW) {
          const float y = axis_coord(i, g.two_over_h, g.bias_h);
          const float dy = y - s.ty;
          float one[1];
          VecIO<T, 1>::load(zt, static_cast<long>(i) * W + j, one);
          st_scalar<T>(dzb + static_cast<long>(i) * W + j, pixel(one[0], gx * ex2(g.k2 * dy * dy) * s.ginv, x, y));
        }
      }
    }
  }
}

template <typename T, int VEC, int GROUP, int REG>
__global__ void __launch_bounds__(stream_block_threads<GROUP>()) head_bwd_fast_kernel(const HeadBwdFastParams ps) {
  head_bwd_fast_body<T, VEC, GROUP, REG>(ps, blockIdx.x);
}

}  // namespace dsnt
