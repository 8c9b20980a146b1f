// head_step.cuh -- the whole training step of the head in ONE pass over the logits (8 B/pixel fp32 instead of 12).
//
// The two-kernel contract (head_fast.cuh / head_stream.cuh) reads z in the forward, then reads it again in the
// backward.  In a training step the only upstream gradient is the scalar d(loss) (src/dsnt/bin/train.py:381,
// `loss.backward()`), and the only thing the backward needs from other heatmaps is the denominator of masked_average
// (src/dsnt/nn.py:81-94), which depends on the MASK alone and is therefore known before the forward.  So each
// heatmap can be brought on chip once, reduced (forward) and overwritten with dL/dz (backward) before it leaves:
//
//   TMA bulk load   cp.async.bulk.shared.global  one heatmap (H*W*sizeof bytes, contiguous) -> the warp's smem buffer,
//                   completion on the warp's mbarrier (one elected lane issues it; no registers are held by the load)
//   forward         max, sum, coordinate moments, variance / windowed divergence: LDS sweeps over the buffer
//   backward        dz = P (a x + b y + rho r - c), written IN PLACE over z in the buffer
//   TMA bulk store  cp.async.bulk.global.shared  buffer -> dz, then the next heatmap's load is issued
//
// One warp owns one buffer and loops over heatmaps (persistent CTAs, one per SM, up to 14 x 16 KiB buffers);
// warps are fully independent -- no CTA barrier -- and the 12-14 loads in flight per SM keep HBM busy while other
// warps compute.  Arithmetic is the same as the two-kernel path (same closed forms, same Gaussian window), so the
// results agree with it to rounding; parity is checked against the same oracle.
#pragma once

#include "finish_common.cuh"
#include "head_stream.cuh"

namespace dsnt {

constexpr int kStepMaxWarps = 16;
constexpr int kStepMaxBufs = 32;
constexpr int kStepSmemBudget = 224 * 1024;   // of the 227 KiB a CTA may opt in to

struct HeadStepParams {
  const void* z;
  void* dz;
  const float* target;   // [N,2] or null
  const float* mask;     // [N] or null
  const float* denom;    // device scalar: max(sum mask, 1) (or the heatmap count without a mask)
  const float* g_loss;   // device scalar d(loss), or null = 1
  float* coords;         // [N,2]
  float* stats;          // [N,8] or null
  float* terms;          // [N,2] or null
  long n;
  int H, W;
  int flags;
  float sigma, reg_coeff;
  Geom g;
  int buf_bytes;         // one buffer (heatmap bytes rounded up to 128)
  int nwarps;            // groups (of GROUP/32 warps) per CTA
  int nbufs;             // shared-memory buffers per CTA, >= nwarps: a ring shared by the groups
  int direct_store;      // 1: dz goes to global memory with 128-bit stores straight from registers; 0: in place + bulk store
  Stacks st;             // head_step2: stacked hourglass (count > 1): n = count * n_per heatmaps, stack s at z + st.z_off[s];
                         //   target / mask are one stack long (indexed inside the stack), coords / stats by heatmap
  PeerXchg xc;           // head_step2, single-launch form with a sharded batch (world > 1): the mask count and the loss sums
                         //   cross the ranks from inside this kernel (finish_common.cuh: peer_exchange_sum3)
  float* out8;           // head_step2, single-launch form (dsnt_head_step_fused): the loss block of dsnt_finish_loss, or null
  float* ws;             //   its workspace (dsnt_finish_workspace_bytes); denom may then be null = computed from the mask here
  int debug;             // head_step2 (DSNT_TUNE_STEP_DEBUG, measurements only): 1 = copy z -> dz with LDS + STG and no
                         // arithmetic: the floor of the data-movement skeleton
  int pace;              // head_step2: minimum SM clocks between two bulk loads issued by a CTA (0 = unpaced)
  int stagger_ns;        // head_step2: warp w starts w * stagger_ns late, so the warps of a CTA are not all in the same sweep
};

// ---------------------------------------------------------------------------------- PTX: mbarrier + bulk copies
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// global -> shared, completion (bytes) signalled on the mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// shared -> global; wait_read returns once the source buffer may be overwritten
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------------------------- shared-memory vector access
template <typename T, int VEC>
__device__ __forceinline__ void lds_vec(const T* buf, int f, float (&v)[VEC]) {
  if constexpr (sizeof(T) * VEC == 16) {
    const uint4 r = *reinterpret_cast<const uint4*>(buf + static_cast<size_t>(f) * VEC);
    if constexpr (sizeof(T) == 4) {
      v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
    } else {
      v[0] = bf16lo(r.x); v[1] = bf16hi(r.x); v[2] = bf16lo(r.y); v[3] = bf16hi(r.y);
      v[4] = bf16lo(r.z); v[5] = bf16hi(r.z); v[6] = bf16lo(r.w); v[7] = bf16hi(r.w);
    }
  } else {
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      if constexpr (sizeof(T) == 4) v[c] = buf[static_cast<size_t>(f) * VEC + c];
      else v[c] = __bfloat162float(buf[static_cast<size_t>(f) * VEC + c]);
    }
  }
}
template <typename T, int VEC>
__device__ __forceinline__ void sts_vec(T* buf, int f, const float (&v)[VEC]) {
  if constexpr (sizeof(T) * VEC == 16) {
    uint4 r;
    if constexpr (sizeof(T) == 4) {
      r.x = __float_as_uint(v[0]); r.y = __float_as_uint(v[1]); r.z = __float_as_uint(v[2]); r.w = __float_as_uint(v[3]);
    } else {
      r.x = pack_bf16(v[0], v[1]); r.y = pack_bf16(v[2], v[3]); r.z = pack_bf16(v[4], v[5]); r.w = pack_bf16(v[6], v[7]);
    }
    *reinterpret_cast<uint4*>(buf + static_cast<size_t>(f) * VEC) = r;
  } else {
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      if constexpr (sizeof(T) == 4) buf[static_cast<size_t>(f) * VEC + c] = v[c];
      else buf[static_cast<size_t>(f) * VEC + c] = __float2bfloat16_rn(v[c]);
    }
  }
}
template <typename T>
__device__ __forceinline__ float lds_one(const T* buf, int i) {
  if constexpr (sizeof(T) == 4) return buf[i];
  else return __bfloat162float(buf[i]);
}

// ---------------------------------------------------------------------------------- group = 1 or 2 warps per heatmap
// GROUP == 64: two warps share a heatmap (halves the dependent-chain latency per heatmap and doubles the warps the
// schedulers can pick from at the same shared-memory footprint); they meet at a named barrier (id = group + 1).
template <int GROUP>
__device__ __forceinline__ void step_group_bar(int grp) {
  if constexpr (GROUP == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(GROUP) : "memory");
}
template <int GROUP>
__device__ __forceinline__ float step_group_max(float v, float* scr, int grp, int gw, int lane) {
  v = warp_max(v);
  if constexpr (GROUP > 32) {
    if (lane == 0) scr[gw] = v;
    step_group_bar<GROUP>(grp);
    v = fmaxf(scr[0], scr[1]);
  }
  return v;
}
template <int GROUP>
__device__ __forceinline__ void step_group_sum4(float& a, float& b, float& c, float& d, float* scr, int grp, int gw, int lane) {
  const float k = warp_sum4_transposed(a, b, c, d, lane);
  if constexpr (GROUP == 32) {
    a = __shfl_sync(kFull, k, 0); b = __shfl_sync(kFull, k, 8); c = __shfl_sync(kFull, k, 16); d = __shfl_sync(kFull, k, 24);
  } else {
    if ((lane & 7) == 0) scr[gw * 4 + (lane >> 3)] = k;
    step_group_bar<GROUP>(grp);
    a = scr[0] + scr[4]; b = scr[1] + scr[5]; c = scr[2] + scr[6]; d = scr[3] + scr[7];
  }
}

// ================================================================================================ the step kernel
// FIXC: 32 lanes cover a whole number of rows (W/VEC divides 32), so a lane always sees the same VEC columns.
// The CTA's heatmaps ("tiles" t = 0, 1, ...; heatmap = t * gridDim.x + blockIdx.x) go round a RING of nbufs buffers:
// tile t lives in buffer t % nbufs and is processed by group t % nwarps; whoever finishes tile t refills its buffer with
// tile t + nbufs.  nbufs == nwarps: every group owns one buffer (its next load starts when it is done); nbufs > nwarps:
// nbufs - nwarps further loads are in flight while every group computes.
template <typename T, int VEC, int REG, bool FIXC, int GROUP>
__global__ void __launch_bounds__(kStepMaxWarps * 32, 1) head_step_kernel(const HeadStepParams p) {
  constexpr bool kKL = REG == DSNT_REG_KL;
  constexpr bool kJS = REG == DSNT_REG_JS;
  constexpr bool kVar = REG == DSNT_REG_VAR;
  constexpr bool kMSE = REG == DSNT_REG_MSE;
  constexpr bool kWin = kKL || kJS || kMSE;
  extern __shared__ __align__(128) unsigned char step_smem[];
  __shared__ __align__(8) unsigned long long bars[kStepMaxBufs];
  // loads issued into each buffer so far.  A parity wait cannot tell "round r+1 not armed yet" from "done" when the
  // waiter is a whole phase ahead, and with nbufs > nwarps the group that consumes round r+1 of a buffer is not the one
  // that consumed round r -- so it first waits (on this counter) until that group has re-armed the barrier.
  __shared__ volatile int issued[kStepMaxBufs];
  __shared__ float scr_all[kStepMaxWarps][3][8];   // cross-warp reductions of a group: three rotating slots

  // p.nwarps counts GROUPS (of GROUP/32 warps) here
  const int grp = threadIdx.x / GROUP, tg = threadIdx.x % GROUP, gw = tg >> 5, lane = threadIdx.x & 31;
  if (grp >= p.nwarps) return;
  float (*scr)[8] = scr_all[grp];
  const Geom& g = p.g;
  const int H = p.H, W = p.W;
  const int wv = g.wv, nvec = g.nvec;
  const uint32_t hm_bytes = static_cast<uint32_t>(H) * W * sizeof(T);
  const int NW = p.nwarps, NB = p.nbufs;
  const uint32_t smem0_s = smem_u32(step_smem), bars0_s = smem_u32(&bars[0]);
  const char* zsrc = static_cast<const char*>(p.z);
  char* dzdst = static_cast<char*>(p.dz);
  // tiles of this CTA
  const long ntiles = p.n > blockIdx.x ? (p.n - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  // the group that will process tile t < nbufs also arms its barrier and issues its first load
  if (tg == 0) {
    for (int t = grp; t < NB; t += NW) {
      const uint32_t bar = bars0_s + 8 * t;
      mbar_init(bar, 1);
      if (t < ntiles) {
        mbar_expect_tx(bar, hm_bytes);
        bulk_load(smem0_s + static_cast<uint32_t>(t) * p.buf_bytes, zsrc + (t * static_cast<long>(gridDim.x) + blockIdx.x) * hm_bytes,
                  hm_bytes, bar);
      }
      issued[t] = 1;
    }
  }
  __syncthreads();   // every barrier is initialised before any group looks at a buffer another group armed
  const float gl = p.g_loss ? __ldg(p.g_loss) : 1.0f;
  const float inv_denom = 1.0f / __ldg(p.denom);
  const float s2 = p.sigma * p.sigma;

  // lane geometry when the columns are fixed
  float xs[VEC];
  const int cv0 = FIXC ? (tg & (wv - 1)) : 0;
  const int row0 = FIXC ? tg / wv : 0;
  const int rstep = FIXC ? GROUP / wv : 0;
  if constexpr (FIXC) {
#pragma unroll
    for (int c = 0; c < VEC; ++c) xs[c] = axis_coord(cv0 * VEC + c, g.two_over_w, g.bias_w);
  }

  for (long t = grp; t < ntiles; t += NW) {
    const long hm = t * static_cast<long>(gridDim.x) + blockIdx.x;
    const uint32_t round = static_cast<uint32_t>(t / NB);
    const uint32_t bi = static_cast<uint32_t>(t - static_cast<long>(round) * NB);
    const uint32_t phase = round & 1u;
    T* buf = reinterpret_cast<T*>(step_smem + static_cast<size_t>(bi) * p.buf_bytes);
    const uint32_t buf_s = smem0_s + bi * static_cast<uint32_t>(p.buf_bytes), bar_s = bars0_s + 8 * bi;
    // per-heatmap scalars while the load is in flight
    float tx = 0.f, ty = 0.f;
    if (p.target) {
      const float2 t = __ldg(reinterpret_cast<const float2*>(p.target) + hm);
      tx = t.x; ty = t.y;
    }
    const float wgt = (p.mask ? __ldg(p.mask + hm) : 1.0f) * inv_denom;
    Window win{1, 0, 1, 0};
    if constexpr (kWin) win = make_window(g, H, W, tx, ty);

    if (NB != NW) {
      while (issued[bi] < static_cast<int>(round) + 1) { }   // armed for this round by the consumer of the previous one
    }
    mbar_wait(bar_s, phase);

    // ---------------------------------------------------------------- forward: max
    float mloc = -INFINITY;
#pragma unroll 4
    for (int f = tg; f < nvec; f += GROUP) {
      float v[VEC];
      lds_vec<T, VEC>(buf, f, v);
#pragma unroll
      for (int c = 0; c < VEC; ++c) mloc = fmaxf(mloc, v[c]);
    }
    const float m2 = step_group_max<GROUP>(mloc, scr[0], grp, gw, lane) * kLog2e;

    // ---------------------------------------------------------------- forward: S, S_x, S_y (+ sum e t / sum e^2)
    float S = 0.f, Sx = 0.f, Sy = 0.f, Tt = 0.f;
    {
      VecWalker wk(tg, GROUP, wv);
      float y = FIXC ? axis_coord(row0, g.two_over_h, g.bias_h) : 0.f;
      const float dy_step = static_cast<float>(rstep) * g.two_over_h;
#pragma unroll 4
      for (int f = tg; f < nvec; f += GROUP) {
        float v[VEC];
        lds_vec<T, VEC>(buf, f, v);
        if constexpr (!FIXC) {
#pragma unroll
          for (int c = 0; c < VEC; ++c) xs[c] = axis_coord(wk.cv * VEC + c, g.two_over_w, g.bias_w);
          y = axis_coord(wk.row, g.two_over_h, g.bias_h);
        }
        float rs = 0.f;
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const float t = fmaf(v[c], kLog2e, -m2);
          const float e = ex2(t);
          rs += e;
          Sx = fmaf(e, xs[c], Sx);
          if (kKL) Tt = fmaf(e, fmaxf(t, -1e30f), Tt);
          if (kMSE) Tt = fmaf(e, e, Tt);
        }
        S += rs;
        Sy = fmaf(rs, y, Sy);
        if constexpr (FIXC) y += dy_step; else wk.next();
      }
    }
    step_group_sum4<GROUP>(S, Sx, Sy, Tt, scr[1], grp, gw, lane);
    const float invS = 1.0f / S;
    const float mux = Sx * invS, muy = Sy * invS;

    float D = 0.f, creg = 0.f, ginv = 0.f, vx = 0.f, vy = 0.f;

    // ---------------------------------------------------------------- forward: variance (second sweep about the mean)
    if constexpr (kVar) {
      float ax = 0.f, ay = 0.f;
      VecWalker wk(tg, GROUP, wv);
      float y = FIXC ? axis_coord(row0, g.two_over_h, g.bias_h) : 0.f;
      const float dy_step = static_cast<float>(rstep) * g.two_over_h;
      float dx2[VEC];
      if constexpr (FIXC) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) { const float d = xs[c] - mux; dx2[c] = d * d; }
      }
      for (int f = tg; f < nvec; f += GROUP) {
        float v[VEC];
        lds_vec<T, VEC>(buf, f, v);
        if constexpr (!FIXC) {
#pragma unroll
          for (int c = 0; c < VEC; ++c) { const float d = axis_coord(wk.cv * VEC + c, g.two_over_w, g.bias_w) - mux; dx2[c] = d * d; }
          y = axis_coord(wk.row, g.two_over_h, g.bias_h);
        }
        float rs = 0.f;
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const float e = ex2(fmaf(v[c], kLog2e, -m2));
          rs += e;
          ax = fmaf(e, dx2[c], ax);
        }
        const float dy = y - muy;
        ay = fmaf(rs * dy, dy, ay);
        if constexpr (FIXC) y += dy_step; else wk.next();
      }
      float z0 = 0.f, z1 = 0.f;
      step_group_sum4<GROUP>(ax, ay, z0, z1, scr[2], grp, gw, lane);
      vx = ax * invS;
      vy = ay * invS;
      const float ex = vx - s2, ey = vy - s2;
      D = ex * ex + ey * ey;
      creg = 2.f * (ex * vx + ey * vy);
    }

    // ---------------------------------------------------------------- forward: divergence on the Gaussian window
    if constexpr (kWin) {
      float qa = 0.f, qb = 0.f, qc = 0.f, qd = 0.f;
      if (!win.empty()) {
        float sx = 0.f, sy = 0.f;
        for (int j = win.j_lo + lane; j <= win.j_hi; j += 32) {
          const float d = axis_coord(j, g.two_over_w, g.bias_w) - tx;
          sx += ex2(g.k2 * d * d);
        }
        for (int i = win.i_lo + lane; i <= win.i_hi; i += 32) {
          const float d = axis_coord(i, g.two_over_h, g.bias_h) - ty;
          sy += ex2(g.k2 * d * d);
        }
        {
          const float k = warp_sum2_transposed(sx, sy, lane);
          sx = __shfl_sync(kFull, k, 0);
          sy = __shfl_sync(kFull, k, 16);
        }
        ginv = 1.0f / (sx * sy + kEps);
        const float l2ginv = log2f(ginv);
        const float tlm1 = -log2f(S) - 1.0f;       // log2 P - 1 = t + tlm1
        const float hinvS = 0.5f * invS;
        const int wcols = win.j_hi - win.j_lo + 1;
        const int npx = (win.i_hi - win.i_lo + 1) * wcols;
        const float inv_wc = 1.0f / static_cast<float>(wcols);
#pragma unroll 2
        for (int idx = tg; idx < npx; idx += GROUP) {
          const int r = static_cast<int>((static_cast<float>(idx) + 0.5f) * inv_wc);
          const int i = win.i_lo + r, j = win.j_lo + (idx - r * wcols);
          const float dx = axis_coord(j, g.two_over_w, g.bias_w) - tx;
          const float dy = axis_coord(i, g.two_over_h, g.bias_h) - ty;
          const float lgG = fmaf(g.k2 * dx, dx, g.k2 * dy * dy) + l2ginv;
          const float G = ex2(lgG);
          const float t = fmaf(lds_one<T>(buf, i * W + j), kLog2e, -m2);
          const float e = ex2(t);
          if (kJS) {
            const float Mp = fmaf(e, hinvS, fmaf(0.5f, G, kEps));
            const float L = lg2(Mp);
            qa = fmaf(e * invS, (t + tlm1) - L, qa);
            qb = fmaf(G, lgG - L, qb);
          } else if (kKL) {
            qa = fmaf(e * invS, lg2(G + kEps), qa);     // see head_fast.cuh: keeps D free of the -79.7 offset
            qb += e * invS;
          } else {
            const float P = e * invS, df = P - G;
            qa = fmaf(df, df, qa);
            qb = fmaf(P, P, qb);
            qc = fmaf(P, df, qc);
          }
        }
      }
      step_group_sum4<GROUP>(qa, qb, qc, qd, scr[2], grp, gw, lane);
      if (kMSE) {
        const float outside = fmaxf(fmaf(Tt * invS, invS, -qb), 0.f);
        D = outside + qa;
        creg = 2.f * (outside + qc);
      } else if (kJS) {
        creg = 0.5f * kLn2 * (1.0f + qa);
        D = fmaf(0.5f * kLn2, qb, creg);
      } else {
        const float plnp = fmaf(kLn2 * invS, Tt, -logf(S));
        D = plnp - kLn2 * fmaf(kLog2Eps, 1.0f - qb, qa);   // outside the window G + eps = eps exactly
        creg = D + 1.0f;
      }
    }

    // ---------------------------------------------------------------- outputs + the scalars of the backward
    float dist = 0.f, a = 0.f, b = 0.f;
    if (p.target) {
      const float dx = mux - tx, dy = muy - ty;
      const float d2 = dx * dx + dy * dy;
      dist = sqrtf(d2);
      const float invd = d2 > 0.f ? rsqrtf(d2) : ((p.flags & DSNT_FLAG_STRICT_NAN) ? INFINITY : 0.f);
      a = gl * wgt * (dx * invd);
      b = gl * wgt * (dy * invd);
    }
    const float rho = gl * wgt * p.reg_coeff;
    if (tg == 0) {
      reinterpret_cast<float2*>(p.coords)[hm] = make_float2(mux, muy);
      if (p.stats) {
        float4* st = reinterpret_cast<float4*>(p.stats + hm * kStatsK);
        st[0] = make_float4(m2, invS, mux, muy);
        st[1] = make_float4(vx, vy, creg, ginv);
      }
      if (p.terms) reinterpret_cast<float2*>(p.terms)[hm] = make_float2(dist, D);
    }

    // ---------------------------------------------------------------- backward: dz in place
    const float cc = fmaf(a, mux, fmaf(b, muy, rho * creg));
    float cbase = -cc;
    if (kJS) cbase = fmaf(0.5f * kLn2, rho, cbase);
    if (kKL) cbase = fmaf(rho, 1.0f - kLnEps + kLn2 * __log2f(invS), cbase);
    const float rho_t = kKL ? rho * kLn2 : 0.f;
    const float rho_p = kMSE ? 2.f * rho : 0.f;
    const float kx = kVar ? rho * 2.f * (vx - s2) : 0.f;
    const float ky = kVar ? rho * 2.f * (vy - s2) : 0.f;
    {
      float acol[VEC], gxs[VEC];
      bool anycol = false;
      auto init_cols = [&](int col0) {
        anycol = false;
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const float x = axis_coord(col0 + c, g.two_over_w, g.bias_w);
          float av = a * x;
          if (kVar) { const float d = x - mux; av = fmaf(kx * d, d, av); }
          acol[c] = av;
          gxs[c] = 0.f;
          if (kWin) {
            const bool in = col0 + c >= win.j_lo && col0 + c <= win.j_hi;
            const float d = x - tx;
            gxs[c] = in ? ex2(g.k2 * d * d) : 0.f;
            anycol |= in;
          }
        }
      };
      VecWalker wk(tg, GROUP, wv);
      int row = row0;
      if constexpr (FIXC) init_cols(cv0 * VEC);
#pragma unroll 4
      for (int f = tg; f < nvec; f += GROUP) {
        float v[VEC];
        lds_vec<T, VEC>(buf, f, v);
        if constexpr (!FIXC) { init_cols(wk.cv * VEC); row = wk.row; }
        const float y = axis_coord(row, g.two_over_h, g.bias_h);
        float rowc = fmaf(b, y, cbase);
        if (kVar) { const float d = y - muy; rowc = fmaf(ky * d, d, rowc); }
        float out[VEC], Pv[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const float t = fmaf(v[c], kLog2e, -m2);
          Pv[c] = ex2(t) * invS;
          float gmc = acol[c] + rowc;
          if (kKL) gmc = fmaf(rho_t, fmaxf(t, -1e30f), gmc);
          if (kMSE) gmc = fmaf(rho_p, Pv[c], gmc);
          out[c] = gmc;
        }
        if constexpr (kWin) {
          // A WARP-UNIFORM branch (ptxas if-converts a per-lane one and then every pixel pays the rcp + lg2): taken when
          // any lane's vector touches the window.  Lanes outside it have G = 0, for which the window term vanishes by
          // itself (lg2(1 + 2eps/P), lg2(eps) - lg2(eps), 0), so no per-lane select is needed.
          const bool heavy = anycol && row >= win.i_lo && row <= win.i_hi;
          if (__any_sync(kFull, heavy)) {
            const float d = y - ty;
            const float gyn = heavy ? ex2(g.k2 * d * d) * ginv : 0.f;
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
              const float G = gxs[c] * gyn;
              if (kJS) {
                const float q = (G + 2.f * kEps) * rcp(fmaxf(Pv[c], 1e-37f));
                out[c] = fmaf(-0.5f * kLn2 * rho, lg2(1.0f + q), out[c]);
              } else if (kKL) {
                out[c] = fmaf(-kLn2 * rho, lg2(G + kEps) - kLog2Eps, out[c]);
              } else {
                out[c] = fmaf(-rho_p, G, out[c]);
              }
            }
          }
        }
#pragma unroll
        for (int c = 0; c < VEC; ++c) out[c] *= Pv[c];
        if (p.direct_store) VecIO<T, VEC>::store(reinterpret_cast<T*>(dzdst + hm * hm_bytes), static_cast<long>(f) * VEC, out);
        else sts_vec<T, VEC>(buf, f, out);
        if constexpr (FIXC) row += rstep; else wk.next();
      }
    }

    // ---------------------------------------------------------------- hand the buffer back / to the copy engine
    if (!p.direct_store) fence_async_smem();      // this lane's generic-proxy writes -> visible to the async proxy
    step_group_bar<GROUP>(grp);
    if (tg == 0) {
      if (!p.direct_store) {
        bulk_store(dzdst + hm * hm_bytes, buf_s, hm_bytes);
        bulk_wait_read();      // the buffer has been read out: it may be refilled
      }
      const long nt = t + NB;  // the tile that takes this buffer over
      if (nt < ntiles) {
        mbar_expect_tx(bar_s, hm_bytes);
        bulk_load(buf_s, zsrc + (nt * static_cast<long>(gridDim.x) + blockIdx.x) * hm_bytes, hm_bytes, bar_s);
        __threadfence_block();
        issued[bi] = static_cast<int>(round) + 2;
      }
    }
    step_group_bar<GROUP>(grp);
  }
  if (tg == 0) bulk_wait_all();   // every store has landed before the warp retires
}

}  // namespace dsnt
