// step_pair.cu -- the one-pass training step for 256x256 heatmaps (BASELINE config 5) on a CLUSTER of SMs.
//
// A 256x256 fp32 heatmap is 256 KiB: more than the 227 KiB of shared memory one CTA can have, which is why config 5 ran the
// two-kernel path at 12 bytes per pixel.  A thread-block CLUSTER of CS CTAs (CS = 2 or 4) has CS x 227 KiB: each CTA takes
// 256 / CS rows of the heatmap (TMA bulk loads into its own shared memory), reads them ONCE into the registers of its
// 2048 / CS threads (32 pixels per thread) taking the maximum on the way, turns them into e = 2^(z log2e - max) there, and
// writes the gradient from there; the CTAs exchange their partial results through DISTRIBUTED SHARED MEMORY (each CTA stores
// into the shared memory of the others, all merge in rank order, so all hold bit-identical totals).  HBM sees every logit
// once: 8 bytes per pixel (4 for bf16); shared memory is written once (by the TMA) and read once per heatmap.
//
//   CS = 2: 1024 threads, one CTA per SM: a heatmap per pair of SMs (round 1).
//   CS = 4:  512 threads, TWO CTAs per SM: four SMs hold two heatmaps in flight, so while one heatmap sits in its
//            reduction / exchange / window phase (latency, nothing issued) the other CTA of the SM sweeps.  The register
//            file holds the same number of heatmaps as before (one per two SMs); the phases interleave.
//
// The shared memory of a CTA is a ring of NS chunk buffers; a part of a heatmap takes NCH of them (NS > NCH), so while
// heatmap k is being processed NS - NCH chunks of heatmap k+1 (and k+2) are already loading, and as soon as k has been read
// into registers its buffers take the next NCH chunks of the stream: the loads of the next heatmaps overlap all the arithmetic.
//
// Same mathematics as head_step2.cuh (src/dsnt/nn.py:25-116,274-298, src/dsnt/model.py:24-63,145):
// column accumulators + one row sum per sweep step give S, S_x, S_y and the variance about a pivot without a second look;
// packed fp32 pairs (FFMA2 / FADD2 / FMUL2).  Regularisers: none, variance, JS and MSE (KL keeps the two-kernel path).
// bf16 heatmaps: e stays fp32 in the registers between the sweeps (no fp16 stash as in the 64x64 kernel).  The
// denominator of masked_average is an input, as for dsnt_head_step.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "capi_util.cuh"
#include "f32x2.cuh"
#include "head_step2.cuh"

namespace cg = cooperative_groups;

namespace dsnt {

constexpr int kPairH = 256, kPairW = 256;
constexpr int kPairWV = kPairW / 4;                             // 64 vectors of four pixels per row
constexpr int kPairIters = 8;                                   // sweep steps per heatmap part: 32 pixels per thread

template <int CS, typename T>
struct PairCfg {
  static constexpr int ES = sizeof(T);
  static constexpr int NT = 2048 / CS;                          // threads per CTA
  static constexpr int NW = NT / 32;
  static constexpr int ROWS = kPairH / CS;                      // rows per CTA
  static constexpr int RPS = NT / kPairWV;                      // rows per sweep step
  static constexpr int STEP_BYTES = RPS * kPairW * ES;          // bytes one sweep step reads
  static constexpr int PART_BYTES = ROWS * kPairW * ES;
  static constexpr int NCH = 8;                                 // chunks per part: one per sweep step
  static constexpr int IPC = kPairIters / NCH;                  // sweep steps per chunk
  static constexpr int CHUNK = PART_BYTES / NCH;
  // ring slots: fp32 CS=2: 13 x 16 KiB = 208 KiB (one CTA per SM); fp32 CS=4: 13 x 8 KiB = 104 KiB (two per SM);
  // bf16 CS=2: 24 x 8 KiB = 192 KiB; bf16 CS=4: 24 x 4 KiB = 96 KiB (two per SM)
  static constexpr int NS = ES == 4 ? 13 : 24;
  static constexpr int SMEM = NS * CHUNK;
  static constexpr int CTAS_PER_SM = CS == 2 ? 1 : 2;
  static_assert(ROWS % RPS == 0 && ROWS / RPS == kPairIters && kPairIters % NCH == 0 && NS > NCH, "geometry");
};

struct PairParams {
  const void* z;
  void* dz;
  const float* target;
  const float* mask;
  const float* denom;
  const float* g_loss;
  float* coords;
  float* stats;
  float* terms;
  long n;
  int flags;
  float sigma, reg_coeff;
  float k2, r2_win;       // Gaussian window (JS / MSE): -0.5 / sigma^2 * log2(e) and the window radius^2 (head_stream.cuh: make_geom)
  int tune;               // DSNT_TUNE_STEP_PAIR_FLAGS (measurements only): 8 = phase trace
};

// DSNT_TUNE_STEP_PAIR_FLAGS & 8 (measurements only): SM clocks thread 0 of CTA 0 spends in each phase of a heatmap, summed over
// its heatmaps; [15] counts the heatmaps
__device__ unsigned long long g_pair_trace[16];

constexpr int kPairSlotsX = 64;       // warps per heatmap (2048 threads): one exchange slot each
constexpr int kPairSlotF = 12;        // floats per slot: eight used, 48 bytes apart (128-bit reads of 8 consecutive slots hit 32 banks)

// What the 64 warps of a heatmap exchange: every warp stores its partial results (relative to ITS OWN maximum, about the
// common pivot) into slot [rank * NW + warp] of every CTA of the cluster; after the wait lane l of every warp holds the
// slots l and l + 32, rescales them like blocks of an online softmax and the lanes add up with the same butterflies in every
// warp of every CTA: bit-identical totals everywhere, no block barrier, no cluster barrier.
struct PairSums {
  float m, S, Sx, Sy, a, b, scme;     // a, b: second moments (variance) | a: sum e^2 (MSE)
};

template <int REG, int CS, typename T>
__global__ void __launch_bounds__(PairCfg<CS, T>::NT, PairCfg<CS, T>::CTAS_PER_SM) head_step_pair_kernel(const PairParams p) {
  using C = PairCfg<CS, T>;
  constexpr int NW = C::NW, NS = C::NS, NCH = C::NCH, RPS = C::RPS, ES = C::ES;
  constexpr bool kVar = REG == DSNT_REG_VAR;
  constexpr bool kJS = REG == DSNT_REG_JS;
  constexpr bool kMSE = REG == DSNT_REG_MSE;
  constexpr bool kWin = kJS || kMSE;
  static_assert(CS * NW == kPairSlotsX, "one slot per warp of the heatmap");
  // JS / MSE: the Gaussian window (16 x 16 pixels at sigma = 1 px) lies in the registers of the few threads that hold its
  // vectors.  Its terms need P = e / S, i.e. the merged sums: they are evaluated AFTER the first exchange and exchanged in a
  // second message; the backward evaluates them again (nothing per-pixel is kept: the 64 registers of a thread hold its 32
  // values of e).  Outside the window the closed forms of head_step2.cuh apply.
  constexpr float tow = 2.0f / kPairW, bw = 1.0f / kPairW - 1.0f, toh = 2.0f / kPairH, bh = 1.0f / kPairH - 1.0f;
  extern __shared__ __align__(128) unsigned char pair_smem[];
  __shared__ __align__(8) unsigned long long bars[NS];
  __shared__ __align__(16) float xin[2][kPairSlotsX][kPairSlotF];  // message number n lands in xin[n & 1]: a warp can be one message ahead
  __shared__ __align__(8) unsigned long long xbar[2];     // ... completing 64 x 32 bytes on xbar[n & 1]
  __shared__ __align__(16) float geo[2][8];               // JS / MSE: window geometry of heatmap k in geo[k & 1] (warp 0, one heatmap ahead)
  __shared__ unsigned read_count;                         // warps that have their part of the current heatmap in registers
#ifdef DSNT_PAIR_TRACE
  __shared__ long long trace_last;
  const bool tracing = (p.tune & 8) && blockIdx.x == 0 && threadIdx.x == 0;
  auto stamp = [&](int phase) {
    if (tracing) {
      const long long t = clock64();
      g_pair_trace[phase] += static_cast<unsigned long long>(t - trace_last);
      trace_last = t;
    }
  };
#else
  auto stamp = [](int) {};
#endif

  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();            // rows rank * ROWS ...
  const int cluster_id = blockIdx.x / CS, n_clusters = gridDim.x / CS;      // (32-bit: n < 2^31, checked by the entry point)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // The exchange: lane r < CS of every warp stores the warp's eight floats into slot [rank * NW + warp] of CTA r with
  // st.async, which completes the bytes on THAT CTA's barrier (the own CTA included); thread 0 arms the local barrier for
  // the 64 x 32 bytes of a message; everybody waits on the local barrier and reads the local xin.  Two barriers in turn:
  // message n+1 of a fast warp must not complete bytes of the phase that still waits for message n of a slow one.
  const uint32_t xin_s = smem_u32(&xin[0][0][0]), xbar_s = smem_u32(&xbar[0]);
  uint32_t xn = 0;          // messages exchanged so far: the same on all warps of the cluster (same branches on bit-identical totals)
  // post message number xn (values identical on every lane of the warp)
  auto post = [&](float a, float b, float c, float d, float e, float f, float g7) {
    const uint32_t par = xn & 1u;
    if (tid == 0) mbar_expect_tx(xbar_s + 8 * par, 32 * kPairSlotsX);
    if (lane < CS) {
      uint32_t dst, bar;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(xin_s + (par * kPairSlotsX + rank * NW + warp) * (4u * kPairSlotF)), "r"(lane));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar) : "r"(xbar_s + 8 * par), "r"(lane));
      asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
                   "f"(a), "f"(b), "f"(c), "f"(d), "r"(bar)
                   : "memory");
      asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst + 16),
                   "f"(e), "f"(f), "f"(g7), "f"(0.f), "r"(bar)
                   : "memory");
    }
  };
  // wait for message number xn of all 64 warps; lane l reads the slots l and l + 32 of what is returned
  auto collect = [&]() -> const float (*)[kPairSlotF] {
    const uint32_t par = xn & 1u;
    mbar_wait(xbar_s + 8 * par, (xn >> 1) & 1u);
    ++xn;
    return xin[par];
  };
  const uint32_t bars_s = smem_u32(&bars[0]), buf_s = smem_u32(pair_smem);
  const long hm_bytes = static_cast<long>(kPairH) * kPairW * ES;
  // heatmaps of this cluster: hm = cluster_id + k * n_clusters, k = 0 .. nk-1; chunk c of heatmap k is chunk number
  // g = NCH k + c of this CTA's stream and lives in ring slot g mod NS on that slot's (g / NS)-th use
  const int nk = p.n > cluster_id ? static_cast<int>((p.n - cluster_id + n_clusters - 1) / n_clusters) : 0;
  auto issue = [&](unsigned g) {
    const int k = g / NCH;
    const unsigned c = g % NCH;
    if (k >= nk) return;
    const long hm = cluster_id + k * n_clusters;
    const unsigned slot = g % NS;
    const char* src = static_cast<const char*>(p.z) + hm * hm_bytes + static_cast<size_t>(rank) * C::PART_BYTES +
                      static_cast<size_t>(c) * C::CHUNK;
    mbar_expect_tx(bars_s + 8 * slot, C::CHUNK);
    bulk_load(buf_s + slot * C::CHUNK, src, C::CHUNK, bars_s + 8 * slot);
  };
  // JS / MSE: where the Gaussian window of a heatmap lies and its normalisation -- functions of the target alone; ONE warp
  // works them out, a heatmap ahead (while it would wait for the exchange), the others read six numbers
  auto geometry = [&](int k) {      // warp 0, all lanes
    if (k >= nk) return;
    const float2 tt = __ldg(reinterpret_cast<const float2*>(p.target) + (cluster_id + k * n_clusters));
    int j_lo, j_hi, i_lo, i_hi;
    axis_window_fast(tt.x, kPairW, 0.5f * kPairW, tow, bw, p.r2_win, j_lo, j_hi);
    axis_window_fast(tt.y, kPairH, 0.5f * kPairH, toh, bh, p.r2_win, i_lo, i_hi);
    float ginv = 0.f, l2ginv = 0.f;
    if (j_lo <= j_hi && i_lo <= i_hi) {
      float sx = 0.f, sy = 0.f;
      for (int j = j_lo + lane; j <= j_hi; j += 32) {
        const float d = fmaf(static_cast<float>(j), tow, bw) - tt.x;
        sx += ex2(p.k2 * d * d);
      }
      for (int i = i_lo + lane; i <= i_hi; i += 32) {
        const float d = fmaf(static_cast<float>(i), toh, bh) - tt.y;
        sy += ex2(p.k2 * d * d);
      }
      const float ks = warp_sum2_transposed(sx, sy, lane);
      sx = __shfl_sync(kFull, ks, 0);
      sy = __shfl_sync(kFull, ks, 16);
      ginv = rcp(sx * sy + kEps);
      l2ginv = lg2(ginv);
    } else {
      j_lo = 1; j_hi = 0; i_lo = 1; i_hi = 0;
    }
    if (lane == 0) {
      float4* gq = reinterpret_cast<float4*>(&geo[k & 1][0]);
      gq[0] = make_float4(ginv, l2ginv, __int_as_float(j_lo), __int_as_float(j_hi));
      gq[1] = make_float4(__int_as_float(i_lo), __int_as_float(i_hi), 0.f, 0.f);
    }
    __syncwarp();
  };
  if (tid == 0) {
    for (int sl = 0; sl < NS; ++sl) mbar_init(bars_s + 8 * sl, 1);
    mbar_init(xbar_s, 1);
    mbar_init(xbar_s + 8, 1);
    read_count = 0;
    for (unsigned g = 0; g < NS; ++g) issue(g);
  }
  if (kWin && warp == 0) geometry(0);
  cluster.sync();       // every CTA's barriers exist before the first remote store (and geo[0] is there)
  const float s2 = p.sigma * p.sigma;

  // thread geometry: vector column cv (4 pixels), rows r0 + RPS * it inside this CTA's part
  const int cv = tid & (kPairWV - 1), r0 = tid >> 6;
  const float xs0 = fmaf(static_cast<float>(cv * 4), tow, bw);      // this thread's pixels: x = xs0 + c * tow
  const float y0 = fmaf(static_cast<float>(rank * C::ROWS + r0), toh, bh);
  constexpr float dyi = RPS * toh;
  const f2 l2e2 = pk1(kLog2e);

#ifdef DSNT_PAIR_TRACE
  if (tracing) trace_last = clock64();
#endif
  for (int k = 0; k < nk; ++k) {
    const int hm = cluster_id + k * n_clusters;
    const unsigned g0 = static_cast<unsigned>(NCH * k);       // chunk number of this heatmap's first chunk
    const float xk = opaque(xs0);
    const float xs[4] = {xk, xk + tow, xk + 2.f * tow, xk + 3.f * tow};
    float tx = 0.f, ty = 0.f;
    if (p.target) {
      const float2 tt = __ldg(reinterpret_cast<const float2*>(p.target) + hm);
      tx = tt.x; ty = tt.y;
    }

    // ---------------------------------------------------------------- the part comes into REGISTERS (32 pixels per thread)
    // and its maximum is taken on the way; shared memory is read exactly once, so its buffers are free for the next loads
    // as soon as every warp is through this sweep
    f2 ev[kPairIters][2];
    float mloc;
    if constexpr (ES == 4) {
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int it = 0; it < kPairIters; ++it) {
        const unsigned g = g0 + it / C::IPC, slot = g % NS;
        if (it % C::IPC == 0) mbar_wait(bars_s + 8 * slot, (g / NS) & 1u);
        const uint4 r = *reinterpret_cast<const uint4*>(pair_smem + slot * C::CHUNK + (it % C::IPC) * C::STEP_BYTES + tid * 16);
        unpack_pairs<float>(r, ev[it]);
        m0 = fmaxf(m0, fmaxf(__uint_as_float(r.x), __uint_as_float(r.y)));
        m1 = fmaxf(m1, fmaxf(__uint_as_float(r.z), __uint_as_float(r.w)));
      }
      mloc = fmaxf(m0, m1);
    } else {
      uint32_t mb = 0xff80ff80u;      // (-inf, -inf)
#pragma unroll
      for (int it = 0; it < kPairIters; ++it) {
        const unsigned g = g0 + it / C::IPC, slot = g % NS;
        if (it % C::IPC == 0) mbar_wait(bars_s + 8 * slot, (g / NS) & 1u);
        const uint2 r = *reinterpret_cast<const uint2*>(pair_smem + slot * C::CHUNK + (it % C::IPC) * C::STEP_BYTES + tid * 8);
        ev[it][0] = bf16x2_to_f2(r.x);
        ev[it][1] = bf16x2_to_f2(r.y);
        mb = max_bf16x2(mb, max_bf16x2(r.x, r.y));
      }
      mloc = fmaxf(bf16lo(mb), bf16hi(mb));
    }
    // Every WARP sums relative to ITS OWN maximum (no block-wide maximum, no block barrier); the 64 warps of the heatmap are
    // merged afterwards like the blocks of an online softmax (S = sum_w S_w 2^(m_w - m)): ONE exchange carries everything.
    mloc = warp_max_redux(mloc);       // (depends on every value the warp has read)
    {
      // The warp's data is in registers.  The LAST warp of the CTA to get here hands this heatmap's buffers to the next NCH
      // chunks of the stream (nothing was written to them by the generic proxy, so no proxy fence); nobody waits.
      unsigned last = 0;
      if (lane == 0) {
        unsigned old;
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(&read_count)), "f"(mloc) : "memory");
        last = (old + 1) % NW == 0 ? 1u : 0u;
      }
      last = __shfl_sync(kFull, last, 0);
      if (last && lane < NCH) issue(g0 + NS + lane);
    }
    stamp(0);                                  // wait for the loads, shared memory -> registers, hand-over of the buffers
    const float m2h = mloc * kLog2e;
    const f2 nm2 = pk1(-m2h);

    // ---------------------------------------------------------------- sums; e replaces z in the registers
    // First and second moments are taken about a PIVOT known before the sweep: the target (a trained model predicts near
    // it; a diffuse heatmap has a variance far larger than any offset).  var = M2_c / S - (mu - c)^2 then loses
    // (mu - c)^2 / var digits: checked below, and the rare ill-conditioned heatmap takes the exact second walk about the mean.
    const float pcx = kVar ? tx : 0.f, pcy = kVar ? ty : 0.f;
    f2 colE[2] = {pk1(0.f), pk1(0.f)};
    f2 tt2 = pk1(0.f);                       // MSE: sum e^2 (outside the window (P - G)^2 = P^2)
    float Syc = 0.f, Syy = 0.f;
    const float dyb = y0 - pcy;
#pragma unroll
    for (int it = 0; it < kPairIters; ++it) {
      ev[it][0] = ex2_2(fma2(ev[it][0], l2e2, nm2));
      ev[it][1] = ex2_2(fma2(ev[it][1], l2e2, nm2));
      colE[0] = add2(colE[0], ev[it][0]);
      colE[1] = add2(colE[1], ev[it][1]);
      const float rs = hsum(add2(ev[it][0], ev[it][1]));
      const float dy = dyb + static_cast<float>(it) * dyi;
      Syc = fmaf(rs, dy, Syc);
      if constexpr (kVar) Syy = fmaf(rs * dy, dy, Syy);
      if constexpr (kMSE) tt2 = fma2(ev[it][0], ev[it][0], fma2(ev[it][1], ev[it][1], tt2));
    }
    PairSums t;
    {
      float c0, c1, c2, c3;
      upk(colE[0], c0, c1);
      upk(colE[1], c2, c3);
      float Sh = (c0 + c1) + (c2 + c3);
      const float dx0 = xs[0] - pcx, dx1 = xs[1] - pcx, dx2 = xs[2] - pcx, dx3 = xs[3] - pcx;
      float Sxh = fmaf(c0, dx0, fmaf(c1, dx1, fmaf(c2, dx2, c3 * dx3)));
      float fourth = 0.f, fifth = 0.f;
      if constexpr (kVar) {
        fourth = fmaf(c0 * dx0, dx0, fmaf(c1 * dx1, dx1, fmaf(c2 * dx2, dx2, c3 * dx3 * dx3)));
        fifth = warp_sum(Syy);
      }
      if constexpr (kMSE) fourth = hsum(tt2);
      const float kq = warp_sum4_transposed(Sh, Sxh, Syc, fourth, lane);
      Sh = __shfl_sync(kFull, kq, 0); Sxh = __shfl_sync(kFull, kq, 8);
      const float Syh = __shfl_sync(kFull, kq, 16);
      fourth = __shfl_sync(kFull, kq, 24);
      stamp(1);                                // exponential sweep + the warp's sums
      post(m2h, Sh, Sxh, Syh, fourth, fifth, 0.f);
    }
    if (kWin && warp == 0) geometry(k + 1);    // ... for the next heatmap, while this one's message is on its way

    // ---------------------------------------------------------------- merge: the same arithmetic in every warp of the cluster
    {
      const float (*xm)[kPairSlotF] = collect();
      float s0, s1, S, Sx, Sy, fourth = 0.f, fifth = 0.f;
      {
        const float4 q0 = *reinterpret_cast<const float4*>(&xm[lane][0]), q1 = *reinterpret_cast<const float4*>(&xm[lane + 32][0]);
        t.m = warp_max_redux(fmaxf(q0.x, q1.x));
        s0 = ex2(q0.x - t.m); s1 = ex2(q1.x - t.m);
        // each warp summed relative to ITS OWN maximum: rescaled like the blocks of an online softmax; everything is about
        // the same pivot, so the parts simply add
        S = fmaf(q0.y, s0, q1.y * s1); Sx = fmaf(q0.z, s0, q1.z * s1); Sy = fmaf(q0.w, s0, q1.w * s1);
      }
      if constexpr (kVar) {
        const float2 qa = *reinterpret_cast<const float2*>(&xm[lane][4]), qb = *reinterpret_cast<const float2*>(&xm[lane + 32][4]);
        fourth = fmaf(qa.x, s0, qb.x * s1);
        fifth = fmaf(qa.y, s0, qb.y * s1);
      }
      if constexpr (kMSE) fourth = fmaf(xm[lane][4], s0 * s0, xm[lane + 32][4] * s1 * s1);      // sum e^2 over all parts
      const float kq = warp_sum4_transposed(S, Sx, Sy, fourth, lane);
      t.S = __shfl_sync(kFull, kq, 0); t.Sx = __shfl_sync(kFull, kq, 8); t.Sy = __shfl_sync(kFull, kq, 16);
      t.a = __shfl_sync(kFull, kq, 24);
      t.b = 0.f;
      if constexpr (kVar) t.b = warp_sum(fifth);
      t.scme = ex2(m2h - t.m);
    }
    stamp(2);                                  // post, wait for the 64 warps, merge
    const float m2 = t.m, S = t.S;
    const float invS = rcp(S);
    const float mxc = t.Sx * invS, myc = t.Sy * invS;
    const float mux = pcx + mxc, muy = pcy + myc;
    const float invSh = t.scme * invS;       // this warp's e (relative to its own maximum) -> probability

    float D = 0.f, creg = 0.f, vx = 0.f, vy = 0.f;
    if constexpr (kVar) {
      vx = t.a * invS - mxc * mxc;
      vy = t.b * invS - myc * myc;
      // conditioning of the pivot form: (mu - c)^2 <= 16 var keeps the cancellation below 17 fp32 roundings (1e-6 relative)
      const bool ill = !(mxc * mxc <= 16.f * vx) || !(myc * myc <= 16.f * vy);
      if (ill) {
        // exact: second moments about the mean itself, from the registers (column sums are still there, the row sums are
        // formed again) and one more exchange; all warps get here together (identical totals)
        f2 ce[2] = {pk1(0.f), pk1(0.f)};
        float aye = 0.f;
        const float eyb = y0 - muy;
#pragma unroll
        for (int it = 0; it < kPairIters; ++it) {
          const float d = eyb + static_cast<float>(it) * dyi;
          ce[0] = add2(ce[0], ev[it][0]);
          ce[1] = add2(ce[1], ev[it][1]);
          aye = fmaf(hsum(add2(ev[it][0], ev[it][1])) * d, d, aye);
        }
        float c0, c1, c2, c3;
        upk(ce[0], c0, c1);
        upk(ce[1], c2, c3);
        const float ex0 = xs[0] - mux, ex1 = xs[1] - mux, ex2v = xs[2] - mux, ex3 = xs[3] - mux;
        float axe = fmaf(c0 * ex0, ex0, fmaf(c1 * ex1, ex1, fmaf(c2 * ex2v, ex2v, c3 * ex3 * ex3)));
        const float kq = warp_sum2_transposed(axe, aye, lane);
        post(m2h, __shfl_sync(kFull, kq, 0), __shfl_sync(kFull, kq, 16), 0.f, 0.f, 0.f, 0.f);
        const float (*xe)[kPairSlotF] = collect();
        const float4 lo0 = *reinterpret_cast<const float4*>(&xe[lane][0]), lo1 = *reinterpret_cast<const float4*>(&xe[lane + 32][0]);
        const float s0 = ex2(lo0.x - m2), s1 = ex2(lo1.x - m2);
        const float kr = warp_sum2_transposed(fmaf(lo0.y, s0, lo1.y * s1), fmaf(lo0.z, s0, lo1.z * s1), lane);
        vx = __shfl_sync(kFull, kr, 0) * invS;
        vy = __shfl_sync(kFull, kr, 16) * invS;
      }
      const float ex = vx - s2, ey = vy - s2;
      D = ex * ex + ey * ey;
      creg = 2.f * (ex * vx + ey * vy);
    }

    // ---------------------------------------------------------------- JS / MSE: the Gaussian window, from the registers
    float ginv = 0.f, l2ginv = 0.f;
    int wi_lo = 1, wi_hi = 0;
    bool colin = false;
    if constexpr (kWin) {
      const float4* gq = reinterpret_cast<const float4*>(&geo[k & 1][0]);      // (written before warp 0 posted this heatmap's message)
      const float4 ga = gq[0], gb = gq[1];
      ginv = ga.x; l2ginv = ga.y;
      wi_lo = __float_as_int(gb.x); wi_hi = __float_as_int(gb.y);
      colin = cv >= (__float_as_int(ga.z) >> 2) && cv <= (__float_as_int(ga.w) >> 2) && __float_as_int(ga.z) <= __float_as_int(ga.w);
    }
    const int rowb = static_cast<int>(rank) * C::ROWS + r0;       // this thread's rows: rowb + RPS it
    const f2 k2p = pk1(p.k2), eps2 = pk1(kEps), half2 = pk1(0.5f), invSh2 = pk1(invSh);
    const f2 wdx[2] = {pk(xs[0] - tx, xs[1] - tx), pk(xs[2] - tx, xs[3] - tx)};
    // the window terms of one pair of pixels: P, and d = log2(P + eps) - 1 - log2(M + eps) (JS) or G (MSE); JS also lgG - L
    auto win_pair = [&](int it, int c, float dyw, f2& P, f2& d, f2& gl) {
      const f2 lgG = fma2(mul2(wdx[c], k2p), wdx[c], pk1(fmaf(p.k2 * dyw, dyw, l2ginv)));
      const f2 G = ex2_2(lgG);
      P = mul2(ev[it][c], invSh2);
      if constexpr (kJS) {
        const f2 L = lg2_2(fma2(half2, P, fma2(half2, G, eps2)));
        d = sub2(sub2(lg2_2(add2(P, eps2)), pk1(1.0f)), L);
        gl = mul2(G, sub2(lgG, L));
      } else {
        d = G;
        gl = G;
      }
    };
    if constexpr (kWin) {
      f2 qa = pk1(0.f), qb = pk1(0.f), qc = pk1(0.f);
      if (colin) {
#pragma unroll
        for (int it = 0; it < kPairIters; ++it) {
          const int row = rowb + RPS * it;
          if (row >= wi_lo && row <= wi_hi) {
            const float dyw = (y0 + static_cast<float>(it) * dyi) - ty;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              f2 P, d, gl;
              win_pair(it, c, dyw, P, d, gl);
              if constexpr (kJS) {
                qa = fma2(P, d, qa);
                qb = add2(qb, gl);
              } else {
                const f2 df = sub2(P, d);
                qa = fma2(df, df, qa);
                qb = fma2(P, P, qb);
                qc = fma2(P, df, qc);
              }
            }
          }
        }
      }
      float a0 = hsum(qa), a1 = hsum(qb), a2 = hsum(qc);
      if (__any_sync(kFull, colin)) {          // (most warps hold no window pixel: they post zeros at once)
        const float kq = warp_sum4_transposed(a0, a1, a2, 0.f, lane);
        a0 = __shfl_sync(kFull, kq, 0); a1 = __shfl_sync(kFull, kq, 8); a2 = __shfl_sync(kFull, kq, 16);
      }
      stamp(3);                                // window terms
      post(a0, a1, a2, 0.f, 0.f, 0.f, 0.f);
      const float (*xw)[kPairSlotF] = collect();
      const float4 lo0 = *reinterpret_cast<const float4*>(&xw[lane][0]), lo1 = *reinterpret_cast<const float4*>(&xw[lane + 32][0]);
      const float kq = warp_sum4_transposed(lo0.x + lo1.x, lo0.y + lo1.y, lo0.z + lo1.z, 0.f, lane);
      const float w0 = __shfl_sync(kFull, kq, 0), w1 = __shfl_sync(kFull, kq, 8), w2 = __shfl_sync(kFull, kq, 16);
      if constexpr (kMSE) {
        const float outside = fmaxf(fmaf(t.a * invS, invS, -w1), 0.f);
        D = outside + w0;
        creg = 2.f * (outside + w2);
      } else {
        creg = 0.5f * kLn2 * (1.0f + w0);
        D = fmaf(0.5f * kLn2, w1, creg);
      }
      stamp(4);                                // second exchange
    }

    // ---------------------------------------------------------------- outputs + the scalars of the backward
    float dist = 0.f, a = 0.f, b = 0.f;
    // (per-launch scalars are read where they are used: no register holds them through the sweeps)
    const float gl = p.g_loss ? __ldg(p.g_loss) : 1.0f;
    const float wgt = (p.mask ? __ldg(p.mask + hm) : 1.0f) * (1.0f / __ldg(p.denom));
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
    if (rank == 0 && tid == 0) {
      reinterpret_cast<float2*>(p.coords)[hm] = make_float2(mux, muy);
      if (p.stats) {
        float4* st = reinterpret_cast<float4*>(p.stats + static_cast<long>(hm) * kStatsK);
        st[0] = make_float4(m2, invS, mux, muy);
        st[1] = make_float4(vx, vy, creg, ginv);
      }
      if (p.terms) reinterpret_cast<float2*>(p.terms)[hm] = make_float2(dist, D);
    }
    float cbase = -fmaf(a, mux, fmaf(b, muy, rho * creg));
    if (kJS) cbase = fmaf(0.5f * kLn2, rho, cbase);

    // ---------------------------------------------------------------- backward: dz = e * (A_col + R_row) / S
    {
      f2 acol[2];
      const float kx = kVar ? rho * 2.f * (vx - s2) : 0.f;
      float av[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v0 = a * xs[c];
        if (kVar) { const float d = xs[c] - mux; v0 = fmaf(kx * d, d, v0); }
        av[c] = v0 * invSh;
      }
      acol[0] = pk(av[0], av[1]);
      acol[1] = pk(av[2], av[3]);
      const float bS = b * invSh, cbS = cbase * invSh;
      const float kyS = kVar ? rho * 2.f * (vy - s2) * invSh : 0.f;
      const f2 rpS = pk1(kMSE ? 2.f * rho * invSh * invSh : 0.f);
      const f2 kwS = pk1(kJS ? 0.5f * kLn2 * rho * invSh : (kMSE ? -2.f * rho * invSh : 0.f));   // JS: rho (ln2/2) d;  MSE: -2 rho G
      char* dzb = static_cast<char*>(p.dz) + static_cast<long>(hm) * hm_bytes + static_cast<size_t>(rank) * C::PART_BYTES + tid * (4 * ES);
#pragma unroll
      for (int it = 0; it < kPairIters; ++it) {
        f2 o[2];
        const float y = y0 + static_cast<float>(it) * dyi;
        float rc = fmaf(bS, y, cbS);
        if (kVar) { const float d = y - muy; rc = fmaf(kyS * d, d, rc); }
        const f2 rc2 = pk1(rc);
        f2 g0v = add2(acol[0], rc2), g1v = add2(acol[1], rc2);
        if constexpr (kMSE) { g0v = fma2(rpS, ev[it][0], g0v); g1v = fma2(rpS, ev[it][1], g1v); }      // 2 rho P
        if constexpr (kWin) {
          const int row = rowb + RPS * it;
          if (colin && row >= wi_lo && row <= wi_hi) {                 // window pixels: the G-dependent term, evaluated again
            f2 P, d, glw;
            win_pair(it, 0, y - ty, P, d, glw);
            g0v = fma2(kwS, d, g0v);
            win_pair(it, 1, y - ty, P, d, glw);
            g1v = fma2(kwS, d, g1v);
          }
        }
        o[0] = mul2(ev[it][0], g0v);
        o[1] = mul2(ev[it][1], g1v);
        if constexpr (ES == 4) {
          *reinterpret_cast<uint4*>(dzb + it * C::STEP_BYTES) = pack_pairs<float>(o);
        } else {
          float l0, h0, l1, h1;
          upk(o[0], l0, h0);
          upk(o[1], l1, h1);
          *reinterpret_cast<uint2*>(dzb + it * C::STEP_BYTES) = make_uint2(pack_bf16(l0, h0), pack_bf16(l1, h1));
        }
      }
    }
    stamp(5);                                  // outputs + backward sweep (stores issued)
#ifdef DSNT_PAIR_TRACE
    if (tracing) g_pair_trace[15] += 1;
#endif
  }
  cluster.sync();      // no CTA leaves while another may still store into its shared memory
}

static int pair_enabled() {
  static const int v = [] { const char* e = std::getenv("DSNT_TUNE_STEP_PAIR"); return e ? std::atoi(e) : 1; }();
  return v;
}

// cluster size: DSNT_TUNE_STEP_PAIR_CS = 2 | 4 overrides the measured default (profiles/r02_v6_kbench_cfg5.txt); read on every
// call (a launch of 0.3 ms and more), so that the tests can take both forms in one process
static int pair_cluster_size(int dtype, int reg) {
  const char* e = std::getenv("DSNT_TUNE_STEP_PAIR_CS");
  const int forced = e ? std::atoi(e) : 0;
  if (forced == 2 || forced == 4) return forced;
  (void)dtype; (void)reg;
  return 4;
}

bool step_pair_supported(int dtype, int H, int W, int reg) {
  static const int win = [] { const char* e = std::getenv("DSNT_TUNE_STEP_PAIR_WIN"); return e ? std::atoi(e) : 1; }();
  return pair_enabled() && (dtype == DSNT_DTYPE_F32 || dtype == DSNT_DTYPE_BF16) && H == kPairH && W == kPairW &&
         (reg == DSNT_REG_NONE || reg == DSNT_REG_VAR || (win && (reg == DSNT_REG_JS || reg == DSNT_REG_MSE)));
}

// returns 1 when clusters of this shape cannot run on the device
template <int REG, int CS, typename T>
static int launch_pair(const PairParams& p, cudaStream_t stream) {
  using C = PairCfg<CS, T>;
  auto kern = head_step_pair_kernel<REG, CS, T>;
  static int max_clusters_of[kMaxDevices] = {};     // per device: 0 = not asked yet, -1 = clusters of this size cannot run
  int& max_clusters = max_clusters_of[current_device()];
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  if (max_clusters == 0) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess)
      return check_launch("head_step_pair_kernel (shared-memory opt-in)");
    cudaLaunchConfig_t probe = {};
    probe.gridDim = dim3(CS * C::CTAS_PER_SM * sm_count_of_current_device()); probe.blockDim = dim3(C::NT);
    probe.dynamicSmemBytes = C::SMEM;
    probe.attrs = at; probe.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, kern, &probe) != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = -1; }
    max_clusters = nc;
    if (std::getenv("DSNT_TUNE_STEP_PAIR_VERBOSE"))
      fprintf(stderr, "head_step_pair_kernel<reg %d, cluster %d, %d B>: %d co-resident clusters\n", REG, CS, C::ES, nc);
  }
  if (max_clusters < 0) return 1;      // clusters of this size cannot run here
  static const int cap = [] { const char* e = std::getenv("DSNT_TUNE_STEP_PAIR_CLUSTERS"); return e ? std::atoi(e) : 0; }();
  long clusters = p.n < max_clusters ? p.n : max_clusters;
  if (cap > 0 && clusters > cap) clusters = cap;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(CS * clusters)); cfg.blockDim = dim3(C::NT);
  cfg.dynamicSmemBytes = C::SMEM; cfg.stream = stream;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (p.tune & 8) {
    unsigned long long zero[16] = {};
    cudaMemcpyToSymbol(g_pair_trace, zero, sizeof(zero));
  }
  if (cudaLaunchKernelEx(&cfg, kern, p) != cudaSuccess) {
    // no room for such a cluster on this device / partition: not an error, the caller takes another kernel
    cudaGetLastError();
    return 1;
  }
  if (p.tune & 8) {
    unsigned long long t[16];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(t, g_pair_trace, sizeof(t));
    const double nh = t[15] ? static_cast<double>(t[15]) : 1.0;
    double tot = 0;
    for (int i = 0; i < 10; ++i) tot += t[i] / nh;
    fprintf(stderr, "pair trace <reg %d, cluster %d, %d B> clocks per heatmap (CTA 0, %llu heatmaps, %.0f in all): load+lds %.0f | "
            "exp sweep + warp sums %.0f | exchange + merge %.0f | window %.0f | exchange 2 %.0f | backward %.0f\n", REG, CS, C::ES, t[15], tot,
            t[0] / nh, t[1] / nh, t[2] / nh, t[3] / nh, t[4] / nh, t[5] / nh);
  }
  return check_launch("head_step_pair_kernel");
}

template <int CS, typename T>
static int launch_pair_reg(const PairParams& p, int reg, cudaStream_t stream) {
  switch (reg) {
    case DSNT_REG_VAR: return launch_pair<DSNT_REG_VAR, CS, T>(p, stream);
    case DSNT_REG_JS: return launch_pair<DSNT_REG_JS, CS, T>(p, stream);
    case DSNT_REG_MSE: return launch_pair<DSNT_REG_MSE, CS, T>(p, stream);
    default: return launch_pair<DSNT_REG_NONE, CS, T>(p, stream);
  }
}

// returns 1 when the case is not served
int launch_step_pair(const void* z, int dtype, long n, int H, int W, const float* target, const float* mask,
                     const float* denom, const float* g_loss, float reg_coeff, int reg, float sigma, int flags,
                     float* coords, float* stats, float* terms, void* dz, cudaStream_t stream) {
  if (!step_pair_supported(dtype, H, W, reg) || !denom) return 1;
  PairParams p;
  p.z = z; p.dz = dz; p.target = target; p.mask = mask; p.denom = denom;
  p.g_loss = g_loss; p.coords = coords; p.stats = stats; p.terms = terms; p.n = n; p.flags = flags; p.sigma = sigma;
  p.reg_coeff = reg_coeff;
  const Geom g = make_geom(H, W, 4, 32, sigma > 0.f ? sigma : 1.f, reg);
  p.k2 = g.k2; p.r2_win = g.r2_win;
  { const char* e = std::getenv("DSNT_TUNE_STEP_PAIR_FLAGS"); p.tune = e ? std::atoi(e) : 0; }
  const bool f32 = dtype == DSNT_DTYPE_F32;
  int rc = 1;
  if (pair_cluster_size(dtype, reg) == 4)
    rc = f32 ? launch_pair_reg<4, float>(p, reg, stream) : launch_pair_reg<4, __nv_bfloat16>(p, reg, stream);
  if (rc == 1)      // (also the fall-back of a device on which clusters of four do not fit)
    rc = f32 ? launch_pair_reg<2, float>(p, reg, stream) : launch_pair_reg<2, __nv_bfloat16>(p, reg, stream);
  return rc;
}

}  // namespace dsnt
