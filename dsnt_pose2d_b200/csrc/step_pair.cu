// step_pair.cu -- the one-pass training step for 256x256 fp32 heatmaps (BASELINE config 5) on a PAIR of SMs.
//
// A 256x256 fp32 heatmap is 256 KiB: more than the 227 KiB of shared memory one CTA can have, which is why config 5 ran the
// two-kernel path at 12 bytes per pixel.  A thread-block CLUSTER of two CTAs has 2 x 227 KiB: each CTA of the pair takes
// one half of the heatmap (128 rows, 128 KiB, TMA bulk loads into its own shared memory), reads it ONCE into the registers
// of its 1024 threads (32 per thread) taking the maximum on the way, turns it into e = 2^(z log2e - max) there, and writes
// the gradient from there; the pair exchanges its per-half partial results through DISTRIBUTED SHARED MEMORY (each CTA
// stores into the other's shared memory, both merge in rank order, so both hold bit-identical totals).  HBM sees every
// logit once: 8 bytes per pixel; shared memory is written once (by the TMA) and read once per heatmap.
//
// The shared memory of a CTA is a ring of seven 32 KiB chunk buffers; a half heatmap takes four, so while heatmap k is
// being processed three chunks of heatmap k+1 are already loading, and as soon as k has been read into registers its four
// buffers take the last chunk of k+1 and the first three of k+2: the loads of the next heatmaps overlap all the arithmetic.
//
// Same mathematics as head_step2.cuh (src/dsnt/nn.py:25-116,274-298, src/dsnt/model.py:24-63,145):
// column accumulators + one row sum per sweep step give S, S_x, S_y and the variance about the mean without a second look;
// packed fp32 pairs (FFMA2 / FADD2 / FMUL2).  Regularisers: none and variance (what config 5 uses); the others keep the
// two-kernel path.  The denominator of masked_average is an input, as for dsnt_head_step.
#include <cooperative_groups.h>

#include <cstdlib>

#include "capi_util.cuh"
#include "f32x2.cuh"
#include "head_step2.cuh"

namespace cg = cooperative_groups;

namespace dsnt {

constexpr int kPairH = 256, kPairW = 256;
constexpr int kPairThreads = 1024;
constexpr int kPairHalfRows = kPairH / 2;                       // rows per CTA
constexpr int kPairHalfBytes = kPairHalfRows * kPairW * 4;      // 128 KiB
constexpr int kPairWV = kPairW / 4;                             // 64 vectors per row
constexpr int kPairRowsPerStep = kPairThreads / kPairWV;        // 16 rows per sweep step
constexpr int kPairIters = kPairHalfRows / kPairRowsPerStep;    // 8 sweep steps
constexpr int kPairChunks = 4;                                  // a half heatmap = 4 chunks of 32 KiB
constexpr int kPairChunkBytes = kPairHalfBytes / kPairChunks;   // 32 KiB = two sweep steps
constexpr int kPairSlots = 7;                                   // ring of chunk buffers: 224 KiB of the 227 KiB a CTA may have
constexpr int kPairSmemBytes = kPairSlots * kPairChunkBytes;

struct PairParams {
  const float* z;
  float* dz;
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
};

// sum over the 1024 threads of one CTA of up to four values, identical on every thread, fixed order
__device__ __forceinline__ void pair_block_sum4(float& a, float& b, float& c, float& d, float (*red)[8], int warp, int lane) {
  const float k = warp_sum4_transposed(a, b, c, d, lane);
  if ((lane & 7) == 0) red[warp][lane >> 3] = k;
  __syncthreads();
  // 32 warps = 32 lanes: every warp adds the per-warp partials with the same butterfly
  const float k2 = warp_sum4_transposed(red[lane][0], red[lane][1], red[lane][2], red[lane][3], lane);
  a = __shfl_sync(kFull, k2, 0); b = __shfl_sync(kFull, k2, 8); c = __shfl_sync(kFull, k2, 16); d = __shfl_sync(kFull, k2, 24);
  __syncthreads();      // red may be reused
}

// ... and of six (the variance path: S, the first and the second moments about the pivot)
__device__ __forceinline__ void pair_block_sum6(float& a, float& b, float& c, float& d, float& e, float& f, float (*red)[8],
                                                int warp, int lane) {
  const float k = warp_sum4_transposed(a, b, c, d, lane);
  const float k2 = warp_sum2_transposed(e, f, lane);
  if ((lane & 7) == 0) red[warp][lane >> 3] = k;
  if ((lane & 15) == 0) red[warp][4 + (lane >> 4)] = k2;
  __syncthreads();
  const float q = warp_sum4_transposed(red[lane][0], red[lane][1], red[lane][2], red[lane][3], lane);
  const float q2 = warp_sum2_transposed(red[lane][4], red[lane][5], lane);
  a = __shfl_sync(kFull, q, 0); b = __shfl_sync(kFull, q, 8); c = __shfl_sync(kFull, q, 16); d = __shfl_sync(kFull, q, 24);
  e = __shfl_sync(kFull, q2, 0); f = __shfl_sync(kFull, q2, 16);
  __syncthreads();      // red may be reused
}

template <int REG>
__global__ void __launch_bounds__(kPairThreads, 1) head_step_pair_kernel(const PairParams p) {
  constexpr bool kVar = REG == DSNT_REG_VAR;
  constexpr bool kJS = REG == DSNT_REG_JS;
  constexpr bool kMSE = REG == DSNT_REG_MSE;
  constexpr bool kWin = kJS || kMSE;
  // JS / MSE: the Gaussian window (16 x 16 pixels at sigma = 1 px) lies in the registers of the <= 80 threads that hold its
  // vectors.  Its terms need P = e / S, i.e. the merged sums: they are evaluated AFTER the first exchange, block-reduced
  // and exchanged in a second message; the backward evaluates them again (nothing per-pixel is kept: the 64 registers of a
  // thread hold its 32 values of e).  Outside the window the closed forms of head_step2.cuh apply.
  constexpr float tow = 2.0f / kPairW, bw = 1.0f / kPairW - 1.0f, toh = 2.0f / kPairH, bh = 1.0f / kPairH - 1.0f;
  extern __shared__ __align__(128) unsigned char pair_smem[];
  __shared__ __align__(8) unsigned long long bars[kPairSlots];
  __shared__ float red[32][8];
  __shared__ __align__(16) float xin[2][2][4];            // the PEER's partial results, stored here by the peer (st.async over DSMEM);
                                                          // message number n lands in xin[n & 1]: the peer can be one message ahead
  __shared__ __align__(8) unsigned long long xbar[1];     // ... the two stores completing 32 bytes on this mbarrier

  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();            // 0: rows 0..127, 1: rows 128..255
  const unsigned peer = rank ^ 1u;
  const long cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // The exchange of a heatmap: thread 0 arms its own xbar for 32 bytes and stores its partial results into the peer's xin
  // with st.async, which completes the bytes on the PEER's xbar; everybody then waits on the local barrier and reads the
  // local xin.  No cluster-wide barrier (whose release fence makes all 1024 threads wait for their global stores: the
  // first version of this kernel, 1053 us at config 5, against 878 us with three such exchanges and less with one).
  const uint32_t xin_s = smem_u32(&xin[0][0][0]), xbar_s = smem_u32(&xbar[0]);
  uint32_t peer_xin_s, peer_xbar_s;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_xin_s) : "r"(xin_s), "r"(peer));
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_xbar_s) : "r"(xbar_s), "r"(peer));
  uint32_t xn = 0;          // messages exchanged so far: the same on both CTAs (they take the same branches on bit-identical totals)
  auto send = [&](float a, float b, float c, float d, float e, float f, float g7 = 0.f) {      // thread 0; message number xn
    const uint32_t dst = peer_xin_s + (xn & 1u) * 32u;
    mbar_expect_tx(xbar_s, 32);
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
                 "f"(a), "f"(b), "f"(c), "f"(d), "r"(peer_xbar_s)
                 : "memory");
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst + 16),
                 "f"(e), "f"(f), "f"(g7), "f"(0.f), "r"(peer_xbar_s)
                 : "memory");
  };
  const uint32_t bars_s = smem_u32(&bars[0]), buf_s = smem_u32(pair_smem);
  const long hm_floats = static_cast<long>(kPairH) * kPairW;
  // heatmaps of this cluster: hm = cluster_id + k * n_clusters, k = 0 .. nk-1; chunk c of heatmap k is chunk number
  // g = 4k + c of this CTA's stream and lives in ring slot g mod 7 on that slot's (g / 7)-th use
  const long nk = p.n > cluster_id ? (p.n - cluster_id + n_clusters - 1) / n_clusters : 0;
  auto issue = [&](long k, int c) {      // thread 0
    if (k >= nk) return;
    const long hm = cluster_id + k * n_clusters;
    const unsigned g = static_cast<unsigned>(4 * k + c), slot = g % kPairSlots;
    const char* src = reinterpret_cast<const char*>(p.z + hm * hm_floats) + static_cast<size_t>(rank) * kPairHalfBytes +
                      static_cast<size_t>(c) * kPairChunkBytes;
    mbar_expect_tx(bars_s + 8 * slot, kPairChunkBytes);
    bulk_load(buf_s + slot * kPairChunkBytes, src, kPairChunkBytes, bars_s + 8 * slot);
  };
  if (tid == 0) {
    for (int sl = 0; sl < kPairSlots; ++sl) mbar_init(bars_s + 8 * sl, 1);
    mbar_init(xbar_s, 1);
    for (int c = 0; c < 4; ++c) issue(0, c);
    for (int c = 0; c < 3; ++c) issue(1, c);
  }
  cluster.sync();       // both CTAs' barriers exist before the first remote store
  const float gl = p.g_loss ? __ldg(p.g_loss) : 1.0f;
  const float inv_denom = 1.0f / __ldg(p.denom);
  const float s2 = p.sigma * p.sigma;

  // thread geometry: vector column cv (4 pixels), rows row_base + 16 * it inside this CTA's half
  const int cv = tid & (kPairWV - 1), r0 = tid >> 6;
  float xs[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) xs[c] = fmaf(static_cast<float>(cv * 4 + c), tow, bw);
  const float y0 = fmaf(static_cast<float>(rank * kPairHalfRows + r0), toh, bh);
  constexpr float dyi = kPairRowsPerStep * toh;
  const f2 l2e2 = pk1(kLog2e);

  for (long k = 0; k < nk; ++k) {
    const long hm = cluster_id + k * n_clusters;
    const unsigned g0 = static_cast<unsigned>(4 * k);       // chunk number of this heatmap's first chunk
    float tx = 0.f, ty = 0.f;
    if (p.target) {
      const float2 tt = __ldg(reinterpret_cast<const float2*>(p.target) + hm);
      tx = tt.x; ty = tt.y;
    }
    const float wgt = (p.mask ? __ldg(p.mask + hm) : 1.0f) * inv_denom;

    // ---------------------------------------------------------------- the half heatmap comes into REGISTERS (32 per thread)
    // and its maximum is taken on the way; shared memory is read exactly once, so its buffers are free for the next loads
    // as soon as this sweep is over
    f2 ev[kPairIters][2];
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int it = 0; it < kPairIters; ++it) {
      const unsigned g = g0 + it / 2, slot = g % kPairSlots;
      if ((it & 1) == 0) mbar_wait(bars_s + 8 * slot, (g / kPairSlots) & 1u);
      const uint4 r = *reinterpret_cast<const uint4*>(pair_smem + slot * kPairChunkBytes + (it & 1) * (kPairChunkBytes / 2) + tid * 16);
      unpack_pairs<float>(r, ev[it]);
      m0 = fmaxf(m0, fmaxf(__uint_as_float(r.x), __uint_as_float(r.y)));
      m1 = fmaxf(m1, fmaxf(__uint_as_float(r.z), __uint_as_float(r.w)));
    }
    float mloc = warp_max_redux(fmaxf(m0, m1));
    if (lane == 0) red[warp][0] = mloc;
    __syncthreads();
    // every thread has its data: this heatmap's four buffers take the last chunk of the next heatmap and the first three of
    // the one after it (nothing was written to them by the generic proxy, so no proxy fence)
    if (tid == 0) {
      issue(k + 1, 3);
      for (int c = 0; c < 3; ++c) issue(k + 2, c);
    }
    mloc = warp_max_redux(red[lane][0]);
    __syncthreads();                                  // red is free again
    // Each half is summed relative to ITS OWN maximum; the halves are merged afterwards like two blocks of an online
    // softmax (S = S_0 2^(m_0 - m) + S_1 2^(m_1 - m)), so ONE exchange per heatmap carries everything.
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
    float c0, c1, c2, c3;
    upk(colE[0], c0, c1);
    upk(colE[1], c2, c3);
    float Sh = (c0 + c1) + (c2 + c3);
    const float dx0 = xs[0] - pcx, dx1 = xs[1] - pcx, dx2 = xs[2] - pcx, dx3 = xs[3] - pcx;
    float Sxh = fmaf(c0, dx0, fmaf(c1, dx1, fmaf(c2, dx2, c3 * dx3)));
    float Syh = Syc;
    float axh = 0.f, ayh = 0.f;
    float Tth = 0.f;
    if constexpr (!kVar) {
      Tth = hsum(tt2);
      pair_block_sum4(Sh, Sxh, Syh, Tth, red, warp, lane);
    } else {
      axh = fmaf(c0 * dx0, dx0, fmaf(c1 * dx1, dx1, fmaf(c2 * dx2, dx2, c3 * dx3 * dx3)));
      ayh = Syy;
      pair_block_sum6(Sh, Sxh, Syh, axh, ayh, Syy, red, warp, lane);      // (the sixth value rides along unused)
    }
    if (tid == 0) send(m2h, Sh, Sxh, Syh, axh, ayh, Tth);
    mbar_wait(xbar_s, xn & 1u);
    // merge in rank order on both CTAs: bit-identical totals
    const bool first = rank == 0;
    const float (*xm)[4] = xin[xn & 1u];
    ++xn;
    const float p_m = xm[0][0], p_S = xm[0][1], p_Sx = xm[0][2], p_Sy = xm[0][3], p_ax = xm[1][0], p_ay = xm[1][1], p_Tt = xm[1][2];
    const float h_m[2] = {first ? m2h : p_m, first ? p_m : m2h};
    const float h_S[2] = {first ? Sh : p_S, first ? p_S : Sh};
    const float h_Sx[2] = {first ? Sxh : p_Sx, first ? p_Sx : Sxh};
    const float h_Sy[2] = {first ? Syh : p_Sy, first ? p_Sy : Syh};
    const float h_ax[2] = {first ? axh : p_ax, first ? p_ax : axh};
    const float h_ay[2] = {first ? ayh : p_ay, first ? p_ay : ayh};
    const float m2 = fmaxf(h_m[0], h_m[1]);
    const float sc0 = ex2(h_m[0] - m2), sc1 = ex2(h_m[1] - m2);
    // each half was summed relative to ITS OWN maximum: rescaled like two blocks of an online softmax; everything is about
    // the same pivot, so the halves simply add
    const float S = fmaf(h_S[0], sc0, h_S[1] * sc1);
    const float Sxc = fmaf(h_Sx[0], sc0, h_Sx[1] * sc1), Sycm = fmaf(h_Sy[0], sc0, h_Sy[1] * sc1);
    const float invS = rcp(S);
    const float mxc = Sxc * invS, myc = Sycm * invS;
    const float mux = pcx + mxc, muy = pcy + myc;
    const float invSh = (first ? sc0 : sc1) * invS;       // this half's e (relative to its own maximum) -> probability

    float D = 0.f, creg = 0.f, vx = 0.f, vy = 0.f;
    if constexpr (kVar) {
      vx = fmaf(h_ax[0], sc0, h_ax[1] * sc1) * invS - mxc * mxc;
      vy = fmaf(h_ay[0], sc0, h_ay[1] * sc1) * invS - myc * myc;
      // conditioning of the pivot form: (mu - c)^2 <= 16 var keeps the cancellation below 17 fp32 roundings (1e-6 relative)
      const bool ill = !(mxc * mxc <= 16.f * vx) || !(myc * myc <= 16.f * vy);
      if (ill) {
        // exact: second moments about the mean itself, from the registers (column sums are still there, the row sums are
        // formed again), one more block reduction and one more exchange; both CTAs get here together (identical totals)
        const float ex0 = xs[0] - mux, ex1 = xs[1] - mux, ex2v = xs[2] - mux, ex3 = xs[3] - mux;
        float axe = fmaf(c0 * ex0, ex0, fmaf(c1 * ex1, ex1, fmaf(c2 * ex2v, ex2v, c3 * ex3 * ex3)));
        float aye = 0.f;
        const float eyb = y0 - muy;
#pragma unroll
        for (int it = 0; it < kPairIters; ++it) {
          const float d = eyb + static_cast<float>(it) * dyi;
          aye = fmaf(hsum(add2(ev[it][0], ev[it][1])) * d, d, aye);
        }
        float z0 = 0.f, z1 = 0.f;
        pair_block_sum4(axe, aye, z0, z1, red, warp, lane);
        if (tid == 0) send(axe, aye, 0.f, 0.f, 0.f, 0.f);
        mbar_wait(xbar_s, xn & 1u);
        const float (*xe)[4] = xin[xn & 1u];
        ++xn;
        const float q_ax = xe[0][0], q_ay = xe[0][1];
        vx = fmaf(first ? axe : q_ax, sc0, (first ? q_ax : axe) * sc1) * invS;
        vy = fmaf(first ? aye : q_ay, sc0, (first ? q_ay : aye) * sc1) * invS;
      }
      const float ex = vx - s2, ey = vy - s2;
      D = ex * ex + ey * ey;
      creg = 2.f * (ex * vx + ey * vy);
    }

    // ---------------------------------------------------------------- JS / MSE: the Gaussian window, from the registers
    float ginv = 0.f, l2ginv = 0.f;
    int wi_lo = 0, wi_hi = -1;
    bool colin = false;
    const int rowb = static_cast<int>(rank) * kPairHalfRows + r0;       // this thread's rows: rowb + 16 it
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
      int j_lo, j_hi;
      axis_window_fast(tx, kPairW, 0.5f * kPairW, tow, bw, p.r2_win, j_lo, j_hi);
      axis_window_fast(ty, kPairH, 0.5f * kPairH, toh, bh, p.r2_win, wi_lo, wi_hi);
      const bool has = j_lo <= j_hi && wi_lo <= wi_hi;
      f2 qa = pk1(0.f), qb = pk1(0.f), qc = pk1(0.f);
      if (has) {
        float sx = 0.f, sy = 0.f;                    // every warp for itself: 2 x <= 17 exponentials, no block barrier
        for (int j = j_lo + lane; j <= j_hi; j += 32) {
          const float d = fmaf(static_cast<float>(j), tow, bw) - tx;
          sx += ex2(p.k2 * d * d);
        }
        for (int i = wi_lo + lane; i <= wi_hi; i += 32) {
          const float d = fmaf(static_cast<float>(i), toh, bh) - ty;
          sy += ex2(p.k2 * d * d);
        }
        const float k = warp_sum2_transposed(sx, sy, lane);
        sx = __shfl_sync(kFull, k, 0);
        sy = __shfl_sync(kFull, k, 16);
        ginv = rcp(sx * sy + kEps);
        l2ginv = lg2(ginv);
        colin = cv >= (j_lo >> 2) && cv <= (j_hi >> 2);
      } else {
        wi_hi = wi_lo - 1;
      }
      if (colin) {
#pragma unroll
        for (int it = 0; it < kPairIters; ++it) {
          const int row = rowb + kPairRowsPerStep * it;
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
      float a0 = hsum(qa), a1 = hsum(qb), a2 = hsum(qc), a3 = 0.f;
      pair_block_sum4(a0, a1, a2, a3, red, warp, lane);
      if (tid == 0) send(a0, a1, a2, 0.f, 0.f, 0.f);
      mbar_wait(xbar_s, xn & 1u);
      const float (*xw)[4] = xin[xn & 1u];
      ++xn;
      const float w0 = first ? a0 + xw[0][0] : xw[0][0] + a0;       // rank order on both CTAs
      const float w1 = first ? a1 + xw[0][1] : xw[0][1] + a1;
      const float w2 = first ? a2 + xw[0][2] : xw[0][2] + a2;
      if constexpr (kMSE) {
        const float Tt = fmaf(first ? Tth : p_Tt, sc0 * sc0, (first ? p_Tt : Tth) * sc1 * sc1);      // sum e^2 over both halves
        const float outside = fmaxf(fmaf(Tt * invS, invS, -w1), 0.f);
        D = outside + w0;
        creg = 2.f * (outside + w2);
      } else {
        creg = 0.5f * kLn2 * (1.0f + w0);
        D = fmaf(0.5f * kLn2, w1, creg);
      }
    }

    // ---------------------------------------------------------------- outputs + the scalars of the backward
    float dist = 0.f, a = 0.f, b = 0.f;
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
        float4* st = reinterpret_cast<float4*>(p.stats + hm * kStatsK);
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
      uint4* dzv = reinterpret_cast<uint4*>(p.dz + hm * hm_floats) + static_cast<size_t>(rank) * (kPairHalfBytes / 16);
#pragma unroll
      for (int it = 0; it < kPairIters; ++it) {
        f2 o[2];
        const float y = y0 + static_cast<float>(it) * dyi;
        float rc = fmaf(bS, y, cbS);
        if (kVar) { const float d = y - muy; rc = fmaf(kyS * d, d, rc); }
        const f2 rc2 = pk1(rc);
        f2 g0 = add2(acol[0], rc2), g1 = add2(acol[1], rc2);
        if constexpr (kMSE) { g0 = fma2(rpS, ev[it][0], g0); g1 = fma2(rpS, ev[it][1], g1); }      // 2 rho P
        if constexpr (kWin) {
          const int row = rowb + kPairRowsPerStep * it;
          if (colin && row >= wi_lo && row <= wi_hi) {                 // window pixels: the G-dependent term, evaluated again
            f2 P, d, gl;
            win_pair(it, 0, y - ty, P, d, gl);
            g0 = fma2(kwS, d, g0);
            win_pair(it, 1, y - ty, P, d, gl);
            g1 = fma2(kwS, d, g1);
          }
        }
        o[0] = mul2(ev[it][0], g0);
        o[1] = mul2(ev[it][1], g1);
        dzv[it * kPairThreads + tid] = pack_pairs<float>(o);
      }
    }

  }
  cluster.sync();      // neither CTA leaves while the other may still store into its shared memory
}

static int pair_enabled() {
  static const int v = [] { const char* e = std::getenv("DSNT_TUNE_STEP_PAIR"); return e ? std::atoi(e) : 1; }();
  return v;
}

bool step_pair_supported(int dtype, int H, int W, int reg) {
  static const int win = [] { const char* e = std::getenv("DSNT_TUNE_STEP_PAIR_WIN"); return e ? std::atoi(e) : 1; }();
  return pair_enabled() && dtype == DSNT_DTYPE_F32 && H == kPairH && W == kPairW &&
         (reg == DSNT_REG_NONE || reg == DSNT_REG_VAR || (win && (reg == DSNT_REG_JS || reg == DSNT_REG_MSE)));
}

template <int REG>
static int launch_pair(const PairParams& p, cudaStream_t stream) {
  auto kern = head_step_pair_kernel<REG>;
  static int max_clusters_of[kMaxDevices] = {};     // per device: 0 = not asked yet, -1 = clusters of this size cannot run
  int& max_clusters = max_clusters_of[current_device()];
  if (max_clusters == 0) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmemBytes) != cudaSuccess)
      return check_launch("head_step_pair_kernel (shared-memory opt-in)");
    cudaLaunchConfig_t probe = {};
    probe.gridDim = dim3(2 * sm_count_of_current_device()); probe.blockDim = dim3(kPairThreads); probe.dynamicSmemBytes = kPairSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    probe.attrs = at; probe.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, kern, &probe) != cudaSuccess || nc <= 0) { cudaGetLastError(); nc = -1; }
    max_clusters = nc;
  }
  if (max_clusters < 0) return 1;      // clusters of this size cannot run here
  long clusters = p.n < max_clusters ? p.n : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(2 * clusters)); cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = kPairSmemBytes; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (cudaLaunchKernelEx(&cfg, kern, p) != cudaSuccess) {
    // no room for a cluster of two 224 KiB CTAs on this device / partition: not an error, the caller takes another kernel
    cudaGetLastError();
    return 1;
  }
  return check_launch("head_step_pair_kernel");
}

// returns 1 when the case is not served
int launch_step_pair(const void* z, int dtype, long n, int H, int W, const float* target, const float* mask,
                     const float* denom, const float* g_loss, float reg_coeff, int reg, float sigma, int flags,
                     float* coords, float* stats, float* terms, void* dz, cudaStream_t stream) {
  if (!step_pair_supported(dtype, H, W, reg) || !denom) return 1;
  PairParams p;
  p.z = static_cast<const float*>(z); p.dz = static_cast<float*>(dz); p.target = target; p.mask = mask; p.denom = denom;
  p.g_loss = g_loss; p.coords = coords; p.stats = stats; p.terms = terms; p.n = n; p.flags = flags; p.sigma = sigma;
  p.reg_coeff = reg_coeff;
  const Geom g = make_geom(H, W, 4, 32, sigma > 0.f ? sigma : 1.f, reg);
  p.k2 = g.k2; p.r2_win = g.r2_win;
  switch (reg) {
    case DSNT_REG_VAR: return launch_pair<DSNT_REG_VAR>(p, stream);
    case DSNT_REG_JS: return launch_pair<DSNT_REG_JS>(p, stream);
    case DSNT_REG_MSE: return launch_pair<DSNT_REG_MSE>(p, stream);
    default: return launch_pair<DSNT_REG_NONE>(p, stream);
  }
}

}  // namespace dsnt
