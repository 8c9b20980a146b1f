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
//   CS = 2: 1024 threads, one CTA per SM: a heatmap per pair of SMs.  The default.
//   CS = 4:  512 threads, TWO CTAs per SM: four SMs hold two heatmaps in flight.  Measured slower on B200 for every
//            regulariser (profiles/r02_v6_pair_variants.txt): a cluster moves at the pace of the slowest of its four SMs in
//            every phase, and each of them also hosts a CTA of another cluster somewhere else in ITS cycle.  Kept as the
//            form a device takes on which two 1024-thread CTAs cannot be co-scheduled (DSNT_TUNE_STEP_PAIR_CS=4 forces it).
//
// The CTA moves in lockstep through the phases of a heatmap (block barriers; letting the warps run free with a warp-granular
// exchange was measured slower: STG bursts of one warp delay the shared-memory sweeps of another, same file).  A phase that
// keeps a pipe busy while every warp waits is therefore exposed, and the longest one was the write-back: the path from an SM
// to L2 takes 32 bytes per clock (tools/probe/store_probe.cu: 128 KiB = 4000 clocks for STG and for TMA alike), and a warp
// that stores from registers waits for it.  So the gradient of five of the eight sweep steps goes back through the chunk
// buffers it came in (STS at 90 B/clk, then one asynchronous bulk store per chunk) and only three steps are stored with STG:
// config 5, variance: 722 -> 666 us (0.92 -> 1.00 of the HBM peak), JS 900 -> 788 us, MSE 914 -> 772 us.
//
// The shared memory of a CTA is a ring of NS chunk buffers, one chunk per sweep step; a part of a heatmap takes NCH = 8 of
// them (NS > NCH), so while heatmap k is being processed NS - NCH chunks of heatmap k+1 are already there.  The buffers of
// k's first STAGE0 steps take their next loads as soon as k has been read into registers; the others carry k's gradient
// first and take theirs when the bulk stores have read them (thread 0, a heatmap later, while it waits for the exchange).
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
#ifdef DSNT_PAIR_NS_F32        // (measurements)
  static constexpr int NS = ES == 4 ? (CS == 2 ? DSNT_PAIR_NS_F32 : 13) : 24;
#else
  static constexpr int NS = ES == 4 ? 13 : 24;
#endif
  // The gradient of the sweep steps >= STAGE0 goes back through the chunk buffers it came in (STS + one bulk store per
  // chunk, asynchronous), that of the first STAGE0 steps straight from the registers (STG): the SM's path to L2 takes
  // 32 B/clk (tools/probe/store_probe.cu: 128 KiB = 4000 clocks, whoever sends them), and an STG waits for it.  A staged
  // buffer is handed to the next load only when its bulk store has read it, an unstaged one right after the sweep:
  // fp32 keeps three early buffers so that the last chunks of the next heatmap are on their way in time.
#ifdef DSNT_PAIR_STAGE0      // (measurements: -DDSNT_PAIR_STAGE0=8 is "no staging")
  static constexpr int STAGE0 = ES == 4 ? (CS == 2 ? DSNT_PAIR_STAGE0 : 3) : 0;
#else
  static constexpr int STAGE0 = ES == 4 ? 3 : 0;
#endif
  static constexpr int SMEM = NS * CHUNK;
  static constexpr int CTAS_PER_SM = CS == 2 ? 1 : 2;
  static_assert(ROWS % RPS == 0 && ROWS / RPS == kPairIters && kPairIters % NCH == 0 && NS > NCH, "geometry");
  // chunk c of heatmap k+1 lives in the buffer of chunk c - (NS - NCH) of heatmap k: if that one were staged, it would be
  // released a heatmap later, after the sweep that waits for the load -- a dead-lock by construction
  static_assert(STAGE0 == NCH || NS >= 2 * NCH || STAGE0 > 2 * NCH - 1 - NS, "a staged buffer the next sweep waits for");
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
  int tune;               // DSNT_TUNE_STEP_PAIR_FLAGS (measurements only; a build with -DDSNT_PAIR_TRACE): 8 = phase trace
};

// sum over the threads of one CTA of up to four values, identical on every thread, fixed order.  `red` has 32 rows; the
// rows of warps that do not exist stay zero (cleared once at kernel start), so 16 warps reduce like 32.  ONE block
// barrier: the caller alternates between two `red` arrays, see the kernel.
__device__ __forceinline__ void pair_block_sum4(float& a, float& b, float& c, float& d, float (*red)[8], int warp, int lane) {
  const float k = warp_sum4_transposed(a, b, c, d, lane);
  if ((lane & 7) == 0) red[warp][lane >> 3] = k;
  __syncthreads();
  // (up to) 32 warps = 32 lanes: every warp adds the per-warp partials with the same butterfly
  const float4 r = *reinterpret_cast<const float4*>(&red[lane][0]);
  const float k2 = warp_sum4_transposed(r.x, r.y, r.z, r.w, lane);
  a = __shfl_sync(kFull, k2, 0); b = __shfl_sync(kFull, k2, 8); c = __shfl_sync(kFull, k2, 16); d = __shfl_sync(kFull, k2, 24);
}

// ... and of six (the variance path: S, the first and the second moments about the pivot)
__device__ __forceinline__ void pair_block_sum6(float& a, float& b, float& c, float& d, float& e, float& f, float (*red)[8],
                                                int warp, int lane) {
  const float k = warp_sum4_transposed(a, b, c, d, lane);
  const float k2 = warp_sum2_transposed(e, f, lane);
  if ((lane & 7) == 0) red[warp][lane >> 3] = k;
  if ((lane & 15) == 0) red[warp][4 + (lane >> 4)] = k2;
  __syncthreads();
  const float4 r = *reinterpret_cast<const float4*>(&red[lane][0]);
  const float2 r2 = *reinterpret_cast<const float2*>(&red[lane][4]);
  const float q = warp_sum4_transposed(r.x, r.y, r.z, r.w, lane);
  const float q2 = warp_sum2_transposed(r2.x, r2.y, lane);
  a = __shfl_sync(kFull, q, 0); b = __shfl_sync(kFull, q, 8); c = __shfl_sync(kFull, q, 16); d = __shfl_sync(kFull, q, 24);
  e = __shfl_sync(kFull, q2, 0); f = __shfl_sync(kFull, q2, 16);
}

// DSNT_TUNE_STEP_PAIR_FLAGS & 8 (measurements only): SM clocks thread 0 of CTA 0 spends in each phase of a heatmap, summed over
// its heatmaps; [15] counts the heatmaps
__device__ unsigned long long g_pair_trace[16];

template <int REG, int CS, typename T>
__global__ void __launch_bounds__(PairCfg<CS, T>::NT, PairCfg<CS, T>::CTAS_PER_SM) head_step_pair_kernel(const PairParams p) {
  using C = PairCfg<CS, T>;
  constexpr int NT = C::NT, NW = C::NW, NS = C::NS, NCH = C::NCH, RPS = C::RPS, ES = C::ES;
  constexpr int kEarly = C::STAGE0;      // buffers handed to the next loads right after the sweep
  constexpr bool kVar = REG == DSNT_REG_VAR;
  constexpr bool kJS = REG == DSNT_REG_JS;
  constexpr bool kMSE = REG == DSNT_REG_MSE;
  constexpr bool kWin = kJS || kMSE;
  // JS / MSE: the Gaussian window (16 x 16 pixels at sigma = 1 px) lies in the registers of the few threads that hold its
  // vectors.  Its terms need P = e / S, i.e. the merged sums: they are evaluated AFTER the first exchange, block-reduced
  // and exchanged in a second message; the backward evaluates them again (nothing per-pixel is kept: the 64 registers of a
  // thread hold its 32 values of e).  Outside the window the closed forms of head_step2.cuh apply.
  constexpr float tow = 2.0f / kPairW, bw = 1.0f / kPairW - 1.0f, toh = 2.0f / kPairH, bh = 1.0f / kPairH - 1.0f;
  extern __shared__ __align__(128) unsigned char pair_smem[];
  __shared__ __align__(8) unsigned long long bars[NS];
  __shared__ __align__(16) float red[2][32][8];           // two scratch arrays taken in turn: a reduction needs ONE block barrier
  __shared__ float redm[32];
  __shared__ __align__(16) float xin[2][CS][8];           // the partial results of every CTA of the cluster, slot [r] stored by CTA r
                                                          // (st.async over DSMEM; the own slot locally); message number n lands in
                                                          // xin[n & 1]: a peer can be one message ahead
  __shared__ __align__(8) unsigned long long xbar[2];     // ... completing 32 (CS - 1) bytes on xbar[n & 1]
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
  // The exchange of a heatmap: thread 0 puts its CTA's partial results into the own slot and arms the local barrier for
  // the bytes of the CS - 1 peers; threads 1 .. CS-1 store them into the slot [rank] of one peer each with st.async, which
  // completes the bytes on THAT CTA's barrier; everybody then waits on the local barrier and reads the local xin.  No
  // cluster-wide barrier (whose release fence makes all threads wait for their global stores: the first version of this
  // kernel, 1053 us at config 5, against 878 us with three such exchanges and less with one).  Two barriers in turn: with
  // more than one peer, message n+1 of a fast peer must not complete bytes of the phase that still waits for message n of
  // a slow one.
  const uint32_t xin_s = smem_u32(&xin[0][0][0]), xbar_s = smem_u32(&xbar[0]);
  uint32_t peer_xin_s = 0, peer_xbar_s = 0;
  if (tid >= 1 && tid < CS) {
    const unsigned peer = (rank + static_cast<unsigned>(tid)) % CS;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_xin_s) : "r"(xin_s), "r"(peer));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer_xbar_s) : "r"(xbar_s), "r"(peer));
  }
  uint32_t xn = 0;          // messages exchanged so far: the same on all CTAs (they take the same branches on bit-identical totals)
  // post message number xn (values identical on every thread of the CTA)
  auto post = [&](float a, float b, float c, float d, float e, float f, float g7) {
    const uint32_t par = xn & 1u;
    if (tid == 0) {
      float4* own = reinterpret_cast<float4*>(&xin[par][rank][0]);
      own[0] = make_float4(a, b, c, d);
      own[1] = make_float4(e, f, g7, 0.f);
      mbar_expect_tx(xbar_s + 8 * par, 32 * (CS - 1));      // (release: the own slot is visible to whoever passes the barrier)
    } else if (tid < CS) {
      const uint32_t dst = peer_xin_s + (par * CS + rank) * 32u, bar = peer_xbar_s + 8 * par;
      asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst),
                   "f"(a), "f"(b), "f"(c), "f"(d), "r"(bar)
                   : "memory");
      asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst + 16),
                   "f"(e), "f"(f), "f"(g7), "f"(0.f), "r"(bar)
                   : "memory");
    }
  };
  // wait for message number xn of every peer; returns the CS slots
  auto collect = [&]() -> const float (*)[8] {
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
  if (tid == 0) {
    for (int sl = 0; sl < NS; ++sl) mbar_init(bars_s + 8 * sl, 1);
    mbar_init(xbar_s, 1);
    mbar_init(xbar_s + 8, 1);
    for (unsigned g = 0; g < NS; ++g) issue(g);
  }
  for (int i = tid; i < 2 * 32 * 8; i += NT) (&red[0][0][0])[i] = 0.f;
  cluster.sync();       // every CTA's barriers exist before the first remote store (and red is cleared)
  const float gl = p.g_loss ? __ldg(p.g_loss) : 1.0f;
  const float inv_denom = 1.0f / __ldg(p.denom);
  const float s2 = p.sigma * p.sigma;

  // thread geometry: vector column cv (4 pixels), rows r0 + RPS * it inside this CTA's part
  const int cv = tid & (kPairWV - 1), r0 = tid >> 6;
  float xs[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) xs[c] = fmaf(static_cast<float>(cv * 4 + c), tow, bw);
  const float y0 = fmaf(static_cast<float>(rank * C::ROWS + r0), toh, bh);
  constexpr float dyi = RPS * toh;
  const f2 l2e2 = pk1(kLog2e);

#ifdef DSNT_PAIR_TRACE
  if (tracing) trace_last = clock64();
#endif
  for (int k = 0; k < nk; ++k) {
    const int hm = cluster_id + k * n_clusters;
    const unsigned g0 = static_cast<unsigned>(NCH * k);       // chunk number of this heatmap's first chunk
    float tx = 0.f, ty = 0.f;
    if (p.target) {
      const float2 tt = __ldg(reinterpret_cast<const float2*>(p.target) + hm);
      tx = tt.x; ty = tt.y;
    }
    const float wgt = (p.mask ? __ldg(p.mask + hm) : 1.0f) * inv_denom;

    // ---------------------------------------------------------------- the part comes into REGISTERS (32 pixels per thread)
    // and its maximum is taken on the way; shared memory is read exactly once, so its buffers are free for the next loads
    // as soon as this sweep is over
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
    stamp(0);                                  // wait for the loads + shared memory -> registers
    mloc = warp_max_redux(mloc);
    if (lane == 0) redm[warp] = mloc;
    __syncthreads();
    stamp(1);                                  // block barrier (the slowest warp's loads)
    // every thread has its data: this heatmap's buffers take the next NCH chunks of the stream (nothing was written to
    // them by the generic proxy, so no proxy fence)
    if (tid < kEarly) issue(g0 + NS + tid);
    mloc = warp_max_redux(redm[lane & (NW - 1)]);
    // Each part is summed relative to ITS OWN maximum; the parts are merged afterwards like the blocks of an online
    // softmax (S = sum_r S_r 2^(m_r - m)), so ONE exchange per heatmap carries everything.
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
    stamp(2);                                  // issue of the next loads, maximum, exponential sweep
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
      pair_block_sum4(Sh, Sxh, Syh, Tth, red[0], warp, lane);
    } else {
      axh = fmaf(c0 * dx0, dx0, fmaf(c1 * dx1, dx1, fmaf(c2 * dx2, dx2, c3 * dx3 * dx3)));
      ayh = Syy;
      pair_block_sum6(Sh, Sxh, Syh, axh, ayh, Syy, red[0], warp, lane);      // (the sixth value rides along unused)
    }
    stamp(3);                                  // block sum
    post(m2h, Sh, Sxh, Syh, axh, ayh, Tth);

    // ---------------------------------------------------------------- while the first message is on its way: the buffers whose
    // bulk stores (previous heatmap) have been read take their next loads
    if (kEarly < NCH && tid == 0 && k > 0) {
      bulk_wait_read();
      for (int c = kEarly; c < NCH; ++c) issue(g0 - NCH + NS + c);
    }
    // JS / MSE: where the Gaussian window lies and its normalisation -- functions of the target alone
    float ginv = 0.f, l2ginv = 0.f;
    int wi_lo = 0, wi_hi = -1;
    bool colin = false;
    if constexpr (kWin) {
      int j_lo, j_hi;
      axis_window_fast(tx, kPairW, 0.5f * kPairW, tow, bw, p.r2_win, j_lo, j_hi);
      axis_window_fast(ty, kPairH, 0.5f * kPairH, toh, bh, p.r2_win, wi_lo, wi_hi);
      if (j_lo <= j_hi && wi_lo <= wi_hi) {
        // Normalisation of the target Gaussian (src/dsnt/nn.py:168-180 divides by its sum over the image).  A window that
        // the image border does not clip holds the whole sum to theta, and a sum of Gaussian samples at unit spacing is
        // sigma sqrt(2 pi) (1 + 2 e^(-2 pi^2 sigma^2) cos(..) + ...): to 5.4e-9 relative from sigma = 1 px on (Poisson
        // summation).  Clipped windows and narrower Gaussians are summed: every warp for itself, no block barrier.
        const float spx = p.sigma * (0.5f * kPairW);       // sigma in pixels (H == W)
        float sx = spx * 2.5066282746310002f, sy = sx;
        if (!(spx >= 1.0f && j_lo > 0 && j_hi < kPairW - 1 && wi_lo > 0 && wi_hi < kPairH - 1)) {
          sx = 0.f; sy = 0.f;
          for (int j = j_lo + lane; j <= j_hi; j += 32) {
            const float d = fmaf(static_cast<float>(j), tow, bw) - tx;
            sx += ex2(p.k2 * d * d);
          }
          for (int i = wi_lo + lane; i <= wi_hi; i += 32) {
            const float d = fmaf(static_cast<float>(i), toh, bh) - ty;
            sy += ex2(p.k2 * d * d);
          }
          const float ks = warp_sum2_transposed(sx, sy, lane);
          sx = __shfl_sync(kFull, ks, 0);
          sy = __shfl_sync(kFull, ks, 16);
        }
        ginv = rcp(sx * sy + kEps);
        l2ginv = lg2(ginv);
        colin = cv >= (j_lo >> 2) && cv <= (j_hi >> 2);
      } else {
        wi_hi = wi_lo - 1;
      }
    }

    stamp(4);                                  // post + window geometry

    // ---------------------------------------------------------------- merge in rank order on every CTA: bit-identical totals
    float sc[CS];
    float m2, S, Sxc, Sycm, sax = 0.f, say = 0.f, Tt = 0.f;
    {
      const float (*xm)[8] = collect();
      float4 lo[CS], hi[CS];
#pragma unroll
      for (int r = 0; r < CS; ++r) {
        lo[r] = *reinterpret_cast<const float4*>(&xm[r][0]);
        hi[r] = *reinterpret_cast<const float4*>(&xm[r][4]);
      }
      m2 = lo[0].x;
#pragma unroll
      for (int r = 1; r < CS; ++r) m2 = fmaxf(m2, lo[r].x);
#pragma unroll
      for (int r = 0; r < CS; ++r) sc[r] = ex2(lo[r].x - m2);
      // each part was summed relative to ITS OWN maximum: rescaled like the blocks of an online softmax; everything is
      // about the same pivot, so the parts simply add
      S = lo[CS - 1].y * sc[CS - 1]; Sxc = lo[CS - 1].z * sc[CS - 1]; Sycm = lo[CS - 1].w * sc[CS - 1];
      if (kVar) { sax = hi[CS - 1].x * sc[CS - 1]; say = hi[CS - 1].y * sc[CS - 1]; }
      if (kMSE) Tt = hi[CS - 1].z * sc[CS - 1] * sc[CS - 1];
#pragma unroll
      for (int r = CS - 2; r >= 0; --r) {
        S = fmaf(lo[r].y, sc[r], S); Sxc = fmaf(lo[r].z, sc[r], Sxc); Sycm = fmaf(lo[r].w, sc[r], Sycm);
        if (kVar) { sax = fmaf(hi[r].x, sc[r], sax); say = fmaf(hi[r].y, sc[r], say); }
        if (kMSE) Tt = fmaf(hi[r].z, sc[r] * sc[r], Tt);                       // sum e^2 over all parts
      }
    }
    stamp(5);                                  // wait for the peers' message + merge
    const float invS = rcp(S);
    const float mxc = Sxc * invS, myc = Sycm * invS;
    const float mux = pcx + mxc, muy = pcy + myc;
    float scme = sc[0];
#pragma unroll
    for (int r = 1; r < CS; ++r) scme = rank == r ? sc[r] : scme;
    const float invSh = scme * invS;       // this part's e (relative to its own maximum) -> probability

    float D = 0.f, creg = 0.f, vx = 0.f, vy = 0.f;
    if constexpr (kVar) {
      vx = sax * invS - mxc * mxc;
      vy = say * invS - myc * myc;
      // conditioning of the pivot form: (mu - c)^2 <= 16 var keeps the cancellation below 17 fp32 roundings (1e-6 relative)
      const bool ill = !(mxc * mxc <= 16.f * vx) || !(myc * myc <= 16.f * vy);
      if (ill) {
        // exact: second moments about the mean itself, from the registers (column sums are still there, the row sums are
        // formed again), one more block reduction and one more exchange; all CTAs get here together (identical totals)
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
        pair_block_sum4(axe, aye, z0, z1, red[1], warp, lane);
        post(axe, aye, 0.f, 0.f, 0.f, 0.f, 0.f);
        const float (*xe)[8] = collect();
        float qx = xe[CS - 1][0] * sc[CS - 1], qy = xe[CS - 1][1] * sc[CS - 1];
#pragma unroll
        for (int r = CS - 2; r >= 0; --r) { qx = fmaf(xe[r][0], sc[r], qx); qy = fmaf(xe[r][1], sc[r], qy); }
        vx = qx * invS;
        vy = qy * invS;
      }
      const float ex = vx - s2, ey = vy - s2;
      D = ex * ex + ey * ey;
      creg = 2.f * (ex * vx + ey * vy);
    }

    // ---------------------------------------------------------------- JS / MSE: the Gaussian window, from the registers
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
      float a0 = hsum(qa), a1 = hsum(qb), a2 = hsum(qc), a3 = 0.f;
      stamp(6);                                // window terms (thread 0 holds none unless the window is at the left border)
      pair_block_sum4(a0, a1, a2, a3, red[1], warp, lane);
        stamp(7);                                // block sum of the window terms: waits for the threads that hold the window
      post(a0, a1, a2, 0.f, 0.f, 0.f, 0.f);
      const float (*xw)[8] = collect();
      stamp(8);                                // second exchange
      float w0 = xw[0][0], w1 = xw[0][1], w2 = xw[0][2];       // rank order on every CTA
#pragma unroll
      for (int r = 1; r < CS; ++r) { w0 += xw[r][0]; w1 += xw[r][1]; w2 += xw[r][2]; }
      if constexpr (kMSE) {
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
      char* dzp = static_cast<char*>(p.dz) + static_cast<long>(hm) * hm_bytes + static_cast<size_t>(rank) * C::PART_BYTES;
      char* dzb = dzp + tid * (4 * ES);
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
        // sweep steps < kEarly: straight to global memory; the others back into the chunk buffer they came from
        char* dst = it < kEarly ? dzb + it * C::STEP_BYTES
                                : reinterpret_cast<char*>(pair_smem) + ((g0 + it) % NS) * C::CHUNK + tid * (4 * ES);
        if constexpr (ES == 4) {
          *reinterpret_cast<uint4*>(dst) = pack_pairs<float>(o);
        } else {
          float l0, h0, l1, h1;
          upk(o[0], l0, h0);
          upk(o[1], l1, h1);
          *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(l0, h0), pack_bf16(l1, h1));
        }
      }
      if constexpr (kEarly < NCH) {
        fence_async_smem();           // the bulk stores (async proxy) read what this thread wrote
        __syncthreads();
        if (tid == 0) {
          for (int c = kEarly; c < NCH; ++c)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dzp + c * C::CHUNK),
                         "r"(buf_s + ((g0 + c) % NS) * C::CHUNK), "r"(C::CHUNK)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    stamp(9);                                  // outputs + backward sweep (stores issued)
#ifdef DSNT_PAIR_TRACE
    if (tracing) g_pair_trace[15] += 1;
#endif
  }
  if (kEarly < NCH && tid == 0) bulk_wait_all();      // the last heatmap's bulk stores still read this CTA's shared memory
  cluster.sync();      // no CTA leaves while another may still store into its shared memory
}

static int pair_enabled() {
  static const int v = [] { const char* e = std::getenv("DSNT_TUNE_STEP_PAIR"); return e ? std::atoi(e) : 1; }();
  return v;
}

// cluster size: DSNT_TUNE_STEP_PAIR_CS = 2 | 4 overrides the measured default (profiles/r02_v6_pair_variants.txt: two CTAs
// of 1024 threads, one per SM, beat four CTAs of 512 threads, two per SM, for every regulariser and both dtypes); read on
// every call (a launch of 0.3 ms and more), so that the tests can take both forms in one process
static int pair_cluster_size(int dtype, int reg) {
  const char* e = std::getenv("DSNT_TUNE_STEP_PAIR_CS");
  const int forced = e ? std::atoi(e) : 0;
  if (forced == 2 || forced == 4) return forced;
  (void)dtype; (void)reg;
  return 2;
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
    fprintf(stderr, "pair trace <reg %d, cluster %d, %d B> clocks per heatmap (CTA 0, %llu heatmaps, %.0f in all): load+lds %.0f | barrier %.0f | "
            "exp sweep %.0f | block sum %.0f | post+geometry %.0f | exchange+merge %.0f | window %.0f | window sum %.0f | exchange 2 %.0f | "
            "backward %.0f\n", REG, CS, C::ES, t[15], tot, t[0] / nh, t[1] / nh, t[2] / nh, t[3] / nh, t[4] / nh, t[5] / nh, t[6] / nh,
            t[7] / nh, t[8] / nh, t[9] / nh);
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
  p.tune = 0;
#ifdef DSNT_PAIR_TRACE
  { const char* e = std::getenv("DSNT_TUNE_STEP_PAIR_FLAGS"); p.tune = e ? std::atoi(e) : 0; }
#endif
  const bool f32 = dtype == DSNT_DTYPE_F32;
  const int cs = pair_cluster_size(dtype, reg);
  int rc = 1;
  for (int attempt = 0; attempt < 2 && rc == 1; ++attempt) {      // (a device on which the preferred clusters do not fit takes the other size)
    if ((cs == 4) == (attempt == 0))
      rc = f32 ? launch_pair_reg<4, float>(p, reg, stream) : launch_pair_reg<4, __nv_bfloat16>(p, reg, stream);
    else
      rc = f32 ? launch_pair_reg<2, float>(p, reg, stream) : launch_pair_reg<2, __nv_bfloat16>(p, reg, stream);
  }
  return rc;
}

}  // namespace dsnt
