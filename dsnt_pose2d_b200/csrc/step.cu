// step.cu -- launcher and extern "C" entry point of the one-pass training step (head_step.cuh).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "capi_util.cuh"
#include "head_step.cuh"
#include "head_step2.cuh"

namespace dsnt {

// step_pair.cu: the one-pass step for 256x256 fp32 heatmaps on a cluster of two CTAs (distributed shared memory)
bool step_pair_supported(int dtype, int H, int W, int reg);
int launch_step_pair(const void* z, int dtype, long n, int H, int W, const float* target, const float* mask, const float* denom,
                     const float* g_loss, float reg_coeff, int reg, float sigma, int flags, float* coords, float* stats,
                     float* terms, void* dz, cudaStream_t stream);

static int sm_count() { return sm_count_of_current_device(); }

static int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}
static int step_direct_store() { static const int v = env_int("DSNT_TUNE_STEP_STG", 1); return v; }
static int step_group() { static const int v = env_int("DSNT_TUNE_STEP_GROUP", 32); return v; }
static int step_warps() { static const int v = env_int("DSNT_TUNE_STEP_WARPS", 0); return v; }
static int step_v2() { static const int v = env_int("DSNT_TUNE_STEP_V2", 1); return v; }
static int step_spare() { static const int v = env_int("DSNT_TUNE_STEP_SPARE", -1); return v; }

// ---------------------------------------------------------------------------------- shape-specialised kernel (head_step2.cuh)
// The compact window mapping holds 32 * MAXS vectors in registers: the window of ANY target (upper bound as in
// launch.cuh:stash_fits -- floor(n sqrt(r2 + 1/n^2)) + 4 pixels per axis, widened to whole vectors along x) must fit.
// Returns the number of slots the launch needs: 0 = no window, -1 = does not fit the largest build.
template <typename T>
static int step2_window_slots(int H, int W, int vec, int reg, float r2_win) {
  if (!reg_needs_gauss(reg)) return 0;
  const double r2 = r2_win;
  const int mc = static_cast<int>(std::floor(W * std::sqrt(r2 + 1.0 / (static_cast<double>(W) * W)))) + 4;
  const int mr = static_cast<int>(std::floor(H * std::sqrt(r2 + 1.0 / (static_cast<double>(H) * H)))) + 4;
  const int vecs = std::min(W / vec, (mc + vec - 2) / vec + 1);
  const int rows = std::min(H, mr);
  if (rows * vecs <= 32 * step2_slots_small<T>()) return step2_slots_small<T>();
  if (rows * vecs <= 32 * step2_slots_large<T>()) return step2_slots_large<T>();
  return -1;
}

// Load pacing (head_step2.cuh:pace_wait): SM clocks between two bulk loads of a CTA = the time one heatmap (read + write)
// takes at a whole-GPU bandwidth of DSNT_TUNE_STEP_PACE_GBS (0 = unpaced); DSNT_TUNE_STEP_PACE gives the clocks directly.
static int step2_pace_cycles(long hm_bytes, int ctas) {
  static const int fixed = env_int("DSNT_TUNE_STEP_PACE", -1);
  if (fixed >= 0) return fixed;
  static const int gbs = env_int("DSNT_TUNE_STEP_PACE_GBS", 6800);
  if (gbs <= 0) return 0;
  static double ghz_of[kMaxDevices] = {};
  double& ghz = ghz_of[current_device()];
  if (ghz == 0.0) {
    int dev = 0, khz = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev) != cudaSuccess || khz <= 0)
      khz = 1965000;
    ghz = khz * 1e-6;
  }
  const double ns = 2.0 * hm_bytes * ctas / static_cast<double>(gbs);     // bytes / (GB/s) = ns
  return static_cast<int>(ns * ghz + 0.5);
}

template <typename T, int REG, int H, int W, int NWMAX, bool PACED, int MAXS>
static int launch_step2_nw(HeadStepParams p, cudaStream_t stream) {
  auto kern = head_step2_kernel<T, REG, H, W, NWMAX, PACED, MAXS>;
  p.nbufs = kStepSmemBudget / p.buf_bytes;
  if (p.nbufs > kStepMaxBufs) p.nbufs = kStepMaxBufs;
  static const int nbuf_cap = env_int("DSNT_TUNE_STEP_NBUF2", 0);
  if (nbuf_cap > 0 && p.nbufs > nbuf_cap) p.nbufs = nbuf_cap;
  const int spare = step_spare() >= 0 ? step_spare() : 2;   // loads in flight while every warp computes
  p.nwarps = step_warps() > 0 ? step_warps() : p.nbufs - spare;
  if (p.nwarps > NWMAX) p.nwarps = NWMAX;
  if (p.nwarps > p.nbufs) p.nwarps = p.nbufs;
  if (p.nwarps < 1) p.nwarps = 1;
  const size_t smem = static_cast<size_t>(p.nbufs) * p.buf_bytes;
  static size_t configured[kMaxDevices] = {};     // per instantiation and device: the opt-in only ever needs to grow
  size_t& conf = configured[current_device()];
  if (smem > conf) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return check_launch("head_step2_kernel (shared-memory opt-in)");
    conf = smem;
  }
  long ctas = p.n < sm_count() ? p.n : sm_count();
  if (ctas > kFinishSlots) ctas = kFinishSlots;             // one workspace slot per CTA in the single-launch form
  p.pace = PACED ? step2_pace_cycles(static_cast<long>(H) * W * sizeof(T), static_cast<int>(ctas)) : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(ctas)); cfg.blockDim = dim3(p.nwarps * 32);
  cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  // Single-launch form: the CTAs wait for a count that the LAST of them publishes, so the whole grid must be resident at
  // once.  The cooperative attribute makes the driver guarantee that (or refuse the launch) instead of the kernel assuming it.
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  static const int coop = env_int("DSNT_TUNE_STEP_COOP", 1);     // 0: measurements only (no co-residency guarantee)
  cfg.attrs = at; cfg.numAttrs = (!p.denom && coop) ? 1 : 0;
  if (cudaLaunchKernelEx(&cfg, kern, p) != cudaSuccess) return check_launch("head_step2_kernel");
  return check_launch("head_step2_kernel");
}

// NWMAX only sets the register budget (__launch_bounds__): 65536 / (32 NWMAX) registers per thread.  bf16 heatmaps with a
// Gaussian window are bound by arithmetic, not by HBM: they run unpaced (DSNT_TUNE_STEP_PACED=1 forces the paced build).
// MAXS: window slots per lane (head_step2.cuh); regularisers without a window have one build.
template <typename T, int REG, int H, int W, int MAXS>
static int launch_step2_slots(const HeadStepParams& p, cudaStream_t stream) {
  static const int paced = env_int("DSNT_TUNE_STEP_PACED", -1);
  static const int nwmax = env_int("DSNT_TUNE_STEP_NWMAX", 0);
  if constexpr (sizeof(T) == 2) {
    if constexpr (REG == DSNT_REG_JS || REG == DSNT_REG_MSE) {
      if constexpr (MAXS == step2_slots_small<T>()) {
        if (nwmax != 16) return launch_step2_nw<T, REG, H, W, 20, false, MAXS>(p, stream);    // fewer window registers: 20 warps
      }
      if (paced != 1) return launch_step2_nw<T, REG, H, W, 16, false, MAXS>(p, stream);
    }
    if constexpr (REG == DSNT_REG_KL) {                // bound by MUFU / issue, not by HBM: more warps, no pacing
      if (nwmax != 16) return launch_step2_nw<T, REG, H, W, 20, false, MAXS>(p, stream);
      return launch_step2_nw<T, REG, H, W, 16, false, MAXS>(p, stream);
    }
    return launch_step2_nw<T, REG, H, W, 16, true, MAXS>(p, stream);
  } else {
    return launch_step2_nw<T, REG, H, W, 12, true, MAXS>(p, stream);
  }
}

template <typename T, int REG, int H, int W>
static int launch_step2(const HeadStepParams& p, int slots, cudaStream_t stream) {
  if constexpr (REG == DSNT_REG_JS || REG == DSNT_REG_MSE) {
    if (slots == step2_slots_small<T>()) return launch_step2_slots<T, REG, H, W, step2_slots_small<T>()>(p, stream);
    return launch_step2_slots<T, REG, H, W, step2_slots_large<T>()>(p, stream);
  } else {
    return launch_step2_slots<T, REG, H, W, 1>(p, stream);
  }
}

// Measured on B200 at cfg 4 (profiles/r01_v4_step_sweep.txt, r01_v5_step_ring.txt): one warp per heatmap and direct
// 128-bit stores; DSNT_TUNE_STEP_{WARPS,GROUP,STG} override for experiments.
template <typename T, int VEC, int REG, bool FIXC, int GROUP>
static int launch_step_one(HeadStepParams p, cudaStream_t stream) {
  auto kern = head_step_kernel<T, VEC, REG, FIXC, GROUP>;
  constexpr int kMaxGroups = kStepMaxWarps * 32 / GROUP;
  p.nbufs = kStepSmemBudget / p.buf_bytes;
  if (p.nbufs > kStepMaxBufs) p.nbufs = kStepMaxBufs;
  // two spare buffers keep loads in flight while every group computes (64x64 fp32: 14 buffers, 12 warps)
  p.nwarps = step_warps() > 0 ? step_warps() : (p.nbufs > 4 ? p.nbufs - 2 : p.nbufs);
  if (p.nwarps > kMaxGroups) p.nwarps = kMaxGroups;
  if (p.nwarps > p.nbufs) p.nwarps = p.nbufs;
  const size_t smem = static_cast<size_t>(p.nbufs) * p.buf_bytes;
  static size_t configured[kMaxDevices] = {};     // per instantiation and device: the opt-in only ever needs to grow
  size_t& conf = configured[current_device()];
  if (smem > conf) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return check_launch("head_step_kernel (shared-memory opt-in)");
    conf = smem;
  }
  long ctas = p.n < sm_count() ? p.n : sm_count();   // persistent: one CTA per SM, tiles interleaved across CTAs
  kern<<<static_cast<unsigned>(ctas), p.nwarps * GROUP, smem, stream>>>(p);
  return check_launch("head_step_kernel");
}

// does this launch take the shape-specialised kernel (head_step2.cuh)?  Returns the window slots it needs, -1 = no.
static int step2_eligible(int dtype, int H, int W, int reg, float sigma) {
  if (H != 64 || W != 64 || !step_v2() || !step_direct_store() || step_group() == 64) return -1;
  static const int kl_v2 = env_int("DSNT_TUNE_STEP_KL_V2", 1);      // 0: KL on the generic kernel (head_step.cuh), as in round 1
  if (reg == DSNT_REG_KL) return kl_v2 ? 0 : -1;                    // walks its window slot by slot: any sigma
  const int vec = dtype == DSNT_DTYPE_F32 ? 4 : 8;
  const Geom g = make_geom(H, W, vec, 32, sigma > 0.f ? sigma : 1.f, reg);
  return dtype == DSNT_DTYPE_F32 ? step2_window_slots<float>(H, W, vec, reg, g.r2_win)
                                 : step2_window_slots<__nv_bfloat16>(H, W, vec, reg, g.r2_win);
}

// one warp per heatmap by default (DSNT_TUNE_STEP_GROUP=64 selects two)
template <typename T, int VEC, int REG>
static int launch_step_fixc(HeadStepParams p, cudaStream_t stream) {
  {
    const int slots = step2_eligible(sizeof(T) == 4 ? DSNT_DTYPE_F32 : DSNT_DTYPE_BF16, p.H, p.W, REG, p.sigma);
    if (slots >= 0) {
      p.g = make_geom(p.H, p.W, VEC, 32, p.sigma > 0.f ? p.sigma : 1.f, REG);
      return launch_step2<T, REG, 64, 64>(p, slots, stream);
    }
  }
  if (p.out8 || !p.denom) { set_error("single-launch step: this shape / regulariser takes the generic kernel"); return DSNT_ERR_UNSUPPORTED; }
  if (step_group() != 64) {
    p.g = make_geom(p.H, p.W, VEC, 32, p.sigma > 0.f ? p.sigma : 1.f, REG);
    return (32 % p.g.wv == 0) ? launch_step_one<T, VEC, REG, true, 32>(p, stream) : launch_step_one<T, VEC, REG, false, 32>(p, stream);
  }
  p.g = make_geom(p.H, p.W, VEC, 64, p.sigma > 0.f ? p.sigma : 1.f, REG);
  return (64 % p.g.wv == 0) ? launch_step_one<T, VEC, REG, true, 64>(p, stream) : launch_step_one<T, VEC, REG, false, 64>(p, stream);
}

template <typename T, int VEC>
static int launch_step_reg(const HeadStepParams& p, int reg, cudaStream_t stream) {
  switch (reg) {
    case DSNT_REG_NONE: return launch_step_fixc<T, VEC, DSNT_REG_NONE>(p, stream);
    case DSNT_REG_VAR: return launch_step_fixc<T, VEC, DSNT_REG_VAR>(p, stream);
    case DSNT_REG_KL: return launch_step_fixc<T, VEC, DSNT_REG_KL>(p, stream);
    case DSNT_REG_JS: return launch_step_fixc<T, VEC, DSNT_REG_JS>(p, stream);
    case DSNT_REG_MSE: return launch_step_fixc<T, VEC, DSNT_REG_MSE>(p, stream);
  }
  set_error("bad reg %d", reg);
  return DSNT_ERR_BAD_ARG;
}

}  // namespace dsnt

using namespace dsnt;

extern "C" {

DSNT_API int dsnt_head_step_supported(int dtype, int H, int W) {
  if (dtype != DSNT_DTYPE_F32 && dtype != DSNT_DTYPE_BF16) return 0;
  if (H <= 0 || W <= 0) return 0;
  const int vec = dtype == DSNT_DTYPE_F32 ? 4 : 8;
  const long bytes = static_cast<long>(H) * W * (dtype == DSNT_DTYPE_F32 ? 4 : 2);
  if (W % vec != 0 || bytes % 16 != 0) return 0;
  const long buf = (bytes + 127) / 128 * 128;
  return kStepSmemBudget / buf >= 4 ? 1 : 0;
}

static int head_step_impl(const void* z, int dtype, long n, int H, int W, const float* target, const float* mask,
                          const float* denom, const float* g_loss, float reg_coeff, int reg, float sigma, int flags,
                          float* coords, float* stats, float* terms, void* dz, float* out8, float* ws, void* stream,
                          const Stacks* st = nullptr, const PeerXchg* xc = nullptr) {
  int rc = check_common(z, dtype, n, H, W, reg);
  if (rc) return rc;
  if (n == 0) return DSNT_OK;
  if (!dz || !coords || (!denom && !out8)) { set_error("dsnt_head_step: z, dz, coords and denom are required"); return DSNT_ERR_BAD_ARG; }
  if (reg_needs_gauss(reg) && !target) { set_error("reg %d needs a target", reg); return DSNT_ERR_BAD_ARG; }
  if (reg != DSNT_REG_NONE && !(sigma > 0.f)) { set_error("sigma must be > 0"); return DSNT_ERR_BAD_ARG; }
  if (!aligned(coords, 8) || (stats && !aligned(stats, 16)) || (terms && !aligned(terms, 8)) || (target && !aligned(target, 8))) {
    set_error("per-heatmap buffers must be naturally aligned (coords/terms/target 8 B, stats 16 B)");
    return DSNT_ERR_BAD_ARG;
  }
  if (!out8 && denom && !dsnt_head_step_supported(dtype, H, W) && aligned(z, 16) && aligned(dz, 16)) {
    // 256x256 fp32: each heatmap in the shared memory of a PAIR of CTAs (thread-block cluster), exchanged through DSMEM
    const int prc = launch_step_pair(z, dtype, n, H, W, target, mask, denom, g_loss, reg_coeff, reg, sigma, flags, coords, stats,
                                     terms, dz, static_cast<cudaStream_t>(stream));
    if (prc != 1) return prc;
  }
  if (!dsnt_head_step_supported(dtype, H, W) || !aligned(z, 16) || !aligned(dz, 16)) {
    set_error("dsnt_head_step: heatmap %dx%d (dtype %d) does not fit the one-pass kernel (needs W %% %d == 0, 16-byte "
              "aligned bases and at least 4 heatmaps in shared memory); use dsnt_head_fwd + dsnt_head_bwd",
              H, W, dtype, dtype == DSNT_DTYPE_F32 ? 4 : 8);
    return DSNT_ERR_UNSUPPORTED;
  }
  const int es = dtype == DSNT_DTYPE_F32 ? 4 : 2;
  const int vec = dtype == DSNT_DTYPE_F32 ? 4 : 8;
  HeadStepParams p;
  p.z = z; p.dz = dz; p.target = target; p.mask = mask; p.denom = denom; p.g_loss = g_loss;
  p.coords = coords; p.stats = stats; p.terms = terms;
  p.n = n; p.H = H; p.W = W; p.flags = flags; p.sigma = sigma; p.reg_coeff = reg_coeff;
  p.g = make_geom(H, W, vec, 32, sigma > 0.f ? sigma : 1.f, reg);
  p.buf_bytes = (H * W * es + 127) / 128 * 128;
  p.nwarps = 0; p.nbufs = 0;   // chosen at launch
  p.direct_store = step_direct_store();
  static const int stagger = env_int("DSNT_TUNE_STEP_STAGGER", 0);
  p.stagger_ns = stagger;
  static const int debug = env_int("DSNT_TUNE_STEP_DEBUG", 0);
  p.debug = debug;
  p.pace = 0;   // set per kernel in launch_step2_nw
  p.out8 = out8; p.ws = ws;
  p.xc = xc ? *xc : no_peers();
  if (st) {
    p.st = *st;
  } else {
    p.st.count = 1; p.st.n_per = n;
    for (int k = 0; k < kMaxStacks; ++k) { p.st.z_off[k] = 0; p.st.dz_off[k] = 0; }
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return dtype == DSNT_DTYPE_F32 ? launch_step_reg<float, 4>(p, reg, s) : launch_step_reg<__nv_bfloat16, 8>(p, reg, s);
}

DSNT_API int dsnt_head_step(const void* z, int dtype, long n, int H, int W, const float* target, const float* mask,
                            const float* denom, const float* g_loss, float reg_coeff, int reg, float sigma, int flags,
                            float* coords, float* stats, float* terms, void* dz, void* stream) {
  if (!denom && n != 0) { set_error("dsnt_head_step: denom is required (dsnt_mask_count)"); return DSNT_ERR_BAD_ARG; }
  return head_step_impl(z, dtype, n, H, W, target, mask, denom, g_loss, reg_coeff, reg, sigma, flags, coords, stats, terms, dz,
                        nullptr, nullptr, stream);
}

DSNT_API int dsnt_head_step_pair_supported(int dtype, int H, int W, int reg) {
  if (dtype != DSNT_DTYPE_F32 && dtype != DSNT_DTYPE_BF16) return 0;
  return step_pair_supported(dtype, H, W, reg) ? 1 : 0;
}

DSNT_API int dsnt_head_step_supported_reg(int dtype, int H, int W, int reg) {
  if (reg < DSNT_REG_NONE || reg > DSNT_REG_MSE) return 0;
  if (dsnt_head_step_supported(dtype, H, W)) return 1;
  return step_pair_supported(dtype, H, W, reg) ? 1 : 0;
}

DSNT_API int dsnt_head_step_fused_supported(int dtype, int H, int W, int reg, float sigma) {
  return dsnt_head_step_supported(dtype, H, W) && step2_eligible(dtype, H, W, reg, sigma) >= 0 ? 1 : 0;
}

DSNT_API int dsnt_head_step_fused(const void* z, int dtype, long n, int H, int W, const float* target, const float* mask,
                                  const float* g_loss, float reg_coeff, int reg, float sigma, int flags, float* coords,
                                  float* stats, void* dz, float* out, float* workspace, void* stream) {
  if (!out || !workspace || !aligned(workspace, 16)) { set_error("dsnt_head_step_fused: out and a 16-byte aligned workspace are required"); return DSNT_ERR_BAD_ARG; }
  if (n == 0) {     // empty batch: the loss block of an empty reduction
    return dsnt_finish_loss(nullptr, mask, 0, reg_coeff, out, workspace, stream);
  }
  if (!dsnt_head_step_fused_supported(dtype, H, W, reg, sigma)) {
    set_error("dsnt_head_step_fused: %dx%d, dtype %d, reg %d is not served by the single-launch kernel; use dsnt_mask_count + "
              "dsnt_head_step + dsnt_finish_loss", H, W, dtype, reg);
    return DSNT_ERR_UNSUPPORTED;
  }
  return head_step_impl(z, dtype, n, H, W, target, mask, nullptr, g_loss, reg_coeff, reg, sigma, flags, coords, stats, nullptr, dz,
                        out, workspace, stream);
}

DSNT_API int dsnt_head_step_fused_peer(const void* z, int dtype, long n, int H, int W, const float* target, const float* mask,
                                       const float* g_loss, float reg_coeff, int reg, float sigma, int flags, float* coords,
                                       float* stats, void* dz, float* out, float* workspace, const void* const* peers,
                                       int rank, int world, unsigned* epoch, int* error, void* stream) {
  if (!out || !workspace || !aligned(workspace, 16)) { set_error("dsnt_head_step_fused_peer: out and a 16-byte aligned workspace are required"); return DSNT_ERR_BAD_ARG; }
  PeerXchg xc;
  const int rc = make_peers(peers, rank, world, epoch, error, xc);
  if (rc) return rc;
  if (n <= 0) { set_error("dsnt_head_step_fused_peer: an empty shard must take the three-launch form (every rank exchanges)"); return DSNT_ERR_UNSUPPORTED; }
  if (!dsnt_head_step_fused_supported(dtype, H, W, reg, sigma)) {
    set_error("dsnt_head_step_fused_peer: %dx%d, dtype %d, reg %d is not served by the single-launch kernel", H, W, dtype, reg);
    return DSNT_ERR_UNSUPPORTED;
  }
  return head_step_impl(z, dtype, n, H, W, target, mask, nullptr, g_loss, reg_coeff, reg, sigma, flags, coords, stats, nullptr, dz,
                        out, workspace, stream, nullptr, &xc);
}

DSNT_API int dsnt_head_step_fused_stacked(const void* const* z, void* const* dz, int n_stacks, int dtype, long n_per_stack,
                                          int H, int W, const float* target, const float* mask, const float* g_loss,
                                          float reg_coeff, int reg, float sigma, int flags, float* coords, float* stats,
                                          float* out, float* workspace, void* stream) {
  if (!out || !workspace || !aligned(workspace, 16)) { set_error("dsnt_head_step_fused_stacked: out and a 16-byte aligned workspace are required"); return DSNT_ERR_BAD_ARG; }
  if (n_stacks < 1 || n_stacks > kMaxStacks || n_per_stack < 0 || !z || !dz) { set_error("dsnt_head_step_fused_stacked: bad arguments"); return DSNT_ERR_BAD_ARG; }
  if (n_per_stack == 0) return dsnt_finish_loss(nullptr, mask, 0, reg_coeff, out, workspace, stream);
  if (!dsnt_head_step_fused_supported(dtype, H, W, reg, sigma)) {
    set_error("dsnt_head_step_fused_stacked: %dx%d, dtype %d, reg %d is not served by the single-launch kernel; use the "
              "dsnt_head_fwd_stacked / dsnt_head_bwd_stacked form", H, W, dtype, reg);
    return DSNT_ERR_UNSUPPORTED;
  }
  Stacks st;
  st.count = n_stacks; st.n_per = n_per_stack;
  for (int k = 0; k < kMaxStacks; ++k) { st.z_off[k] = 0; st.dz_off[k] = 0; }
  for (int k = 0; k < n_stacks; ++k) {
    if (!z[k] || !dz[k] || !aligned(z[k], 16) || !aligned(dz[k], 16)) { set_error("stack %d: null or misaligned heatmap pointer", k); return DSNT_ERR_BAD_ARG; }
    st.z_off[k] = static_cast<const char*>(z[k]) - static_cast<const char*>(z[0]);
    st.dz_off[k] = static_cast<char*>(dz[k]) - static_cast<char*>(dz[0]);
  }
  if (n_per_stack * n_stacks > 0x7fffffffL) { set_error("too many heatmaps for one stacked launch"); return DSNT_ERR_BAD_ARG; }
  return head_step_impl(z[0], dtype, n_per_stack * n_stacks, H, W, target, mask, nullptr, g_loss, reg_coeff, reg, sigma, flags,
                        coords, stats, nullptr, dz[0], out, workspace, stream, &st);
}

}  // extern "C"
