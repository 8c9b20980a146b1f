// preact.cu -- launchers and extern "C" entry points of the fused head for the non-softmax pre-activations.
#include "capi_util.cuh"
#include "head_preact.cuh"
#include "launch.cuh"

namespace dsnt {

constexpr size_t kPreactMaxDynSmem = 48 * 1024;

template <typename T, int VEC, int GROUP, int NV>
static int launch_preact_fwd_one(const HeadPreactFwdParams& ps, cudaStream_t stream) {
  constexpr int BLOCK = fwd_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  const HeadFwdParams& p = ps.base;
  const size_t smem = reg_needs_gauss(p.reg) ? sizeof(float) * GPB * table_floats(p.H, p.W) : 0;
  if (smem > kPreactMaxDynSmem) { set_error("heatmap %dx%d: Gaussian tables exceed shared memory", p.H, p.W); return DSNT_ERR_UNSUPPORTED; }
  const long grid = (p.n + GPB - 1) / GPB;
  head_preact_fwd_kernel<T, VEC, GROUP, NV><<<static_cast<unsigned>(grid), BLOCK, smem, stream>>>(ps);
  return check_launch("head_preact_fwd_kernel");
}

// register-resident up to 4096 vectors (128x128 fp32), three L2-served passes beyond
template <typename T, int VEC>
static int launch_preact_fwd_shape(const HeadPreactFwdParams& ps, cudaStream_t stream) {
  const long nvec = static_cast<long>(ps.base.H) * ps.base.W / VEC;
  if (nvec <= 32 * 2) return launch_preact_fwd_one<T, VEC, 32, 2>(ps, stream);
  if (nvec <= 32 * 8) return launch_preact_fwd_one<T, VEC, 32, 8>(ps, stream);
  if (nvec <= 256 * 4) return launch_preact_fwd_one<T, VEC, 256, 4>(ps, stream);
  if (nvec <= 512 * 8) return launch_preact_fwd_one<T, VEC, 512, 8>(ps, stream);
  return launch_preact_fwd_one<T, VEC, 512, 0>(ps, stream);
}

template <typename T, int VEC, int GROUP, int NV>
static int launch_preact_bwd_one(const HeadPreactBwdParams& ps, cudaStream_t stream) {
  constexpr int BLOCK = fwd_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  const HeadBwdParams& p = ps.base;
  const size_t smem = reg_needs_gauss(p.reg) ? sizeof(float) * GPB * table_floats(p.H, p.W) : 0;
  if (smem > kPreactMaxDynSmem) { set_error("heatmap %dx%d: Gaussian tables exceed shared memory", p.H, p.W); return DSNT_ERR_UNSUPPORTED; }
  const long nvec = static_cast<long>(p.H) * p.W / VEC;
  const long chunks = (nvec + GROUP * NV - 1) / (GROUP * NV);
  if (chunks > 65535) { set_error("heatmap %dx%d too large", p.H, p.W); return DSNT_ERR_UNSUPPORTED; }
  dim3 grid(static_cast<unsigned>((p.n + GPB - 1) / GPB), static_cast<unsigned>(chunks));
  head_preact_bwd_kernel<T, VEC, GROUP, NV><<<grid, BLOCK, smem, stream>>>(ps);
  return check_launch("head_preact_bwd_kernel");
}

template <typename T, int VEC>
static int launch_preact_bwd_shape(const HeadPreactBwdParams& ps, cudaStream_t stream) {
  const long nvec = static_cast<long>(ps.base.H) * ps.base.W / VEC;
  if (nvec <= 32 * 2) return launch_preact_bwd_one<T, VEC, 32, 2>(ps, stream);
  if (nvec <= 32 * 8) return launch_preact_bwd_one<T, VEC, 32, 8>(ps, stream);
  return launch_preact_bwd_one<T, VEC, 256, 4>(ps, stream);  // chunked over blockIdx.y beyond 1024 vectors
}

static int check_preact(int preact, float eps) {
  if (preact < DSNT_PREACT_SOFTMAX || preact > DSNT_PREACT_SIGMOID) { set_error("bad preact %d", preact); return DSNT_ERR_BAD_ARG; }
  if (!(eps >= 0.f)) { set_error("eps must be >= 0"); return DSNT_ERR_BAD_ARG; }
  return DSNT_OK;
}

static void single_stack(Stacks& st, long n) {
  st.count = 1;
  st.n_per = n;
  for (int s = 0; s < kMaxStacks; ++s) { st.z_off[s] = 0; st.dz_off[s] = 0; }
}

}  // namespace dsnt

using namespace dsnt;
using bf16_t = __nv_bfloat16;

extern "C" {

DSNT_API int dsnt_head_preact_fwd(const void* z, int dtype, int preact, float threshold, float eps, long n, int H, int W,
                                  const float* target, int reg, float sigma, float* coords, float* stats, float* terms,
                                  int variant, void* stream) {
  int rc = check_common(z, dtype, n, H, W, reg);
  if (rc) return rc;
  rc = check_preact(preact, eps);
  if (rc) return rc;
  if (n == 0) return DSNT_OK;
  if (!coords) { set_error("coords output is required"); return DSNT_ERR_BAD_ARG; }
  if (reg_needs_gauss(reg) && !target) { set_error("reg %d needs a target", reg); return DSNT_ERR_BAD_ARG; }
  if (reg != DSNT_REG_NONE && !(sigma > 0.f)) { set_error("sigma must be > 0"); return DSNT_ERR_BAD_ARG; }
  if (!aligned(coords, 8) || (stats && !aligned(stats, 16)) || (terms && !aligned(terms, 8)) || (target && !aligned(target, 8))) {
    set_error("per-heatmap buffers must be naturally aligned (coords/terms/target 8 B, stats 16 B)");
    return DSNT_ERR_BAD_ARG;
  }
  HeadPreactFwdParams ps;
  HeadFwdParams& p = ps.base;
  p.z = z; p.target = target; p.coords = coords; p.stats = stats; p.terms = terms;
  p.n = n; p.H = H; p.W = W; p.reg = reg; p.sigma = sigma;
  single_stack(p.st, n);
  ps.pc.preact = preact; ps.pc.threshold = threshold; ps.pc.eps = eps;
  const int vec = pick_vec(dtype, W, z, nullptr);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ps.fl = FlipCfg{nullptr, nullptr, 0};
  if (variant == 0) {   // the tuned streaming kernels whenever the layout qualifies (returns 1 otherwise)
    rc = dtype == DSNT_DTYPE_F32 ? launch_preact_fast_fwd_f32(p, ps.pc, vec, s) : launch_preact_fast_fwd_bf16(p, ps.pc, vec, s);
    if (rc != 1) return rc;
  }
  if (dtype == DSNT_DTYPE_F32)
    return vec == 4 ? launch_preact_fwd_shape<float, 4>(ps, s) : launch_preact_fwd_shape<float, 1>(ps, s);
  return vec == 8   ? launch_preact_fwd_shape<bf16_t, 8>(ps, s)
         : vec == 4 ? launch_preact_fwd_shape<bf16_t, 4>(ps, s)
                    : launch_preact_fwd_shape<bf16_t, 1>(ps, s);
}

DSNT_API int dsnt_head_preact_bwd(const void* z, int dtype, int preact, float threshold, long n, int H, int W,
                                  const float* target, const float* mask, const float* stats, const float* g_coords,
                                  const float* g_reg, const float* g_loss, const float* denom, float reg_coeff, int reg,
                                  float sigma, int flags, void* dz, int variant, void* stream) {
  int rc = check_common(dz, dtype, n, H, W, reg);
  if (rc) return rc;
  rc = check_preact(preact, 0.f);
  if (rc) return rc;
  if (n == 0) return DSNT_OK;
  if (!z) { set_error("null heatmap pointer"); return DSNT_ERR_BAD_ARG; }
  if (!stats) { set_error("stats from the forward are required"); return DSNT_ERR_BAD_ARG; }
  if (reg_needs_gauss(reg) && !target) { set_error("reg %d needs a target", reg); return DSNT_ERR_BAD_ARG; }
  if ((g_loss == nullptr) != (denom == nullptr)) { set_error("g_loss and denom go together"); return DSNT_ERR_BAD_ARG; }
  if (reg != DSNT_REG_NONE && !(sigma > 0.f)) { set_error("sigma must be > 0"); return DSNT_ERR_BAD_ARG; }
  if (!aligned(stats, 16) || (target && !aligned(target, 8)) || (g_coords && !aligned(g_coords, 8))) {
    set_error("per-heatmap buffers must be naturally aligned (target/g_coords 8 B, stats 16 B)");
    return DSNT_ERR_BAD_ARG;
  }
  HeadPreactBwdParams ps;
  HeadBwdParams& p = ps.base;
  p.z = z; p.target = target; p.mask = mask; p.stats = stats; p.g_coords = g_coords; p.g_reg = g_reg;
  p.g_loss = g_loss; p.denom = denom; p.dz = dz; p.n = n; p.H = H; p.W = W; p.reg = reg; p.flags = flags;
  p.sigma = sigma; p.reg_coeff = reg_coeff;
  single_stack(p.st, n);
  ps.pc.preact = preact; ps.pc.threshold = threshold; ps.pc.eps = 0.f;
  const int vec = pick_vec(dtype, W, dz, z);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (variant == 0) {
    rc = dtype == DSNT_DTYPE_F32 ? launch_preact_fast_bwd_f32(p, ps.pc, vec, s) : launch_preact_fast_bwd_bf16(p, ps.pc, vec, s);
    if (rc != 1) return rc;
  }
  if (dtype == DSNT_DTYPE_F32)
    return vec == 4 ? launch_preact_bwd_shape<float, 4>(ps, s) : launch_preact_bwd_shape<float, 1>(ps, s);
  return vec == 8   ? launch_preact_bwd_shape<bf16_t, 8>(ps, s)
         : vec == 4 ? launch_preact_bwd_shape<bf16_t, 4>(ps, s)
                    : launch_preact_bwd_shape<bf16_t, 1>(ps, s);
}

}  // extern "C"
