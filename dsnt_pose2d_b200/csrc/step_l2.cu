// step_l2.cu -- the one-pass training step for heatmaps that do not fit the shared-memory ring of head_step*.cuh
// (256x256 fp32 is 256 KiB, more than the 227 KiB a CTA can have): the L2 is the staging buffer instead.
//
// The two-kernel contract runs the forward over ALL heatmaps and then the backward over all of them, so the backward's
// read of the logits comes from HBM again (2 GiB of logits at BASELINE config 5 against 126 MB of L2): 12 B per fp32
// pixel.  Here ONE persistent kernel runs forward and backward of the SAME heatmap back to back in the same CTA
// (head_fwd_fast_body, then head_bwd_fast_body of head_fast.cuh -- the tuned streaming kernels, unchanged arithmetic):
// the working set in flight is (CTAs resident) x (one heatmap), a few tens of MB, so the backward's read hits L2 and
// HBM sees 8 B per pixel.  As for dsnt_head_step, the denominator of masked_average is an input (dsnt_mask_count).
#include <cstdlib>

#include "launch.cuh"
#include "capi_util.cuh"

namespace dsnt {

__device__ float g_unit_gradient = 1.0f;     // d(loss) when the caller passes none (the backward reads it from memory)

template <typename T, int VEC, int REG>
__global__ void __launch_bounds__(256) head_step_l2_kernel(const HeadFwdFastParams pf, const HeadBwdFastParams pb) {
  for (long hb = blockIdx.x; hb < pf.base.n; hb += gridDim.x) {
    head_fwd_fast_body<T, VEC, 256, REG>(pf, hb);
    __syncthreads();      // the statistics this CTA wrote are visible to all of its threads; shared scratch is free again
    head_bwd_fast_body<T, VEC, 256, REG>(pb, hb);
    __syncthreads();
  }
}

static int l2_ctas_per_sm() {
  static const int v = [] { const char* e = std::getenv("DSNT_TUNE_STEP_L2_CTAS"); return e ? std::atoi(e) : 2; }();
  return v < 1 ? 1 : v;
}

template <typename T, int VEC, int REG>
static int launch_l2(const HeadFwdParams& f, const HeadBwdParams& b, cudaStream_t stream) {
  FastGeom fg;
  if (!make_fast_geom(f.H, f.W, VEC, 256, fg)) return 1;
  HeadFwdFastParams pf;
  pf.base = f;
  pf.g = make_geom(f.H, f.W, VEC, 256, f.sigma, REG);
  pf.f = fg;
  pf.fl = FlipCfg{nullptr, nullptr, 0};
  pf.pc = PreactCfg{DSNT_PREACT_SOFTMAX, 0.f, 0.f};
  if (!stash_fits(f.H, f.W, VEC, REG, pf.g.r2_win)) return 1;
  HeadBwdFastParams pb;
  pb.base = b;
  pb.g = pf.g;
  pb.f = fg;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  long ctas = static_cast<long>(sms) * l2_ctas_per_sm();
  if (ctas > f.n) ctas = f.n;
  head_step_l2_kernel<T, VEC, REG><<<static_cast<unsigned>(ctas), 256, 0, stream>>>(pf, pb);
  return check_launch("head_step_l2_kernel");
}

template <typename T, int VEC>
static int launch_l2_reg(const HeadFwdParams& f, const HeadBwdParams& b, int reg, cudaStream_t stream) {
  switch (reg) {
    case DSNT_REG_NONE: return launch_l2<T, VEC, DSNT_REG_NONE>(f, b, stream);
    case DSNT_REG_VAR: return launch_l2<T, VEC, DSNT_REG_VAR>(f, b, stream);
    default: break;
  }
  if constexpr (sizeof(T) == 2) {      // fp32 with a Gaussian window takes head_bwd_stream_kernel in the two-kernel path
    switch (reg) {
      case DSNT_REG_KL: return launch_l2<T, VEC, DSNT_REG_KL>(f, b, stream);
      case DSNT_REG_JS: return launch_l2<T, VEC, DSNT_REG_JS>(f, b, stream);
      case DSNT_REG_MSE: return launch_l2<T, VEC, DSNT_REG_MSE>(f, b, stream);
      default: break;
    }
  }
  return 1;
}

// Does the L2-staged step serve this case?  (CTA-per-heatmap layouts of the tuned kernels: more than 2048 vectors per
// heatmap, power-of-two vectors per row; fp32 only without a Gaussian window.)
bool step_l2_supported(int dtype, int H, int W, int reg) {
  const int vec = dtype == DSNT_DTYPE_F32 ? 4 : 8;
  if (W % vec != 0) return false;
  const long nvec = static_cast<long>(H) * W / vec;
  if (!stream_group_is_cta(nvec)) return false;
  FastGeom fg;
  if (!make_fast_geom(H, W, vec, 256, fg)) return false;
  if (dtype == DSNT_DTYPE_F32 && reg_needs_gauss(reg)) return false;
  return true;
}

// returns 1 when the case is not served (the caller reports DSNT_ERR_UNSUPPORTED)
int launch_step_l2(const void* z, int dtype, long n, int H, int W, const float* target, const float* mask,
                   const float* denom, const float* g_loss, float reg_coeff, int reg, float sigma, int flags,
                   float* coords, float* stats, float* terms, void* dz, cudaStream_t stream) {
  if (!step_l2_supported(dtype, H, W, reg) || !stats) return 1;
  if (!g_loss) {
    void* unit = nullptr;
    if (cudaGetSymbolAddress(&unit, g_unit_gradient) != cudaSuccess) return check_launch("head_step_l2_kernel (unit gradient)");
    g_loss = static_cast<const float*>(unit);
  }
  HeadFwdParams f;
  f.z = z; f.target = target; f.coords = coords; f.stats = stats; f.terms = terms;
  f.n = n; f.H = H; f.W = W; f.reg = reg; f.sigma = sigma;
  f.st.count = 1; f.st.n_per = n;
  for (int k = 0; k < kMaxStacks; ++k) { f.st.z_off[k] = 0; f.st.dz_off[k] = 0; }
  HeadBwdParams b;
  b.z = z; b.target = target; b.mask = mask; b.stats = stats; b.g_coords = nullptr; b.g_reg = nullptr;
  b.g_loss = g_loss; b.denom = denom; b.dz = dz; b.n = n; b.H = H; b.W = W; b.reg = reg; b.flags = flags;
  b.sigma = sigma; b.reg_coeff = reg_coeff; b.st = f.st;
  return dtype == DSNT_DTYPE_F32 ? launch_l2_reg<float, 4>(f, b, reg, stream)
                                 : launch_l2_reg<__nv_bfloat16, 8>(f, b, reg, stream);
}

}  // namespace dsnt
