// capi.cu -- the extern "C" surface of libdsnt_b200.so (see include/dsnt_b200.h for the contract).
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "aux_kernels.cuh"
#include "launch.cuh"
#include "capi_util.cuh"

namespace dsnt {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DSNT_ERR_LAUNCH;
  }
  return DSNT_OK;
}

}  // namespace dsnt

using namespace dsnt;

template <typename T>
static int launch_reg_dmu(const void* z, int input_is_logits, const float* stats, const float* mu, const float* mask,
                          const float* g_loss, const float* denom, float reg_coeff, long n, int W, int H, float sigma, int reg,
                          float* dmu, size_t smem, cudaStream_t s) {
  const T* zt = static_cast<const T*>(z);
  const unsigned grid = static_cast<unsigned>(n);
  switch (reg) {
    case DSNT_REG_KL:
      reg_dmu_kernel<T, DSNT_REG_KL><<<grid, kGaussBlock, smem, s>>>(zt, input_is_logits, stats, mu, mask, g_loss, denom, reg_coeff, W, H, sigma, dmu);
      break;
    case DSNT_REG_JS:
      reg_dmu_kernel<T, DSNT_REG_JS><<<grid, kGaussBlock, smem, s>>>(zt, input_is_logits, stats, mu, mask, g_loss, denom, reg_coeff, W, H, sigma, dmu);
      break;
    default:
      reg_dmu_kernel<T, DSNT_REG_MSE><<<grid, kGaussBlock, smem, s>>>(zt, input_is_logits, stats, mu, mask, g_loss, denom, reg_coeff, W, H, sigma, dmu);
      break;
  }
  return check_launch("reg_dmu_kernel");
}

extern "C" {

DSNT_API int dsnt_b200_version(void) { return DSNT_B200_VERSION; }

DSNT_API const char* dsnt_b200_last_error(void) { return g_err; }

// Fills the stack view (byte offsets relative to stack 0) and returns the widest vector every stack allows.
static int fill_stacks(Stacks& st, const void* const* z, void* const* dz, int n_stacks, long n_per, int dtype, int W,
                       bool need_z, int& vec) {
  if (n_stacks < 1 || n_stacks > kMaxStacks) { set_error("n_stacks must be in [1, %d], got %d", kMaxStacks, n_stacks); return DSNT_ERR_BAD_ARG; }
  st.count = n_stacks;
  st.n_per = n_per;
  vec = 8;
  for (int s = 0; s < kMaxStacks; ++s) { st.z_off[s] = 0; st.dz_off[s] = 0; }
  for (int s = 0; s < n_stacks; ++s) {
    const void* zs = z ? z[s] : nullptr;
    void* ds = dz ? dz[s] : nullptr;
    if ((need_z && !zs) || (dz && !ds)) { set_error("null heatmap pointer (stack %d)", s); return DSNT_ERR_BAD_ARG; }
    if (zs) st.z_off[s] = static_cast<const char*>(zs) - static_cast<const char*>(z[0]);
    if (ds) st.dz_off[s] = static_cast<char*>(ds) - static_cast<char*>(dz[0]);
    const int v = pick_vec(dtype, W, dz ? static_cast<const void*>(ds) : zs, dz && need_z ? zs : nullptr);
    if (v < vec) vec = v;
  }
  return DSNT_OK;
}

DSNT_API int dsnt_head_fwd_stacked(const void* const* z, int n_stacks, int dtype, int input_is_logits, long n_per_stack,
                                   int H, int W, const float* target, int reg, float sigma, float* coords, float* stats,
                                   float* terms, int variant, void* stream) {
  if (!z) { set_error("null pointer array"); return DSNT_ERR_BAD_ARG; }
  const long n = n_per_stack * static_cast<long>(n_stacks > 0 ? n_stacks : 1);
  int rc = check_common(n_per_stack == 0 ? reinterpret_cast<const void*>(1) : z[0], dtype, n, H, W, reg);
  if (rc) return rc;
  if (n == 0) return DSNT_OK;
  if (!coords) { set_error("coords output is required"); return DSNT_ERR_BAD_ARG; }
  if (reg_needs_gauss(reg) && !target) { set_error("reg %d needs a target", reg); return DSNT_ERR_BAD_ARG; }
  if (reg != DSNT_REG_NONE && !(sigma > 0.f)) { set_error("sigma must be > 0"); return DSNT_ERR_BAD_ARG; }
  if (!aligned(coords, 8) || (stats && !aligned(stats, 16)) || (terms && !aligned(terms, 8)) || (target && !aligned(target, 8))) {
    set_error("per-heatmap buffers must be naturally aligned (coords/terms/target 8 B, stats 16 B)");
    return DSNT_ERR_BAD_ARG;
  }
  HeadFwdParams p;
  p.z = z[0]; p.target = target; p.coords = coords; p.stats = stats; p.terms = terms;
  p.n = n; p.H = H; p.W = W; p.reg = reg; p.sigma = sigma;
  int vec = 1;
  rc = fill_stacks(p.st, z, nullptr, n_stacks, n_per_stack, dtype, W, true, vec);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return dtype == DSNT_DTYPE_F32 ? launch_head_fwd_f32(p, vec, input_is_logits != 0, variant, s)
                                 : launch_head_fwd_bf16(p, vec, input_is_logits != 0, variant, s);
}

DSNT_API int dsnt_head_fwd(const void* z, int dtype, int input_is_logits, long n, int H, int W, const float* target, int reg,
                  float sigma, float* coords, float* stats, float* terms, int variant, void* stream) {
  if (!z && n != 0) { set_error("null heatmap pointer"); return DSNT_ERR_BAD_ARG; }
  const void* zs[1] = {z};
  return dsnt_head_fwd_stacked(zs, 1, dtype, input_is_logits, n, H, W, target, reg, sigma, coords, stats, terms, variant, stream);
}

DSNT_API int dsnt_head_bwd_stacked(const void* const* z, void* const* dz, int n_stacks, int dtype, int input_is_logits,
                                   long n_per_stack, int H, int W, const float* target, const float* mask,
                                   const float* stats, const float* g_coords, const float* g_reg, const float* g_loss,
                                   const float* denom, float reg_coeff, int reg, float sigma, int flags, int variant,
                                   void* stream) {
  if (!dz) { set_error("null pointer array"); return DSNT_ERR_BAD_ARG; }
  const long n = n_per_stack * static_cast<long>(n_stacks > 0 ? n_stacks : 1);
  int rc = check_common(n_per_stack == 0 ? reinterpret_cast<const void*>(1) : dz[0], dtype, n, H, W, reg);
  if (rc) return rc;
  if (n == 0) return DSNT_OK;
  const bool need_z = input_is_logits || reg_needs_gauss(reg);
  if (need_z && !z) { set_error("null heatmap pointer"); return DSNT_ERR_BAD_ARG; }
  if (!stats) { set_error("stats from the forward are required"); return DSNT_ERR_BAD_ARG; }
  if (reg_needs_gauss(reg) && !target) { set_error("reg %d needs a target", reg); return DSNT_ERR_BAD_ARG; }
  if ((g_loss == nullptr) != (denom == nullptr)) { set_error("g_loss and denom go together"); return DSNT_ERR_BAD_ARG; }
  if (reg != DSNT_REG_NONE && !(sigma > 0.f)) { set_error("sigma must be > 0"); return DSNT_ERR_BAD_ARG; }
  if (!aligned(stats, 16) || (target && !aligned(target, 8)) || (g_coords && !aligned(g_coords, 8))) {
    set_error("per-heatmap buffers must be naturally aligned (target/g_coords 8 B, stats 16 B)");
    return DSNT_ERR_BAD_ARG;
  }
  HeadBwdParams p;
  p.z = need_z ? z[0] : nullptr; p.target = target; p.mask = mask; p.stats = stats; p.g_coords = g_coords; p.g_reg = g_reg;
  p.g_loss = g_loss; p.denom = denom; p.dz = dz[0]; p.n = n; p.H = H; p.W = W; p.reg = reg; p.flags = flags;
  p.sigma = sigma; p.reg_coeff = reg_coeff;
  int vec = 1;
  rc = fill_stacks(p.st, need_z ? z : nullptr, dz, n_stacks, n_per_stack, dtype, W, need_z, vec);
  if (rc) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return dtype == DSNT_DTYPE_F32 ? launch_head_bwd_f32(p, vec, input_is_logits != 0, variant, s)
                                 : launch_head_bwd_bf16(p, vec, input_is_logits != 0, variant, s);
}

DSNT_API int dsnt_head_bwd(const void* z, int dtype, int input_is_logits, long n, int H, int W, const float* target,
                  const float* mask, const float* stats, const float* g_coords, const float* g_reg,
                  const float* g_loss, const float* denom, float reg_coeff, int reg, float sigma, int flags, void* dz,
                  int variant, void* stream) {
  if (!dz && n != 0) { set_error("null heatmap pointer"); return DSNT_ERR_BAD_ARG; }
  const void* zs[1] = {z};
  void* dzs[1] = {dz};
  return dsnt_head_bwd_stacked(z ? zs : nullptr, dzs, 1, dtype, input_is_logits, n, H, W, target, mask, stats, g_coords,
                               g_reg, g_loss, denom, reg_coeff, reg, sigma, flags, variant, stream);
}

DSNT_API int dsnt_finish_workspace_bytes(void) { return static_cast<int>(sizeof(float) * kFinishWorkspaceFloats); }

DSNT_API int dsnt_finish_trace_offset_bytes(void) { return static_cast<int>(sizeof(float) * kFinishTrace); }

DSNT_API int dsnt_finish_loss_stacked(const float* terms, const float* mask, long n_per_stack, int n_stacks, float reg_coeff,
                                      float* out, float* workspace, void* stream) {
  if (n_stacks < 1 || n_per_stack < 0) { set_error("dsnt_finish_loss: bad arguments"); return DSNT_ERR_BAD_ARG; }
  const long n = n_per_stack * n_stacks;
  if ((!terms && n > 0) || !out || !workspace) { set_error("dsnt_finish_loss: bad arguments"); return DSNT_ERR_BAD_ARG; }
  if (!aligned(terms, 8) || !aligned(workspace, 16)) { set_error("dsnt_finish_loss: misaligned buffers"); return DSNT_ERR_BAD_ARG; }
  long ctas = (n + 1023) / 1024;  // >= 4 heatmaps per thread before adding CTAs
  if (ctas < 1) ctas = 1;
  if (ctas > kFinishMaxCtas) ctas = kFinishMaxCtas;
  finish_loss_kernel<<<static_cast<unsigned>(ctas), kFinishBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      terms, mask, n, n_per_stack > 0 ? n_per_stack : 1, reg_coeff, out, workspace, no_peers());
  return check_launch("finish_loss_kernel");
}

DSNT_API int dsnt_peer_exchange_bytes(void) { return kPeerExchangeBytes; }

DSNT_API int dsnt_finish_loss_peer(const float* terms, const float* mask, long n_per_stack, int n_stacks, float reg_coeff,
                                   float* out, float* workspace, const void* const* peers, int rank, int world,
                                   unsigned* epoch, int* error, void* stream) {
  if (n_stacks < 1 || n_per_stack < 0) { set_error("dsnt_finish_loss_peer: bad arguments"); return DSNT_ERR_BAD_ARG; }
  const long n = n_per_stack * n_stacks;
  if ((!terms && n > 0) || !out || !workspace) { set_error("dsnt_finish_loss_peer: bad arguments"); return DSNT_ERR_BAD_ARG; }
  if (!aligned(terms, 8) || !aligned(workspace, 16)) { set_error("dsnt_finish_loss_peer: misaligned buffers"); return DSNT_ERR_BAD_ARG; }
  PeerXchg xc;
  const int rc = make_peers(peers, rank, world, epoch, error, xc);
  if (rc) return rc;
  long ctas = (n + 1023) / 1024;
  if (ctas < 1) ctas = 1;       // an empty shard still takes part in the exchange
  if (ctas > kFinishMaxCtas) ctas = kFinishMaxCtas;
  finish_loss_kernel<<<static_cast<unsigned>(ctas), kFinishBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      terms, mask, n, n_per_stack > 0 ? n_per_stack : 1, reg_coeff, out, workspace, xc);
  return check_launch("finish_loss_kernel<peer>");
}

DSNT_API int dsnt_mask_count_peer(const float* mask, long n, float* out, float* workspace, const void* const* peers,
                                  int rank, int world, unsigned* epoch, int* error, void* stream) {
  if (n < 0 || !out || !workspace || !aligned(workspace, 16)) { set_error("dsnt_mask_count_peer: bad arguments"); return DSNT_ERR_BAD_ARG; }
  PeerXchg xc;
  const int rc = make_peers(peers, rank, world, epoch, error, xc);
  if (rc) return rc;
  long ctas = (n + 4095) / 4096;
  if (ctas < 1) ctas = 1;
  if (ctas > kFinishMaxCtas) ctas = kFinishMaxCtas;
  finish_loss_kernel<<<static_cast<unsigned>(ctas), kFinishBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      nullptr, mask, n, n > 0 ? n : 1, 0.f, out, workspace, xc);
  return check_launch("finish_loss_kernel<count, peer>");
}

DSNT_API int dsnt_finish_loss(const float* terms, const float* mask, long n, float reg_coeff, float* out, float* workspace,
                     void* stream) {
  if (n < 0) { set_error("dsnt_finish_loss: bad arguments"); return DSNT_ERR_BAD_ARG; }
  return dsnt_finish_loss_stacked(terms, mask, n, 1, reg_coeff, out, workspace, stream);
}

DSNT_API int dsnt_mask_count(const float* mask, long n, float* out, float* workspace, void* stream) {
  if (n < 0 || !out || !workspace || !aligned(workspace, 16)) { set_error("dsnt_mask_count: bad arguments"); return DSNT_ERR_BAD_ARG; }
  long ctas = (n + 4095) / 4096;
  if (ctas < 1) ctas = 1;
  if (ctas > kFinishMaxCtas) ctas = kFinishMaxCtas;
  finish_loss_kernel<<<static_cast<unsigned>(ctas), kFinishBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      nullptr, mask, n, n > 0 ? n : 1, 0.f, out, workspace, no_peers());
  return check_launch("finish_loss_kernel<count>");
}

DSNT_API int dsnt_scale_unless_one(void* x, int dtype, long numel, const float* g, void* stream) {
  if (numel < 0 || !g || (!x && numel > 0)) { set_error("dsnt_scale_unless_one: bad arguments"); return DSNT_ERR_BAD_ARG; }
  if (numel == 0) return DSNT_OK;
  const bool al = aligned(x, 16);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // One CTA of 1024 threads per SM: in the usual case (*g == 1) the kernel is pure launch overhead, which grows with the
  // number of CTAs; when it does scale, 4 independent 128-bit loads per thread keep 64 KiB per SM in flight.
  const unsigned ctas = static_cast<unsigned>(sm_count_of_current_device());
  if (dtype == DSNT_DTYPE_F32) {
    const long nv = al ? numel / 4 : 0;
    scale_unless_one_kernel<float><<<ctas, kScaleBlock, 0, s>>>(static_cast<float*>(x), nv, numel, g);
  } else if (dtype == DSNT_DTYPE_BF16) {
    const long nv = al ? numel / 8 : 0;
    scale_unless_one_kernel<__nv_bfloat16><<<ctas, kScaleBlock, 0, s>>>(static_cast<__nv_bfloat16*>(x), nv, numel, g);
  } else { set_error("unsupported dtype %d", dtype); return DSNT_ERR_UNSUPPORTED; }
  return check_launch("scale_unless_one_kernel");
}

DSNT_API int dsnt_combine_loss(float* out, float reg_coeff, void* stream) {
  if (!out) { set_error("dsnt_combine_loss: null"); return DSNT_ERR_BAD_ARG; }
  combine_loss_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(out, reg_coeff);
  return check_launch("combine_loss_kernel");
}

DSNT_API int dsnt_euclid_fwd(const float* actual, const float* target, long n, int d, float* terms, void* stream) {
  if (!actual || !target || !terms || n < 0 || d <= 0 || !aligned(terms, 8)) { set_error("dsnt_euclid_fwd: bad arguments"); return DSNT_ERR_BAD_ARG; }
  if (n == 0) return DSNT_OK;
  euclid_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(actual, target, n, d, terms);
  return check_launch("euclid_fwd_kernel");
}

DSNT_API int dsnt_euclid_bwd(const float* actual, const float* target, const float* terms, const float* mask,
                             const float* g_loss, const float* denom, long n, int d, int flags, float* g_actual,
                             void* stream) {
  if (!actual || !target || !terms || !g_loss || !denom || !g_actual || n < 0 || d <= 0) { set_error("dsnt_euclid_bwd: bad arguments"); return DSNT_ERR_BAD_ARG; }
  if (n == 0) return DSNT_OK;
  euclid_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      actual, target, terms, mask, g_loss, denom, n, d, flags, g_actual);
  return check_launch("euclid_bwd_kernel");
}

DSNT_API int dsnt_tsoftmax_fwd(const void* x, int dtype, long rows, long len, float threshold, float eps, void* out,
                      void* stream) {
  if (!x || !out || rows < 0 || len <= 0 || rows > 0x7fffffffL) { set_error("dsnt_tsoftmax_fwd: bad arguments"); return DSNT_ERR_BAD_ARG; }
  if (rows == 0) return DSNT_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == DSNT_DTYPE_F32)
    tsoftmax_fwd_kernel<float><<<static_cast<unsigned>(rows), kRowBlock, 0, s>>>(static_cast<const float*>(x), len, threshold, eps, static_cast<float*>(out));
  else if (dtype == DSNT_DTYPE_BF16)
    tsoftmax_fwd_kernel<__nv_bfloat16><<<static_cast<unsigned>(rows), kRowBlock, 0, s>>>(static_cast<const __nv_bfloat16*>(x), len, threshold, eps, static_cast<__nv_bfloat16*>(out));
  else { set_error("unsupported dtype %d", dtype); return DSNT_ERR_UNSUPPORTED; }
  return check_launch("tsoftmax_fwd_kernel");
}

DSNT_API int dsnt_tsoftmax_bwd(const void* out, const void* g, int dtype, long rows, long len, void* dx, void* stream) {
  if (!out || !g || !dx || rows < 0 || len <= 0 || rows > 0x7fffffffL) { set_error("dsnt_tsoftmax_bwd: bad arguments"); return DSNT_ERR_BAD_ARG; }
  if (rows == 0) return DSNT_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == DSNT_DTYPE_F32)
    tsoftmax_bwd_kernel<float><<<static_cast<unsigned>(rows), kRowBlock, 0, s>>>(static_cast<const float*>(out), static_cast<const float*>(g), len, static_cast<float*>(dx));
  else if (dtype == DSNT_DTYPE_BF16)
    tsoftmax_bwd_kernel<__nv_bfloat16><<<static_cast<unsigned>(rows), kRowBlock, 0, s>>>(static_cast<const __nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(g), len, static_cast<__nv_bfloat16*>(dx));
  else { set_error("unsupported dtype %d", dtype); return DSNT_ERR_UNSUPPORTED; }
  return check_launch("tsoftmax_bwd_kernel");
}

DSNT_API int dsnt_make_gauss_fwd(const float* mu, long n, int W, int H, float sigma, float* out, void* stream) {
  if (!mu || !out || n < 0 || W <= 0 || H <= 0 || !(sigma > 0.f) || n > 0x7fffffffL) { set_error("dsnt_make_gauss_fwd: bad arguments"); return DSNT_ERR_BAD_ARG; }
  if (n == 0) return DSNT_OK;
  const size_t smem = sizeof(float) * table_floats(H, W);
  if (smem > kMaxDynSmem) { set_error("gaussian %dx%d too large", H, W); return DSNT_ERR_UNSUPPORTED; }
  make_gauss_fwd_kernel<<<static_cast<unsigned>(n), kGaussBlock, smem, static_cast<cudaStream_t>(stream)>>>(mu, W, H, sigma, out);
  return check_launch("make_gauss_fwd_kernel");
}

DSNT_API int dsnt_make_gauss_bwd(const float* mu, const float* g, long n, int W, int H, float sigma, float* dmu, void* stream) {
  if (!mu || !g || !dmu || n < 0 || W <= 0 || H <= 0 || !(sigma > 0.f) || n > 0x7fffffffL) { set_error("dsnt_make_gauss_bwd: bad arguments"); return DSNT_ERR_BAD_ARG; }
  if (n == 0) return DSNT_OK;
  const size_t smem = sizeof(float) * table_floats(H, W);
  if (smem > kMaxDynSmem) { set_error("gaussian %dx%d too large", H, W); return DSNT_ERR_UNSUPPORTED; }
  make_gauss_bwd_kernel<<<static_cast<unsigned>(n), kGaussBlock, smem, static_cast<cudaStream_t>(stream)>>>(mu, g, W, H, sigma, dmu);
  return check_launch("make_gauss_bwd_kernel");
}

DSNT_API int dsnt_reg_dmu(const void* z, int dtype, int input_is_logits, long n, int H, int W, const float* stats,
                          const float* mu, const float* mask, const float* g_loss, const float* denom, float reg_coeff,
                          int reg, float sigma, float* dmu, void* stream) {
  if (!z || !mu || !g_loss || !denom || !dmu || n < 0 || W <= 0 || H <= 0 || !(sigma > 0.f) || n > 0x7fffffffL) { set_error("dsnt_reg_dmu: bad arguments"); return DSNT_ERR_BAD_ARG; }
  if (reg != DSNT_REG_KL && reg != DSNT_REG_JS && reg != DSNT_REG_MSE) { set_error("dsnt_reg_dmu: reg %d has no Gaussian target", reg); return DSNT_ERR_BAD_ARG; }
  if (input_is_logits && !stats) { set_error("dsnt_reg_dmu: logits need the statistics of the forward"); return DSNT_ERR_BAD_ARG; }
  if (n == 0) return DSNT_OK;
  const size_t smem = sizeof(float) * table_floats(H, W);
  if (smem > kMaxDynSmem) { set_error("gaussian %dx%d too large", H, W); return DSNT_ERR_UNSUPPORTED; }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == DSNT_DTYPE_F32) return launch_reg_dmu<float>(z, input_is_logits, stats, mu, mask, g_loss, denom, reg_coeff, n, W, H, sigma, reg, dmu, smem, s);
  if (dtype == DSNT_DTYPE_BF16) return launch_reg_dmu<__nv_bfloat16>(z, input_is_logits, stats, mu, mask, g_loss, denom, reg_coeff, n, W, H, sigma, reg, dmu, smem, s);
  set_error("unsupported dtype %d", dtype);
  return DSNT_ERR_UNSUPPORTED;
}

}  // extern "C"
