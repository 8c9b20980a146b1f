// infer.cu -- the inference path in front of the head: flip test-time augmentation fused into
// the forward-only head.  Replaces src/dsnt/inference.py:36-48 (reverse_tensor + index_select(HFLIP_INDICES) + mean
// of the two raw heatmap sets, then forward_part2) with ONE launch that reads both heatmap sets once.
#include "capi_util.cuh"
#include "head_preact.cuh"
#include "launch.cuh"

namespace dsnt {

template <typename T, int VEC, int GROUP, int NV>
static int launch_flip_one(const HeadPreactFwdParams& ps, cudaStream_t stream) {
  constexpr int BLOCK = fwd_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  const long grid = (ps.base.n + GPB - 1) / GPB;
  head_preact_fwd_kernel<T, VEC, GROUP, NV, true><<<static_cast<unsigned>(grid), BLOCK, 0, stream>>>(ps);
  return check_launch("head_preact_fwd_kernel<flip>");
}

template <typename T, int VEC>
static int launch_flip_shape(const HeadPreactFwdParams& ps, cudaStream_t stream) {
  const long nvec = static_cast<long>(ps.base.H) * ps.base.W / VEC;
  if (nvec <= 32 * 2) return launch_flip_one<T, VEC, 32, 2>(ps, stream);
  if (nvec <= 32 * 8) return launch_flip_one<T, VEC, 32, 8>(ps, stream);
  if (nvec <= 256 * 4) return launch_flip_one<T, VEC, 256, 4>(ps, stream);
  if (nvec <= 512 * 8) return launch_flip_one<T, VEC, 512, 8>(ps, stream);
  return launch_flip_one<T, VEC, 512, 0>(ps, stream);
}

}  // namespace dsnt

using namespace dsnt;
using bf16_t = __nv_bfloat16;

extern "C" {

DSNT_API int dsnt_flip_tta_fwd(const void* z, int dtype, long batch, int C, int H, int W, const int* flip_perm, int preact,
                               float threshold, float eps, float* coords, void* avg_out, void* stream) {
  if (C <= 0 || batch < 0) { set_error("bad shape batch=%ld C=%d", batch, C); return DSNT_ERR_BAD_ARG; }
  const long n = batch * C;
  int rc = check_common(z, dtype, n, H, W, DSNT_REG_NONE);
  if (rc) return rc;
  if (2 * n > 0x7fffffffL) { set_error("too many heatmaps: %ld", 2 * n); return DSNT_ERR_UNSUPPORTED; }
  if (preact < DSNT_PREACT_SOFTMAX || preact > DSNT_PREACT_SIGMOID) { set_error("bad preact %d", preact); return DSNT_ERR_BAD_ARG; }
  if (!(eps >= 0.f)) { set_error("eps must be >= 0"); return DSNT_ERR_BAD_ARG; }
  if (n == 0) return DSNT_OK;
  if (!coords || !aligned(coords, 8)) { set_error("coords output is required (8-byte aligned)"); return DSNT_ERR_BAD_ARG; }
  HeadPreactFwdParams ps;
  HeadFwdParams& p = ps.base;
  p.z = z; p.target = nullptr; p.coords = coords; p.stats = nullptr; p.terms = nullptr;
  p.n = n; p.H = H; p.W = W; p.reg = DSNT_REG_NONE; p.sigma = 1.f;
  p.st.count = 1; p.st.n_per = n;
  for (int s = 0; s < kMaxStacks; ++s) { p.st.z_off[s] = 0; p.st.dz_off[s] = 0; }
  ps.pc.preact = preact; ps.pc.threshold = threshold; ps.pc.eps = eps;
  ps.fl.perm = flip_perm; ps.fl.avg_out = avg_out; ps.fl.C = C;
  const int vec = pick_vec(dtype, W, z, avg_out);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // plain softmax (every model of the reference's experiments): the tuned streaming forward with the flip fused in
  if (preact == DSNT_PREACT_SOFTMAX && eps == 0.f && threshold == -INFINITY) {
    rc = 1;
    if (dtype == DSNT_DTYPE_F32 && vec == 4) rc = try_launch_fwd_fast_flip<float, 4>(p, ps.fl, s);
    if (dtype == DSNT_DTYPE_BF16 && vec == 8) rc = try_launch_fwd_fast_flip<bf16_t, 8>(p, ps.fl, s);
    if (rc != 1) return rc;
  }
  if (dtype == DSNT_DTYPE_F32)
    return vec == 4 ? launch_flip_shape<float, 4>(ps, s) : launch_flip_shape<float, 1>(ps, s);
  return vec == 8   ? launch_flip_shape<bf16_t, 8>(ps, s)
         : vec == 4 ? launch_flip_shape<bf16_t, 4>(ps, s)
                    : launch_flip_shape<bf16_t, 1>(ps, s);
}

}  // extern "C"
