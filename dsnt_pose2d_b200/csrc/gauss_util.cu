// gauss_util.cu -- the 'gauss' output-strategy helpers of the reference on the GPU.
//
//   dsnt_draw_gaussians    replaces encode_heatmaps + draw_gaussian (src/dsnt/util.py:70-148): a CPU double loop over
//                          (sample, joint) followed by an H2D copy of the whole target tensor on every training step
//                          (src/dsnt/model.py:148-154,247-256).  Here: one CTA per heatmap, a pure streaming write.
//   dsnt_decode_heatmaps   replaces get_preds + decode_heatmaps (src/dsnt/util.py:151-198): D2H of the whole heatmap
//                          tensor + CPU argmax + a Python loop for the quarter-pixel offset (src/dsnt/model.py:165,269).
//                          Here: one streaming read with an (value, first index) arg-max reduction per heatmap.
//
// Both follow the reference's float32 arithmetic operation by operation (separately rounded add/mul, Python's
// round-half-to-even, int() truncation, and the y = idx / HEIGHT quirk of get_preds, util.py:163), so pixel
// choices and rounding decisions are identical; only exp() may differ in the last bit.
#include "capi_util.cuh"

namespace dsnt {

constexpr int kDrawBlock = 256;

struct DrawParams {
  const float* centres;  // [N,2]
  float* out;            // [N,H,W]
  long n;
  int H, W;
  int mode;              // 0: centres are NORMALISED coords (encode_heatmaps); 1: pixel coordinates (draw_gaussian)
  int normalize;
  float kf;              // -0.5 (1/sigma)^2 rounded to float32 (tensor.mul_(python float), util.py:116,119)
  float radius;          // clip_size/2, or max(W,H) when unclipped (util.py:96-98)
  float half_w, half_h;  // W/2, H/2 (util.py:134-135)
};

// Window of the reference's draw_gaussian (util.py:100-109); returns false when nothing is drawn.
__device__ __forceinline__ bool draw_window(const DrawParams& p, int x, int y, int& sx, int& ex, int& sy, int& ey) {
  const float r = p.radius, xf = static_cast<float>(x), yf = static_cast<float>(y);
  if (r < 0.5f || xf <= -r || yf <= -r || xf >= static_cast<float>(p.W - 1) + r || yf >= static_cast<float>(p.H - 1) + r)
    return false;
  sx = max(0, static_cast<int>(ceilf(xf - r)));
  ex = min(p.W, static_cast<int>(xf + r + 1.0f));     // int(): truncation; the argument is positive here
  sy = max(0, static_cast<int>(ceilf(yf - r)));
  ey = min(p.H, static_cast<int>(yf + r + 1.0f));
  return true;
}

__device__ __forceinline__ float draw_value(const DrawParams& p, int i, int j, int x, int y) {
  const int dx = j - x, dy = i - y;
  // (xs - x)^2 + (ys - y)^2 is exact in float32; then ONE rounded multiply by k and exp (util.py:117-120)
  return expf(__fmul_rn(static_cast<float>(dx * dx + dy * dy), p.kf));
}

template <int VEC>
__global__ void __launch_bounds__(kDrawBlock) draw_gaussians_kernel(const DrawParams p) {
  __shared__ float red[kDrawBlock / 32];
  const long hm = blockIdx.x;
  const float2 c = __ldg(reinterpret_cast<const float2*>(p.centres) + hm);
  int x, y;
  if (p.mode == 0) {
    // coords.add_(1); [:, :, 0].mul_(W/2); [:, :, 1].mul_(H/2); coords.add_(-0.5)  -- four separately rounded float32
    // ops (util.py:133-136), then Python round(): half to even (util.py:142-143), then int() in draw_gaussian
    x = static_cast<int>(rintf(__fadd_rn(__fmul_rn(__fadd_rn(c.x, 1.0f), p.half_w), -0.5f)));
    y = static_cast<int>(rintf(__fadd_rn(__fmul_rn(__fadd_rn(c.y, 1.0f), p.half_h), -0.5f)));
  } else {
    x = static_cast<int>(c.x);   // int(x): truncation toward zero (util.py:84-85)
    y = static_cast<int>(c.y);
  }
  int sx = 0, ex = 0, sy = 0, ey = 0;
  const bool drawn = draw_window(p, x, y, sx, ex, sy, ey);

  float tot = 0.f;
  if (p.normalize && drawn) {   // val_sum = subimg.sum(); if val_sum > 0: subimg.div_(val_sum)  (util.py:122-125)
    const int ww = ex - sx, npx = ww * (ey - sy);
    float s = 0.f;
    for (int k = threadIdx.x; k < npx; k += kDrawBlock) {
      const int r = k / ww;
      s += draw_value(p, sy + r, sx + (k - r * ww), x, y);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kDrawBlock / 32; ++w) tot += red[w];
  }

  float* ob = p.out + hm * static_cast<long>(p.H) * p.W;
  const int wv = p.W / VEC, nvec = p.H * wv;
  VecWalker wk(threadIdx.x, kDrawBlock, wv);
  for (int f = threadIdx.x; f < nvec; f += kDrawBlock) {
    const int i = wk.row, j0 = wk.cv * VEC;
    float v[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int j = j0 + e;
      const bool in = drawn && i >= sy && i < ey && j >= sx && j < ex;
      float val = 0.f;
      if (in) {
        val = draw_value(p, i, j, x, y);
        if (tot > 0.f) val = __fdiv_rn(val, tot);   // subimg.div_(val_sum), only when normalising and the sum is > 0
      }
      v[e] = val;
    }
    VecIO<float, VEC>::store(ob, static_cast<long>(f) * VEC, v);
    wk.next();
  }
}

// ------------------------------------------------------------------------------------------------ decode
struct DecodeParams {
  const void* hm;
  float* coords;    // [N,2]
  long n;
  int H, W;
  int use_neighbours;
  float two_over_w, two_over_h;   // 2/W, 2/H rounded to float32 (tensor.mul_(python float), util.py:196-197)
};

template <typename T>
__device__ __forceinline__ float load_one(const T* p, long i);
template <>
__device__ __forceinline__ float load_one<float>(const float* p, long i) { return __ldg(p + i); }
template <>
__device__ __forceinline__ float load_one<__nv_bfloat16>(const __nv_bfloat16* p, long i) { return __bfloat162float(p[i]); }

// (value, index) arg-max where the FIRST maximal index wins (torch.max, util.py:154).
__device__ __forceinline__ void argmax_merge(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

template <typename T, int VEC, int GROUP>
__global__ void __launch_bounds__(GROUP >= 64 ? GROUP : 128) decode_heatmaps_kernel(const DecodeParams p) {
  constexpr int BLOCK = GROUP >= 64 ? GROUP : 128;
  constexpr int GPB = BLOCK / GROUP;
  constexpr int NW = GROUP / 32;
  __shared__ float red_v[GPB * NW];
  __shared__ int red_i[GPB * NW];
  const int tid = threadIdx.x;
  const int gid = tid / GROUP, lane_g = tid % GROUP, warp_g = lane_g >> 5, lane = tid & 31;
  const long hm = static_cast<long>(blockIdx.x) * GPB + gid;
  if (hm >= p.n) return;   // GROUP == 32 only
  const int H = p.H, W = p.W;
  const long npx = static_cast<long>(H) * W;
  const T* zb = reinterpret_cast<const T*>(p.hm) + hm * npx;

  float best = -INFINITY;
  int bidx = 0x7fffffff;
  const int nvec = static_cast<int>(npx / VEC);
  constexpr int U = 4;
  int f = lane_g;
  for (; f + (U - 1) * GROUP < nvec; f += U * GROUP) {
    float v[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) VecIO<T, VEC>::load(zb, static_cast<long>(f + u * GROUP) * VEC, v[u]);
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int e = 0; e < VEC; ++e)
        if (v[u][e] > best) { best = v[u][e]; bidx = (f + u * GROUP) * VEC + e; }   // ascending order: strict > keeps the first
  }
  for (; f < nvec; f += GROUP) {
    float v[VEC];
    VecIO<T, VEC>::load(zb, static_cast<long>(f) * VEC, v);
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      if (v[e] > best) { best = v[e]; bidx = f * VEC + e; }
  }
  // a heatmap of -inf only never takes the branch above: index 0, like torch.max
  if (bidx == 0x7fffffff && lane_g == 0) bidx = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(kFull, best, o);
    const int oi = __shfl_xor_sync(kFull, bidx, o);
    argmax_merge(best, bidx, ov, oi);
  }
  if constexpr (GROUP > 32) {
    if (lane == 0) { red_v[warp_g] = best; red_i[warp_g] = bidx; }
    __syncthreads();
    if (lane_g == 0) {
      for (int w = 1; w < NW; ++w) argmax_merge(best, bidx, red_v[w], red_i[w]);
    }
  }
  if (lane_g != 0) return;

  // get_preds (util.py:159-168): x = idx % width, y = idx / HEIGHT (sic), (0, 0) unless the maximum is > 0
  int x = bidx % W, y = bidx / H;
  if (!(best > 0.f)) { x = 0; y = 0; }
  float cx = static_cast<float>(x), cy = static_cast<float>(y);
  if (p.use_neighbours && x > 0 && x < W - 1 && y > 0 && y < H - 1) {   // util.py:187-192
    const float l = load_one<T>(zb, static_cast<long>(y) * W + x - 1), r = load_one<T>(zb, static_cast<long>(y) * W + x + 1);
    const float u = load_one<T>(zb, static_cast<long>(y - 1) * W + x), d = load_one<T>(zb, static_cast<long>(y + 1) * W + x);
    cx += 0.25f * static_cast<float>((r > l) - (r < l));   // sign(hm[y, x+1] - hm[y, x-1])
    cy += 0.25f * static_cast<float>((d > u) - (d < u));
  }
  // coords.add_(0.5); [:, :, 0].mul_(2/W); [:, :, 1].mul_(2/H); coords.add_(-1)  -- separately rounded (util.py:195-198)
  cx = __fadd_rn(__fmul_rn(__fadd_rn(cx, 0.5f), p.two_over_w), -1.0f);
  cy = __fadd_rn(__fmul_rn(__fadd_rn(cy, 0.5f), p.two_over_h), -1.0f);
  reinterpret_cast<float2*>(p.coords)[hm] = make_float2(cx, cy);
}

template <typename T, int VEC>
static int launch_decode(const DecodeParams& p, cudaStream_t s) {
  const long npx = static_cast<long>(p.H) * p.W;
  if (npx / VEC <= 512) {
    decode_heatmaps_kernel<T, VEC, 32><<<static_cast<unsigned>((p.n + 3) / 4), 128, 0, s>>>(p);
  } else {
    decode_heatmaps_kernel<T, VEC, 256><<<static_cast<unsigned>(p.n), 256, 0, s>>>(p);
  }
  return check_launch("decode_heatmaps_kernel");
}

}  // namespace dsnt

using namespace dsnt;

extern "C" {

DSNT_API int dsnt_draw_gaussians(const float* centres, int centres_are_pixels, long n, int W, int H, double sigma,
                                 double clip_size, int normalize, float* out, void* stream) {
  if (n < 0 || W <= 0 || H <= 0 || !(sigma > 0.0) || n > 0x7fffffffL) { set_error("dsnt_draw_gaussians: bad arguments"); return DSNT_ERR_BAD_ARG; }
  if (n == 0) return DSNT_OK;
  if (!centres || !out || !aligned(centres, 8)) { set_error("dsnt_draw_gaussians: null or misaligned buffers"); return DSNT_ERR_BAD_ARG; }
  DrawParams p;
  p.centres = centres; p.out = out; p.n = n; p.H = H; p.W = W;
  p.mode = centres_are_pixels ? 1 : 0;
  p.normalize = normalize ? 1 : 0;
  p.kf = static_cast<float>(-0.5 * (1.0 / sigma) * (1.0 / sigma));
  p.radius = static_cast<float>(clip_size > 0.0 ? clip_size / 2.0 : static_cast<double>(W > H ? W : H));
  p.half_w = static_cast<float>(W / 2.0);
  p.half_h = static_cast<float>(H / 2.0);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (W % 4 == 0 && aligned(out, 16))
    draw_gaussians_kernel<4><<<static_cast<unsigned>(n), kDrawBlock, 0, s>>>(p);
  else
    draw_gaussians_kernel<1><<<static_cast<unsigned>(n), kDrawBlock, 0, s>>>(p);
  return check_launch("draw_gaussians_kernel");
}

DSNT_API int dsnt_decode_heatmaps(const void* hm, int dtype, long n, int H, int W, int use_neighbours, float* coords,
                                  void* stream) {
  int rc = check_common(hm, dtype, n, H, W, DSNT_REG_NONE);
  if (rc) return rc;
  if (static_cast<long>(H) * W > 0x7fffffffL) { set_error("heatmap %dx%d too large", H, W); return DSNT_ERR_UNSUPPORTED; }
  if (n == 0) return DSNT_OK;
  if (!coords || !aligned(coords, 8)) { set_error("coords output is required (8-byte aligned)"); return DSNT_ERR_BAD_ARG; }
  DecodeParams p;
  p.hm = hm; p.coords = coords; p.n = n; p.H = H; p.W = W; p.use_neighbours = use_neighbours;
  p.two_over_w = static_cast<float>(2.0 / W);
  p.two_over_h = static_cast<float>(2.0 / H);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // vectors may straddle rows here (the reduction is over the flat heatmap), so only the size and base alignment matter
  const long npx = static_cast<long>(H) * W;
  if (dtype == DSNT_DTYPE_F32)
    return (npx % 4 == 0 && aligned(hm, 16)) ? launch_decode<float, 4>(p, s) : launch_decode<float, 1>(p, s);
  return (npx % 8 == 0 && aligned(hm, 16)) ? launch_decode<__nv_bfloat16, 8>(p, s) : launch_decode<__nv_bfloat16, 1>(p, s);
}

}  // extern "C"
