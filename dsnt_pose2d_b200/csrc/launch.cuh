// launch.cuh -- host-side kernel selection for the fused head (shared by the per-dtype translation units).
#pragma once

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#include "head_bwd.cuh"
#include "head_fwd.cuh"
#include "head_fast.cuh"
#include "head_stream.cuh"

namespace dsnt {

void set_error(const char* fmt, ...);  // capi.cu
int check_launch(const char* what);    // capi.cu: cudaGetLastError -> DSNT_ERR_LAUNCH

// per-dtype entry points (one translation unit each so they compile in parallel)
int launch_head_fwd_f32(const HeadFwdParams& p, int vec, bool logits, int variant, cudaStream_t stream);
int launch_head_fwd_bf16(const HeadFwdParams& p, int vec, bool logits, int variant, cudaStream_t stream);
int launch_head_bwd_f32(const HeadBwdParams& p, int vec, bool logits, int variant, cudaStream_t stream);
int launch_head_bwd_bf16(const HeadBwdParams& p, int vec, bool logits, int variant, cudaStream_t stream);

constexpr size_t kMaxDynSmem = 48 * 1024;  // stay under the no-opt-in limit: tables are (W+H+8) floats

// ------------------------------------------------------------------------------------------------ forward
template <typename T, int VEC, int GROUP, int NV, int REG, bool LOGITS>
int launch_fwd_resident(const HeadFwdParams& p, cudaStream_t stream) {
  constexpr int BLOCK = fwd_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  const bool gauss = REG >= 0 ? reg_needs_gauss(REG) : reg_needs_gauss(p.reg);
  const size_t smem = gauss ? sizeof(float) * GPB * table_floats(p.H, p.W) : 0;
  if (smem > kMaxDynSmem) { set_error("heatmap %dx%d: Gaussian tables exceed shared memory", p.H, p.W); return DSNT_ERR_UNSUPPORTED; }
  const long grid = (p.n + GPB - 1) / GPB;
  head_fwd_kernel<T, VEC, GROUP, NV, REG, LOGITS><<<static_cast<unsigned>(grid), BLOCK, smem, stream>>>(p);
  return check_launch("head_fwd_kernel");
}

template <typename T, int VEC, int REG, bool LOGITS>
int launch_fwd_large(const HeadFwdParams& p, cudaStream_t stream) {
  const bool gauss = REG >= 0 ? reg_needs_gauss(REG) : reg_needs_gauss(p.reg);
  const size_t smem = gauss ? sizeof(float) * table_floats(p.H, p.W) : 0;
  if (smem > kMaxDynSmem) { set_error("heatmap %dx%d: Gaussian tables exceed shared memory", p.H, p.W); return DSNT_ERR_UNSUPPORTED; }
  head_fwd_large_kernel<T, VEC, REG, LOGITS><<<static_cast<unsigned>(p.n), kLargeBlock, smem, stream>>>(p);
  return check_launch("head_fwd_large_kernel");
}

// variant: 0 auto; 1 force the streaming two-pass kernel (for tests / comparison)
template <typename T, int VEC, int REG, bool LOGITS>
int launch_fwd_shape(const HeadFwdParams& p, int variant, cudaStream_t stream) {
  const long nvec = static_cast<long>(p.H) * p.W / VEC;
  if (variant == 1) return launch_fwd_large<T, VEC, REG, LOGITS>(p, stream);
  if (nvec <= 32 * 2) return launch_fwd_resident<T, VEC, 32, 2, REG, LOGITS>(p, stream);
  if (nvec <= 32 * 8) return launch_fwd_resident<T, VEC, 32, 8, REG, LOGITS>(p, stream);
  if (nvec <= 256 * 4) return launch_fwd_resident<T, VEC, 256, 4, REG, LOGITS>(p, stream);
  if (nvec <= 512 * 8) return launch_fwd_resident<T, VEC, 512, 8, REG, LOGITS>(p, stream);
  return launch_fwd_large<T, VEC, REG, LOGITS>(p, stream);
}

// ---- streaming kernels (logits, vectorised): variant 0 picks them; variant 2 forces the v0 register-resident path
inline bool stream_group_is_cta(long nvec) { return nvec > 2048; }

// ---- tuned kernels (head_fast.cuh): variant 0 picks them whenever the layout qualifies; variant 3 forces v1
inline bool make_fast_geom(int H, int W, int vec, int group, FastGeom& f, int u = kFastU) {
  const int wv = W / vec;
  if (wv <= 0 || (wv & (wv - 1)) != 0 || group % wv != 0) return false;
  const long nvec = static_cast<long>(H) * wv;
  if (nvec % (static_cast<long>(group) * u) != 0) return false;
  f.wv_shift = 0;
  while ((1 << f.wv_shift) < wv) ++f.wv_shift;
  f.nbatch = static_cast<int>(nvec / (static_cast<long>(group) * u));
  f.rstep = group / wv;
  f.dy_step = static_cast<float>(f.rstep) * (2.0f / static_cast<float>(H));
  f.dy_batch = static_cast<float>(u) * f.dy_step;
  return true;
}

// Upper bound of the Gaussian window (make_window) for any target: floor(n sqrt(r2 + 1/n^2)) + 3 pixels per axis,
// widened to whole vectors along x.  When that cannot fit the shared-memory stash the generic kernels run instead.
inline bool stash_fits(int H, int W, int vec, int reg, float r2_win) {
  if (!reg_needs_gauss(reg)) return true;
  const double r2 = r2_win;
  const int mc = static_cast<int>(std::floor(W * std::sqrt(r2 + 1.0 / (static_cast<double>(W) * W)))) + 4;
  const int mr = static_cast<int>(std::floor(H * std::sqrt(r2 + 1.0 / (static_cast<double>(H) * H)))) + 4;
  const int cols = std::min(W, ((mc + vec - 2) / vec + 1) * vec);
  const int rows = std::min(H, mr);
  return rows <= kTabN && cols <= kTabN && rows * cols <= kStashFloats;
}

inline size_t tune_smem(const char* name) {
  const char* v = std::getenv(name);
  return v ? static_cast<size_t>(std::atol(v)) : 0;
}

template <typename T, int VEC, int GROUP, int REG>
int launch_fwd_fast_group(const HeadFwdParams& p, const FastGeom& f, cudaStream_t stream) {
  constexpr int BLOCK = stream_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  HeadFwdFastParams ps;
  ps.base = p;
  ps.g = make_geom(p.H, p.W, VEC, GROUP, p.sigma, REG);
  ps.f = f;
  ps.fl = FlipCfg{nullptr, nullptr, 0};
  ps.pc = PreactCfg{DSNT_PREACT_SOFTMAX, 0.f, 0.f};
  if (!stash_fits(p.H, p.W, VEC, REG, ps.g.r2_win)) return 1;
  const unsigned grid = static_cast<unsigned>((p.n + GPB - 1) / GPB);
  head_fwd_fast_kernel<T, VEC, GROUP, REG><<<grid, BLOCK, 0, stream>>>(ps);
  return check_launch("head_fwd_fast_kernel");
}

// returns 1 when the layout does not qualify (caller falls through to the generic kernels)
template <typename T, int VEC, int REG>
int try_launch_fwd_fast(const HeadFwdParams& p, cudaStream_t stream) {
  if constexpr (sizeof(T) * VEC != 16) {
    return 1;
  } else {
    const long nvec = static_cast<long>(p.H) * p.W / VEC;
    FastGeom f;
    if (stream_group_is_cta(nvec)) {
      if (!make_fast_geom(p.H, p.W, VEC, 256, f)) return 1;
      return launch_fwd_fast_group<T, VEC, 256, REG>(p, f, stream);
    }
    if (!make_fast_geom(p.H, p.W, VEC, 32, f)) return 1;
    return launch_fwd_fast_group<T, VEC, 32, REG>(p, f, stream);
  }
}

// flip test-time augmentation on the tuned forward (infer.cu); returns 1 when the layout does not qualify
template <typename T, int VEC, int GROUP>
int launch_fwd_fast_flip_group(const HeadFwdParams& p, const FlipCfg& fl, const FastGeom& f, cudaStream_t stream) {
  constexpr int BLOCK = stream_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  HeadFwdFastParams ps;
  ps.base = p;
  ps.g = make_geom(p.H, p.W, VEC, GROUP, p.sigma, DSNT_REG_NONE);
  ps.f = f;
  ps.fl = fl;
  ps.pc = PreactCfg{DSNT_PREACT_SOFTMAX, 0.f, 0.f};
  const unsigned grid = static_cast<unsigned>((p.n + GPB - 1) / GPB);
  head_fwd_fast_kernel<T, VEC, GROUP, DSNT_REG_NONE, true><<<grid, BLOCK, 0, stream>>>(ps);
  return check_launch("head_fwd_fast_kernel<flip>");
}

template <typename T, int VEC>
int try_launch_fwd_fast_flip(const HeadFwdParams& p, const FlipCfg& fl, cudaStream_t stream) {
  if constexpr (sizeof(T) * VEC != 16) {
    return 1;
  } else {
    const long nvec = static_cast<long>(p.H) * p.W / VEC;
    FastGeom f;
    if (stream_group_is_cta(nvec)) {
      if (!make_fast_geom(p.H, p.W, VEC, 256, f, kFlipU)) return 1;
      return launch_fwd_fast_flip_group<T, VEC, 256>(p, fl, f, stream);
    }
    if (!make_fast_geom(p.H, p.W, VEC, 32, f, kFlipU)) return 1;
    return launch_fwd_fast_flip_group<T, VEC, 32>(p, fl, f, stream);
  }
}

template <typename T, int VEC, int GROUP, int REG>
int launch_bwd_fast_group(const HeadBwdParams& p, const FastGeom& f, cudaStream_t stream) {
  constexpr int BLOCK = stream_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  HeadBwdFastParams ps;
  ps.base = p;
  ps.g = make_geom(p.H, p.W, VEC, GROUP, p.sigma, REG);
  ps.f = f;
  if (!stash_fits(p.H, p.W, VEC, REG, ps.g.r2_win)) return 1;
  const unsigned grid = static_cast<unsigned>((p.n + GPB - 1) / GPB);
  static const size_t extra_smem = tune_smem("DSNT_TUNE_BWD_SMEM");   // developer knob: caps CTAs per SM
  head_bwd_fast_kernel<T, VEC, GROUP, REG><<<grid, BLOCK, extra_smem, stream>>>(ps);
  return check_launch("head_bwd_fast_kernel");
}

template <typename T, int VEC, int REG>
int try_launch_bwd_fast(const HeadBwdParams& p, cudaStream_t stream) {
  // fp32 with a Gaussian window: the generic kernel's in-loop window arithmetic hides completely under the
  // 8 B/pixel of traffic (0.98 of HBM peak), while the stash epilogue adds an exposed tail (0.89) -- measured in
  // profiles/r01_v2_kbench.txt.  bf16 has half the bytes per pixel to hide work under, there the stash wins.
  if constexpr (sizeof(T) * VEC != 16 || (sizeof(T) == 4 && reg_needs_gauss(REG))) {
    return 1;
  } else {
    const long nvec = static_cast<long>(p.H) * p.W / VEC;
    FastGeom f;
    if (stream_group_is_cta(nvec)) {
      if (!make_fast_geom(p.H, p.W, VEC, 256, f)) return 1;
      return launch_bwd_fast_group<T, VEC, 256, REG>(p, f, stream);
    }
    if (!make_fast_geom(p.H, p.W, VEC, 32, f)) return 1;
    return launch_bwd_fast_group<T, VEC, 32, REG>(p, f, stream);
  }
}

template <typename T, int VEC, int GROUP, int REG>
int launch_fwd_stream_group(const HeadFwdParams& p, cudaStream_t stream) {
  constexpr int BLOCK = stream_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  HeadFwdStreamParams ps;
  ps.base = p;
  ps.g = make_geom(p.H, p.W, VEC, GROUP, p.sigma, REG);
  const unsigned grid = static_cast<unsigned>((p.n + GPB - 1) / GPB);
  if (GROUP % ps.g.wv == 0) {
    head_fwd_stream_kernel<T, VEC, GROUP, REG, true><<<grid, BLOCK, 0, stream>>>(ps);
  } else if constexpr (REG != DSNT_REG_VAR) {
    head_fwd_stream_kernel<T, VEC, GROUP, REG, false><<<grid, BLOCK, 0, stream>>>(ps);
  } else {
    // the single-pass variance needs thread-fixed columns; odd widths keep the two-pass kernels
    return launch_fwd_shape<T, VEC, REG, true>(p, 0, stream);
  }
  return check_launch("head_fwd_stream_kernel");
}

template <typename T, int VEC, int REG>
int launch_fwd_stream(const HeadFwdParams& p, cudaStream_t stream) {
  const long nvec = static_cast<long>(p.H) * p.W / VEC;
  return stream_group_is_cta(nvec) ? launch_fwd_stream_group<T, VEC, 256, REG>(p, stream)
                                   : launch_fwd_stream_group<T, VEC, 32, REG>(p, stream);
}

template <typename T, int VEC, bool LOGITS>
int launch_fwd_reg(const HeadFwdParams& p, int variant, cudaStream_t stream) {
  if constexpr (VEC == 1) {
    return launch_fwd_shape<T, VEC, -1, LOGITS>(p, variant, stream);  // odd sizes: regulariser chosen at run time
  } else {
    if constexpr (LOGITS) {
      if (variant == 0) {
        int rc = 1;
        switch (p.reg) {
          case DSNT_REG_NONE: rc = try_launch_fwd_fast<T, VEC, DSNT_REG_NONE>(p, stream); break;
          case DSNT_REG_VAR: rc = try_launch_fwd_fast<T, VEC, DSNT_REG_VAR>(p, stream); break;
          case DSNT_REG_KL: rc = try_launch_fwd_fast<T, VEC, DSNT_REG_KL>(p, stream); break;
          case DSNT_REG_JS: rc = try_launch_fwd_fast<T, VEC, DSNT_REG_JS>(p, stream); break;
          case DSNT_REG_MSE: rc = try_launch_fwd_fast<T, VEC, DSNT_REG_MSE>(p, stream); break;
        }
        if (rc != 1) return rc;
      }
      if (variant == 0 || variant == 3) {
        if (p.reg == DSNT_REG_NONE) return launch_fwd_stream<T, VEC, DSNT_REG_NONE>(p, stream);
        if (p.reg == DSNT_REG_VAR) return launch_fwd_stream<T, VEC, DSNT_REG_VAR>(p, stream);
        if (p.reg == DSNT_REG_KL) return launch_fwd_stream<T, VEC, DSNT_REG_KL>(p, stream);
        if (p.reg == DSNT_REG_JS) return launch_fwd_stream<T, VEC, DSNT_REG_JS>(p, stream);
        if (p.reg == DSNT_REG_MSE) return launch_fwd_stream<T, VEC, DSNT_REG_MSE>(p, stream);
      }
      if (variant == 2) variant = 0;
    }
    switch (p.reg) {
      case DSNT_REG_NONE: return launch_fwd_shape<T, VEC, DSNT_REG_NONE, LOGITS>(p, variant, stream);
      case DSNT_REG_VAR: return launch_fwd_shape<T, VEC, DSNT_REG_VAR, LOGITS>(p, variant, stream);
      case DSNT_REG_KL: return launch_fwd_shape<T, VEC, DSNT_REG_KL, LOGITS>(p, variant, stream);
      case DSNT_REG_JS: return launch_fwd_shape<T, VEC, DSNT_REG_JS, LOGITS>(p, variant, stream);
      case DSNT_REG_MSE: return launch_fwd_shape<T, VEC, DSNT_REG_MSE, LOGITS>(p, variant, stream);
    }
    set_error("bad reg %d", p.reg);
    return DSNT_ERR_BAD_ARG;
  }
}

// ---- the tuned kernels for the other pre-activations (preact_*.cu); return 1 = layout does not qualify, the caller
// then runs the generic kernels of head_preact.cuh
int launch_preact_fast_fwd_f32(const HeadFwdParams& p, const PreactCfg& pc, int vec, cudaStream_t stream);
int launch_preact_fast_fwd_bf16(const HeadFwdParams& p, const PreactCfg& pc, int vec, cudaStream_t stream);
int launch_preact_fast_bwd_f32(const HeadBwdParams& p, const PreactCfg& pc, int vec, cudaStream_t stream);
int launch_preact_fast_bwd_bf16(const HeadBwdParams& p, const PreactCfg& pc, int vec, cudaStream_t stream);

template <typename T, int VEC, int GROUP, int REG, int PA>
int launch_preact_fwd_fast_group(const HeadFwdParams& p, const PreactCfg& pc, const FastGeom& f, cudaStream_t stream) {
  constexpr int BLOCK = stream_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  HeadFwdFastParams ps;
  ps.base = p;
  ps.g = make_geom(p.H, p.W, VEC, GROUP, p.sigma, REG);
  ps.f = f;
  ps.fl = FlipCfg{nullptr, nullptr, 0};
  ps.pc = pc;
  if (!stash_fits(p.H, p.W, VEC, REG, ps.g.r2_win)) return 1;
  const unsigned grid = static_cast<unsigned>((p.n + GPB - 1) / GPB);
  head_fwd_fast_kernel<T, VEC, GROUP, REG, false, PA><<<grid, BLOCK, 0, stream>>>(ps);
  return check_launch("head_fwd_fast_kernel<preact>");
}

template <typename T, int VEC, int REG, int PA>
int try_launch_preact_fwd_fast(const HeadFwdParams& p, const PreactCfg& pc, cudaStream_t stream) {
  const long nvec = static_cast<long>(p.H) * p.W / VEC;
  FastGeom f;
  if (stream_group_is_cta(nvec)) {
    if (!make_fast_geom(p.H, p.W, VEC, 256, f)) return 1;
    return launch_preact_fwd_fast_group<T, VEC, 256, REG, PA>(p, pc, f, stream);
  }
  if (!make_fast_geom(p.H, p.W, VEC, 32, f)) return 1;
  return launch_preact_fwd_fast_group<T, VEC, 32, REG, PA>(p, pc, f, stream);
}

template <typename T, int VEC, int PA>
int launch_preact_fwd_fast_reg(const HeadFwdParams& p, const PreactCfg& pc, cudaStream_t stream) {
  switch (p.reg) {
    case DSNT_REG_NONE: return try_launch_preact_fwd_fast<T, VEC, DSNT_REG_NONE, PA>(p, pc, stream);
    case DSNT_REG_VAR: return try_launch_preact_fwd_fast<T, VEC, DSNT_REG_VAR, PA>(p, pc, stream);
    case DSNT_REG_KL: return try_launch_preact_fwd_fast<T, VEC, DSNT_REG_KL, PA>(p, pc, stream);
    case DSNT_REG_JS: return try_launch_preact_fwd_fast<T, VEC, DSNT_REG_JS, PA>(p, pc, stream);
    case DSNT_REG_MSE: return try_launch_preact_fwd_fast<T, VEC, DSNT_REG_MSE, PA>(p, pc, stream);
  }
  return 1;
}

template <typename T, int VEC>
int launch_preact_fwd_fast(const HeadFwdParams& p, const PreactCfg& pc, cudaStream_t stream) {
  switch (pc.preact) {
    case DSNT_PREACT_TSOFTMAX: return launch_preact_fwd_fast_reg<T, VEC, DSNT_PREACT_TSOFTMAX>(p, pc, stream);
    case DSNT_PREACT_ABS: return launch_preact_fwd_fast_reg<T, VEC, DSNT_PREACT_ABS>(p, pc, stream);
    case DSNT_PREACT_RELU: return launch_preact_fwd_fast_reg<T, VEC, DSNT_PREACT_RELU>(p, pc, stream);
    case DSNT_PREACT_SIGMOID: return launch_preact_fwd_fast_reg<T, VEC, DSNT_PREACT_SIGMOID>(p, pc, stream);
  }
  return 1;   // DSNT_PREACT_SOFTMAX with an epsilon: the generic epsilon-exact kernel
}

template <typename T, int VEC, int GROUP, int REG, int PA>
int launch_preact_bwd_stream_group(const HeadBwdParams& p, const PreactCfg& pc, cudaStream_t stream) {
  constexpr int BLOCK = stream_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  HeadBwdStreamParams ps;
  ps.base = p;
  ps.g = make_geom(p.H, p.W, VEC, GROUP, p.sigma, REG);
  ps.pc = pc;
  const unsigned grid = static_cast<unsigned>((p.n + GPB - 1) / GPB);
  if (GROUP % ps.g.wv == 0)
    head_bwd_stream_kernel<T, VEC, GROUP, REG, true, PA><<<grid, BLOCK, 0, stream>>>(ps);
  else
    head_bwd_stream_kernel<T, VEC, GROUP, REG, false, PA><<<grid, BLOCK, 0, stream>>>(ps);
  return check_launch("head_bwd_stream_kernel<preact>");
}

template <typename T, int VEC, int REG, int PA>
int launch_preact_bwd_stream(const HeadBwdParams& p, const PreactCfg& pc, cudaStream_t stream) {
  const long nvec = static_cast<long>(p.H) * p.W / VEC;
  return stream_group_is_cta(nvec) ? launch_preact_bwd_stream_group<T, VEC, 256, REG, PA>(p, pc, stream)
                                   : launch_preact_bwd_stream_group<T, VEC, 32, REG, PA>(p, pc, stream);
}

template <typename T, int VEC, int PA>
int launch_preact_bwd_fast_reg(const HeadBwdParams& p, const PreactCfg& pc, cudaStream_t stream) {
  switch (p.reg) {
    case DSNT_REG_NONE: return launch_preact_bwd_stream<T, VEC, DSNT_REG_NONE, PA>(p, pc, stream);
    case DSNT_REG_VAR: return launch_preact_bwd_stream<T, VEC, DSNT_REG_VAR, PA>(p, pc, stream);
    case DSNT_REG_KL: return launch_preact_bwd_stream<T, VEC, DSNT_REG_KL, PA>(p, pc, stream);
    case DSNT_REG_JS: return launch_preact_bwd_stream<T, VEC, DSNT_REG_JS, PA>(p, pc, stream);
    case DSNT_REG_MSE: return launch_preact_bwd_stream<T, VEC, DSNT_REG_MSE, PA>(p, pc, stream);
  }
  return 1;
}

template <typename T, int VEC>
int launch_preact_bwd_fast(const HeadBwdParams& p, const PreactCfg& pc, cudaStream_t stream) {
  switch (pc.preact) {
    case DSNT_PREACT_TSOFTMAX: return launch_preact_bwd_fast_reg<T, VEC, DSNT_PREACT_TSOFTMAX>(p, pc, stream);
    case DSNT_PREACT_ABS: return launch_preact_bwd_fast_reg<T, VEC, DSNT_PREACT_ABS>(p, pc, stream);
    case DSNT_PREACT_RELU: return launch_preact_bwd_fast_reg<T, VEC, DSNT_PREACT_RELU>(p, pc, stream);
    case DSNT_PREACT_SIGMOID: return launch_preact_bwd_fast_reg<T, VEC, DSNT_PREACT_SIGMOID>(p, pc, stream);
  }
  return 1;
}

// ------------------------------------------------------------------------------------------------ backward
template <typename T, int VEC, int GROUP, int NV, int REG, bool LOGITS>
int launch_bwd_one(const HeadBwdParams& p, cudaStream_t stream) {
  constexpr int BLOCK = fwd_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  const bool gauss = REG >= 0 ? reg_needs_gauss(REG) : reg_needs_gauss(p.reg);
  const size_t smem = gauss ? sizeof(float) * GPB * table_floats(p.H, p.W) : 0;
  if (smem > kMaxDynSmem) { set_error("heatmap %dx%d: Gaussian tables exceed shared memory", p.H, p.W); return DSNT_ERR_UNSUPPORTED; }
  const long nvec = static_cast<long>(p.H) * p.W / VEC;
  const long chunks = (nvec + GROUP * NV - 1) / (GROUP * NV);
  if (chunks > 65535) { set_error("heatmap %dx%d too large", p.H, p.W); return DSNT_ERR_UNSUPPORTED; }
  dim3 grid(static_cast<unsigned>((p.n + GPB - 1) / GPB), static_cast<unsigned>(chunks));
  head_bwd_kernel<T, VEC, GROUP, NV, REG, LOGITS><<<grid, BLOCK, smem, stream>>>(p);
  return check_launch("head_bwd_kernel");
}

// variant: 0 = streaming kernel (logits, vectorised) else the v0 register-resident kernel
template <typename T, int VEC, int REG, bool LOGITS>
int launch_bwd_shape(const HeadBwdParams& p, int variant, cudaStream_t stream) {
  (void)variant;
  const long nvec = static_cast<long>(p.H) * p.W / VEC;
  if (nvec <= 32 * 2) return launch_bwd_one<T, VEC, 32, 2, REG, LOGITS>(p, stream);
  if (nvec <= 32 * 8) return launch_bwd_one<T, VEC, 32, 8, REG, LOGITS>(p, stream);
  return launch_bwd_one<T, VEC, 256, 4, REG, LOGITS>(p, stream);  // chunked over blockIdx.y beyond 1024 vectors
}

template <typename T, int VEC, int GROUP, int REG>
int launch_bwd_stream_group(const HeadBwdParams& p, cudaStream_t stream) {
  constexpr int BLOCK = stream_block_threads<GROUP>();
  constexpr int GPB = BLOCK / GROUP;
  HeadBwdStreamParams ps;
  ps.base = p;
  ps.g = make_geom(p.H, p.W, VEC, GROUP, p.sigma, REG);
  ps.pc = PreactCfg{DSNT_PREACT_SOFTMAX, 0.f, 0.f};
  const unsigned grid = static_cast<unsigned>((p.n + GPB - 1) / GPB);
  if (GROUP % ps.g.wv == 0)
    head_bwd_stream_kernel<T, VEC, GROUP, REG, true><<<grid, BLOCK, 0, stream>>>(ps);
  else
    head_bwd_stream_kernel<T, VEC, GROUP, REG, false><<<grid, BLOCK, 0, stream>>>(ps);
  return check_launch("head_bwd_stream_kernel");
}

template <typename T, int VEC, int REG>
int launch_bwd_stream(const HeadBwdParams& p, cudaStream_t stream) {
  const long nvec = static_cast<long>(p.H) * p.W / VEC;
  return stream_group_is_cta(nvec) ? launch_bwd_stream_group<T, VEC, 256, REG>(p, stream)
                                   : launch_bwd_stream_group<T, VEC, 32, REG>(p, stream);
}

template <typename T, int VEC, bool LOGITS>
int launch_bwd_reg(const HeadBwdParams& p, int variant, cudaStream_t stream) {
  if constexpr (VEC == 1) {
    return launch_bwd_shape<T, VEC, -1, LOGITS>(p, variant, stream);
  } else {
    if constexpr (LOGITS) {
      if (variant == 0) {
        int rc = 1;
        switch (p.reg) {
          case DSNT_REG_NONE: rc = try_launch_bwd_fast<T, VEC, DSNT_REG_NONE>(p, stream); break;
          case DSNT_REG_VAR: rc = try_launch_bwd_fast<T, VEC, DSNT_REG_VAR>(p, stream); break;
          case DSNT_REG_KL: rc = try_launch_bwd_fast<T, VEC, DSNT_REG_KL>(p, stream); break;
          case DSNT_REG_JS: rc = try_launch_bwd_fast<T, VEC, DSNT_REG_JS>(p, stream); break;
          case DSNT_REG_MSE: rc = try_launch_bwd_fast<T, VEC, DSNT_REG_MSE>(p, stream); break;
        }
        if (rc != 1) return rc;
      }
      if (variant == 0 || variant == 3) {
        switch (p.reg) {
          case DSNT_REG_NONE: return launch_bwd_stream<T, VEC, DSNT_REG_NONE>(p, stream);
          case DSNT_REG_VAR: return launch_bwd_stream<T, VEC, DSNT_REG_VAR>(p, stream);
          case DSNT_REG_KL: return launch_bwd_stream<T, VEC, DSNT_REG_KL>(p, stream);
          case DSNT_REG_JS: return launch_bwd_stream<T, VEC, DSNT_REG_JS>(p, stream);
          case DSNT_REG_MSE: return launch_bwd_stream<T, VEC, DSNT_REG_MSE>(p, stream);
        }
      }
    }
    switch (p.reg) {
      case DSNT_REG_NONE: return launch_bwd_shape<T, VEC, DSNT_REG_NONE, LOGITS>(p, variant, stream);
      case DSNT_REG_VAR: return launch_bwd_shape<T, VEC, DSNT_REG_VAR, LOGITS>(p, variant, stream);
      case DSNT_REG_KL: return launch_bwd_shape<T, VEC, DSNT_REG_KL, LOGITS>(p, variant, stream);
      case DSNT_REG_JS: return launch_bwd_shape<T, VEC, DSNT_REG_JS, LOGITS>(p, variant, stream);
      case DSNT_REG_MSE: return launch_bwd_shape<T, VEC, DSNT_REG_MSE, LOGITS>(p, variant, stream);
    }
    set_error("bad reg %d", p.reg);
    return DSNT_ERR_BAD_ARG;
  }
}

}  // namespace dsnt
