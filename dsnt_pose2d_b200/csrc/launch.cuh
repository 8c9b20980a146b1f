// launch.cuh -- host-side kernel selection for the fused head (shared by the per-dtype translation units).
#pragma once

#include <cstdarg>
#include <cstdio>

#include "head_bwd.cuh"
#include "head_fwd.cuh"
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
