// capi_util.cuh -- argument checks shared by the translation units that define extern "C" entry points.
#pragma once

#include <stdint.h>

#include "common.cuh"
#include "finish_common.cuh"

namespace dsnt {

void set_error(const char* fmt, ...);  // capi.cu
int check_launch(const char* what);    // capi.cu: cudaGetLastError -> DSNT_ERR_LAUNCH

// Launch attributes (shared-memory opt-in, SM count, cluster occupancy) are properties of a DEVICE, and one process may
// drive several: every cache of them is an array indexed by the current device ordinal.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}
inline int sm_count_of_current_device() {
  static int cached[kMaxDevices] = {};
  const int dev = current_device();
  if (cached[dev] == 0) {
    int n = 0;
    cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
  }
  return cached[dev];
}

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// Widest vector the layout allows: a vector must not straddle a row (W % VEC == 0) and every heatmap
// base must be VEC*sizeof aligned (then H*W*sizeof is a multiple of it as well).
inline int pick_vec(int dtype, int W, const void* a, const void* b) {
  if (dtype == DSNT_DTYPE_F32) {
    if (W % 4 == 0 && aligned(a, 16) && (!b || aligned(b, 16))) return 4;
    return 1;
  }
  if (W % 8 == 0 && aligned(a, 16) && (!b || aligned(b, 16))) return 8;
  if (W % 4 == 0 && aligned(a, 8) && (!b || aligned(b, 8))) return 4;
  return 1;
}

inline int check_common(const void* z, int dtype, long n, int H, int W, int reg) {
  if (!z && n != 0) { set_error("null heatmap pointer"); return DSNT_ERR_BAD_ARG; }
  if (dtype != DSNT_DTYPE_F32 && dtype != DSNT_DTYPE_BF16) { set_error("unsupported dtype %d (fp32 and bf16 only)", dtype); return DSNT_ERR_UNSUPPORTED; }
  if (n < 0 || H <= 0 || W <= 0) { set_error("bad shape n=%ld H=%d W=%d", n, H, W); return DSNT_ERR_BAD_ARG; }
  if (static_cast<long>(H) * W > (1L << 28)) { set_error("heatmap %dx%d too large", H, W); return DSNT_ERR_UNSUPPORTED; }
  if (n > 0x7fffffffL) { set_error("too many heatmaps: %ld", n); return DSNT_ERR_UNSUPPORTED; }
  if (reg < DSNT_REG_NONE || reg > DSNT_REG_MSE) { set_error("bad reg %d", reg); return DSNT_ERR_BAD_ARG; }
  return DSNT_OK;
}

// ---- peer exchange block of the *_peer entry points (include/dsnt_b200.h)
inline PeerXchg no_peers() {
  PeerXchg xc;
  for (int r = 0; r < kMaxRanks; ++r) xc.peers[r] = nullptr;
  xc.epoch = nullptr; xc.error = nullptr; xc.rank = 0; xc.world = 1;
  return xc;
}

inline int make_peers(const void* const* peers, int rank, int world, unsigned* epoch, int* error, PeerXchg& xc) {
  if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world || !peers || !epoch || !error) {
    set_error("peer exchange: bad arguments (world %d, rank %d, at most %d ranks)", world, rank, kMaxRanks);
    return DSNT_ERR_BAD_ARG;
  }
  xc = no_peers();
  for (int r = 0; r < world; ++r) {
    if (!peers[r] || !aligned(peers[r], 16)) { set_error("peer exchange: buffer of rank %d is null or misaligned", r); return DSNT_ERR_BAD_ARG; }
    xc.peers[r] = static_cast<unsigned long long*>(const_cast<void*>(peers[r]));
  }
  xc.epoch = epoch; xc.error = error; xc.rank = rank; xc.world = world;
  return DSNT_OK;
}

}  // namespace dsnt
