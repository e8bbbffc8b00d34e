// common.cuh — shared helpers for the sm_100a kernels behind include/prometheus_gs.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/prometheus_gs.h"

namespace pm {

void set_last_cuda_error(cudaError_t e);

inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_cuda_error(e);
    return PM_ERR_CUDA;
  }
  return PM_OK;
}

inline cudaStream_t as_stream(pm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kTile = 16;  // the only tile size the reference ever uses (gsplat default)

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Effective element count: the device-side count (if given) clamped to the host capacity.
__device__ __forceinline__ int64_t effective_n(int64_t n_host, const int32_t *d_n) {
  if (d_n == nullptr) return n_host;
  int64_t v = (int64_t)(*d_n);
  if (v < 0) v = 0;
  return v < n_host ? v : n_host;
}

// floor(log2(n)) + 1, the bit widths gsplat packs tile and camera ids with.
__host__ __device__ inline int n_bits_for(int n) {
  int b = 0;
  while (n > 0) {
    ++b;
    n >>= 1;
  }
  return b;
}

// Tile AABB of a projected Gaussian. Written with _rn
// intrinsics so no compiler flag can contract or reorder it: the result must be
// bit-identical between the count (project) and the emit kernels and the oracle.
__device__ __forceinline__ int clamp_tile(float v, int hi) {
  if (!(v > 0.0f)) return 0;
  if (v >= (float)hi) return hi;
  return (int)v;
}
__device__ __forceinline__ void tile_aabb(float mx, float my, int radius, int tile_width, int tile_height,
                                          int &x0, int &x1, int &y0, int &y1) {
  // x / 16 == x * 0.0625 exactly (power of two), and the multiply is one instruction
  const float inv_ts = 1.0f / (float)kTile;
  const float tr = __fmul_rn((float)radius, inv_ts);
  const float tx = __fmul_rn(mx, inv_ts), ty = __fmul_rn(my, inv_ts);
  x0 = clamp_tile(floorf(__fsub_rn(tx, tr)), tile_width);
  x1 = clamp_tile(ceilf(__fadd_rn(tx, tr)), tile_width);
  y0 = clamp_tile(floorf(__fsub_rn(ty, tr)), tile_height);
  y1 = clamp_tile(ceilf(__fadd_rn(ty, tr)), tile_height);
}

// Axis-aligned half extents of {alpha >= 1/255} = {sigma <= tau}, tau = ln(255 o): sqrt(2 tau Sigma_xx),
// sqrt(2 tau Sigma_yy) with Sigma = conic^-1, inflated so that fp rounding can only over-include. ONE expression for
// the compositor's culling (composite_common.cuh: gather_finish) and for the fast path's binning (project.cu:
// usable_tiles). -1e30: alpha < 1/255 everywhere; 3e38: degenerate conic, no culling.
__device__ __forceinline__ void alpha_extents(float q0, float q1, float q2, float o, float &hx, float &hy) {
  // det(conic) with Kahan's fma compensation: anisotropic splats cancel badly in q0*q2 - q1*q1
  const float w = q1 * q1;
  const float det = fmaf(q0, q2, -w) + fmaf(-q1, q1, w);
  if (!(o * 255.0f > 1.0f)) {
    hx = hy = -1e30f;  // never evaluated (as upstream would skip it)
  } else if (!(det > 0.0f) || !(q0 > 0.0f) || !(q2 > 0.0f)) {
    hx = hy = 3e38f;
  } else {
    const float two_tau = 2.0f * __logf(o * 255.0f) * 1.001f + 1e-3f;
    hx = sqrtf(two_tau * q2 / det) * 1.005f + 0.05f;
    hy = sqrtf(two_tau * q0 / det) * 1.005f + 0.05f;
    if (!(hx == hx) || !(hy == hy)) hx = hy = 3e38f;
  }
}

}  // namespace pm
