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

// Fast-path binning bins the tiles a splat can actually be SEEN in: gsplat's tile AABB is the square of 3 sigma of the
// LARGER axis, but the compositor skips a splat wherever alpha < 1/255, i.e. outside the ellipse sigma <= ln(255 o),
// whose half extents along x / y are sqrt(2 ln(255 o) Sigma_xx / yy) — 19-20 % fewer (Gaussian, tile) pairs on the
// anisotropic splats of C2 / C5 (tools/isect_stats.py). A tile is kept iff it lies in gsplat's AABB (never a tile more
// than upstream) AND the span of its pixel centres meets that extent. The extent is the compositor's OWN culling
// extent (alpha_extents, common.cuh: the value its footprint / live-rectangle tests use, already inflated against fp
// rounding), widened once more, so a dropped tile is one where the compositor would have culled the splat on every
// pixel anyway: images, depths, masks and gradients are unchanged bit for bit (checked at full size on C1 / C2 / C3 /
// C5, tests/test_gpu_parity_full.py: check_usable_tiles); only the fast path's private intersection lists get shorter.
// Written with _rn intrinsics (like tile_aabb): the projection kernel (compiled without FMA contraction) and the scatter
// pass of the two-pass binning (tile_sort.cu, compiled with it) must reach the same tiles bit for bit — the histogram of
// one sizes the segments the other fills.
__device__ __forceinline__ void usable_tiles(float mx, float my, float q0, float q1, float q2, float o, int &x0, int &x1,
                                             int &y0, int &y1) {
  // alpha_extents (above) with approximate division / square root (a few ulp; the projection kernel is issue-bound
  // and the IEEE forms cost 0.04 ms on C2), widened by 1e-4 relative + 0.01 px + the rounding of the closed form below
  const float o255 = __fmul_rn(o, 255.0f);
  if (!(o255 > 1.0f)) { x1 = x0; y1 = y0; return; }   // alpha < 1/255 everywhere
  const float w = __fmul_rn(q1, q1);
  const float det = __fadd_rn(__fmaf_rn(q0, q2, -w), __fmaf_rn(-q1, q1, w));
  if (!(det > 0.0f) || !(q0 > 0.0f) || !(q2 > 0.0f)) return;  // degenerate conic: keep gsplat's box
  const float two_tau = __fadd_rn(__fmul_rn(__fmul_rn(2.0f, __logf(o255)), 1.001f), 1e-3f);
  const float r = __fdividef(two_tau, det);
  const float ex = __fmul_rn(r, q2), ey = __fmul_rn(r, q0);
  float hx = __fadd_rn(__fmul_rn(__fmul_rn(ex, rsqrtf(ex)), 1.005f), 0.05f);
  float hy = __fadd_rn(__fmul_rn(__fmul_rn(ey, rsqrtf(ey)), 1.005f), 0.05f);
  if (!(hx < 1e6f) || !(hy < 1e6f)) return;           // NaN / huge: keep gsplat's box
  hx = __fadd_rn(__fadd_rn(__fmul_rn(hx, 1.0001f), 0.01f), __fmul_rn(1e-6f, fabsf(mx)));
  hy = __fadd_rn(__fadd_rn(__fmul_rn(hy, 1.0001f), 0.01f), __fmul_rn(1e-6f, fabsf(my)));
  // tile t holds the pixel centres 16 t + 0.5 .. 16 t + 15.5 (composite_common.cuh: footprint_mask keeps it iff
  // 16 t + 0.5 - m <= h and 16 t + 15.5 - m >= -h)
  const float inv_ts = 1.0f / (float)kTile, last = (float)kTile - 0.5f;
  const float lo_x = fmaxf(__fmul_rn(__fsub_rn(__fsub_rn(mx, hx), last), inv_ts), -1.0f);
  const float hi_x = fminf(__fmul_rn(__fsub_rn(__fadd_rn(mx, hx), 0.5f), inv_ts), 1e6f);
  const float lo_y = fmaxf(__fmul_rn(__fsub_rn(__fsub_rn(my, hy), last), inv_ts), -1.0f);
  const float hi_y = fminf(__fmul_rn(__fsub_rn(__fadd_rn(my, hy), 0.5f), inv_ts), 1e6f);
  x0 = max(x0, (int)ceilf(lo_x)); x1 = min(x1, (int)floorf(hi_x) + 1);
  y0 = max(y0, (int)ceilf(lo_y)); y1 = min(y1, (int)floorf(hi_y) + 1);
  if (x1 <= x0 || y1 <= y0) { x1 = x0; y1 = y0; }
}

}  // namespace pm
