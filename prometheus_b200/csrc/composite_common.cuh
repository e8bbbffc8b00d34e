// composite_common.cuh — pieces shared by the compositing kernels (composite.cu) and the fused
// sort-as-you-go renderer (tile_render.cu): the staged Gaussian records, the gather that fills them,
// the explicitly rounded alpha evaluation and the per-batch forward loop. Keeping ONE copy of the
// arithmetic is what makes the fused path's images bit-identical to the separate-stage path.
#pragma once
#include "common.cuh"

namespace pm {

constexpr int kBatch = 256;
constexpr float kAlphaMin = 1.0f / 255.0f;
constexpr float kTStop = 1e-4f;

struct __align__(16) GaussA { float mx, my, hx, hy; };      // mean + half extents of the alpha>=1/255 box
struct __align__(16) GaussB { float cx, cy, cz, opac; };    // conic pre-scaled to base-2 exponent units + opacity
struct __align__(16) GaussC { float c0, c1, c2, c3; };      // colour channels (+ depth), zero padded

constexpr float kLog2e = 1.4426950408889634f;

// Shared-memory accesses of the hot loops go through explicit 32-bit shared addresses: with plain
// array indexing nvcc re-derived the shared window base (S2R SR_CgaCtaId + LEA) on every iteration.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// -sigma * log2(e) from the pre-scaled conic; written with explicit roundings so the forward and the
// backward kernel evaluate bit-identical alphas (their skip / stop decisions must agree).
__device__ __forceinline__ float power2(float Q0, float Q1, float Q2, float dx, float dy) {
  return __fmaf_rn(Q0, __fmul_rn(dx, dx), __fmaf_rn(Q2, __fmul_rn(dy, dy), __fmul_rn(Q1, __fmul_rn(dx, dy))));
}

// Gather one Gaussian of the sorted list into the staging records.
__device__ __forceinline__ void gather(int32_t f, int N, const float *__restrict__ means2d,
                                       const float *__restrict__ conics, const float *__restrict__ opacities,
                                       const float *__restrict__ colors, int colors_per_camera, int D,
                                       const float *__restrict__ depths, GaussA &a, GaussB &b, GaussC &c) {
  const int32_t g = f % N;
  const float2 m = __ldg(reinterpret_cast<const float2 *>(means2d) + f);
  const float q0 = __ldg(conics + (size_t)f * 3 + 0), q1 = __ldg(conics + (size_t)f * 3 + 1),
              q2 = __ldg(conics + (size_t)f * 3 + 2);
  const float o = __ldg(opacities + g);
  // extent of {alpha >= 1/255} = {sigma <= tau}, tau = ln(255*o): half widths sqrt(2 tau Sigma_xx),
  // sqrt(2 tau Sigma_yy) with Sigma = conic^-1. Inflated so fp rounding can only over-include.
  float hx, hy;
  // det(conic) with Kahan's fma compensation: anisotropic splats cancel badly in q0*q2 - q1*q1
  const float w = q1 * q1;
  const float det = fmaf(q0, q2, -w) + fmaf(-q1, q1, w);
  if (!(o * 255.0f > 1.0f)) {
    hx = hy = -1e30f;  // alpha < 1/255 everywhere: never evaluated (as upstream would skip it)
  } else if (!(det > 0.0f) || !(q0 > 0.0f) || !(q2 > 0.0f)) {
    hx = hy = 3e38f;   // degenerate conic: no culling
  } else {
    const float two_tau = 2.0f * __logf(o * 255.0f) * 1.001f + 1e-3f;
    hx = sqrtf(two_tau * q2 / det) * 1.005f + 0.05f;
    hy = sqrtf(two_tau * q0 / det) * 1.005f + 0.05f;
    if (!(hx == hx) || !(hy == hy)) hx = hy = 3e38f;
  }
  a = GaussA{m.x, m.y, hx, hy};
  // exp(-sigma) = 2^(Q0 dx^2 + Q1 dx dy + Q2 dy^2) with Q0 = -q0/2 * log2e, Q1 = -q1 * log2e, Q2 = -q2/2 * log2e
  b = GaussB{q0 * (-0.5f * kLog2e), q1 * (-kLog2e), q2 * (-0.5f * kLog2e), o};
  float ch[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  const float *col = colors + (size_t)(colors_per_camera ? f : g) * D;
  for (int k = 0; k < D; ++k) ch[k] = __ldg(col + k);
  if (depths) ch[D] = __ldg(depths + f);
  c = GaussC{ch[0], ch[1], ch[2], ch[3]};
}


// Per-pixel forward state of one lane.
struct PixelState {
  float T, o0, o1, o2, o3;
  int last;
  bool done;
};

// Footprint of one warp: 8 x 4 pixels of the 16 x 16 tile.
struct WarpFootprint {
  float px, py;    // this lane's pixel centre
  float wcx, wcy;  // footprint centre (pixel centres span +-3.5, +-1.5 around it)
};

// Composite one staged batch (records in shared memory at aA/aB/aC, `batch_size` valid entries whose
// position in the tile's sorted list starts at `list_pos`) into the lane's pixel. Warp-level culling via
// ballot, branch-free body. Returns with st.done updated; callers handle the CTA-level early exit.
__device__ __forceinline__ void composite_batch_fwd(uint32_t aA, uint32_t aB, uint32_t aC, int batch_size,
                                                    int list_pos, const WarpFootprint &fp, PixelState &st) {
  const int lane = threadIdx.x & 31;
  for (int sub = 0; sub < batch_size; sub += 32) {
    const int j = sub + lane;
    bool hit = false;
    if (j < batch_size) {
      const float4 a = lds128(aA + j * 16);
      hit = fabsf(a.x - fp.wcx) <= a.z + 3.5f && fabsf(a.y - fp.wcy) <= a.w + 1.5f;
    }
    unsigned mask = __ballot_sync(0xffffffffu, hit);
    while (mask) {  // branch-free body: every lane evaluates, selects decide what sticks
      const int t = sub + __ffs(mask) - 1;
      mask &= mask - 1;
      const float2 m = lds64(aA + t * 16);
      const float4 b = lds128(aB + t * 16);
      const float4 c = lds128(aC + t * 16);
      const float dx = m.x - fp.px, dy = m.y - fp.py;
      const float p = power2(b.x, b.y, b.z, dx, dy);  // = -sigma * log2e
      const float alpha = fminf(0.999f, b.w * ex2_approx(p));
      const float next_T = st.T * (1.0f - alpha);
      const bool contrib = !st.done && !(p > 0.0f) && alpha >= kAlphaMin;  // p > 0 <=> sigma < 0
      const bool stop = contrib && next_T <= kTStop;
      const bool add = contrib && !stop;
      st.done = st.done || stop;
      const float w = add ? alpha * st.T : 0.0f;
      st.o0 = fmaf(c.x, w, st.o0); st.o1 = fmaf(c.y, w, st.o1); st.o2 = fmaf(c.z, w, st.o2); st.o3 = fmaf(c.w, w, st.o3);
      st.T = add ? next_T : st.T;
      st.last = add ? list_pos + t : st.last;
    }
    if (__all_sync(0xffffffffu, st.done)) break;
  }
}

// Write one pixel's results (shared epilogue).
__device__ __forceinline__ void store_pixel(const PixelState &st, size_t pix, int cam, int D, int CD,
                                            const float *__restrict__ backgrounds, float *__restrict__ render,
                                            float *__restrict__ alphas, int32_t *__restrict__ last_ids) {
  alphas[pix] = 1.0f - st.T;
  last_ids[pix] = st.last;
  float out[4] = {st.o0, st.o1, st.o2, st.o3};
  if (backgrounds) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < D) out[k] += st.T * __ldg(backgrounds + cam * D + k);
  }
  if (CD == 4) {
    reinterpret_cast<float4 *>(render)[pix] = make_float4(out[0], out[1], out[2], out[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < CD) render[pix * CD + k] = out[k];
  }
}

}  // namespace pm
