// composite_common.cuh — pieces shared by the compositing kernels (composite.cu) and the fused
// sort-as-you-go renderer (tile_render.cu): the staged Gaussian records, the gather that fills them,
// the explicitly rounded alpha evaluation and the per-batch forward loop. Keeping ONE copy of the
// arithmetic is what makes the fused path's images bit-identical to the separate-stage path.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace pm {

#ifdef PM_STATS
// developer build (tools/build_variant.sh stats -DPM_STATS): culling statistics of the forward compositor:
// [0] warp-chunks with >= 1 hit, [1] warp-level hits, [2] lane evaluations above the alpha threshold,
// [3] those on a live pixel, [4] hits with >= 1 such lane, [5] hits with >= 1 live such lane; fused renderer:
// [6] full passes over a tile's records, [7] groups, [8] level-2 refinements, [9] radix-sorted ("fat") groups,
// [10] staged Gaussians, [11] footprint candidates, [12] dense cull rounds, [13] non-empty tiles, [14] records
__device__ unsigned long long g_pm_stats[16];
#define PM_STAT_ADD(i, v) do { atomicAdd(&g_pm_stats[i], (unsigned long long)(v)); } while (0)
#else
#define PM_STAT_ADD(i, v) do { } while (0)
#endif

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

// The same value from the packed pair (mean - pixel): sm_100's two-wide fp32 instructions (FADD2 / FMUL2) form
// (dx, dy) and (dx^2, dy^2) in one issue slot each; every lane of the pair is an individually rounded IEEE operation,
// so the result is bit-identical to power2() — forward and backward both go through this one helper.
#ifndef PM_F32X2
#define PM_F32X2 1
#endif
__device__ __forceinline__ float power2_at(float Q0, float Q1, float Q2, float2 mean, float2 neg_pixel) {
#if PM_F32X2
  const float2 d = __fadd2_rn(mean, neg_pixel);
  const float2 sq = __fmul2_rn(d, d);
  return __fmaf_rn(Q0, sq.x, __fmaf_rn(Q2, sq.y, __fmul_rn(Q1, __fmul_rn(d.x, d.y))));
#else
  return power2(Q0, Q1, Q2, mean.x + neg_pixel.x, mean.y + neg_pixel.y);
#endif
}

// @region cc_gather
// Gather one Gaussian of the sorted list into the staging records, in two halves so that callers can
// put the global loads of several Gaussians in flight before the first use (gather_load) and do the
// arithmetic afterwards (gather_finish). f = flatten id (camera * N + g), g = Gaussian id.
struct GaussRaw {
  float2 m;
  float q0, q1, q2, o;
  float ch[4];
};
__device__ __forceinline__ GaussRaw gather_load(int32_t f, int32_t g, const float *__restrict__ means2d,
                                                const float *__restrict__ conics, const float *__restrict__ opacities,
                                                const float *__restrict__ colors, int colors_per_camera, int D,
                                                const float *__restrict__ depths) {
  GaussRaw r;
  r.m = __ldg(reinterpret_cast<const float2 *>(means2d) + f);
  r.q0 = __ldg(conics + (size_t)f * 3 + 0);
  r.q1 = __ldg(conics + (size_t)f * 3 + 1);
  r.q2 = __ldg(conics + (size_t)f * 3 + 2);
  r.o = __ldg(opacities + g);
  r.ch[0] = r.ch[1] = r.ch[2] = r.ch[3] = 0.0f;
  const float *col = colors + (size_t)(colors_per_camera ? f : g) * D;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (k < D) r.ch[k] = __ldg(col + k);
  if (depths) {
    const float d = __ldg(depths + f);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k == D) r.ch[k] = d;
  }
  return r;
}
__device__ __forceinline__ void gather_finish(const GaussRaw &r, GaussA &a, GaussB &b, GaussC &c) {
  const float q0 = r.q0, q1 = r.q1, q2 = r.q2, o = r.o;
  float hx, hy;  // extent of {alpha >= 1/255} (common.cuh)
  alpha_extents(q0, q1, q2, o, hx, hy);
  a = GaussA{r.m.x, r.m.y, hx, hy};
  // exp(-sigma) = 2^(Q0 dx^2 + Q1 dx dy + Q2 dy^2) with Q0 = -q0/2 * log2e, Q1 = -q1 * log2e, Q2 = -q2/2 * log2e
  b = GaussB{q0 * (-0.5f * kLog2e), q1 * (-kLog2e), q2 * (-0.5f * kLog2e), o};
  c = GaussC{r.ch[0], r.ch[1], r.ch[2], r.ch[3]};
}
__device__ __forceinline__ void gather(int32_t f, int32_t g, const float *__restrict__ means2d,
                                       const float *__restrict__ conics, const float *__restrict__ opacities,
                                       const float *__restrict__ colors, int colors_per_camera, int D,
                                       const float *__restrict__ depths, GaussA &a, GaussB &b, GaussC &c) {
  gather_finish(gather_load(f, g, means2d, conics, opacities, colors, colors_per_camera, D, depths), a, b, c);
}

// @region cc_stream
// ---- the splat stream: what the backward replays -------------------------------------------------------------
// When a backward will follow, the forward kernels write every Gaussian they stage — in the tile's sorted order, at
// its list position — as one 48-byte record {mean x, mean y, half extents (2 x fp16, rounded up), flatten id |
// scaled conic, opacity | colour, depth}. A tile's records are contiguous, so the backward streams a chunk of 32 of
// them (1536 bytes) into shared memory with ONE cp.async.bulk (TMA) and an mbarrier instead of 32 lanes x 7
// scattered loads through the L1 data pipe (which ran at 74 % of its peak doing that, profiles/r2d_*), and never
// recomputes log / sqrt / div for the extents.
constexpr int kStreamFloats = 12;
__device__ __forceinline__ float pack_extents(float hx, float hy) {
  // conservative: rounded towards +inf; 3e38 ("no culling") becomes +inf, -1e30 ("never visible") becomes -inf
  const unsigned lo = __half_as_ushort(__float2half_ru(hx)), hi = __half_as_ushort(__float2half_ru(hy));
  return __uint_as_float(lo | (hi << 16));
}
__device__ __forceinline__ float2 unpack_extents(float packed) {
  const unsigned u = __float_as_uint(packed);
  return make_float2(__half2float(__ushort_as_half((unsigned short)(u & 0xffffu))),
                     __half2float(__ushort_as_half((unsigned short)(u >> 16))));
}
__device__ __forceinline__ void stream_store(float *__restrict__ stream, size_t pos, int32_t f, const struct GaussA &a,
                                             const struct GaussB &b, const struct GaussC &c);

__device__ __forceinline__ void stream_store(float *__restrict__ stream, size_t pos, int32_t f, const GaussA &a,
                                             const GaussB &b, const GaussC &c) {
  float4 *dst = reinterpret_cast<float4 *>(stream) + pos * 3;
  dst[0] = make_float4(a.mx, a.my, pack_extents(a.hx, a.hy), __int_as_float(f));
  dst[1] = make_float4(b.cx, b.cy, b.cz, b.opac);
  dst[2] = make_float4(c.c0, c.c1, c.c2, c.c3);
}

// @region cc_eval_blend
// Per-pixel forward state of one lane. The sign of T carries the "done" flag: T > 0 while the pixel
// is live (the stop rule keeps a live T above 1e-4), T < 0 once it stopped or if it lies outside the
// image; |T| is the transmittance the epilogue writes. A done pixel can never add again because
// next_T = T (1 - alpha) stays negative, so no separate flag has to be tested in the hot loop.
struct PixelState {
  float T, o0, o1, o2, o3;
  int last;
};
__device__ __forceinline__ bool pixel_done(const PixelState &st) { return st.T < 0.0f; }

// Footprint of one warp: 8 x 4 pixels of the 16 x 16 tile.
struct WarpFootprint {
  float px, py;    // this lane's pixel centre
  float wcx, wcy;  // footprint centre (pixel centres span +-3.5, +-1.5 around it)
};

// Staged Gaussians of one batch / group, structure of arrays with compile-time strides so that the hot
// loop addresses all three records of a Gaussian from ONE register (+ immediates).
template <int CAP>
struct Staged {
  GaussA sA[CAP + 1];  // slot CAP is a null Gaussian (opacity 0): pads odd hit counts in the pairwise loop
  GaussB sB[CAP + 1];
  GaussC sC[CAP + 1];
};
constexpr int kQueueWords = 34;  // per-warp hit queue: up to 32 hits + pad, as shared-memory addresses

template <int CAP>
__device__ __forceinline__ void staged_init_null(Staged<CAP> &s) {
  if (threadIdx.x == 0) {
    s.sA[CAP] = GaussA{0.0f, 0.0f, -1e30f, -1e30f};
    s.sB[CAP] = GaussB{0.0f, 0.0f, 0.0f, 0.0f};
    s.sC[CAP] = GaussC{0.0f, 0.0f, 0.0f, 0.0f};
  }
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32u(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds64u(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}

// alpha of the staged Gaussian whose A record sits at shared address `ga`, at the lane's pixel, and
// whether it may contribute (sigma >= 0, alpha >= 1/255)
template <int CAP>
__device__ __forceinline__ void eval_alpha(uint32_t ga, const WarpFootprint &fp, float &alpha, bool &contrib) {
  const float2 m = lds64(ga);
  const float4 b = lds128(ga + (CAP + 1) * 16);
  const float p = power2_at(b.x, b.y, b.z, m, make_float2(-fp.px, -fp.py));  // = -sigma * log2e
  alpha = fminf(0.999f, b.w * ex2_approx(p));
  contrib = !(p > 0.0f) && alpha >= kAlphaMin;  // p > 0 <=> sigma < 0
}

// fold one evaluated Gaussian into the pixel (branch-free: selects decide what sticks)
__device__ __forceinline__ void blend(const float4 &c, float alpha, bool contrib, uint32_t ga, uint32_t &last_ga,
                                      PixelState &st) {
  const float next_T = st.T * (1.0f - alpha);
  const bool stop = contrib && next_T <= kTStop;  // also true for a done pixel (next_T < 0): harmless
  const bool add = contrib && !stop;
  const float w = add ? alpha * st.T : 0.0f;
  st.o0 = fmaf(c.x, w, st.o0); st.o1 = fmaf(c.y, w, st.o1); st.o2 = fmaf(c.z, w, st.o2); st.o3 = fmaf(c.w, w, st.o3);
  st.T = add ? next_T : st.T;
  st.T = stop ? -fabsf(st.T) : st.T;
  last_ga = add ? ga : last_ga;
}

// @region cc_cull
// Composite one staged run (`batch_size` valid entries of the Staged<CAP> block at shared address aS,
// whose position in the tile's sorted list starts at `list_pos`) into the lane's pixel.
//   * warp-level culling: 32 lanes test 32 Gaussians against the rectangle of the warp's LIVE pixels — first the
//     axis-aligned extent of {alpha >= 1/255}, then exactly (peak of the exponent over the rectangle) — ballot.
//     Measured before (C2): 16 % of the AABB hits reached the threshold on no pixel, 14 % only on finished ones;
//   * the hits' shared-memory addresses are compacted into the warp's queue (aQ, kQueueWords words), so
//     the evaluation loop walks a dense list — one LDS.64 per two hits instead of a bit scan, shift,
//     mask update and address computation per hit (10 of 38 instructions per hit, profiles/r1f_*);
//   * hits are evaluated two at a time (their alphas are independent, only the transmittance chain is
//     serial), which hides the LDS -> FMA -> MUFU latency with fewer resident warps. An odd count is
//     padded with the null Gaussian.
// Returns when the run is exhausted or every pixel of the warp is done; callers handle CTA-level exits.
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#ifndef PM_EXACT_CULL
#define PM_EXACT_CULL 1
#endif

// Rectangle spanned by the pixel centres of the lanes in `live` (lane = 8 * row + column of the warp's 8x4 footprint).
struct LiveRect {
  float x0, x1, y0, y1;
};
__device__ __forceinline__ LiveRect live_rect(unsigned live, float wcx, float wcy) {
  const unsigned cols = (live | (live >> 8) | (live >> 16) | (live >> 24)) & 0xffu;
  LiveRect r;
  r.x0 = wcx - 3.5f + (float)(__ffs(cols) - 1);
  r.x1 = wcx - 3.5f + (float)(31 - __clz(cols));
  r.y0 = wcy - 1.5f + (float)((__ffs(live) - 1) >> 3);
  r.y1 = wcy - 1.5f + (float)((31 - __clz(live)) >> 3);
  return r;
}

// Can the staged Gaussian (A record `a`, B record at shared address aB_j) reach alpha >= 1/255 on any point of
// the rectangle? First the axis-aligned extent of {alpha >= 1/255}, then exactly: the exponent
// p = Q0 dx^2 + Q1 dx dy + Q2 dy^2 is concave with its maximum (0) at the mean, so over a rectangle that does
// not contain the mean it peaks on the edge(s) facing the mean; the splat can contribute only if that peak is
// >= -log2(255 opacity). The margins make rounding err on the side of keeping a splat, so culling never changes
// a result.
template <typename LoadB>
__device__ __forceinline__ bool cull_hit_impl(const float4 &a, LoadB load_b, const LiveRect &r) {
  const float dx0 = r.x0 - a.x, dx1 = r.x1 - a.x, dy0 = r.y0 - a.y, dy1 = r.y1 - a.y;
  bool hit = dx0 <= a.z && dx1 >= -a.z && dy0 <= a.w && dy1 >= -a.w;
#if PM_EXACT_CULL
  if (hit && a.z < 1e37f) {
    const float ax = fminf(fmaxf(0.0f, dx0), dx1), by = fminf(fmaxf(0.0f, dy0), dy1);
    if (ax != 0.0f || by != 0.0f) {
      const float4 b = load_b();
      // (slack = a relative bound on the rounding of the three products: strongly anisotropic conics cancel)
      float pmax = -3e38f, slack = 0.0f;
      if (ax != 0.0f) {
        const float dy = fminf(fmaxf(-0.5f * b.y * ax * rcp_approx(b.z), dy0), dy1);
        const float t0 = b.x * ax * ax, t1 = b.y * ax * dy, t2 = b.z * dy * dy;
        pmax = t0 + t1 + t2;
        slack = fabsf(t0) + fabsf(t1) + fabsf(t2);
      }
      if (by != 0.0f) {
        const float dx = fminf(fmaxf(-0.5f * b.y * by * rcp_approx(b.x), dx0), dx1);
        const float t0 = b.x * dx * dx, t1 = b.y * dx * by, t2 = b.z * by * by;
        pmax = fmaxf(pmax, t0 + t1 + t2);
        slack = fmaxf(slack, fabsf(t0) + fabsf(t1) + fabsf(t2));
      }
      const float tau2 = lg2_approx(255.0f * b.w);
      hit = pmax + 4e-6f * slack >= -tau2 * 1.0005f - 2e-3f;
    }
  }
#endif
  return hit;
}
// B record in shared memory (forward: the whole group is staged) or already in registers (backward: staged only if hit)
__device__ __forceinline__ bool cull_hit(const float4 &a, uint32_t aB_j, const LiveRect &r) {
  return cull_hit_impl(a, [aB_j] { return lds128(aB_j); }, r);
}
__device__ __forceinline__ bool cull_hit_reg(const float4 &a, const float4 &b, const LiveRect &r) {
  return cull_hit_impl(a, [&b] { return b; }, r);
}

// @region cc_batch_loop
template <int CAP>
__device__ __forceinline__ void composite_batch_fwd(uint32_t aS, uint32_t aQ, int batch_size, int list_pos,
                                                    const WarpFootprint &fp, PixelState &st) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  uint32_t last_ga = 0xffffffffu;  // explicit 'nothing blended' sentinel (a staged slot may sit at shared address 0)
  for (int sub = 0; sub < batch_size; sub += 32) {
    // Gaussians are culled against the rectangle of the warp's LIVE pixels, so splats that only cover finished
    // pixels are skipped as well
    const unsigned live = __ballot_sync(0xffffffffu, !pixel_done(st));
    if (live == 0u) break;
    const LiveRect rect = live_rect(live, fp.wcx, fp.wcy);
    const int j = sub + lane;
    bool hit = false;
    if (j < batch_size) hit = cull_hit(lds128(aS + j * 16), aS + j * 16 + (CAP + 1) * 16, rect);
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (bal == 0u) continue;
    const int n = __popc(bal);
#ifdef PM_STATS
    if (lane == 0) { atomicAdd(&g_pm_stats[0], 1ull); atomicAdd(&g_pm_stats[1], (unsigned long long)n); }
#endif
    if (hit) sts32(aQ + __popc(bal & lt) * 4, aS + j * 16);
    if (lane == 0) sts32(aQ + n * 4, aS + CAP * 16);
    __syncwarp();
    for (int i = 0; i < n; i += 2) {
      const uint2 ga = lds64u(aQ + i * 4);
      float alpha0, alpha1;
      bool c0, c1;
      eval_alpha<CAP>(ga.x, fp, alpha0, c0);
      eval_alpha<CAP>(ga.y, fp, alpha1, c1);
#ifdef PM_STATS
      {  // lanes whose alpha clears the 1/255 threshold, and those that are also still live
        const unsigned b0 = __ballot_sync(0xffffffffu, c0), b1 = __ballot_sync(0xffffffffu, c1 && ga.y != aS + CAP * 16);
        const unsigned l0 = __ballot_sync(0xffffffffu, c0 && st.T > 0.0f), l1 = __ballot_sync(0xffffffffu, c1 && st.T > 0.0f && ga.y != aS + CAP * 16);
        if (lane == 0) {
          atomicAdd(&g_pm_stats[2], (unsigned long long)(__popc(b0) + __popc(b1)));
          atomicAdd(&g_pm_stats[3], (unsigned long long)(__popc(l0) + __popc(l1)));
          atomicAdd(&g_pm_stats[4], (unsigned long long)((b0 != 0) + (b1 != 0)));
          atomicAdd(&g_pm_stats[5], (unsigned long long)((l0 != 0) + (l1 != 0)));
        }
      }
#endif
      const float4 col0 = lds128(ga.x + 2 * (CAP + 1) * 16);
      const float4 col1 = lds128(ga.y + 2 * (CAP + 1) * 16);
      blend(col0, alpha0, c0, ga.x, last_ga, st);
      blend(col1, alpha1, c1, ga.y, last_ga, st);
    }
    __syncwarp();
  }
  if (last_ga != 0xffffffffu) st.last = list_pos + (int)((last_ga - aS) >> 4);
}

// @region cc_footprint_mask
// ---- footprint masks (fused renderer) ------------------------------------------------------------------------
// composite_batch_fwd makes every warp test every staged Gaussian: 8 warps x ~120 instructions per 32 entries, of
// which ~20 % survive (the splats of these scenes are a few pixels wide, a warp's footprint is 8 x 4). The fused
// renderer's staging threads hold each Gaussian's extent in registers anyway, so ONE thread tests it against the 8
// static warp footprints of the tile (footprint_mask: two column and four row intervals) and the staging warp
// publishes a 32-bit word per (footprint, 32 staged entries) with 8 ballots. A compositing warp then only walks the
// non-zero words of ITS footprint, gathers the candidates into a ring of 64 shared-memory addresses and runs the
// live-rectangle test (AABB + exact) on 32 DENSE candidates at a time — same decisions, same order, same arithmetic
// as composite_batch_fwd (culling never changes a result), ~3x fewer instructions outside the evaluation itself.
constexpr int kCandWords = 64;  // per-warp candidate ring

// bit w (= 2 * row + column, the warp index) set when the axis-aligned extent of {alpha >= 1/255} overlaps the pixel
// centres of warp w's 8 x 4 footprint; tx0 / ty0 = first pixel centre of the tile
__device__ __forceinline__ unsigned footprint_mask(float mx, float my, float hx, float hy, float tx0, float ty0) {
  unsigned cols = 0u, fm = 0u;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const float x0 = tx0 + (float)(8 * c), x1 = x0 + 7.0f;
    if (x0 - mx <= hx && x1 - mx >= -hx) cols |= 1u << c;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float y0 = ty0 + (float)(4 * r), y1 = y0 + 3.0f;
    if (y0 - my <= hy && y1 - my >= -hy) fm |= cols << (2 * r);
  }
  return fm;
}

// @region cc_masked_loop
// Composite one staged group through the warp's footprint words (aM: ceil(batch_size / 32) words, bit = staged slot).
template <int CAP>
__device__ __forceinline__ void composite_group_masked_fwd(uint32_t aS, uint32_t aQ, uint32_t aCQ, uint32_t aM,
                                                           int batch_size, int list_pos, const WarpFootprint &fp,
                                                           PixelState &st, uint32_t *__restrict__ hit_words = nullptr,
                                                           int rel0 = 0) {
  static_assert(CAP <= 1024, "one footprint word per lane");
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int nch = (batch_size + 31) >> 5;
  const uint32_t words = lane < nch ? lds32u(aM + lane * 4) : 0u;
  unsigned nz = __ballot_sync(0xffffffffu, words != 0u);
  uint32_t last_ga = 0xffffffffu;
  int head = 0, q = 0;  // the ring holds candidates head .. head + q - 1 (mod kCandWords), in list order
  while (nz != 0u || q > 0) {
    // refill: append whole words while fewer than 32 candidates wait (q < 32 before, <= 63 after)
    while (q < 32 && nz != 0u) {
      const int c = __ffs(nz) - 1;
      nz &= nz - 1u;
      const uint32_t word = __shfl_sync(0xffffffffu, words, c);
      if ((word >> lane) & 1u)
        sts32(aCQ + ((head + q + __popc(word & lt)) & (kCandWords - 1)) * 4, aS + (c * 32 + lane) * 16);
      q += __popc(word);
    }
    __syncwarp();
    const unsigned live = __ballot_sync(0xffffffffu, !pixel_done(st));
    if (live == 0u) break;
    const LiveRect rect = live_rect(live, fp.wcx, fp.wcy);
    const int nb = min(q, 32);
    if (lane == 0) { PM_STAT_ADD(11, nb); PM_STAT_ADD(12, 1); }
    bool hit = false;
    uint32_t ga = 0u;
    if (lane < nb) {
      ga = lds32u(aCQ + ((head + lane) & (kCandWords - 1)) * 4);
      hit = cull_hit(lds128(ga), ga + (CAP + 1) * 16, rect);
    }
    head = (head + nb) & (kCandWords - 1);
    q -= nb;
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (bal == 0u) continue;
    const int n = __popc(bal);
#ifdef PM_STATS
    if (lane == 0) { atomicAdd(&g_pm_stats[0], 1ull); atomicAdd(&g_pm_stats[1], (unsigned long long)n); }
#endif
    if (hit) {
      sts32(aQ + __popc(bal & lt) * 4, ga);
      // hit map for the backward (one bit per (warp footprint, list entry), zero-filled by the caller): the entries this
      // warp evaluates are exactly the ones its backward has to replay — it then skips its own culling
      if (hit_words) {
        const int p = rel0 + (int)((ga - aS) >> 4);  // position in the tile's list
        atomicOr(hit_words + (p >> 5), 1u << (p & 31));
      }
    }
    if (lane == 0) sts32(aQ + n * 4, aS + CAP * 16);
    __syncwarp();
    for (int i = 0; i < n; i += 2) {
      const uint2 g2 = lds64u(aQ + i * 4);
      float alpha0, alpha1;
      bool c0, c1;
      eval_alpha<CAP>(g2.x, fp, alpha0, c0);
      eval_alpha<CAP>(g2.y, fp, alpha1, c1);
      const float4 col0 = lds128(g2.x + 2 * (CAP + 1) * 16);
      const float4 col1 = lds128(g2.y + 2 * (CAP + 1) * 16);
      blend(col0, alpha0, c0, g2.x, last_ga, st);
      blend(col1, alpha1, c1, g2.y, last_ga, st);
    }
    __syncwarp();
  }
  if (last_ga != 0xffffffffu) st.last = list_pos + (int)((last_ga - aS) >> 4);
}

// @region cc_epilogue
// Write one pixel's results (shared epilogue).
__device__ __forceinline__ void store_pixel(const PixelState &st, size_t pix, int cam, int D, int CD,
                                            const float *__restrict__ backgrounds, float *__restrict__ render,
                                            float *__restrict__ alphas, int32_t *__restrict__ last_ids,
                                            float *__restrict__ transmittance) {
  const float T = fabsf(st.T);
  alphas[pix] = 1.0f - T;
  last_ids[pix] = st.last;
  // The backward needs the final transmittance. Upstream rebuilds it as 1 - render_alpha, which loses up to
  // ulp(alpha) / T ~ 3e-4 relative on an opaque pixel; the exact fp32 T is 4 bytes per pixel.
  if (transmittance) transmittance[pix] = T;
  float out[4] = {st.o0, st.o1, st.o2, st.o3};
  if (backgrounds) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < D) out[k] += T * __ldg(backgrounds + cam * D + k);
  }
  if (CD == 4) {
    reinterpret_cast<float4 *>(render)[pix] = make_float4(out[0], out[1], out[2], out[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < CD) render[pix * CD + k] = out[k];
  }
}

// Packed epilogue of the module fast path (RGB + expected depth, D = 3): writes what
// GaussianRenderer.forward returns (gaussians_renderer.py:437-440, :507-509) straight from the
// compositor's registers — images [C,3,H,W] = 2 rgb - 1, depths [C,1,H,W] = sum(depth alpha T) /
// max(alpha, 1e-10) (gsplat's "ED" normalisation), masks [C,1,H,W] = alpha clamped to [0,1] — instead of
// an interleaved render tensor that six elementwise / copy kernels then slice, divide, permute and rescale.
// Same roundings as those ATen kernels, so the images are bit-identical to the unpacked path.
__device__ __forceinline__ void store_pixel_packed(const PixelState &st, int cam, int iy, int ix, int width, int height,
                                                   const float *__restrict__ backgrounds, float *__restrict__ images,
                                                   float *__restrict__ depth_out, float *__restrict__ masks,
                                                   int32_t *__restrict__ last_ids, float *__restrict__ transmittance) {
  const float T = fabsf(st.T);
  const float alpha = 1.0f - T;
  const size_t plane = (size_t)height * width;
  const size_t in_plane = (size_t)iy * width + ix;
  const size_t pix = (size_t)cam * plane + in_plane;
  float out[3] = {st.o0, st.o1, st.o2};
  if (backgrounds) {
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] += T * __ldg(backgrounds + cam * 3 + k);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) images[((size_t)cam * 3 + k) * plane + in_plane] = __fmaf_rn(2.0f, out[k], -1.0f);
  depth_out[pix] = __fdiv_rn(st.o3, fmaxf(alpha, 1e-10f));
  masks[pix] = fminf(fmaxf(alpha, 0.0f), 1.0f);
  last_ids[pix] = st.last;
  if (transmittance) transmittance[pix] = T;
}

}  // namespace pm
