// composite.cu — stage 4: front-to-back alpha compositing of the depth-sorted
// per-tile Gaussian lists, forward and backward.
//
// Replaces gsplat 1.4.0 rasterize_to_pixels_{fwd,bwd} (packed=False, tile 16),
// reached by the reference through gsplat.rasterization
// (prometheus/modules/renderers/gaussians_renderer.py:413-427); semantics per
// SURVEY.md §8c steps 10 and 12.
//
// Design (B200): one 256-thread CTA per 16x16 tile; each of its 8 warps owns an
// 8x4-pixel footprint. The tile's sorted list is staged in shared memory in
// batches of 256 Gaussians (each thread gathers one record: flatten id -> mean,
// conic, opacity, colour, depth = 44 algorithmic bytes). Instead of letting all
// 256 pixels evaluate all 256 Gaussians (upstream), every warp first CULLS the
// batch against its footprint: 32 lanes test 32 Gaussians in parallel against
// the exact alpha >= 1/255 extent of each Gaussian (an axis-aligned bound of the
// ellipse sigma <= ln(255*opacity)), __ballot_sync turns the hits into a bitmask,
// and only the set bits are evaluated, in order. The bound is conservative, so
// the surviving evaluations — and therefore the result — are identical to the
// un-culled loop. Early termination is tracked per warp (__all_sync) and per CTA
// (__syncthreads_count), so batches past the last needed intersection are never
// loaded. The backward pass replays the same culled lists back to front and reduces
// each Gaussian's gradient over the warp's pixels with a transposed (splat-parallel)
// pass through shared memory, then issues fire-and-forget global reductions per
// (warp, Gaussian, component) — no per-pixel global atomics. The warps of the backward
// are AUTONOMOUS: each gathers its own 32-entry chunks (the other warps of the tile
// hit the same lines in L1/L2) and walks only the prefix its own pixels need, so there
// is no CTA barrier at all — the per-round barrier of the CTA-wide version was 36 % of
// its stall samples (profiles/r1j_composite_bwd_lines.txt).
#include "composite_common.cuh"

namespace pm {

__global__ void __launch_bounds__(256)
composite_fwd_kernel(const float *__restrict__ means2d, const float *__restrict__ conics,
                     const float *__restrict__ opacities, const float *__restrict__ colors,
                     int colors_per_camera, int D, const float *__restrict__ depths,
                     const float *__restrict__ backgrounds, const int32_t *__restrict__ offsets,
                     const int32_t *__restrict__ flatten_ids, int64_t n_host,
                     const int32_t *__restrict__ d_n, int N, int width, int height, int tile_width,
                     int tile_height, float *__restrict__ render, float *__restrict__ alphas,
                     int32_t *__restrict__ last_ids, float *__restrict__ transmittance,
                     float *__restrict__ stream) {
  __shared__ Staged<kBatch> stg;
  __shared__ uint32_t queue[8][kQueueWords];
  staged_init_null(stg);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_tiles = tile_width * tile_height;
  const int ct = blockIdx.x;  // (camera, tile) flattened
  const int cam = ct / n_tiles, tile = ct % n_tiles;
  const int tile_y = tile / tile_width, tile_x = tile % tile_width;
  // warp footprint: 8 wide x 4 high
  const int wx0 = tile_x * kTile + (warp & 1) * 8, wy0 = tile_y * kTile + (warp >> 1) * 4;
  const int ix = wx0 + (lane & 7), iy = wy0 + (lane >> 3);
  const float px = (float)ix + 0.5f, py = (float)iy + 0.5f;
  const float wcx = (float)wx0 + 4.0f, wcy = (float)wy0 + 2.0f;  // footprint centre (pixel centres span +-3.5, +-1.5)
  const bool inside = ix < width && iy < height;
  const int CD = D + (depths ? 1 : 0);

  const int64_t n_isects = effective_n(n_host, d_n);
  const int range_start = offsets[ct];
  const int range_end = (ct == (int)gridDim.x - 1) ? (int)n_isects : offsets[ct + 1];

  PixelState st{inside ? 1.0f : -1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0};
  const WarpFootprint fp{px, py, wcx, wcy};
  const uint32_t aS = smem_u32(&stg), aQ = smem_u32(queue[warp]);

  for (int batch_start = range_start; batch_start < range_end; batch_start += kBatch) {
    if (__syncthreads_count(pixel_done(st)) == kBatch) break;
    const int idx = batch_start + tid;
    if (idx < range_end) {
      const int32_t f = flatten_ids[idx];
      GaussA ga; GaussB gb; GaussC gc;
      gather(f, f - cam * N, means2d, conics, opacities, colors, colors_per_camera, D, depths, ga, gb, gc);
      stg.sA[tid] = ga; stg.sB[tid] = gb; stg.sC[tid] = gc;
      if (stream) stream_store(stream, (size_t)idx, f, ga, gb, gc);  // what the backward replays
    }
    __syncthreads();
    const int batch_size = min(kBatch, range_end - batch_start);
    if (__all_sync(0xffffffffu, pixel_done(st))) continue;  // this warp is finished; keep hitting the barriers
    composite_batch_fwd<kBatch>(aS, aQ, batch_size, batch_start, fp, st);
  }

  if (inside) store_pixel(st, ((size_t)cam * height + iy) * width + ix, cam, D, CD, backgrounds, render, alphas, last_ids, transmittance);
}

// ---------------------------------------------------------------------------
// Backward.
// ---------------------------------------------------------------------------
// Warp-autonomous replay with a TRANSPOSED reduction. Each warp owns one 8x4-pixel footprint and walks the prefix of
// the tile's sorted list its pixels need, back to front, 32 list entries at a time:
//   stage   lane j gathers entry j into REGISTERS, culls it against the rectangle of the warp's live pixels (same
//           exact test as the forward) and, if it may contribute, appends its 48-byte record to the warp's hit queue
//           in shared memory — entries the warp cannot see are never staged at all;
//   phase 1 (pixel-parallel, lane = pixel) as soon as 16 hits are queued: every pixel replays the 16 hits in order
//           — the only serial part, the transmittance chain — and leaves two scalars per (pixel, hit) in a 32 x 16
//           shared-memory matrix: fac = alpha T and vsig = d loss / d sigma. Every gradient component of a splat is
//           a sum over pixels of one of them times a polynomial in (dx, dy):
//             v_colour_k = sum fac v_k      v_conic = (1/2 sum vsig dx^2, sum vsig dx dy, 1/2 sum vsig dy^2)
//             v_mean2d   = conic (sum vsig dx, sum vsig dy)            v_opacity = -sum vsig / opacity
//   phase 2 (splat-parallel, lane = (hit, half of the pixels)): each lane reads its hit's column of the matrix and
//           accumulates the ten sums in registers over 16 pixels; one shuffle per sum folds the two halves and each
//           lane issues five fire-and-forget global reductions.
// Round 1 reduced every hit across the warp with a 12-shuffle butterfly (~135 instructions per (warp, hit), 2.18 G warp
// instructions on C2, 84 % issue-bound); the transposed form needs ~45 + ~20.
__device__ __forceinline__ bool cull_hit_reg(const float4 &a, const float4 &b, const LiveRect &r);

#ifndef PM_BWD_MINBLOCKS
#define PM_BWD_MINBLOCKS 6
#endif
#ifndef PM_BWD_P1_UNROLL
#define PM_BWD_P1_UNROLL 4
#endif
constexpr int kBwdP1Unroll = PM_BWD_P1_UNROLL;
#ifndef PM_BWD_P2_UNROLL
#define PM_BWD_P2_UNROLL 8
#endif
constexpr int kBwdP2Unroll = PM_BWD_P2_UNROLL;
constexpr int kBwdWarps = 4;   // a CTA holds the upper or the lower half of a tile's eight warp footprints
constexpr int kBwdThreads = 32 * kBwdWarps;
constexpr int kBwdParts = 8 / kBwdWarps;
constexpr int kBwdRound = 16;  // hits per reduction round
constexpr int kBwdQueue = 48;  // hit-queue slots per warp (linear): at most 15 left over + 32 new
constexpr int kBwdRow = 66;    // floats per matrix row: 32 pixels x {fac, vsig} + 2 of padding (conflict-free columns)

struct BwdWarpShared {
  float4 qA[kBwdQueue];  // {mean x, mean y, list position (int bits), flatten id (int bits)}
  float4 qB[kBwdQueue];  // conic in base-2 exponent units + opacity
  float4 qC[kBwdQueue];  // colour channels (+ depth)
  float mat[kBwdRound * kBwdRow];
  float4 vr[32];         // per-pixel cotangents of the (up to) four render channels
  float4 stage[32 * 3];  // TMA landing zone: the next chunk's 32 stream records (48 bytes each)
  unsigned long long bar;  // its mbarrier
};
struct BwdShared {
  BwdWarpShared w[kBwdWarps];
};

// MAP: the forward left a hit map (pm_tile_render_packed_fwd) — no culling code in that instantiation
template <bool MAP>
__global__ void __launch_bounds__(kBwdThreads, PM_BWD_MINBLOCKS)
composite_bwd_kernel(const float *__restrict__ stream, int colors_per_camera, int D, int has_depth,
                     const float *__restrict__ backgrounds, const int32_t *__restrict__ offsets, int64_t n_host,
                     const int32_t *__restrict__ d_n, int N, int width, int height, int tile_width,
                     int tile_height, const float *__restrict__ alphas,
                     const int32_t *__restrict__ last_ids, const float *__restrict__ v_render,
                     const float *__restrict__ v_alphas, float *__restrict__ v_means2d,
                     float *__restrict__ v_conics, float *__restrict__ v_colors,
                     float *__restrict__ v_opacities, float *__restrict__ v_depths,
                     const float *__restrict__ v_images, const float *__restrict__ v_depth_out,
                     const float *__restrict__ v_masks, const float *__restrict__ depth_out, int packed,
                     const int32_t *__restrict__ tile_order, int views_per_scene,
                     const float *__restrict__ transmittance, uint8_t *__restrict__ touched,
                     const uint32_t *__restrict__ hit_words, int64_t hit_stride) {
  extern __shared__ __align__(16) unsigned char bwd_smem[];
  BwdWarpShared &sh = reinterpret_cast<BwdShared *>(bwd_smem)->w[threadIdx.x >> 5];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = (int)(blockIdx.x % kBwdParts) * kBwdWarps + (tid >> 5);  // footprint index inside the tile
  const int n_tiles = tile_width * tile_height;
  const int n_ct = gridDim.x / kBwdParts;
  const int ct = tile_order ? tile_order[blockIdx.x / kBwdParts] : (int)(blockIdx.x / kBwdParts);  // longest lists first
  const int cam = ct / n_tiles, tile = ct % n_tiles;
  const int tile_y = tile / tile_width, tile_x = tile % tile_width;
  const int wx0 = tile_x * kTile + (warp & 1) * 8, wy0 = tile_y * kTile + (warp >> 1) * 4;
  const int ix = wx0 + (lane & 7), iy = wy0 + (lane >> 3);
  const float px = (float)ix + 0.5f, py = (float)iy + 0.5f;
  const float wcx = (float)wx0 + 4.0f, wcy = (float)wy0 + 2.0f;
  const bool inside = ix < width && iy < height;
  const int CD = D + has_depth;
  // flatten id f = cam * N + g; per-Gaussian arrays hold one block of N per scene (scene = cam / views_per_scene)
  const int g_base = cam * N - (cam / views_per_scene) * N;

  const int64_t n_isects = effective_n(n_host, d_n);
  const int range_start = offsets[ct];
  const int range_end = (ct == n_ct - 1) ? (int)n_isects : offsets[ct + 1];
  if (range_end <= range_start) return;

  // @region bwd_prologue
  // per-pixel state
  float T_final = 1.0f, vr[4] = {0.0f, 0.0f, 0.0f, 0.0f}, va = 0.0f;
  int my_last = -1;  // nothing to replay for pixels outside the image
  if (inside && packed) {
    // cotangents of the packed module outputs (store_pixel_packed): images = 2 rgb - 1, depth_out =
    // acc_depth / max(alpha, 1e-10), masks = clamp(alpha, 0, 1); `alphas` is the masks tensor
    const size_t plane = (size_t)height * width, in_plane = (size_t)iy * width + ix;
    const size_t pix = (size_t)cam * plane + in_plane;
    const float alpha = alphas[pix];
    T_final = transmittance ? transmittance[pix] : 1.0f - alpha;  // exact T of the forward when it was kept
    my_last = last_ids[pix];
    if (v_images) {
#pragma unroll
      for (int k = 0; k < 3; ++k) vr[k] = 2.0f * v_images[((size_t)cam * 3 + k) * plane + in_plane];
    }
    if (v_masks) va = v_masks[pix];
    if (v_depth_out) {
      const float vd = v_depth_out[pix];
      const float ac = fmaxf(alpha, 1e-10f);
      vr[3] = vd / ac;
      if (alpha >= 1e-10f) va -= depth_out[pix] * vd / ac;  // d(acc / alpha) / d alpha = -depth_out / alpha
    }
  } else if (inside) {
    const size_t pix = ((size_t)cam * height + iy) * width + ix;
    T_final = transmittance ? transmittance[pix] : 1.0f - alphas[pix];
    my_last = last_ids[pix];
    if (v_render) {
      if (CD == 4) {
        const float4 t = reinterpret_cast<const float4 *>(v_render)[pix];
        vr[0] = t.x; vr[1] = t.y; vr[2] = t.z; vr[3] = t.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < CD) vr[k] = v_render[pix * CD + k];
      }
    }
    if (v_alphas) va = v_alphas[pix];
  }
  // -T_final * sum_k bg_k v_render_k, folded into the alpha-gradient seed
  float bg_dot = 0.0f;
  if (backgrounds) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k < D) bg_dot += __ldg(backgrounds + cam * D + k) * vr[k];
  }
  const float va_seed = T_final * (va - bg_dot);

  int wl = my_last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wl = max(wl, __shfl_xor_sync(0xffffffffu, wl, o));
  if (wl < range_start) return;

  const uint32_t aA = smem_u32(sh.qA), aB = smem_u32(sh.qB), aC = smem_u32(sh.qC), aM = smem_u32(sh.mat),
                 aV = smem_u32(sh.vr);
  sh.vr[lane] = make_float4(vr[0], vr[1], vr[2], vr[3]);
  __syncwarp();

  // phase-2 role of this lane: hit h of the round, pixels 16 * half .. 16 * half + 15 (rows 2 * half, 2 * half + 1)
  const int h = lane & 15, half = lane >> 4;
  const float px_base = (float)wx0 + 0.5f, py_base = (float)(wy0 + 2 * half) + 0.5f;

  float T = T_final;
  float bsum = -va_seed;  // see phase 1

  // One reduction round over the hits in queue slots s0 .. s0 + 15, of which the first n are real (a short last
  // round is padded with null records: opacity 0 -> alpha 0 -> nothing stored).
  auto round = [&](int s0, int n) {
    // @region bwd_phase1
    // ---- phase 1: lane = pixel. The three records of hit i + 1 are fetched while hit i is evaluated.
    float4 a = lds128(aA + s0 * 16), b = lds128(aB + s0 * 16);
#pragma unroll kBwdP1Unroll
    for (int i = 0; i < kBwdRound; ++i) {
      float4 an = a, bn = b;
      if (i + 1 < kBwdRound) {
        an = lds128(aA + (s0 + i + 1) * 16);
        bn = lds128(aB + (s0 + i + 1) * 16);
      }
      const float p = power2_at(b.x, b.y, b.z, make_float2(a.x, a.y), make_float2(-px, -py));
      const float vis = ex2_approx(p);
      const float a_raw = b.w * vis;
      const float alpha = fminf(0.999f, a_raw);
      const bool valid = (__float_as_int(a.z) <= my_last) && !(p > 0.0f) && alpha >= kAlphaMin;
      float fac = 0.0f, vsig = 0.0f;
      if (__any_sync(0xffffffffu, valid)) {
        const float4 c = lds128(aC + (s0 + i) * 16);
        // branch-free: an invalid lane replays alpha = 0 (T, bsum unchanged, zeros stored).
        // d loss / d alpha = sum_k (c_k T - buf_k / (1 - alpha)) v_k + T_final (v_alpha - bg . v) / (1 - alpha) with
        // buf_k = what lies behind this splat in channel k. Only buf . v is ever needed, so ONE running scalar
        // bsum = buf . v - T_final (v_alpha - bg . v) replaces the four buffers: 7 operations instead of 17.
        const float al = valid ? alpha : 0.0f;
        const float ra = rcp_approx(1.0f - al);
        T *= ra;
        fac = al * T;
        const float cv = fmaf(c.x, vr[0], fmaf(c.y, vr[1], fmaf(c.z, vr[2], c.w * vr[3])));
        const float v_alpha = fmaf(T, cv, -(ra * bsum));
        vsig = (valid && a_raw <= 0.999f) ? -a_raw * v_alpha : 0.0f;
        bsum = fmaf(fac, cv, bsum);
      }
      asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(aM + (i * kBwdRow + 2 * lane) * 4), "f"(fac), "f"(vsig) : "memory");
      a = an; b = bn;
    }
    __syncwarp();
    // @region bwd_phase2
    // ---- phase 2: lane = (hit h, half of the pixels)
    const float4 ha = lds128(aA + (s0 + h) * 16);
    const float4 hb = lds128(aB + (s0 + h) * 16);
    const float dx0 = ha.x - px_base, dy0 = ha.y - py_base;
    float c0, c1, c2, c3, sxx, sxy = 0.0f, syy, sx, sy, s00 = 0.0f;
    {
      // two-wide fp32 (FFMA2 / FMUL2 / FADD2): the ten sums as three pairs + two scalars
      float2 c01 = make_float2(0.0f, 0.0f), c23 = c01, sq = c01, sl = c01;
#pragma unroll kBwdP2Unroll
      for (int j = 0; j < 16; ++j) {
        const float2 fv = lds64(aM + (h * kBwdRow + 2 * (16 * half + j)) * 4);
        const float4 v = lds128(aV + (16 * half + j) * 16);
        const float2 d = make_float2(dx0 - (float)(j & 7), dy0 - (float)(j >> 3));
        const float2 ff = make_float2(fv.x, fv.x);
        c01 = __ffma2_rn(ff, make_float2(v.x, v.y), c01);
        c23 = __ffma2_rn(ff, make_float2(v.z, v.w), c23);
        const float2 t = __fmul2_rn(make_float2(fv.y, fv.y), d);  // (vsig dx, vsig dy)
        sq = __ffma2_rn(t, d, sq);                                // (sum vsig dx^2, sum vsig dy^2)
        sxy = fmaf(t.x, d.y, sxy);
        sl = __fadd2_rn(sl, t);
        s00 += fv.y;
      }
      c0 = c01.x; c1 = c01.y; c2 = c23.x; c3 = c23.y; sxx = sq.x; syy = sq.y; sx = sl.x; sy = sl.y;
    }
    // fold the two halves (lanes h and h + 16 hold the same hit): afterwards the lower half owns colours, depth
    // and opacity, the upper half conic and mean
    c0 += __shfl_xor_sync(0xffffffffu, c0, 16); c1 += __shfl_xor_sync(0xffffffffu, c1, 16);
    c2 += __shfl_xor_sync(0xffffffffu, c2, 16); c3 += __shfl_xor_sync(0xffffffffu, c3, 16);
    sxx += __shfl_xor_sync(0xffffffffu, sxx, 16); sxy += __shfl_xor_sync(0xffffffffu, sxy, 16);
    syy += __shfl_xor_sync(0xffffffffu, syy, 16); sx += __shfl_xor_sync(0xffffffffu, sx, 16);
    sy += __shfl_xor_sync(0xffffffffu, sy, 16); s00 += __shfl_xor_sync(0xffffffffu, s00, 16);
    // @region bwd_atomics
    if (h < n) {
      const int32_t f = __float_as_int(ha.w);
      const int32_t g = f - g_base;
      if (half == 0) {
        // one byte per (camera, Gaussian) pair that received a gradient: pm_project_bwd then finds its work from
        // 1 byte per pair instead of 28 (a plain idempotent store; the bit-mask + atomicOr form cost more than it saved)
        if (touched) touched[f] = 1;
        const size_t ci = (size_t)(colors_per_camera ? f : g) * D;
        const float cs[4] = {c0, c1, c2, c3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < D) { if (cs[k] != 0.0f) atomicAdd(v_colors + ci + k, cs[k]); }
          else if (k == D && has_depth) { if (cs[k] != 0.0f) atomicAdd(v_depths + f, cs[k]); }
        }
        if (s00 != 0.0f) atomicAdd(v_opacities + g, -s00 * rcp_approx(hb.w));
      } else {
        if (sxx != 0.0f) atomicAdd(v_conics + (size_t)f * 3 + 0, 0.5f * sxx);
        if (sxy != 0.0f) atomicAdd(v_conics + (size_t)f * 3 + 1, sxy);
        if (syy != 0.0f) atomicAdd(v_conics + (size_t)f * 3 + 2, 0.5f * syy);
        // v_mean2d = conic (sx, sy); the staged conic is pre-scaled: q0 = -2 Q0 / log2e, q1 = -Q1 / log2e, q2 = -2 Q2 / log2e
        const float k2 = -1.0f / kLog2e;
        const float mx_ = k2 * (2.0f * hb.x * sx + hb.y * sy), my_ = k2 * (hb.y * sx + 2.0f * hb.z * sy);
        if (mx_ != 0.0f) atomicAdd(v_means2d + (size_t)f * 2 + 0, mx_);
        if (my_ != 0.0f) atomicAdd(v_means2d + (size_t)f * 2 + 1, my_);
      }
    }
    __syncwarp();
  };

  // @region bwd_stage_cull
  // The list is walked in chunks of 32 entries, back to front. A chunk = 32 consecutive records of the splat stream
  // the forward wrote (1536 contiguous bytes): lane 0 arms the warp's mbarrier with the byte count and issues ONE
  // cp.async.bulk; the copy of chunk k - 1 runs while chunk k is culled and replayed, holds no registers and does
  // not pass through the L1 load pipe.
  const int n_w = min(range_end, wl + 1) - range_start;
  const int first = ((n_w - 1) / 32) * 32;
  const uint32_t aS = smem_u32(sh.stage), aBar = smem_u32(&sh.bar);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(aBar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto stage_chunk = [&](int c0) {  // lane 0 only
    if (c0 < 0) return;
    const uint32_t bytes = (uint32_t)min(32, n_w - c0) * (kStreamFloats * 4);
    const float *src = stream + (size_t)(range_start + c0) * kStreamFloats;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(aBar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(aS), "l"(src), "r"(bytes), "r"(aBar) : "memory");
  };
  // The forward's hit map, when it wrote one (pm_tile_render_packed_fwd): bit p of this (warp footprint, tile)'s words
  // = the forward evaluated list entry p for this footprint. Those are exactly the entries with anything to replay
  // here (an entry the forward culled was blended on no pixel of the footprint), so the map replaces the culling
  // below — and a chunk without a set bit is neither copied nor waited for.
  const uint32_t *hw = MAP ? hit_words + (size_t)warp * (size_t)hit_stride + (size_t)(range_start >> 5) + (size_t)ct : nullptr;
  uint32_t word_next = MAP ? __ldg(hw + (first >> 5)) : 0xffffffffu;
  bool staged_next = word_next != 0u;
  if (lane == 0 && staged_next) stage_chunk(first);
  uint32_t phase = 0;
  int queued = 0;  // hits in queue slots [0, queued), rearmost first
  for (int c0 = first; c0 >= -32; c0 -= 32) {  // (one extra pass with no entries flushes the queue)
    const int chunk_start = range_start + c0;
    const uint32_t word = word_next;
    const bool staged = staged_next && c0 >= 0;
    word_next = (MAP && c0 >= 32) ? __ldg(hw + ((c0 - 32) >> 5)) : 0xffffffffu;
    staged_next = c0 >= 32 && word_next != 0u;
    bool hit = false;
    float4 ra4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), rb4 = ra4;
    if (staged) {
      // wait for this chunk's records (all lanes poll the mbarrier's phase bit)
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(aBar), "r"(phase) : "memory");
      }
      phase ^= 1u;
      if (MAP) {
        hit = c0 + lane < n_w && ((word >> lane) & 1u) != 0u;
        if (hit) {
          ra4 = lds128(aS + lane * 48);
          rb4 = lds128(aS + lane * 48 + 16);
        }
      } else {
        // no map (operator path, streams built from a foreign forward state): cull this chunk against the rectangle of
        // the warp's live pixels (`staged` is warp-uniform: every lane takes part in the ballot)
        const unsigned live = __ballot_sync(0xffffffffu, my_last >= chunk_start);
        const LiveRect rect = live_rect(live, wcx, wcy);
        if (c0 + lane < n_w) {
          ra4 = lds128(aS + lane * 48);
          rb4 = lds128(aS + lane * 48 + 16);
          const float2 ext = unpack_extents(ra4.z);
          hit = cull_hit_reg(make_float4(ra4.x, ra4.y, ext.x, ext.y), rb4, rect);  // (+inf extents = no culling)
        }
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (hit) {  // back to front: the highest lane of the chunk is replayed first
      const float4 rc4 = lds128(aS + lane * 48 + 32);
      const int slot = queued + __popc(bal & ~((2u << lane) - 1u));
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(aA + slot * 16), "f"(ra4.x), "f"(ra4.y),
                   "f"(__int_as_float(chunk_start + lane)), "f"(ra4.w) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(aB + slot * 16), "f"(rb4.x), "f"(rb4.y), "f"(rb4.z), "f"(rb4.w) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(aC + slot * 16), "f"(rc4.x), "f"(rc4.y), "f"(rc4.z), "f"(rc4.w) : "memory");
    }
    // the staged chunk is consumed (every lane's reads above have returned): order them before the async proxy's
    // next write into the same bytes, then start the copy of the next chunk — it runs during the replay below
    __syncwarp();
    if (lane == 0 && staged_next) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      stage_chunk(c0 - 32);
    }
    int n_real = kBwdRound;
    queued += __popc(bal);
    if (c0 < 0 && queued > 0) {
      // the last, short round: pad with null records (opacity 0, list position beyond every last id)
      if (lane >= queued && lane < kBwdRound) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(aA + lane * 16), "f"(0.0f), "f"(0.0f), "f"(__int_as_float(0x7fffffff)), "f"(0.0f) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(aB + lane * 16), "f"(0.0f), "f"(0.0f), "f"(0.0f), "f"(0.0f) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(aC + lane * 16), "f"(0.0f), "f"(0.0f), "f"(0.0f), "f"(0.0f) : "memory");
      }
      n_real = queued;
      queued = kBwdRound;
    }
    if (queued >= kBwdRound) {
      __syncwarp();
      int s0 = 0;
      for (; queued - s0 >= kBwdRound; s0 += kBwdRound) round(s0, n_real);
      // move the < 16 leftover hits to the front of the queue
      const int left = queued - s0;
      float4 ta, tb, tc;
      if (lane < left) { ta = lds128(aA + (s0 + lane) * 16); tb = lds128(aB + (s0 + lane) * 16); tc = lds128(aC + (s0 + lane) * 16); }
      __syncwarp();
      if (lane < left) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(aA + lane * 16), "f"(ta.x), "f"(ta.y), "f"(ta.z), "f"(ta.w) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(aB + lane * 16), "f"(tb.x), "f"(tb.y), "f"(tb.z), "f"(tb.w) : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(aC + lane * 16), "f"(tc.x), "f"(tc.y), "f"(tc.z), "f"(tc.w) : "memory");
      }
      queued = left;
    }
  }
}

static int bwd_set_attr() {
  // per device and cheap: set on every call (a process may drive several GPUs)
  if (cudaFuncSetAttribute(composite_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)sizeof(BwdShared)) != cudaSuccess ||
      cudaFuncSetAttribute(composite_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)sizeof(BwdShared)) != cudaSuccess) {
    set_last_cuda_error(cudaGetLastError());
    return PM_ERR_CUDA;
  }
  return PM_OK;
}

// Gather the whole sorted list into the splat stream (one thread per list entry): lets the backward run on a
// forward state that did not come from our forward kernels (tests feed the oracle's), or on a forward that ran
// without a stream.
__global__ void __launch_bounds__(256)
stream_build_kernel(const float *__restrict__ means2d, const float *__restrict__ conics,
                    const float *__restrict__ opacities, const float *__restrict__ colors, int colors_per_camera,
                    int D, const float *__restrict__ depths, const int32_t *__restrict__ flatten_ids, int64_t n_host,
                    const int32_t *__restrict__ d_n, int N, int views_per_scene, float *__restrict__ stream) {
  const int64_t n = effective_n(n_host, d_n);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t f = flatten_ids[i];
  const int cam = f / N;
  const int g_base = cam * N - (cam / views_per_scene) * N;
  GaussA ga; GaussB gb; GaussC gc;
  gather(f, f - g_base, means2d, conics, opacities, colors, colors_per_camera, D, depths, ga, gb, gc);
  stream_store(stream, (size_t)i, f, ga, gb, gc);
}

static int check_common(const float *means2d, const float *conics, const float *opacities,
                        const float *colors, int D, const float *depths, const int32_t *offsets,
                        const int32_t *flatten_ids, int64_t n_isects, int N, int C, int width,
                        int height, int tile_size) {
  if (N < 0 || C < 0 || width <= 0 || height <= 0 || n_isects < 0 || D < 0) return PM_ERR_BAD_ARG;
  if (tile_size != kTile) return PM_ERR_UNSUPPORTED;
  const int CD = D + (depths ? 1 : 0);
  if (CD < 1 || CD > 4) return PM_ERR_UNSUPPORTED;
  if (!means2d || !conics || !opacities || !offsets) return PM_ERR_BAD_ARG;
  if (D > 0 && !colors) return PM_ERR_BAD_ARG;
  if (n_isects > 0 && !flatten_ids) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(means2d) & 7) != 0) return PM_ERR_BAD_ARG;
  if (n_isects >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  return PM_OK;
}

static int bwd_launch(const float *splat_stream, int colors_per_camera, int D, int has_depth, const float *backgrounds,
                      const int32_t *isect_offsets, int64_t n_isects, const int32_t *d_n, int N, int C, int width,
                      int height, int tile_size, const float *alphas, const int32_t *last_ids, const float *transmittance,
                      const float *v_render, const float *v_alphas, float *v_means2d, float *v_conics, float *v_colors,
                      float *v_opacities, float *v_depths, const float *v_images, const float *v_depth_out,
                      const float *v_masks, const float *depth_out, int packed, const int32_t *tile_order,
                      int views_per_scene, uint8_t *touched, const uint32_t *hit_words, int64_t hit_stride,
                      pm_stream_t stream) {
  if (N < 0 || C < 0 || width <= 0 || height <= 0 || n_isects < 0 || D < 0) return PM_ERR_BAD_ARG;
  if (tile_size != kTile) return PM_ERR_UNSUPPORTED;
  const int CD = D + (has_depth ? 1 : 0);
  if (CD < 1 || CD > 4) return PM_ERR_UNSUPPORTED;
  if (!isect_offsets || !alphas || !last_ids || !v_means2d || !v_conics || !v_opacities) return PM_ERR_BAD_ARG;
  if (D > 0 && !v_colors) return PM_ERR_BAD_ARG;
  if ((has_depth != 0) != (v_depths != nullptr)) return PM_ERR_BAD_ARG;
  if (n_isects > 0 && !splat_stream) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(splat_stream) & 15) != 0) return PM_ERR_BAD_ARG;  // cp.async.bulk source alignment
  if (CD == 4 && v_render && (reinterpret_cast<uintptr_t>(v_render) & 15) != 0) return PM_ERR_BAD_ARG;
  if (n_isects >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  if (C == 0 || n_isects == 0) return PM_OK;
  if (views_per_scene > 0 && C % views_per_scene != 0) return PM_ERR_BAD_ARG;
  const int tw = (width + kTile - 1) / kTile, th = (height + kTile - 1) / kTile;
  const int64_t blocks = (int64_t)C * tw * th;
  if (blocks * kBwdParts >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  if (hit_words && hit_stride < pm_hit_words_stride(n_isects, C, width, height, tile_size)) return PM_ERR_WORKSPACE;
  if (bwd_set_attr() != PM_OK) return PM_ERR_CUDA;
  auto kernel = hit_words ? composite_bwd_kernel<true> : composite_bwd_kernel<false>;
  kernel<<<(unsigned)(blocks * kBwdParts), kBwdThreads, sizeof(BwdShared), as_stream(stream)>>>(
      splat_stream, colors_per_camera, D, has_depth ? 1 : 0, backgrounds, isect_offsets, n_isects, d_n, N, width, height,
      tw, th, alphas, last_ids, v_render, v_alphas, v_means2d, v_conics, v_colors, v_opacities, v_depths, v_images,
      v_depth_out, v_masks, depth_out, packed, tile_order, views_per_scene > 0 ? views_per_scene : C, transmittance, touched,
      hit_words, hit_stride);
  return check_launch();
}

}  // namespace pm

extern "C" {

int pm_composite_fwd(const float *means2d, const float *conics, const float *opacities,
                     const float *colors, int colors_per_camera, int D, const float *depths,
                     const float *backgrounds, const int32_t *isect_offsets,
                     const int32_t *flatten_ids, int64_t n_isects, const int32_t *d_n, int N, int C,
                     int width, int height, int tile_size, float *render, float *alphas,
                     int32_t *last_ids, float *transmittance, float *splat_stream, pm_stream_t stream) {
  const int rc = pm::check_common(means2d, conics, opacities, colors, D, depths, isect_offsets, flatten_ids,
                                  n_isects, N, C, width, height, tile_size);
  if (rc != PM_OK) return rc;
  if (!render || !alphas || !last_ids) return PM_ERR_BAD_ARG;
  const int CD = D + (depths ? 1 : 0);
  if (CD == 4 && (reinterpret_cast<uintptr_t>(render) & 15) != 0) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(splat_stream) & 15) != 0) return PM_ERR_BAD_ARG;
  if (C == 0) return PM_OK;
  const int tw = (width + pm::kTile - 1) / pm::kTile, th = (height + pm::kTile - 1) / pm::kTile;
  const int64_t blocks = (int64_t)C * tw * th;
  if (blocks >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  pm::composite_fwd_kernel<<<(unsigned)blocks, 256, 0, pm::as_stream(stream)>>>(
      means2d, conics, opacities, colors, colors_per_camera, D, depths, backgrounds, isect_offsets,
      flatten_ids, n_isects, d_n, N, width, height, tw, th, render, alphas, last_ids, transmittance, splat_stream);
  return pm::check_launch();
}

int pm_splat_stream_build(const float *means2d, const float *conics, const float *opacities, const float *colors,
                          int colors_per_camera, int D, const float *depths, const int32_t *flatten_ids,
                          int64_t n_isects, const int32_t *d_n, int N, int C, int views_per_scene,
                          float *splat_stream, pm_stream_t stream) {
  if (N <= 0 || C <= 0 || n_isects < 0 || D < 0 || D + (depths ? 1 : 0) > 4) return PM_ERR_BAD_ARG;
  if (!means2d || !conics || !opacities || (D > 0 && !colors)) return PM_ERR_BAD_ARG;
  if (n_isects > 0 && (!flatten_ids || !splat_stream)) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(splat_stream) & 15) != 0 || (reinterpret_cast<uintptr_t>(means2d) & 7) != 0) return PM_ERR_BAD_ARG;
  if (n_isects >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  if (views_per_scene > 0 && C % views_per_scene != 0) return PM_ERR_BAD_ARG;
  if (n_isects == 0) return PM_OK;
  pm::stream_build_kernel<<<(unsigned)((n_isects + 255) / 256), 256, 0, pm::as_stream(stream)>>>(
      means2d, conics, opacities, colors, colors_per_camera, D, depths, flatten_ids, n_isects, d_n, N,
      views_per_scene > 0 ? views_per_scene : C, splat_stream);
  return pm::check_launch();
}

int pm_composite_bwd(const float *splat_stream, int colors_per_camera, int D, int has_depth,
                     const float *backgrounds, const int32_t *isect_offsets, int64_t n_isects, const int32_t *d_n,
                     int N, int C, int width, int height, int tile_size, const float *alphas,
                     const int32_t *last_ids, const float *transmittance, const float *v_render, const float *v_alphas,
                     float *v_means2d, float *v_conics, float *v_colors, float *v_opacities,
                     float *v_depths, uint8_t *touched, pm_stream_t stream) {
  return pm::bwd_launch(splat_stream, colors_per_camera, D, has_depth, backgrounds, isect_offsets, n_isects, d_n, N, C,
                        width, height, tile_size, alphas, last_ids, transmittance, v_render, v_alphas, v_means2d, v_conics,
                        v_colors, v_opacities, v_depths, nullptr, nullptr, nullptr, nullptr, 0, nullptr, 0, touched, nullptr, 0,
                        stream);
}

int pm_composite_bwd_packed(const float *splat_stream, int colors_per_camera, const float *backgrounds,
                            const int32_t *isect_offsets, int64_t n_isects, const int32_t *d_n, int N, int C,
                            int width, int height, int tile_size, const float *masks, const float *depth_out,
                            const int32_t *last_ids, const float *transmittance, const float *v_images,
                            const float *v_depth_out, const float *v_masks, float *v_means2d, float *v_conics,
                            float *v_colors, float *v_opacities, float *v_depths, const int32_t *tile_order,
                            int views_per_scene, uint8_t *touched, const uint32_t *hit_words, int64_t hit_stride,
                            pm_stream_t stream) {
  if (!masks || !depth_out || !v_colors || !v_depths) return PM_ERR_BAD_ARG;
  return pm::bwd_launch(splat_stream, colors_per_camera, 3, 1, backgrounds, isect_offsets, n_isects, d_n, N, C, width,
                        height, tile_size, masks, last_ids, transmittance, nullptr, nullptr, v_means2d, v_conics, v_colors,
                        v_opacities, v_depths, v_images, v_depth_out, v_masks, depth_out, 1, tile_order, views_per_scene,
                        touched, hit_words, hit_stride, stream);
}

}  // extern "C"
