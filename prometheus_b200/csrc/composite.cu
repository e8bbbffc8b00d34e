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
// loaded. The backward pass replays the same culled lists back to front, reduces
// each Gaussian's gradient across the warp with a 12-shuffle halving butterfly and
// issues one fire-and-forget global reduction per (warp, Gaussian, component) —
// no per-pixel global atomics. The warps of the backward are AUTONOMOUS: each stages
// its own 32-Gaussian chunks (lane j gathers Gaussian j; the other warps of the tile
// hit the same lines in L1/L2) and walks only the prefix its own pixels need, so there
// is no CTA barrier at all — the per-round barrier of the CTA-wide version was 36 % of
// its stall samples (profiles/r1j_composite_bwd_lines.txt). (Accumulating the CTA's warps in shared memory first,
// PM_BWD_GLOBAL_RED=0, measured 3.5 % slower: shared atomics return a value and the
// flush needs a barrier.)
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
                     int32_t *__restrict__ last_ids, float *__restrict__ transmittance) {
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
    if (idx < range_end) { const int32_t f = flatten_ids[idx]; gather(f, f - cam * N, means2d, conics, opacities, colors, colors_per_camera, D, depths, stg.sA[tid], stg.sB[tid], stg.sC[tid]); }
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
// Warp reduction of 10 gradient components with a halving butterfly: instead of reducing each
// component over all 32 lanes (10 x 5 = 50 shuffles), every step pairs lanes and SPLITS the
// component set between the two partners (10 -> 5 -> 3 -> 2 -> 1 -> 1), so 5+3+2+1+1 = 12 shuffles
// leave the full sum of component `comp_of_lane()` on the even lanes that own one.
__device__ __forceinline__ int comp_of_lane(int lane) {
  if (lane & 1) return -1;                      // the odd partner holds a copy
  const int b8 = (lane >> 3) & 1, b4 = (lane >> 2) & 1, b2 = (lane >> 1) & 1;
  int sub;
  if (!b8) sub = b4 ? (b2 ? -1 : 2) : b2;       // {0,1,2}
  else sub = b4 ? -1 : 3 + b2;                  // {3,4}
  return sub < 0 ? -1 : sub + ((lane & 16) ? 5 : 0);
}

__device__ __forceinline__ float butterfly_reduce10(const float (&v)[10], int lane) {
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
  float a[5], b[3], c[2];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const float send = h16 ? v[i] : v[i + 5];
    const float keep = h16 ? v[i + 5] : v[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float hi = i < 2 ? a[i + 3] : 0.0f;
    const float send = h8 ? a[i] : hi;
    const float keep = h8 ? hi : a[i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float hi = i < 1 ? b[2] : 0.0f;
    const float send = h4 ? b[i] : hi;
    const float keep = h4 ? hi : b[i];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = h2 ? c[0] : c[1];
  const float keep = h2 ? c[1] : c[0];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}

#ifndef PM_BWD_MINBLOCKS
#define PM_BWD_MINBLOCKS 10
#endif
#ifndef PM_BWD_WARPS
#define PM_BWD_WARPS 4
#endif
// A CTA holds kBwdWarps of the tile's eight 8x4-pixel warp footprints (4 = its upper or lower half). The warps of
// a CTA share nothing but the launch: each has its own staging slots and there is no barrier.
constexpr int kBwdWarps = PM_BWD_WARPS;
constexpr int kBwdThreads = 32 * kBwdWarps;
constexpr int kBwdParts = 8 / kBwdWarps;
constexpr int kBwdStage = 32 * kBwdWarps;  // one private 32-Gaussian chunk per warp

struct BwdShared {
  GaussA sA[kBwdStage];
  GaussB sB[kBwdStage];
  GaussC sC[kBwdStage];
};

__global__ void __launch_bounds__(kBwdThreads, PM_BWD_MINBLOCKS)
composite_bwd_kernel(const float *__restrict__ means2d, const float *__restrict__ conics,
                     const float *__restrict__ opacities, const float *__restrict__ colors,
                     int colors_per_camera, int D, const float *__restrict__ depths,
                     const float *__restrict__ backgrounds, const int32_t *__restrict__ offsets,
                     const int32_t *__restrict__ flatten_ids, int64_t n_host,
                     const int32_t *__restrict__ d_n, int N, int width, int height, int tile_width,
                     int tile_height, const float *__restrict__ alphas,
                     const int32_t *__restrict__ last_ids, const float *__restrict__ v_render,
                     const float *__restrict__ v_alphas, float *__restrict__ v_means2d,
                     float *__restrict__ v_conics, float *__restrict__ v_colors,
                     float *__restrict__ v_opacities, float *__restrict__ v_depths,
                     const float *__restrict__ v_images, const float *__restrict__ v_depth_out,
                     const float *__restrict__ v_masks, const float *__restrict__ depth_out, int packed,
                     const int32_t *__restrict__ tile_order, int views_per_scene,
                     const float *__restrict__ transmittance) {
  extern __shared__ __align__(16) unsigned char bwd_smem[];
  BwdShared &sh = *reinterpret_cast<BwdShared *>(bwd_smem);
  GaussA *sA = sh.sA; GaussB *sB = sh.sB; GaussC *sC = sh.sC;

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
  const int CD = D + (depths ? 1 : 0);
  // flatten id f = cam * N + g; per-Gaussian arrays hold one block of N per scene (scene = cam / views_per_scene)
  const int g_base = cam * N - (cam / views_per_scene) * N;

  const int64_t n_isects = effective_n(n_host, d_n);
  const int range_start = offsets[ct];
  const int range_end = (ct == n_ct - 1) ? (int)n_isects : offsets[ct + 1];
  if (range_end <= range_start) return;

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
  const int my_comp = comp_of_lane(lane);
  // where this lane's component goes: c0..c3 (colour, depth), q0..q2, mx, my, opacity
  float *red_base = nullptr;
  int red_mult = 0;
  bool red_by_pair = true;
  if (my_comp >= 0) {
    if (my_comp < 4) {
      if (my_comp < D) { red_base = v_colors + my_comp; red_mult = D; red_by_pair = colors_per_camera != 0; }
      else if (my_comp == D && depths) { red_base = v_depths; red_mult = 1; }
    } else if (my_comp < 7) { red_base = v_conics + (my_comp - 4); red_mult = 3; }
    else if (my_comp < 9) { red_base = v_means2d + (my_comp - 7); red_mult = 2; }
    else { red_base = v_opacities; red_mult = 1; red_by_pair = false; }
  }
  const int my_comp_live = red_base ? my_comp : -1;

  // Warp-autonomous replay: no CTA barrier anywhere. Each warp walks the prefix of the tile's list ITS pixels
  // need, back to front, 32 Gaussians at a time: lane j gathers Gaussian j into the warp's private staging
  // slots (the gathers of the warps of a tile hit L1/L2), then the warp culls and replays the chunk.
  int wl = my_last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wl = max(wl, __shfl_xor_sync(0xffffffffu, wl, o));
  if (wl < range_start) return;
  const int w = tid >> 5;
  GaussA *wA = sA + w * 32; GaussB *wB = sB + w * 32; GaussC *wC = sC + w * 32;
  const uint32_t qA = smem_u32(wA), qB = smem_u32(wB), qC = smem_u32(wC);
  const int n_w = min(range_end, wl + 1) - range_start;
  float T = T_final;
  float buf0 = 0.0f, buf1 = 0.0f, buf2 = 0.0f, buf3 = 0.0f;
  for (int c0 = ((n_w - 1) / 32) * 32; c0 >= 0; c0 -= 32) {
    const int cnt = min(32, n_w - c0);
    const int chunk_start = range_start + c0;
    int32_t f = 0;
    __syncwarp();
    if (lane < cnt) {
      f = flatten_ids[chunk_start + lane];
      gather(f, f - g_base, means2d, conics, opacities, colors, colors_per_camera, D, depths, wA[lane], wB[lane], wC[lane]);
    }
    __syncwarp();
    const unsigned live = __ballot_sync(0xffffffffu, my_last >= chunk_start);
    const LiveRect rect = live_rect(live, wcx, wcy);
    bool hit = false;
    if (lane < cnt) hit = cull_hit(lds128(qA + lane * 16), qB + lane * 16, rect);
    unsigned mask = __ballot_sync(0xffffffffu, hit);
    while (mask) {
      const int t = 31 - __clz(mask);
      mask &= ~(1u << t);
      const float2 m = lds64(qA + t * 16);
      const float4 b = lds128(qB + t * 16);
      const float dx = m.x - px, dy = m.y - py;
      const float p = power2(b.x, b.y, b.z, dx, dy);
      const float vis = ex2_approx(p);
      const float alpha = fminf(0.999f, b.w * vis);
      const bool valid = (chunk_start + t <= my_last) && !(p > 0.0f) && alpha >= kAlphaMin;
      if (!__any_sync(0xffffffffu, valid)) continue;
      float gv[10] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
      if (valid) {
        const float4 c = lds128(qC + t * 16);
        const float ra = rcp_approx(1.0f - alpha);
        T *= ra;
        const float fac = alpha * T;
        gv[0] = fac * vr[0]; gv[1] = fac * vr[1]; gv[2] = fac * vr[2]; gv[3] = fac * vr[3];
        float v_alpha = (c.x * T - buf0 * ra) * vr[0] + (c.y * T - buf1 * ra) * vr[1] +
                        (c.z * T - buf2 * ra) * vr[2] + (c.w * T - buf3 * ra) * vr[3];
        v_alpha += va_seed * ra;
        if (b.w * vis <= 0.999f) {
          const float v_sigma = -b.w * vis * v_alpha;
          gv[4] = 0.5f * v_sigma * dx * dx;
          gv[5] = v_sigma * dx * dy;
          gv[6] = 0.5f * v_sigma * dy * dy;
          const float s2 = v_sigma * (-1.0f / kLog2e);
          gv[7] = s2 * (2.0f * b.x * dx + b.y * dy);
          gv[8] = s2 * (b.y * dx + 2.0f * b.z * dy);
          gv[9] = vis * v_alpha;
        }
        buf0 += c.x * fac; buf1 += c.y * fac; buf2 += c.z * fac; buf3 += c.w * fac;
      }
      const float total = butterfly_reduce10(gv, lane);
      const int32_t ft = __shfl_sync(0xffffffffu, f, t);
      if (my_comp_live >= 0 && total != 0.0f)
        atomicAdd(red_base + (size_t)(red_by_pair ? ft : ft - g_base) * red_mult, total);
    }
  }
}

static int bwd_set_attr() {
  // per device and cheap: set on every call (a process may drive several GPUs)
  if (cudaFuncSetAttribute(composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)sizeof(BwdShared)) != cudaSuccess) {
    set_last_cuda_error(cudaGetLastError());
    return PM_ERR_CUDA;
  }
  return PM_OK;
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

}  // namespace pm

extern "C" {

int pm_composite_fwd(const float *means2d, const float *conics, const float *opacities,
                     const float *colors, int colors_per_camera, int D, const float *depths,
                     const float *backgrounds, const int32_t *isect_offsets,
                     const int32_t *flatten_ids, int64_t n_isects, const int32_t *d_n, int N, int C,
                     int width, int height, int tile_size, float *render, float *alphas,
                     int32_t *last_ids, float *transmittance, pm_stream_t stream) {
  const int rc = pm::check_common(means2d, conics, opacities, colors, D, depths, isect_offsets, flatten_ids,
                                  n_isects, N, C, width, height, tile_size);
  if (rc != PM_OK) return rc;
  if (!render || !alphas || !last_ids) return PM_ERR_BAD_ARG;
  const int CD = D + (depths ? 1 : 0);
  if (CD == 4 && (reinterpret_cast<uintptr_t>(render) & 15) != 0) return PM_ERR_BAD_ARG;
  if (C == 0) return PM_OK;
  const int tw = (width + pm::kTile - 1) / pm::kTile, th = (height + pm::kTile - 1) / pm::kTile;
  const int64_t blocks = (int64_t)C * tw * th;
  if (blocks >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  pm::composite_fwd_kernel<<<(unsigned)blocks, 256, 0, pm::as_stream(stream)>>>(
      means2d, conics, opacities, colors, colors_per_camera, D, depths, backgrounds, isect_offsets,
      flatten_ids, n_isects, d_n, N, width, height, tw, th, render, alphas, last_ids, transmittance);
  return pm::check_launch();
}

int pm_composite_bwd(const float *means2d, const float *conics, const float *opacities,
                     const float *colors, int colors_per_camera, int D, const float *depths,
                     const float *backgrounds, const int32_t *isect_offsets,
                     const int32_t *flatten_ids, int64_t n_isects, const int32_t *d_n, int N, int C,
                     int width, int height, int tile_size, const float *alphas,
                     const int32_t *last_ids, const float *transmittance, const float *v_render, const float *v_alphas,
                     float *v_means2d, float *v_conics, float *v_colors, float *v_opacities,
                     float *v_depths, pm_stream_t stream) {
  const int rc = pm::check_common(means2d, conics, opacities, colors, D, depths, isect_offsets, flatten_ids,
                                  n_isects, N, C, width, height, tile_size);
  if (rc != PM_OK) return rc;
  if (!alphas || !last_ids || !v_means2d || !v_conics || !v_opacities) return PM_ERR_BAD_ARG;
  if (D > 0 && !v_colors) return PM_ERR_BAD_ARG;
  if ((depths != nullptr) != (v_depths != nullptr)) return PM_ERR_BAD_ARG;
  const int CD = D + (depths ? 1 : 0);
  if (CD == 4 && v_render && (reinterpret_cast<uintptr_t>(v_render) & 15) != 0) return PM_ERR_BAD_ARG;
  if (C == 0 || n_isects == 0) return PM_OK;
  const int tw = (width + pm::kTile - 1) / pm::kTile, th = (height + pm::kTile - 1) / pm::kTile;
  const int64_t blocks = (int64_t)C * tw * th;
  if (blocks >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  if (pm::bwd_set_attr() != PM_OK) return PM_ERR_CUDA;
  pm::composite_bwd_kernel<<<(unsigned)(blocks * pm::kBwdParts), pm::kBwdThreads, sizeof(pm::BwdShared), pm::as_stream(stream)>>>(
      means2d, conics, opacities, colors, colors_per_camera, D, depths, backgrounds, isect_offsets,
      flatten_ids, n_isects, d_n, N, width, height, tw, th, alphas, last_ids, v_render, v_alphas,
      v_means2d, v_conics, v_colors, v_opacities, v_depths, nullptr, nullptr, nullptr, nullptr, 0, nullptr, C, transmittance);
  return pm::check_launch();
}

int pm_composite_bwd_packed(const float *means2d, const float *conics, const float *opacities,
                            const float *colors, int colors_per_camera, const float *depths,
                            const float *backgrounds, const int32_t *isect_offsets,
                            const int32_t *flatten_ids, int64_t n_isects, const int32_t *d_n, int N, int C,
                            int width, int height, int tile_size, const float *masks, const float *depth_out,
                            const int32_t *last_ids, const float *transmittance, const float *v_images, const float *v_depth_out,
                            const float *v_masks, float *v_means2d, float *v_conics, float *v_colors,
                            float *v_opacities, float *v_depths, const int32_t *tile_order, int views_per_scene,
                            pm_stream_t stream) {
  const int rc = pm::check_common(means2d, conics, opacities, colors, 3, depths, isect_offsets, flatten_ids,
                                  n_isects, N, C, width, height, tile_size);
  if (rc != PM_OK) return rc;
  if (!depths || !masks || !depth_out || !last_ids || !v_means2d || !v_conics || !v_opacities || !v_colors || !v_depths)
    return PM_ERR_BAD_ARG;
  if (C == 0 || n_isects == 0) return PM_OK;
  const int tw = (width + pm::kTile - 1) / pm::kTile, th = (height + pm::kTile - 1) / pm::kTile;
  const int64_t blocks = (int64_t)C * tw * th;
  if (blocks >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  if (pm::bwd_set_attr() != PM_OK) return PM_ERR_CUDA;
  pm::composite_bwd_kernel<<<(unsigned)(blocks * pm::kBwdParts), pm::kBwdThreads, sizeof(pm::BwdShared), pm::as_stream(stream)>>>(
      means2d, conics, opacities, colors, colors_per_camera, 3, depths, backgrounds, isect_offsets,
      flatten_ids, n_isects, d_n, N, width, height, tw, th, masks, last_ids, nullptr, nullptr,
      v_means2d, v_conics, v_colors, v_opacities, v_depths, v_images, v_depth_out, v_masks, depth_out, 1, tile_order,
      views_per_scene > 0 ? views_per_scene : C, transmittance);
  return pm::check_launch();
}

}  // extern "C"
