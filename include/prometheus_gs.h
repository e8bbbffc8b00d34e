/*
 * prometheus_gs.h — C-ABI of the B200-native GS-decoder rendering hot path.
 *
 * This is the drop-in boundary: a plain C interface (device pointers, sizes, a
 * CUDA stream) that a host binding — the ctypes loader in prometheus_b200/_lib.py,
 * or the stub shown in INTEGRATION.md — calls in place of the kernels the reference
 * reaches through `gsplat.rasterization(...)`
 * (reference call site: prometheus/modules/renderers/gaussians_renderer.py:413-427)
 * and through the ~40 ATen ops of `GaussianConverter.forward` (same file :210-287).
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer on the current CUDA device unless a
 *     parameter is documented as "host". Tensors are row-major and contiguous.
 *   - fp32 = float, ids = int32_t, sort keys = int64_t.
 *   - The caller owns every buffer (torch's caching allocator in our binding).
 *     Functions never allocate, never synchronise the device and never block the
 *     host; they enqueue work on `stream` (a cudaStream_t passed as void*).
 *   - Return value: PM_OK (0) or a negative PM_ERR_* code. Argument errors are
 *     detected on the host before anything is enqueued.
 *   - "P" below is the number of (camera, Gaussian) pairs, P = C * N, flattened
 *     camera-major: pair index f = c * N + g  ("flatten id").
 *
 * There is no CPU fallback behind this interface: if the library was not built
 * for sm_100a or no GPU is present, the calls fail with PM_ERR_CUDA.
 */
#ifndef PROMETHEUS_GS_H
#define PROMETHEUS_GS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: exact final transmittance kept for the backward; 3: splat stream (TMA-replayed backward), touched-pair map;
 * 4: pm_project_scatter_fwd (single-pass binning), record_stride of pm_tile_render_packed_fwd;
 * 5: pm_project_scatter_fwd(opacities): bins only the tiles a splat can be seen in; histogram-only form;
 * 6: hit map from pm_tile_render_packed_fwd to pm_composite_bwd_packed (pm_hit_words_stride); pm_bin_scatter(conics,
 *    opacities): the usable-tile rule for the two-pass binning as well */
#define PM_ABI_VERSION 6

#if defined(__GNUC__)
#define PM_API __attribute__((visibility("default")))
#else
#define PM_API
#endif

typedef void *pm_stream_t; /* cudaStream_t */

enum pm_status {
  PM_OK = 0,
  PM_ERR_BAD_ARG = -1,     /* null pointer, negative size, misaligned buffer      */
  PM_ERR_UNSUPPORTED = -2, /* e.g. tile_size != 16, more than 4 colour channels   */
  PM_ERR_WORKSPACE = -3,   /* workspace too small for the requested problem size  */
  PM_ERR_CUDA = -4,        /* a CUDA runtime call / launch failed                 */
  PM_ERR_RANGE = -5        /* size exceeds an implementation limit (2^30 isects)  */
};

/* GaussianConverter.gs_convert_mode (gaussians_renderer.py:175, :258-285) */
enum pm_convert_mode {
  PM_MODE_MVSPLAT = 0,
  PM_MODE_GSLRM = 1,
  PM_MODE_DIRECTOR3D = 2
};

/* storage type of the decoder maps handed to the unproject stage (the *_dt entry points) */
enum pm_dtype {
  PM_DTYPE_F32 = 0,
  PM_DTYPE_F16 = 1,
  PM_DTYPE_BF16 = 2
};

PM_API int pm_abi_version(void);
PM_API const char *pm_status_string(int status);
/* Last CUDA error string recorded by a failing call on this host thread. */
PM_API const char *pm_last_cuda_error(void);

/* ------------------------------------------------------------------------- *
 * Stage 1 — unproject: per-pixel 14-channel decoder maps -> 3D Gaussians.
 * Replaces GaussianConverter.forward (gaussians_renderer.py:210-287).
 *
 *   params   [V, 14, h, w]  channel-planar (rgb3, uv-offset2, depth1, opacity1,
 *                           scale3, quat4), V = B*N views
 *   cameras  [V, 18]        c2w 3x4 row-major (OpenGL) + fx, fy, cx, cy, H, W
 *   muls     [14]           per-channel multipliers (`converter.muls` buffer)
 * outputs, G = V*h*w Gaussians in (view, row, col) order:
 *   xyz [G,3]  rgb [G,3]  opacity [G]  scales [G,3]  quats [G,4] (wxyz)
 * ------------------------------------------------------------------------- */
PM_API int pm_unproject_fwd(const float *params, const float *cameras, const float *muls,
                     int mode, float z_near, float z_far, float s_min, float s_max,
                     int n_views, int h, int w,
                     float *xyz, float *rgb, float *opacity, float *scales, float *quats,
                     pm_stream_t stream);

/* pm_unproject_fwd for decoder maps stored as fp32 / fp16 / bf16 (params_dtype = enum pm_dtype): the reference
 * promotes half-precision maps to fp32 by its `* self.muls` (gaussians_renderer.py:219), so the results are
 * those of pm_unproject_fwd on the widened maps, bit for bit, without the widening pass. */
PM_API int pm_unproject_fwd_dt(const void *params, int params_dtype, const float *cameras, const float *muls,
                     int mode, float z_near, float z_far, float s_min, float s_max,
                     int n_views, int h, int w,
                     float *xyz, float *rgb, float *opacity, float *scales, float *quats,
                     pm_stream_t stream);

/* pm_unproject_fwd_dt with n_dst (1..8) destination sets: dst_ptrs is a HOST array of 5 * n_dst device pointers
 * {xyz, rgb, opacity, scales, quats} per destination. With peer-mapped pointers (NVLink P2P / symmetric memory)
 * this is the fused "unproject + all-gather" of the view-sharded decode: every rank stores its Gaussians straight
 * into its slot of every peer's gather buffer; the caller brackets it with cross-rank barriers. */
PM_API int pm_unproject_fwd_multi(const void *params, int params_dtype, const float *cameras, const float *muls,
                     int mode, float z_near, float z_far, float s_min, float s_max,
                     int n_views, int h, int w, int n_dst, float *const *dst_ptrs, pm_stream_t stream);

/* VJP of pm_unproject_fwd w.r.t. params (cameras carry no gradient in the
 * reference: gs_decoder_system.py:200 feeds dataset cameras). Any v_* input may
 * be NULL (treated as zero). v_params [V,14,h,w] is fully overwritten. */
PM_API int pm_unproject_bwd(const float *params, const float *cameras, const float *muls,
                     int mode, float z_near, float z_far, float s_min, float s_max,
                     int n_views, int h, int w,
                     const float *v_xyz, const float *v_rgb, const float *v_opacity,
                     const float *v_scales, const float *v_quats,
                     float *v_params, pm_stream_t stream);

/* pm_unproject_bwd for half-precision maps: v_params has the maps' storage type and receives the fp32 gradient
 * rounded to nearest-even (what autograd's cast back through the reference's promotion produces). */
PM_API int pm_unproject_bwd_dt(const void *params, int params_dtype, const float *cameras, const float *muls,
                     int mode, float z_near, float z_far, float s_min, float s_max,
                     int n_views, int h, int w,
                     const float *v_xyz, const float *v_rgb, const float *v_opacity,
                     const float *v_scales, const float *v_quats, void *v_params,
                     pm_stream_t stream);

/* Camera preparation on the device. Replaces the host round trip of
 * convert_camera_parameters_into_viewpoint_cameras + MiniCam + the K build
 * (gaussians_renderer.py:453-472, :139-169, :404-412): rescale intrinsics to the
 * render size, flip OpenGL->OpenCV, invert c2w, centre the principal point.
 *   cameras [n,18] -> viewmats [n,4,4] (world-to-camera), Ks [n,3,3] */
PM_API int pm_camera_prep(const float *cameras, int n, int h, int w,
                   float *viewmats, float *Ks, pm_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Stage 2 — EWA projection (+ fused tile-count). Replaces gsplat 1.4.0
 * fully_fused_projection fwd/bwd and pass 1 of isect_tiles.
 *   means [N,3] quats [N,4] scales [N,3] viewmats [C,4,4] Ks [C,3,3]
 * outputs (all [C,N,...]): radii i32, means2d [.,2], depths, conics [.,3],
 *   tiles_per_gauss i32 (may be NULL to skip the count).
 * Culled pairs get radius 0 and zeros in the other outputs.
 * ------------------------------------------------------------------------- */
PM_API int pm_project_fwd(const float *means, const float *quats, const float *scales,
                   const float *viewmats, const float *Ks,
                   int N, int C, int width, int height, int tile_size,
                   float eps2d, float near_plane, float far_plane, float radius_clip,
                   int32_t *radii, float *means2d, float *depths, float *conics,
                   int32_t *tiles_per_gauss, pm_stream_t stream);

/* pm_project_fwd + pm_bin_count in one pass: additionally ACCUMULATES the per-(camera, tile)
 * intersection histogram into tile_counts [C, tile_height, tile_width] int32 (caller zero-fills)
 * while each pair's tile AABB is still in registers. Same outputs, bit for bit, as the two calls.
 * SCENE BATCHING (this and the other fast-path entry points): views_per_scene = V > 0 with C = S * V renders
 * S independent scenes in one launch set — means / quats / scales (and opacities / colours downstream) hold S
 * blocks of N Gaussians, camera c sees only block c / V; pair arrays stay [C, N]. Needs N % 256 == 0 when
 * S > 1. views_per_scene = 0 means one scene (V = C). */
PM_API int pm_project_count_fwd(const float *means, const float *quats, const float *scales,
                   const float *viewmats, const float *Ks,
                   int N, int C, int views_per_scene, int width, int height, int tile_size,
                   float eps2d, float near_plane, float far_plane, float radius_clip,
                   int32_t *radii, float *means2d, float *depths, float *conics,
                   int32_t *tiles_per_gauss, int32_t *tile_counts, pm_stream_t stream);

/* Projection + SINGLE-PASS binning (module fast path; replaces pm_project_count_fwd + pm_bin_scatter, i.e. both
 * passes of gsplat's isect_tiles, for callers that do not need the compact record array): besides the pair arrays
 * and the histogram tile_counts [C, tile_height, tile_width] (caller zero-fills) every intersection record
 * {depth bits, flatten id} is written straight into its tile's fixed-capacity region,
 * records[(c * tiles + t) * tile_cap + rank], rank = arrival order < tile_cap. A tile that receives more than
 * tile_cap intersections only keeps the first tile_cap: the caller compares pm_tile_scan's max count with tile_cap
 * afterwards and, if it is larger, renders again (larger capacity, or pm_project_count_fwd + pm_bin_scatter).
 * pm_tile_render_packed_fwd reads this layout with record_stride = tile_cap.
 * records == NULL with tile_cap == 0: histogram only (sizes the regions on the first call of a shape).
 * opacities [S * N] (may be NULL): bin only the tiles a splat can be SEEN in — those of gsplat's AABB whose pixel centres
 * meet the per-axis extent of {alpha >= 1/255} (conservative: the compositor skips the splat on every pixel of a dropped
 * tile, so images and gradients are unchanged bit for bit; the histogram and the records are a SUBSET of gsplat's
 * intersections, ~19 % fewer on the benchmark scenes). NULL: gsplat's AABB exactly. A count-only call and the
 * scatter call that follows it must agree on this argument. */
PM_API int pm_project_scatter_fwd(const float *means, const float *quats, const float *scales,
                   const float *viewmats, const float *Ks,
                   int N, int C, int views_per_scene, int width, int height, int tile_size,
                   float eps2d, float near_plane, float far_plane, float radius_clip,
                   int32_t *radii, float *means2d, float *depths, float *conics,
                   int32_t *tile_counts, uint64_t *records, int tile_cap, const float *opacities,
                   pm_stream_t stream);

/* VJP of the projection. v_means2d [C,N,2], v_depths [C,N], v_conics [C,N,3]
 * (any may be NULL) -> v_means [N,3], v_quats [N,4], v_scales [N,3], summed over
 * cameras, fully overwritten. viewmats/Ks carry no gradient on this path.
 * touched [C,N] uint8 (may be NULL): the array pm_composite_bwd* marked — nonzero for every pair whose cotangents
 * may be nonzero; pairs with a zero byte are skipped unread (most pairs of an opaque scene). */
PM_API int pm_project_bwd(const float *means, const float *quats, const float *scales,
                   const float *viewmats, const float *Ks,
                   int N, int C, int views_per_scene, int width, int height, float eps2d,
                   const int32_t *radii, const float *conics,
                   const float *v_means2d, const float *v_depths, const float *v_conics,
                   float *v_means, float *v_quats, float *v_scales, const uint8_t *touched,
                   pm_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Stage 3 — tile binning and depth sort. Replaces gsplat isect_tiles (cumsum +
 * emit), cub::DeviceRadixSort::SortPairs and isect_offset_encode.
 * ------------------------------------------------------------------------- */

/* Inclusive prefix sum of tiles_per_gauss [P] -> cum [P] (int32); the total is
 * also written to *n_isects_out (device int32). */
PM_API size_t pm_scan_workspace_bytes(int64_t P);
PM_API int pm_scan_tiles(const int32_t *tiles_per_gauss, int64_t P, int32_t *cum,
                  int32_t *n_isects_out, void *workspace, size_t workspace_bytes,
                  pm_stream_t stream);

/* Emit one (key, value) per (pair, tile): key = cam << (32+tile_bits) |
 * tile << 32 | depth_bits, value = flatten id. Output slot of pair f's k-th
 * tile (row-major over its tile AABB) is cum[f-1] + k. */
PM_API int pm_bin_emit(const float *means2d, const int32_t *radii, const float *depths,
                const int32_t *cum, int N, int C, int tile_size,
                int tile_width, int tile_height,
                int64_t *isect_ids, int32_t *flatten_ids, int64_t capacity,
                pm_stream_t stream);

/* Stable LSD radix sort (hand-written onesweep, no CUB) of n (key,value) pairs
 * on key bits [0, end_bit). Double-buffered: input is buffer 0; *result_buffer
 * (host) receives 0 or 1 = which buffer holds the sorted output (decided on the
 * host from the pass count, before launch). If d_n is non-NULL the element
 * count is read on the device from *d_n (must be <= n) so no host sync is
 * needed; n then only sizes the launch. */
PM_API size_t pm_sort_workspace_bytes(int64_t n);
PM_API int pm_sort_pairs(int64_t *keys0, int32_t *vals0, int64_t *keys1, int32_t *vals1,
                  int64_t n, const int32_t *d_n, int end_bit,
                  void *workspace, size_t workspace_bytes, int *result_buffer,
                  pm_stream_t stream);

/* ---- binned ("tile-bucketed") sort: the same sorted output as pm_bin_emit +
 * pm_sort_pairs + pm_tile_offsets, with ~6x less HBM traffic. The key's high bits
 * (camera, tile) are known when an intersection is emitted, so intersections are
 * scattered straight into their tile's segment (offsets = exclusive scan of the
 * per-tile histogram of pm_bin_count) and each segment is depth-sorted
 * inside shared memory by one CTA. Ties on depth bits are ordered by flatten id,
 * which is exactly the order a stable sort of the emission order produces. */

/* Per-(camera,tile) intersection histogram, ACCUMULATED into tile_counts
 * [C, tile_height, tile_width] int32 (caller zero-fills). */
PM_API int pm_bin_count(const float *means2d, const int32_t *radii, int N, int C, int tile_size,
                        int tile_width, int tile_height, int32_t *tile_counts, pm_stream_t stream);

/* Exclusive scan of tile_counts [T] -> isect_offsets [T]; stats_out (device, 2 x
 * int32) receives {n_isects, max tile count}. tile_order [T] (may be NULL) receives the
 * tile indices heaviest first — the launch order of pm_tile_render_*_fwd. */
PM_API int pm_tile_scan(const int32_t *tile_counts, int64_t T, int32_t *isect_offsets,
                        int32_t *stats_out, int32_t *tile_order, pm_stream_t stream);

/* Scatter every (pair, tile) intersection into its tile segment as a packed
 * 8-byte record {depth bits (low 32), flatten id (high 32)}. cursors [T] must be
 * zero-filled by the caller. Order inside a segment is unspecified.
 * conics [C,N,3] + opacities [S * N] (both or neither; views_per_scene as in pm_project_count_fwd): scatter only the
 * tiles a splat can be seen in — the rule of pm_project_scatter_fwd(opacities), whose histogram-only form
 * (records = NULL) then supplies the counts behind isect_offsets. NULL: gsplat's AABB (the counts of
 * pm_project_count_fwd / pm_bin_count). */
PM_API int pm_bin_scatter(const float *means2d, const int32_t *radii, const float *depths,
                          const int32_t *isect_offsets, int32_t *cursors, int N, int C,
                          int tile_size, int tile_width, int tile_height,
                          uint64_t *records, int64_t capacity, const float *conics, const float *opacities,
                          int views_per_scene, pm_stream_t stream);

/* Segment size pm_tile_sort sorts in one shared-memory pass (entries). Larger
 * segments are first split by depth through `scratch` (see pm_tile_sort). */
PM_API int pm_tile_sort_capacity(void);

/* Sort each tile segment by (depth bits, flatten id) and expand to the gsplat
 * layout: isect_ids [n] int64, flatten_ids [n] int32. max_tile_count (host) is
 * the largest segment as reported by pm_tile_scan, or -1 if unknown (no host
 * sync). Segments above pm_tile_sort_capacity() need `scratch` (n x 8 bytes,
 * may be NULL otherwise): they are split into depth-ordered sub-buckets that
 * are sorted group by group. Only a sub-bucket that is itself larger than the
 * capacity (thousands of near-identical depths in one tile) cannot be handled:
 * *overflow_flag (device int32, caller zero-fills) is then set and the caller
 * must use pm_sort_pairs instead. */
PM_API int pm_tile_sort(const uint64_t *records, uint64_t *scratch, const int32_t *isect_offsets,
                        int64_t n, const int32_t *d_n, int C, int tile_width, int tile_height,
                        int max_tile_count, int64_t *isect_ids, int32_t *flatten_ids,
                        int32_t *overflow_flag, pm_stream_t stream);

/* Per-(camera,tile) start offsets into the sorted list: isect_offsets
 * [C, tile_height, tile_width] int32 (lower bound of each tile id). */
PM_API int pm_tile_offsets(const int64_t *sorted_isect_ids, int64_t n, const int32_t *d_n,
                    int C, int tile_width, int tile_height,
                    int32_t *isect_offsets, pm_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Stage 4 — front-to-back alpha compositing, forward and backward. Replaces
 * gsplat rasterize_to_pixels fwd/bwd (packed=False, tile_size 16).
 *
 *   means2d [C,N,2]  conics [C,N,3]
 *   opacities [N]                      (shared by all cameras)
 *   colors  [N,D] (colors_per_camera=0) or [C,N,D] (=1), 0 <= D <= 3 .. 4
 *   depths  [C,N] or NULL: appended as one more channel (render_mode *+D/ED)
 *           channels CD = D + (depths != NULL), 1 <= CD <= 4
 *   backgrounds [C,D] or NULL (the depth channel's background is 0)
 * outputs: render [C,H,W,CD], alphas [C,H,W], last_ids [C,H,W] int32 (index into
 *   the sorted list of the last Gaussian composited into the pixel).
 *   transmittance [C,H,W] (may be NULL): the exact fp32 T = prod(1 - alpha_i) the pixel ended with. Upstream's
 *   backward rebuilds it as 1 - render_alpha, which rounds away up to ulp(alpha) / T ~ 3e-4 relative on an opaque
 *   pixel; pm_composite_bwd takes this array instead when it is given.
 * ------------------------------------------------------------------------- */
PM_API int pm_composite_fwd(const float *means2d, const float *conics, const float *opacities,
                     const float *colors, int colors_per_camera, int D,
                     const float *depths, const float *backgrounds,
                     const int32_t *isect_offsets, const int32_t *flatten_ids,
                     int64_t n_isects, const int32_t *d_n,
                     int N, int C, int width, int height, int tile_size,
                     float *render, float *alphas, int32_t *last_ids, float *transmittance,
                     float *splat_stream, pm_stream_t stream);

/* The SPLAT STREAM: what the compositing backward replays. splat_stream [n_isects, 12] fp32 (16-byte aligned, may be
 * NULL when no backward follows): the forward kernels write every Gaussian they stage, at its position in the sorted
 * list, as one 48-byte record {mean x, mean y, half extents of alpha >= 1/255 (2 x fp16, rounded up), flatten id |
 * conic scaled to base-2 exponent units, opacity | colour channels, depth}. Only the prefix of each tile's list that
 * the compositor consumed is written — all the backward reads. A tile's records are contiguous, so the backward
 * moves 32 of them into shared memory with one cp.async.bulk (TMA) + mbarrier instead of 32 x 7 scattered loads.
 * pm_splat_stream_build gathers the WHOLE sorted list into a stream (for a forward state that did not come from these
 * kernels, e.g. in tests, or a forward that ran with splat_stream = NULL). */
PM_API int pm_splat_stream_build(const float *means2d, const float *conics, const float *opacities,
                     const float *colors, int colors_per_camera, int D, const float *depths,
                     const int32_t *flatten_ids, int64_t n_isects, const int32_t *d_n,
                     int N, int C, int views_per_scene, float *splat_stream, pm_stream_t stream);

/* Fused "sort-as-you-go" forward (fast path of the module-level renderer): takes
 * the UNSORTED per-tile records of pm_bin_scatter and, per tile, sorts only as
 * many depth-ordered groups as the compositor consumes before every pixel of
 * the tile is saturated. render / alphas / last_ids are bit-identical to
 * pm_tile_sort + pm_composite_fwd; flatten_ids [n] is filled for the first
 * sorted_counts[c,ty,tx] entries of each tile only (all that pm_composite_bwd
 * reads). *overflow_flag (device int32, caller zero-fills) is set if a depth
 * second-level sub-bucket of some tile still exceeds the 768-record group capacity (e.g. hundreds of identical depths): the caller
 * must then redo the call with pm_tile_sort + pm_composite_fwd. tile_order (may be NULL): the
 * (camera, tile) index each CTA takes, e.g. heaviest first from pm_tile_scan. */
PM_API int pm_tile_render_fwd(const uint64_t *records, const int32_t *isect_offsets, int64_t n_isects,
                              const int32_t *d_n, const float *means2d, const float *conics,
                              const float *opacities, const float *colors, int colors_per_camera, int D,
                              const float *depths, const float *backgrounds, int N, int C,
                              int width, int height, int tile_size,
                              float *render, float *alphas, int32_t *last_ids,
                              int32_t *flatten_ids, int32_t *sorted_counts, int32_t *overflow_flag,
                              const int32_t *tile_order, float *transmittance, float *splat_stream,
                              pm_stream_t stream);

/* The same kernel with the module-level epilogue fused in (RGB + expected depth, D = 3): writes
 * what GaussianRenderer.forward returns (gaussians_renderer.py:437-440, :507-509) —
 * images [C,3,H,W] = 2 rgb - 1, depth_out [C,1,H,W] = accumulated depth / max(alpha, 1e-10),
 * masks [C,1,H,W] = alpha in [0,1] — instead of the interleaved render / alphas tensors.
 * n_isects may be a speculative CAPACITY of records / flatten_ids with the true count in *d_n:
 * nothing is accessed beyond min(n_isects, *d_n); if *d_n > n_isects the outputs are invalid and
 * the caller must retry with a larger capacity.
 * record_stride = 0: records is the compact array of pm_bin_scatter (tile segments back to back at isect_offsets).
 * record_stride = S > 0: the fixed-capacity layout of pm_project_scatter_fwd — tile (c, t) owns records[(c * tiles +
 * t) * S ...], of which min(count, S) are valid; isect_offsets / n_isects then only position flatten_ids and the
 * splat stream. */
PM_API int pm_tile_render_packed_fwd(const uint64_t *records, int64_t record_stride,
                              const int32_t *isect_offsets, int64_t n_isects,
                              const int32_t *d_n, const float *means2d, const float *conics,
                              const float *opacities, const float *colors, int colors_per_camera,
                              const float *depths, const float *backgrounds, int N, int C,
                              int width, int height, int tile_size,
                              float *images, float *depth_out, float *masks, int32_t *last_ids,
                              int32_t *flatten_ids, int32_t *sorted_counts, int32_t *overflow_flag,
                              const int32_t *tile_order, int views_per_scene, float *transmittance,
                              float *splat_stream, uint32_t *hit_words, int64_t hit_stride, pm_stream_t stream);

/* The HIT MAP (optional, only useful when a backward follows): hit_words [8, hit_stride] uint32, zero-filled by the
 * caller, hit_stride >= pm_hit_words_stride(n_isects, ...). Row w belongs to warp footprint w of every tile (the 8 x 4
 * pixel blocks of a 16 x 16 tile, row-major in steps of (8, 4)); tile (c, t)'s words start at
 * (isect_offsets[c, t] >> 5) + (c * tiles + t) and bit p of them is set iff the forward evaluated list entry p of the
 * tile for that footprint (it passed the footprint's culling). pm_composite_bwd_packed replays exactly those entries
 * and skips its own culling; chunks of 32 entries without a set bit are not even copied. NULL: no map. */
PM_API int64_t pm_hit_words_stride(int64_t n_isects, int C, int width, int height, int tile_size);

/* Backward. Replays splat_stream (see above; colours [N,D] or [C,N,D] per colors_per_camera, has_depth = the forward
 * rendered a depth channel, CD = D + has_depth). v_render [C,H,W,CD], v_alphas [C,H,W] (either may be NULL).
 * transmittance [C,H,W] = the forward's array of that name, or NULL: T_final is then rebuilt as 1 - alphas like
 * upstream. Gradients are ACCUMULATED (+=) into v_means2d [C,N,2], v_conics [C,N,3], v_colors ([N,D] or [C,N,D]),
 * v_opacities [N], v_depths [C,N]: the caller zero-fills them. v_depths must be non-NULL iff has_depth.
 * touched [C,N] uint8 (may be NULL; caller zero-fills): set to 1 for every pair that received a gradient, for
 * pm_project_bwd. */
PM_API int pm_composite_bwd(const float *splat_stream, int colors_per_camera, int D, int has_depth,
                     const float *backgrounds, const int32_t *isect_offsets,
                     int64_t n_isects, const int32_t *d_n,
                     int N, int C, int width, int height, int tile_size,
                     const float *alphas, const int32_t *last_ids, const float *transmittance,
                     const float *v_render, const float *v_alphas,
                     float *v_means2d, float *v_conics, float *v_colors,
                     float *v_opacities, float *v_depths, uint8_t *touched, pm_stream_t stream);

/* Backward of pm_tile_render_packed_fwd: cotangents of the packed outputs (any may be NULL) in,
 * the same accumulated per-Gaussian gradients out. masks / depth_out are the forward's outputs.
 * tile_order (may be NULL): launch order of the tiles, e.g. pm_tile_scan over the forward's sorted_counts.
 * hit_words / hit_stride (may be NULL / 0): the forward's hit map, see pm_tile_render_packed_fwd. */
PM_API int pm_composite_bwd_packed(const float *splat_stream, int colors_per_camera,
                     const float *backgrounds, const int32_t *isect_offsets,
                     int64_t n_isects, const int32_t *d_n,
                     int N, int C, int width, int height, int tile_size,
                     const float *masks, const float *depth_out, const int32_t *last_ids,
                     const float *transmittance,
                     const float *v_images, const float *v_depth_out, const float *v_masks,
                     float *v_means2d, float *v_conics, float *v_colors, float *v_opacities,
                     float *v_depths, const int32_t *tile_order, int views_per_scene, uint8_t *touched,
                     const uint32_t *hit_words, int64_t hit_stride, pm_stream_t stream);

/* ------------------------------------------------------------------------- *
 * Trajectory rendering: pose sampling along a camera path and
 * frame packing on the device. Replaces sample_from_dense_cameras
 * (prometheus/utils/camera.py:134-217, called once per frame by export_video,
 * prometheus/utils/io.py:167-190) and the per-frame host conversion of io.py:181.
 * ------------------------------------------------------------------------- */

/* dense_cameras [B,N,18] (c2w 3x4 row-major + fx fy cx cy H W), t [B,M] in [0,1] ->
 * out [B,M,18]: poses interpolated at arc length t (position lerp + quaternion slerp,
 * distance = sqrt(|dp|^2 + angle^2)), intrinsics lerped. 2 <= N <= 1024. */
PM_API int pm_sample_cameras(const float *dense_cameras, const float *t, int B, int N, int M,
                      float *out, pm_stream_t stream);

/* images [F,3,H,W] fp32 in [-1,1] -> out [F,H,W,3] uint8 = trunc(clamp(x/2+1/2,0,1)*255). */
PM_API int pm_pack_frames_u8(const float *images, int64_t n_frames, int height, int width,
                      uint8_t *out, pm_stream_t stream);

/* Densification statistics of the SDS++ refiner (prometheus/models/sdspp_refiner.py:275-280,
 * prometheus/utils/gs_utils.py:247-252) for all C rendered views in one pass, views in order:
 * vis = radii[c] > 0; is_visible[vis] = 1; max_radii2d[vis] = max(., radii[c]);
 * xyz_gradient_accum[vis] += |v_means2d[c]|; denom[vis] += 1. radii [C,N] i32, v_means2d [C,N,2]
 * (the gradient the compositing backward wrote); the four [N] fp32 state arrays are updated in place. */
PM_API int pm_densify_stats(const int32_t *radii, const float *v_means2d, int C, int N,
                     float *is_visible, float *max_radii2d, float *xyz_gradient_accum, float *denom,
                     pm_stream_t stream);

/* .splat export (prometheus/utils/io.py:25-125, export_ply_for_gaussians): sort keys
 * -exp(sum log scale) / (1 + exp(-logit opacity)) as order-preserving uint32 in keys[i] (value i in vals[i]),
 * to be sorted ascending with pm_sort_pairs(end_bit = 32); then the 32-byte records in that order:
 * xyz 3 x f32, scale 3 x f32, rgba 4 x u8, normalised quaternion * 128 + 128 as 4 x u8. opacity [n], scales [n,3],
 * xyz [n,3], rgb [n,3], rots [n,4] are the ACTIVATED Gaussians the decoder produces; compatible != 0 replays the
 * reference's log / logit round trip. */
PM_API int pm_splat_keys(const float *opacity, const float *scales, int64_t n, int compatible,
                  int64_t *keys, int32_t *vals, pm_stream_t stream);
PM_API int pm_splat_pack(const float *xyz, const float *rgb, const float *opacity, const float *scales,
                  const float *rots, const int32_t *order, int64_t n, int compatible,
                  uint8_t *out, pm_stream_t stream);

/* Loss-side glue of the stage-1 rendering loss (prometheus/systems/gs_decoder_system.py:214-239) on the packed
 * renderer outputs, one pass: images / target [V,3,H,W], depths / masks [V,1,H,W] ->
 *   disparity [V,1,H,W] = 1 / clamp(depth, 0.1, 100)     (:218, input of the scale-and-shift-invariant loss)
 *   valid     [V,1,H,W] = depth <= 0.1 ? 0 : 1           (:216-217)
 *   sums[0] = sum (images - target)^2  (F.mse_loss numerator, :221),  sums[1] = sum (1 - masks)  (:233)
 * sums is 2 doubles on the device, overwritten. */
PM_API int pm_loss_terms_fwd(const float *images, const float *target, const float *depths, const float *masks,
                      int64_t n_views, int height, int width, float *disparity, float *valid, double *sums,
                      pm_stream_t stream);
/* Its VJP, straight into the packed cotangents pm_composite_bwd_packed takes: g_mse / g_entropy are the (device,
 * scalar) cotangents of mean((images - target)^2) and mean(1 - masks), v_disparity [V,1,H,W] that of disparity;
 * any of the three may be NULL (zero). v_images / v_depths / v_masks are fully overwritten. */
PM_API int pm_loss_terms_bwd(const float *images, const float *target, const float *depths, int64_t n_views,
                      int height, int width, const float *g_mse, const float *g_entropy, const float *v_disparity,
                      float *v_images, float *v_depths, float *v_masks, pm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PROMETHEUS_GS_H */
