// tile_render.cu — fused "sort-as-you-go" tile renderer (forward), the fast path behind
// prometheus_b200.GaussianRenderer.
//
// Front-to-back compositing stops as soon as every pixel of a tile is saturated; on the
// pixel-aligned scenes of this workload that happens after ~19 % of the tile's depth-sorted list
// (I_need / I, bench.py). The separate-stage pipeline (tile_sort.cu + composite.cu — what
// gsplat.rasterization needs, because it exposes the fully sorted arrays) still sorts 100 % of every
// tile. This kernel sorts only what the compositor consumes: one CTA per (camera, tile)
//   1. reads the tile's unsorted {depth bits, flatten id} records (written by bin_scatter),
//      finds the depth range and histograms the records into 256 depth-ordered sub-buckets;
//   2. walks the sub-buckets front to back in groups of <= 768 records: collect the group into shared memory in
//      bucket order, rank each bucket's short run by (depth bits, flatten id), write the sorted flatten ids
//      to their final positions (the backward pass replays them), stage the WHOLE group's Gaussian
//      records (mean/extent, conic/opacity, colour: 48 B each, over the sort buffers) and let the 8
//      warps composite it with no CTA barrier in between — each warp runs at its own pace through
//      the staged group and drops out when its 32 pixels are done (the per-256 barrier of the
//      separate-stage kernel was 30 % of this kernel's stall samples, profiles/r1e_*);
//   3. stops at the first group boundary where all 256 pixels are done.
// Because sub-buckets partition the keys by value and ties are ordered by flatten id, position k of
// the tile's list holds the same Gaussian as after a full stable sort, so images, alphas and
// last_ids are BIT-IDENTICAL to the separate-stage path (tests/test_gpu_parity.py) — only the tail of
// flatten_ids that nobody reads is left unsorted.
// Reference semantics: gsplat 1.4.0 isect_tiles + rasterize_to_pixels_fwd (call site
// prometheus/modules/renderers/gaussians_renderer.py:413-427), SURVEY.md §8c steps 6-10.
#include "composite_common.cuh"

namespace pm {

constexpr int kTrThreads = 256;
constexpr int kTrWarps = kTrThreads / 32;
#ifndef PM_TR_MINBLOCKS
#define PM_TR_MINBLOCKS 4
#endif
#ifndef PM_TR_CAP
#define PM_TR_CAP 768
#endif
constexpr int kTrCap = PM_TR_CAP;                                   // records per group
constexpr int kTrItems = (kTrCap + kTrThreads - 1) / kTrThreads;   // records per thread in the sort / staging loops
constexpr int kTrRadix = 256;   // digits of the in-group radix sort
#ifndef PM_TR_BUCKET_BITS
#define PM_TR_BUCKET_BITS 9
#endif
constexpr int kTrBucketBits = PM_TR_BUCKET_BITS;
constexpr int kTrBuckets = 1 << kTrBucketBits;  // depth-ordered buckets per level: shorter runs = cheaper all-pairs ranking (256: runs of ~23 on C2, 9 % of the kernel's instructions)
#ifndef PM_TR_MLP
#define PM_TR_MLP 8  // record loads in flight per thread in the passes over a tile's segment
#endif
constexpr int kTrMlp = PM_TR_MLP;
#ifndef PM_TR_MLP_K
#define PM_TR_MLP_K 8  // key-only (32-bit) loads in flight per thread in the histogram / range passes (16: 1.14, 24: 1.20 ms — L1 thrashes)
#endif
constexpr int kTrMlpK = PM_TR_MLP_K;
constexpr int kTrRankMax = 128;  // longest bucket run ordered by all-pairs ranking (longer: radix sort of the group)

// depth sub-bucket of a key: monotone in the key for any kmin / shift
__device__ __forceinline__ int tr_bucket(uint32_t k, uint32_t kmin, int shift) {
  return k <= kmin ? 0 : (int)min((k - kmin) >> shift, (uint32_t)(kTrBuckets - 1));
}

__device__ __forceinline__ unsigned tr_match_digit8(uint32_t d) {
  unsigned peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const unsigned m = __ballot_sync(0xffffffffu, (d >> b) & 1u);
    peers &= ((d >> b) & 1u) ? m : ~m;
  }
  return peers;
}

struct TrShared {
  union {                                    // a group is first sorted, then staged over the same bytes
    struct {
      uint2 gbuf[kTrItems * kTrThreads];     // group being sorted: {depth bits, flatten id} (radix sort pads to a full round)
      uint16_t warp_hist[kTrWarps][kTrRadix];
      uint32_t digit_start[kTrRadix];
    };
    Staged<kTrCap> stg;                      // (its null slot lies beyond the sort buffers)
  };
  uint32_t queue[kTrWarps][kQueueWords];     // per-warp hit queues of the compositing loop
  uint32_t cand[kTrWarps][kCandWords];       // per-warp candidate rings (composite_group_masked_fwd)
  uint32_t fmask[kTrWarps][kTrItems * kTrWarps];  // footprint words: [compositing warp][32 staged slots]
  uint32_t red[64];
  int bucket[kTrBuckets + 1];                // start of each depth sub-bucket inside the tile's list
  int bucket2[kTrBuckets + 1];               // second level: sub-buckets of one level-1 bucket that exceeds a group
  uint32_t cursor[kTrBuckets];               // per-bucket cursors of the counting-sort placement
  int cnt;
  uint32_t kmin, kmax;
};
constexpr size_t kTrSmemBytes = sizeof(TrShared);

// Sort sh.gbuf[0..m) by (key, flatten id); gmin/gmax = key range of the group. All threads call it.
// @region tr_radix_sort_group
__device__ __forceinline__ void sort_group(TrShared &sh, int m, uint32_t gmin, uint32_t gmax) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk0 = warp * (kTrItems * 32);
  uint32_t dk[kTrItems], fid[kTrItems], rk[kTrItems];
#pragma unroll
  for (int k = 0; k < kTrItems; ++k) {
    const int slot = chunk0 + k * 32 + lane;
    if (slot < m) {
      const uint2 v = sh.gbuf[slot];
      dk[k] = v.x - gmin;
      fid[k] = v.y;
    } else {  // empty slots: all-ones, sorted like real records and therefore kept at the end
      dk[k] = 0xffffffffu;
      fid[k] = 0xffffffffu;
    }
  }
  const uint32_t range = gmax - gmin;
  for (int shift = 0; shift < 32 && (range >> shift) != 0; shift += 8) {
    __syncthreads();
    for (int i = tid; i < kTrWarps * kTrRadix / 2; i += kTrThreads) reinterpret_cast<uint32_t *>(sh.warp_hist)[i] = 0;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTrItems; ++k) {
      const uint32_t d = (dk[k] >> shift) & 0xffu;
      const unsigned peers = tr_match_digit8(d);
      const unsigned lower = peers & ((1u << lane) - 1);
      uint32_t base = 0;
      if (lower == 0) {
        base = sh.warp_hist[warp][d];
        sh.warp_hist[warp][d] = (uint16_t)(base + __popc(peers));
      }
      base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
      rk[k] = base + __popc(lower);
      __syncwarp();
    }
    __syncthreads();
    {  // digit d (thread d): exclusive offsets across warps, then exclusive scan over the 256 digits
      uint32_t count = 0;
#pragma unroll
      for (int w = 0; w < kTrWarps; ++w) {
        const uint32_t c = sh.warp_hist[w][tid];
        sh.warp_hist[w][tid] = (uint16_t)count;
        count += c;
      }
      uint32_t inc = count;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) sh.red[warp] = inc;
      sh.digit_start[tid] = inc - count;
    }
    __syncthreads();
    {
      uint32_t wbase = 0;
      for (int w = 0; w < warp; ++w) wbase += sh.red[w];
      sh.digit_start[tid] += wbase;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTrItems; ++k) {
      const uint32_t d = (dk[k] >> shift) & 0xffu;
      sh.gbuf[sh.digit_start[d] + sh.warp_hist[warp][d] + rk[k]] = make_uint2(dk[k], fid[k]);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kTrItems; ++k) {
      const uint2 v = sh.gbuf[chunk0 + k * 32 + lane];
      dk[k] = v.x;
      fid[k] = v.y;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kTrItems; ++k) {
    const int slot = chunk0 + k * 32 + lane;
    if (slot < m) sh.gbuf[slot] = make_uint2(dk[k] + gmin, fid[k]);
  }
  __syncthreads();
  // ties on the depth bits: order each run by flatten id
#pragma unroll
  for (int k = 0; k < kTrItems; ++k) {
    const int slot = chunk0 + k * 32 + lane;
    if (slot < m - 1) {
      const uint32_t key = sh.gbuf[slot].x;
      const bool run_start = (slot == 0 || sh.gbuf[slot - 1].x != key) && sh.gbuf[slot + 1].x == key;
      if (run_start) {
        int e = slot + 1;
        while (e < m && sh.gbuf[e].x == key) ++e;
        for (int i = slot + 1; i < e; ++i) {
          const uint32_t v = sh.gbuf[i].y;
          int j = i - 1;
          while (j >= slot && sh.gbuf[j].y > v) { sh.gbuf[j + 1].y = sh.gbuf[j].y; --j; }
          sh.gbuf[j + 1].y = v;
        }
      }
    }
  }
  __syncthreads();
}

// @region tr_scan
// Exclusive scan of tbl[1..kTrBuckets] in place (tbl[0] = 0 on entry): tbl[d] = start of bucket d, tbl[kTrBuckets] = total.
// Thread t owns the kTrBuckets / kTrThreads consecutive entries behind tbl[1 + t * per].
__device__ __forceinline__ void tr_scan_buckets(int *tbl, uint32_t *red) {
  constexpr int per = kTrBuckets / kTrThreads;
  static_assert(per >= 1 && per * kTrThreads == kTrBuckets, "bucket count must be a multiple of the CTA size");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int v[per], sum = 0;
#pragma unroll
  for (int i = 0; i < per; ++i) { v[i] = tbl[1 + tid * per + i]; sum += v[i]; }
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) red[warp] = (uint32_t)inc;
  __syncthreads();
  int run = inc - sum;
  for (int w = 0; w < warp; ++w) run += (int)red[w];
#pragma unroll
  for (int i = 0; i < per; ++i) { run += v[i]; tbl[1 + tid * per + i] = run; }
  __syncthreads();
}

// @region tr_prologue
__global__ void __launch_bounds__(kTrThreads, PM_TR_MINBLOCKS)
tile_render_fwd_kernel(const uint64_t *__restrict__ records, int64_t record_stride,
                       const int32_t *__restrict__ offsets, int64_t n_host, const int32_t *__restrict__ d_n, const float *__restrict__ means2d,
                       const float *__restrict__ conics, const float *__restrict__ opacities,
                       const float *__restrict__ colors, int colors_per_camera, int D,
                       const float *__restrict__ depths, const float *__restrict__ backgrounds, int N,
                       int width, int height, int tile_width, int tile_height, float *__restrict__ render,
                       float *__restrict__ alphas, int32_t *__restrict__ last_ids,
                       int32_t *__restrict__ flatten_ids, int32_t *__restrict__ sorted_counts,
                       int32_t *__restrict__ overflow_flag, float *__restrict__ images,
                       float *__restrict__ depth_out, float *__restrict__ masks,
                       const int32_t *__restrict__ tile_order, int views_per_scene,
                       float *__restrict__ transmittance, float *__restrict__ stream,
                       uint32_t *__restrict__ hit_words, int64_t hit_stride) {
  extern __shared__ __align__(16) unsigned char tr_smem[];
  TrShared &sh = *reinterpret_cast<TrShared *>(tr_smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_tiles = tile_width * tile_height;
  const int ct = tile_order ? tile_order[blockIdx.x] : (int)blockIdx.x;  // heaviest tiles first (pm_tile_scan)
  const int cam = ct / n_tiles, tile = ct % n_tiles;
  const int tile_y = tile / tile_width, tile_x = tile % tile_width;
  const int wx0 = tile_x * kTile + (warp & 1) * 8, wy0 = tile_y * kTile + (warp >> 1) * 4;
  const int ix = wx0 + (lane & 7), iy = wy0 + (lane >> 3);
  const bool inside = ix < width && iy < height;
  const int CD = D + (depths ? 1 : 0);
  // flatten id f = cam * N + g; the per-Gaussian arrays hold one block of N per scene (scene = cam / views_per_scene)
  const int cam_base = cam * N - (cam / views_per_scene) * N;
  const WarpFootprint fp{(float)ix + 0.5f, (float)iy + 0.5f, (float)wx0 + 4.0f, (float)wy0 + 2.0f};
  PixelState st{inside ? 1.0f : -1.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0};
  const uint32_t aS = smem_u32(&sh.stg), aQ = smem_u32(sh.queue[warp]);
  const uint32_t aCQ = smem_u32(sh.cand[warp]), aM = smem_u32(sh.fmask[warp]);
  const float tile_px0 = (float)(tile_x * kTile) + 0.5f, tile_py0 = (float)(tile_y * kTile) + 0.5f;

  const int64_t n_isects = effective_n(n_host, d_n);
  // (clamped: with a speculative record capacity smaller than the device-side count the result is
  //  discarded by the caller, but nothing may be read or written out of bounds)
  const int start = min(offsets[ct], (int)n_isects);
  const int end = (ct == (int)gridDim.x - 1) ? (int)n_isects : min(offsets[ct + 1], (int)n_isects);
  // fixed-capacity tile regions (single-pass binning): the tile's records sit at ct * stride, at most `stride` of them
  const int cnt = record_stride > 0 ? min(end - start, (int)record_stride) : end - start;
  const unsigned long long *src =
      reinterpret_cast<const unsigned long long *>(records) + (record_stride > 0 ? (size_t)ct * (size_t)record_stride : (size_t)start);
  const uint32_t *keys = reinterpret_cast<const uint32_t *>(src);  // record i: keys[2 i] = depth bits, keys[2 i + 1] = flatten id
  int sorted = 0;

  // @region tr_histogram_pass
  // Histogram of the tile's records with key in [flo, fhi] over the 256 buckets of (kmin_c, shift_c) into tbl,
  // then the exclusive scan. (Every pass over the tile's records keeps kTrMlp loads per thread in flight.)
  auto build_table = [&](int *tbl, uint32_t kmin_c, int shift_c, uint32_t flo, uint32_t fhi) {
    for (int i = tid; i <= kTrBuckets; i += kTrThreads) tbl[i] = 0;
    if (tid == 0) PM_STAT_ADD(6, 1);
    __syncthreads();
    for (int i0 = 0; i0 < cnt; i0 += kTrMlpK * kTrThreads) {
      uint32_t k[kTrMlpK];
#pragma unroll
      for (int u = 0; u < kTrMlpK; ++u) {
        const int i = i0 + u * kTrThreads + tid;
        k[u] = i < cnt ? __ldg(keys + 2 * (size_t)i) : 0u;
      }
#pragma unroll
      for (int u = 0; u < kTrMlpK; ++u)
        if (i0 + u * kTrThreads + tid < cnt && k[u] >= flo && k[u] <= fhi)
          atomicAdd(&tbl[tr_bucket(k[u], kmin_c, shift_c) + 1], 1);
    }
    __syncthreads();
    tr_scan_buckets(tbl, sh.red);
  };

  // @region tr_collect_pass
  // One group: the records with key in [flo, fhi] whose bucket under (kmin_c, shift_c) lies in [b0, b1) — at most
  // kTrCap of them, occupying list positions list_base + tbl[b0] .. list_base + tbl[b1] of the tile.
  auto process_group = [&](const int *tbl, uint32_t kmin_c, int shift_c, uint32_t flo, uint32_t fhi, int b0, int b1,
                           int list_base) {
    const int gs = tbl[b0], m = tbl[b1] - gs;
    // collect the group straight into bucket order (a counting-sort placement: the position of a record is
    // its bucket's start + a per-bucket cursor), so that only the short runs of equal bucket remain to be ordered
    uint32_t *cur = sh.cursor;
    for (int bb = b0 + tid; bb < b1; bb += kTrThreads) cur[bb] = 0u;
    if (tid == 0) { sh.kmin = 0xffffffffu; sh.kmax = 0u; PM_STAT_ADD(6, 1); PM_STAT_ADD(7, 1); PM_STAT_ADD(10, m); }
    __syncthreads();
    uint32_t gmin = 0xffffffffu, gmax = 0u;
    for (int i0 = 0; i0 < cnt; i0 += kTrMlp * kTrThreads) {
      unsigned long long r[kTrMlp];
#pragma unroll
      for (int u = 0; u < kTrMlp; ++u) {
        const int i = i0 + u * kTrThreads + tid;
        r[u] = i < cnt ? __ldg(src + i) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < kTrMlp; ++u) {
        const uint32_t k = (uint32_t)r[u];
        const int bb = tr_bucket(k, kmin_c, shift_c);
        if (i0 + u * kTrThreads + tid < cnt && k >= flo && k <= fhi && bb >= b0 && bb < b1) {
          const int pos = tbl[bb] - gs + (int)atomicAdd(&cur[bb], 1u);
          sh.gbuf[pos] = make_uint2(k, (uint32_t)(r[u] >> 32));
          gmin = min(gmin, k);
          gmax = max(gmax, k);
        }
      }
    }
    // @region tr_rank
    bool fat = false;  // a run too long for the all-pairs ranking below
    for (int bb = b0 + tid; bb < b1; bb += kTrThreads) fat |= tbl[bb + 1] - tbl[bb] > kTrRankMax;
    gmin = __reduce_min_sync(0xffffffffu, gmin);
    gmax = __reduce_max_sync(0xffffffffu, gmax);
    if (lane == 0) { atomicMin(&sh.kmin, gmin); atomicMax(&sh.kmax, gmax); }
    fat = __syncthreads_or(fat);
    int32_t ids[kTrItems];
    if (fat) {
      if (tid == 0) PM_STAT_ADD(9, 1);
      sort_group(sh, m, sh.kmin, sh.kmax);  // LSD radix sort of the whole group
#pragma unroll
      for (int k = 0; k < kTrItems; ++k) {
        const int slot = k * kTrThreads + tid;
        ids[k] = slot < m ? (int32_t)sh.gbuf[slot].y : -1;
      }
    } else {
      uint32_t *sorted_ids = reinterpret_cast<uint32_t *>(sh.warp_hist);
      // rank every record inside its bucket's run by (depth bits, flatten id): runs are ~cnt/256 long,
      // a warp's 32 consecutive records share one or two runs (broadcast reads), no barrier needed
#pragma unroll
      for (int k = 0; k < kTrItems; ++k) {
        const int e = k * kTrThreads + tid;
        if (e < m) {
          const uint2 me = sh.gbuf[e];
          const int bb = tr_bucket(me.x, kmin_c, shift_c);
          const int rs = tbl[bb] - gs, re = tbl[bb + 1] - gs;
          const unsigned long long mine = ((unsigned long long)me.x << 32) | me.y;
          int rank = rs;
          for (int j = rs; j < re; ++j) {
            const uint2 o = sh.gbuf[j];
            rank += (((unsigned long long)o.x << 32) | o.y) < mine;
          }
          sorted_ids[rank] = me.y;
        }
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kTrItems; ++k) {
        const int slot = k * kTrThreads + tid;
        ids[k] = slot < m ? (int32_t)sorted_ids[slot] : -1;
      }
    }
    // @region tr_ids_gather_stage
    // sorted flatten ids go to their final positions (read again by the backward pass)
#pragma unroll
    for (int k = 0; k < kTrItems; ++k) {
      const int slot = k * kTrThreads + tid;
      if (slot < m) flatten_ids[(size_t)start + list_base + gs + slot] = ids[k];
    }
    sorted = list_base + gs + m;
    __syncthreads();  // everyone holds its ids: the sort buffers may be overwritten by the staged records
    staged_init_null(sh.stg);  // (for small group capacities the null slot overlaps the sort buffers)
#pragma unroll
    for (int k = 0; k < kTrItems; ++k) {
      const int slot = k * kTrThreads + tid;
      unsigned fm = 0u;  // warp footprints of the tile this Gaussian's extent overlaps
      if (slot < m) {
        GaussA ga; GaussB gb; GaussC gc;
        gather(ids[k], ids[k] - cam_base, means2d, conics, opacities, colors, colors_per_camera, D, depths, ga, gb, gc);
        sh.stg.sA[slot] = ga; sh.stg.sB[slot] = gb; sh.stg.sC[slot] = gc;
        // what the backward replays: the staged record, at its position in the tile's sorted list
        if (stream) stream_store(stream, (size_t)start + list_base + gs + slot, ids[k], ga, gb, gc);
        fm = footprint_mask(ga.mx, ga.my, ga.hx, ga.hy, tile_px0, tile_py0);
      }
      // one word per (compositing warp, these 32 slots): lane w keeps the ballot of footprint w
      uint32_t word = 0u;
#pragma unroll
      for (int w = 0; w < kTrWarps; ++w) {
        const uint32_t b = __ballot_sync(0xffffffffu, (fm >> w) & 1u);
        if (lane == w) word = b;
      }
      if (lane < kTrWarps) sh.fmask[lane][k * kTrWarps + warp] = word;
    }
    __syncthreads();
    // @region tr_composite_call
    // composite the whole group; warps do not wait for each other — same arithmetic as composite_fwd_kernel
    if (!__all_sync(0xffffffffu, pixel_done(st))) {
      // hit map of this (warp footprint, tile): bit p of the words at (start >> 5) + ct = list entry p was evaluated
      // (tiles never share a word: (start >> 5) + ct grows by at least ceil(count / 32) from tile to tile); formed
      // here from values that are live anyway — held across the kernel the pointer cost two registers and a spill
      uint32_t *hit_base = hit_words ? hit_words + (size_t)warp * (size_t)hit_stride + (size_t)(start >> 5) + (size_t)ct : nullptr;
      composite_group_masked_fwd<kTrCap>(aS, aQ, aCQ, aM, m, start + list_base + gs, fp, st, hit_base, list_base + gs);
    }
  };

  // @region tr_group_walk
  // largest b1 in (b0, 256] with tbl[b1] - tbl[b0] <= capacity (binary search; tbl[] is monotone), b0 + 1 if none
  auto group_end = [&](const int *tbl, int b0) {
    const int gs = tbl[b0];
    int b1 = b0 + 1, hi = kTrBuckets;
    while (b1 < hi) {
      const int mid = (b1 + hi + 1) >> 1;
      if (tbl[mid] - gs <= kTrCap) b1 = mid; else hi = mid - 1;
    }
    return b1;
  };

  if (cnt > 0) {
    // ---- 1. level-1 table: 256 depth-ordered buckets over the tile's key range. The bucket of a key is a
    // monotone function of it — clamp((key - kmin) >> shift, 0, 255) — so ANY kmin / shift partition the keys by
    // value; they only steer the bucket sizes. kmin / kmax therefore come from a strided sample of <= 1024
    // records (one round of 4 independent loads per thread, exact when the tile has <= 1024 records) instead of
    // a full pass; the few keys outside the sampled range land in buckets 0 / 255.
    if (tid == 0) { sh.kmin = 0xffffffffu; sh.kmax = 0u; PM_STAT_ADD(13, 1); PM_STAT_ADD(14, cnt); }
    __syncthreads();
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    {
      const int nsamp = min(cnt, 4 * kTrThreads);
      const int stride = cnt / nsamp;
      uint32_t k[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int sidx = u * kTrThreads + tid;
        k[u] = sidx < nsamp ? (uint32_t)__ldg(src + (size_t)sidx * stride) : 0u;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u * kTrThreads + tid < nsamp) { kmin = min(kmin, k[u]); kmax = max(kmax, k[u]); }
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) { atomicMin(&sh.kmin, kmin); atomicMax(&sh.kmax, kmax); }
    __syncthreads();
    kmin = sh.kmin;
    const uint32_t range = sh.kmax - kmin;
    const int bits = 32 - __clz(range | 1u);
    const int shift = bits > kTrBucketBits ? bits - kTrBucketBits : 0;
    __syncthreads();  // sh.kmin / sh.kmax are reused below
    build_table(sh.bucket, kmin, shift, 0u, 0xffffffffu);

    // ---- 2. groups of buckets, front to back, until every pixel is done
    bool stop = false;
    int g0 = 0;
    while (g0 < kTrBuckets && !stop) {
      // group boundary barrier: the staged group of the previous round is dead after it
      if (__syncthreads_count(pixel_done(st)) == kTrThreads) break;
      const int g1 = group_end(sh.bucket, g0);
      const int m = sh.bucket[g1] - sh.bucket[g0];
      if (m <= kTrCap) {
        if (m > 0) process_group(sh.bucket, kmin, shift, 0u, 0xffffffffu, g0, g1, 0);
        g0 = g1;
        continue;
      }
      // @region tr_level2
      // ---- 3. bucket g0 alone exceeds a group (a thin, dense depth shell): give it its own 256 sub-buckets
      // over the key range it actually holds and walk those. Only if a sub-bucket still does not fit — more than
      // kTrCap records within 1/65536 of the tile's depth range, e.g. identical depths — the caller re-renders
      // with the separate stages.
      const uint32_t flo = g0 == 0 ? 0u : kmin + ((uint32_t)g0 << shift);
      const uint32_t fhi = g0 == kTrBuckets - 1 ? 0xffffffffu : kmin + (((uint32_t)g0 + 1u) << shift) - 1u;
      if (tid == 0) { sh.kmin = 0xffffffffu; sh.kmax = 0u; PM_STAT_ADD(6, 1); PM_STAT_ADD(8, 1); }
      __syncthreads();
      uint32_t lo2 = 0xffffffffu, hi2 = 0u;
      for (int i0 = 0; i0 < cnt; i0 += kTrMlpK * kTrThreads) {
        uint32_t k[kTrMlpK];
#pragma unroll
        for (int u = 0; u < kTrMlpK; ++u) {
          const int i = i0 + u * kTrThreads + tid;
          k[u] = i < cnt ? __ldg(keys + 2 * (size_t)i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < kTrMlpK; ++u)
          if (i0 + u * kTrThreads + tid < cnt && k[u] >= flo && k[u] <= fhi) { lo2 = min(lo2, k[u]); hi2 = max(hi2, k[u]); }
      }
      lo2 = __reduce_min_sync(0xffffffffu, lo2);
      hi2 = __reduce_max_sync(0xffffffffu, hi2);
      if (lane == 0) { atomicMin(&sh.kmin, lo2); atomicMax(&sh.kmax, hi2); }
      __syncthreads();
      const uint32_t kmin2 = sh.kmin;
      const uint32_t range2 = sh.kmax - kmin2;
      const int bits2 = 32 - __clz(range2 | 1u);
      const int shift2 = bits2 > kTrBucketBits ? bits2 - kTrBucketBits : 0;
      __syncthreads();
      build_table(sh.bucket2, kmin2, shift2, flo, fhi);
      const int base2 = sh.bucket[g0];
      int s0 = 0;
      while (s0 < kTrBuckets) {
        if (s0 > 0 && __syncthreads_count(pixel_done(st)) == kTrThreads) { stop = true; break; }
        const int s1 = group_end(sh.bucket2, s0);
        const int m2 = sh.bucket2[s1] - sh.bucket2[s0];
        if (m2 > kTrCap) {
          if (tid == 0) atomicExch(overflow_flag, 1);
          stop = true;
          break;
        }
        if (m2 > 0) process_group(sh.bucket2, kmin2, shift2, flo, fhi, s0, s1, base2);
        s0 = s1;
      }
      g0 = g0 + 1;
    }
  }
  // @region tr_epilogue
  if (tid == 0) sorted_counts[ct] = sorted;
  if (inside) {
    if (images != nullptr)  // packed module outputs (pm_tile_render_packed_fwd)
      store_pixel_packed(st, cam, iy, ix, width, height, backgrounds, images, depth_out, masks, last_ids,
                         transmittance);
    else
      store_pixel(st, ((size_t)cam * height + iy) * width + ix, cam, D, CD, backgrounds, render, alphas, last_ids, transmittance);
  }
}

}  // namespace pm

extern "C" {

int64_t pm_hit_words_stride(int64_t n_isects, int C, int width, int height, int tile_size) {
  if (n_isects < 0 || C < 0 || width <= 0 || height <= 0 || tile_size <= 0) return 0;
  const int64_t tw = (width + tile_size - 1) / tile_size, th = (height + tile_size - 1) / tile_size;
  return (n_isects >> 5) + (int64_t)C * tw * th + 1;
}

static int tile_render_launch(const uint64_t *records, int64_t record_stride, const int32_t *isect_offsets, int64_t n_isects,
                              const int32_t *d_n, const float *means2d, const float *conics, const float *opacities,
                              const float *colors, int colors_per_camera, int D, const float *depths,
                              const float *backgrounds, int N, int C, int width, int height, int tile_size,
                              float *render, float *alphas, int32_t *last_ids, int32_t *flatten_ids,
                              int32_t *sorted_counts, int32_t *overflow_flag, float *images, float *depth_out,
                              float *masks, const int32_t *tile_order, int views_per_scene, float *transmittance,
                              float *splat_stream, uint32_t *hit_words, int64_t hit_stride, pm_stream_t stream) {
  if (N < 0 || C < 0 || width <= 0 || height <= 0 || n_isects < 0 || D < 0 || record_stride < 0) return PM_ERR_BAD_ARG;
  if (record_stride >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  if (tile_size != pm::kTile) return PM_ERR_UNSUPPORTED;
  const int CD = D + (depths ? 1 : 0);
  if (CD < 1 || CD > 4) return PM_ERR_UNSUPPORTED;
  if (!means2d || !conics || !opacities || !isect_offsets || !last_ids || !sorted_counts || !overflow_flag)
    return PM_ERR_BAD_ARG;
  if (images) {  // packed outputs: RGB + expected depth only
    if (D != 3 || !depths) return PM_ERR_UNSUPPORTED;
    if (!depth_out || !masks) return PM_ERR_BAD_ARG;
  } else {
    if (!render || !alphas) return PM_ERR_BAD_ARG;
    if (CD == 4 && (reinterpret_cast<uintptr_t>(render) & 15) != 0) return PM_ERR_BAD_ARG;
  }
  if (D > 0 && !colors) return PM_ERR_BAD_ARG;
  if (n_isects > 0 && (!records || !flatten_ids)) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(means2d) & 7) != 0) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(splat_stream) & 15) != 0) return PM_ERR_BAD_ARG;
  if (n_isects >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  if (C == 0) return PM_OK;
  if (views_per_scene > 0 && C % views_per_scene != 0) return PM_ERR_BAD_ARG;
  const int tw = (width + pm::kTile - 1) / pm::kTile, th = (height + pm::kTile - 1) / pm::kTile;
  const int64_t blocks = (int64_t)C * tw * th;
  if (blocks >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  if (hit_words && hit_stride < pm_hit_words_stride(n_isects, C, width, height, tile_size)) return PM_ERR_WORKSPACE;
  // per device and cheap: set on every call (a process may drive several GPUs)
  if (cudaFuncSetAttribute(pm::tile_render_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)pm::kTrSmemBytes) != cudaSuccess) {
    pm::set_last_cuda_error(cudaGetLastError());
    return PM_ERR_CUDA;
  }
  pm::tile_render_fwd_kernel<<<(unsigned)blocks, pm::kTrThreads, pm::kTrSmemBytes, pm::as_stream(stream)>>>(
      records, record_stride, isect_offsets, n_isects, d_n, means2d, conics, opacities, colors, colors_per_camera, D, depths,
      backgrounds, N, width, height, tw, th, render, alphas, last_ids, flatten_ids, sorted_counts, overflow_flag,
      images, depth_out, masks, tile_order, views_per_scene > 0 ? views_per_scene : C, transmittance, splat_stream,
      hit_words, hit_stride);
  return pm::check_launch();
}

#ifdef PM_STATS
__attribute__((visibility("default"))) int pm_debug_stats(unsigned long long *out, int reset) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, pm::g_pm_stats, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(pm::g_pm_stats, z, sizeof(z));
  }
  return 0;
}
#endif

int pm_tile_render_fwd(const uint64_t *records, const int32_t *isect_offsets, int64_t n_isects,
                       const int32_t *d_n, const float *means2d, const float *conics,
                       const float *opacities, const float *colors, int colors_per_camera, int D,
                       const float *depths, const float *backgrounds, int N, int C, int width, int height,
                       int tile_size, float *render, float *alphas, int32_t *last_ids,
                       int32_t *flatten_ids, int32_t *sorted_counts, int32_t *overflow_flag,
                       const int32_t *tile_order, float *transmittance, float *splat_stream, pm_stream_t stream) {
  return tile_render_launch(records, 0, isect_offsets, n_isects, d_n, means2d, conics, opacities, colors, colors_per_camera,
                            D, depths, backgrounds, N, C, width, height, tile_size, render, alphas, last_ids,
                            flatten_ids, sorted_counts, overflow_flag, nullptr, nullptr, nullptr, tile_order, 0, transmittance,
                            splat_stream, nullptr, 0, stream);
}

int pm_tile_render_packed_fwd(const uint64_t *records, int64_t record_stride, const int32_t *isect_offsets, int64_t n_isects,
                              const int32_t *d_n, const float *means2d, const float *conics,
                              const float *opacities, const float *colors, int colors_per_camera,
                              const float *depths, const float *backgrounds, int N, int C, int width, int height,
                              int tile_size, float *images, float *depth_out, float *masks, int32_t *last_ids,
                              int32_t *flatten_ids, int32_t *sorted_counts, int32_t *overflow_flag,
                              const int32_t *tile_order, int views_per_scene, float *transmittance,
                              float *splat_stream, uint32_t *hit_words, int64_t hit_stride, pm_stream_t stream) {
  if (!images) return PM_ERR_BAD_ARG;
  return tile_render_launch(records, record_stride, isect_offsets, n_isects, d_n, means2d, conics, opacities, colors, colors_per_camera,
                            3, depths, backgrounds, N, C, width, height, tile_size, nullptr, nullptr, last_ids,
                            flatten_ids, sorted_counts, overflow_flag, images, depth_out, masks, tile_order, views_per_scene,
                            transmittance, splat_stream, hit_words, hit_stride, stream);
}

}  // extern "C"
