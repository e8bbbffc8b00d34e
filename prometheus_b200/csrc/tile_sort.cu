// tile_sort.cu — stage 3, B200-first: binned ("tile-bucketed") depth sort.
//
// Produces exactly what gsplat 1.4.0's isect_tiles + cub::DeviceRadixSort::SortPairs
// + isect_offset_encode produce (reference call site
// prometheus/modules/renderers/gaussians_renderer.py:413-427; SURVEY.md §8c steps
// 6-9) — sorted isect_ids, flatten_ids, isect_offsets, bit for bit — but without
// streaming 47-bit keys through HBM six times. The (camera, tile) part of every key
// is known when the intersection is emitted, so:
//   1. bin_count accumulates a per-(camera,tile) histogram,
//   2. tile_scan turns it into isect_offsets (the histogram's exclusive scan IS the
//      offsets array) and reports n_isects / the largest segment,
//   3. bin_scatter writes each intersection into its tile's segment as an 8-byte
//      {depth bits, flatten id} record,
//   4. tile_sort: one CTA per tile loads its segment into shared memory, radix-sorts
//      it on the depth bits that actually vary inside the tile (LSD, 8-bit digits,
//      ballot-match ranking, per-warp counters), orders ties by flatten id, and
//      writes the 12 B/intersection gsplat layout with coalesced stores.
// HBM traffic: 8 B/I written + 8 B/I read + 12 B/I written = 28 B/I (vs 164 B/I for
// emit + 6-pass onesweep + histogram). Equal keys: a stable sort of the emission
// order (flatten id ascending) puts equal (tile, depth) keys in flatten-id order, so
// ordering ties by flatten id reproduces it exactly, whatever order the atomics
// scattered the records in.
//
// Atomics: neighbouring Gaussians hit the same few tiles at the same time, and
// same-address global atomics from many CTAs serialise in L2 (measured: +2 ms for
// 47.6M plain REDs). So count and scatter privatise per CTA in shared memory (a dense
// window of tiles around the CTA's bounding box, or a 512-slot hash table when the
// input is scattered), which aggregates the CTA's ~770 intersections into ~40 global
// atomics (one per distinct tile), and CTAs are visited in a strided order so that
// concurrently resident CTAs work on far-apart Gaussians.
#include <stdlib.h>

#include "binning_common.cuh"

namespace pm {


// SCATTER=false: histogram only. SCATTER=true: reserve ranges through `global` (= cursors) and write records.
//
// Per-CTA privatisation in two flavours, chosen per CTA (block-uniform):
//   dense  — the tiles touched by the CTA's 256 Gaussians fit a window of kWinCells tiles around their
//            bounding box (always true for pixel-aligned input: half an image row lands on ~50 tiles):
//            one native shared-memory integer atomic per intersection, no probing;
//   hashed — otherwise (scattered input on a large image): the 512-slot hash table.
template <bool SCATTER>
__global__ void __launch_bounds__(256)
bin_kernel(const float *__restrict__ means2d, const int32_t *__restrict__ radii,
           const float *__restrict__ depths, const int32_t *__restrict__ offsets,
           int32_t *__restrict__ global, int N, int tile_width, int tile_height, int chunk_stride,
           uint64_t *__restrict__ records, int64_t capacity, const float *__restrict__ conics,
           const float *__restrict__ opacities, int views_per_scene) {
  __shared__ TileHash h;
  __shared__ int win[kWinCells];
  __shared__ int win_base[SCATTER ? kWinCells : 1];
  __shared__ int box[4];  // min x, min y, max x, max y over the CTA's tile AABBs
  if (threadIdx.x == 0) { box[0] = 0x7fffffff; box[1] = 0x7fffffff; box[2] = 0; box[3] = 0; }
  __syncthreads();
  // strided CTA order: consecutive (= concurrently resident) CTAs touch far-apart chunks
  const int chunk = strided_chunk(blockIdx.x, chunk_stride, gridDim.x);
  const int g = chunk * blockDim.x + threadIdx.x;
  const int cid = blockIdx.y;
  const size_t f = (size_t)cid * N + (g < N ? g : 0);
  int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
  uint64_t rec = 0;
  if (g < N) {
    const int r = radii[f];
    if (r > 0) {
      const float2 m = reinterpret_cast<const float2 *>(means2d)[f];
      tile_aabb(m.x, m.y, r, tile_width, tile_height, x0, x1, y0, y1);
      // only the tiles the splat can be seen in (common.cuh: usable_tiles) — the rule of the histogram this pass fills
      if (opacities != nullptr)
        usable_tiles(m.x, m.y, conics[f * 3 + 0], conics[f * 3 + 1], conics[f * 3 + 2],
                     opacities[(size_t)(cid / views_per_scene) * N + g], x0, x1, y0, y1);
      if (SCATTER) rec = ((uint64_t)(uint32_t)f << 32) | (uint64_t)__float_as_uint(depths[f]);
    }
  }
  const bool nonempty = x1 > x0 && y1 > y0;
  {
    const int bx0 = __reduce_min_sync(0xffffffffu, nonempty ? x0 : 0x7fffffff);
    const int by0 = __reduce_min_sync(0xffffffffu, nonempty ? y0 : 0x7fffffff);
    const int bx1 = __reduce_max_sync(0xffffffffu, nonempty ? x1 : 0);
    const int by1 = __reduce_max_sync(0xffffffffu, nonempty ? y1 : 0);
    if ((threadIdx.x & 31) == 0 && bx1 > 0) {
      atomicMin(&box[0], bx0); atomicMin(&box[1], by0); atomicMax(&box[2], bx1); atomicMax(&box[3], by1);
    }
  }
  __syncthreads();
  const int wx0 = box[0], wy0 = box[1];
  const int ww = box[2] - wx0, wh = box[3] - wy0;
  if (box[2] == 0 || ww <= 0 || wh <= 0) return;  // nothing visible in this CTA
  const int n_tiles = tile_width * tile_height;
  int32_t *gl = global + (size_t)cid * n_tiles;
  const int32_t *off = offsets + (size_t)cid * n_tiles;
  const bool dense = ww * wh <= kWinCells;

  if (dense) {
    const int cells = ww * wh;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) win[i] = 0;
    __syncthreads();
    for_each_tile(x0, x1, y0, y1, [&](int x, int y) { atomicAdd(&win[(y - wy0) * ww + (x - wx0)], 1); });
    __syncthreads();
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
      const int c = win[i];
      if (c > 0) {
        const int t = (wy0 + i / ww) * tile_width + wx0 + i % ww;
        const int b = atomicAdd(gl + t, c);
        if (SCATTER) { win_base[i] = off[t] + b; win[i] = 0; }  // absolute slot of the CTA's first record in this tile
      }
    }
    if (!SCATTER) return;
    __syncthreads();
    for_each_tile(x0, x1, y0, y1, [&](int x, int y) {
      const int cell = (y - wy0) * ww + (x - wx0);
      const int64_t slot = (int64_t)win_base[cell] + atomicAdd(&win[cell], 1);
      if (slot < capacity) records[slot] = rec;
    });
    return;
  }

  for (int i = threadIdx.x; i < kHashSlots; i += blockDim.x) { h.key[i] = -1; h.cnt[i] = 0; }
  __syncthreads();
  // phase 1: count per distinct tile in shared memory
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      const int t = y * tile_width + x;
      const int s = hash_insert(h, t);
      if (s >= 0) atomicAdd(&h.cnt[s], 1);
      else if (!SCATTER) atomicAdd(gl + t, 1);  // table full (huge splats): straight to global
    }
  __syncthreads();
  // phase 2: one global atomic per distinct tile
  for (int i = threadIdx.x; i < kHashSlots; i += blockDim.x) {
    const int t = h.key[i];
    if (t >= 0) {
      const int c = h.cnt[i];
      const int b = atomicAdd(gl + t, c);
      if (SCATTER) { h.base[i] = b; h.cnt[i] = 0; }
    }
  }
  if (!SCATTER) return;
  __syncthreads();
  // phase 3: place the records (rank inside the CTA's reservation from a second shared counter)
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      const int t = y * tile_width + x;
      const int s = hash_find(h, t);
      const int pos = s >= 0 ? h.base[s] + atomicAdd(&h.cnt[s], 1) : atomicAdd(gl + t, 1);
      const int64_t slot = (int64_t)off[t] + pos;
      if (slot < capacity) records[slot] = rec;
    }
}

// ---- exclusive scan over the per-tile histogram (single CTA; T is at most a few 100k) ----
__global__ void __launch_bounds__(1024)
tile_scan_kernel(const int32_t *__restrict__ counts, int64_t T, int32_t *__restrict__ offsets,
                 int32_t *__restrict__ stats, int32_t *__restrict__ order) {
  __shared__ int sm[32];
  __shared__ int sm_max[32];
  __shared__ int bins[1024 + 1];
  __shared__ int s_vmax;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int carry = 0, vmax = 0;
  for (int64_t base = 0; base < T; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int v = i < T ? counts[i] : 0;
    vmax = max(vmax, v);
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int s = sm[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
      }
      sm[lane] = s;
    }
    __syncthreads();
    const int wbase = warp > 0 ? sm[warp - 1] : 0;
    if (i < T) offsets[i] = carry + wbase + inc - v;
    carry += sm[31];
    __syncthreads();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  if (lane == 0) sm_max[warp] = vmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    int m = 0;
    for (int w = 0; w < 32; ++w) m = max(m, sm_max[w]);
    stats[0] = carry;
    stats[1] = m;
    s_vmax = m;
  }
  if (order == nullptr) return;
  // Launch order for the per-tile kernels: heaviest tiles first (counting sort over 1024 count classes). A tile's
  // render time grows with its list length (max ~4x the median on the benchmark scenes); in index order the long
  // tiles that happen to start last leave the other SMs empty for a quarter of the render kernel.
  for (int i = threadIdx.x; i <= 1024; i += 1024) bins[i] = 0;
  __syncthreads();
  const int bits = 32 - __clz(s_vmax | 1);
  const int shift = bits > 10 ? bits - 10 : 0;
  for (int64_t i = threadIdx.x; i < T; i += 1024) atomicAdd(&bins[1023 - min(1023, counts[i] >> shift) + 1], 1);
  __syncthreads();
  {  // exclusive scan of the 1024 classes (bins[c + 1] holds the count of class c)
    const int v = bins[threadIdx.x + 1];
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int t2 = sm[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, t2, o);
        if (lane >= o) t2 += t;
      }
      sm[lane] = t2;
    }
    __syncthreads();
    const int start = (warp > 0 ? sm[warp - 1] : 0) + inc - v;
    __syncthreads();
    bins[threadIdx.x] = start;  // class c starts at bins[c]
    __syncthreads();
  }
  for (int64_t i = threadIdx.x; i < T; i += 1024)
    order[atomicAdd(&bins[1023 - min(1023, counts[i] >> shift)], 1)] = (int32_t)i;
}

// ---- per-tile shared-memory sort ----
constexpr int kTsRadix = 256;
constexpr int kTsThreads = 512;
constexpr int kTsItems = 16;
constexpr int kTsWarps = kTsThreads / 32;
constexpr int kTsCap = kTsThreads * kTsItems;  // 8192 entries sorted in one go
constexpr size_t kTsSmemBytes = sizeof(uint2) * kTsCap + sizeof(uint16_t) * kTsWarps * kTsRadix +
                                sizeof(uint32_t) * (kTsRadix + 64);

struct TsSmem {
  uint2 *buf;                         // [kTsCap] {key, local index | flatten id}
  uint16_t (*warp_hist)[kTsRadix];    // [warps][256]
  uint32_t *digit_start;              // [256]
  uint32_t *red;                      // [64]
};
__device__ __forceinline__ TsSmem carve_ts_smem(unsigned char *base) {
  TsSmem m;
  m.buf = reinterpret_cast<uint2 *>(base);
  m.warp_hist = reinterpret_cast<uint16_t(*)[kTsRadix]>(m.buf + kTsCap);
  m.digit_start = reinterpret_cast<uint32_t *>(m.warp_hist + kTsWarps);
  m.red = m.digit_start + kTsRadix;
  return m;
}

// 8-bit digit match with ballots: mask of the lanes holding the same digit.
__device__ __forceinline__ unsigned match_digit8(uint32_t d) {
  unsigned peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const unsigned m = __ballot_sync(0xffffffffu, (d >> b) & 1u);
    peers &= ((d >> b) & 1u) ? m : ~m;
  }
  return peers;
}

// Sort one segment of cnt <= kTsCap {depth bits, flatten id} records by (depth bits, flatten id) in
// shared memory and write it out in the gsplat layout. Called by all kTsThreads threads of the CTA.
// Registers hold, per item, the tile-relative depth key and a packed (rank | local index); the flatten
// id stays in global memory (L2-resident: the segment was just read) and is fetched once at the end.
__device__ __forceinline__ void sort_segment(const uint64_t *src, int cnt, int64_t key_hi, const TsSmem &sm,
                                             int64_t *__restrict__ out_ids, int32_t *__restrict__ out_flat) {
  constexpr int ITEMS = kTsItems;
  uint2 *buf = sm.buf;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // load, warp-striped inside a per-warp contiguous chunk: slot(w, k, l) = w*ITEMS*32 + k*32 + l
  uint32_t dk[ITEMS], pk[ITEMS];  // pk: low 16 bits = local index in the unsorted segment, high 16 = rank scratch
  const int chunk0 = warp * (ITEMS * 32);
  uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int slot = chunk0 + k * 32 + lane;
    if (slot < cnt) {
      dk[k] = (uint32_t)src[slot];
      kmin = min(kmin, dk[k]);
      kmax = max(kmax, dk[k]);
    } else {
      dk[k] = 0xffffffffu;
    }
    pk[k] = (uint32_t)slot;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  __syncthreads();  // previous use of the shared buffers (a previous group) is over
  if (lane == 0) { sm.red[warp] = kmin; sm.red[32 + warp] = kmax; }
  __syncthreads();
  kmin = lane < kTsWarps ? sm.red[lane] : 0xffffffffu;
  kmax = lane < kTsWarps ? sm.red[32 + lane] : 0u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  const uint32_t range = kmax - kmin;  // only the bits of (key - min) that vary need sorting
  // Work on tile-relative keys. Empty slots become all-ones: digit 255 in every pass, and since they
  // start behind every real element a stable sort keeps them there — no validity test in the ranking.
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) dk[k] = (chunk0 + k * 32 + lane < cnt) ? dk[k] - kmin : 0xffffffffu;

  for (int shift = 0; shift < 32 && (range >> shift) != 0; shift += 8) {
    for (int i = tid; i < kTsWarps * kTsRadix / 2; i += kTsThreads) reinterpret_cast<uint32_t *>(sm.warp_hist)[i] = 0;
    __syncthreads();
    // rank inside the warp's chunk (stable: item order is (k, lane))
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const uint32_t d = (dk[k] >> shift) & 0xffu;
      const unsigned peers = match_digit8(d);
      const unsigned lower = peers & ((1u << lane) - 1);
      uint32_t base = 0;
      if (lower == 0) {
        base = sm.warp_hist[warp][d];
        sm.warp_hist[warp][d] = (uint16_t)(base + __popc(peers));
      }
      base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
      pk[k] = (pk[k] & 0xffffu) | ((base + __popc(lower)) << 16);
      __syncwarp();
    }
    __syncthreads();
    // digit d (thread d): exclusive offsets across warps, then exclusive scan over the 256 digits
    uint32_t count = 0;
    if (tid < kTsRadix) {
#pragma unroll
      for (int w = 0; w < kTsWarps; ++w) {
        const uint32_t c = sm.warp_hist[w][tid];
        sm.warp_hist[w][tid] = (uint16_t)count;
        count += c;
      }
      uint32_t inc = count;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) sm.red[warp] = inc;
      sm.digit_start[tid] = inc - count;
    }
    __syncthreads();
    if (tid < kTsRadix) {
      uint32_t wbase = 0;
      for (int w = 0; w < warp; ++w) wbase += sm.red[w];
      sm.digit_start[tid] += wbase;
    }
    __syncthreads();
    // scatter to the sorted slot, then read back in chunk order
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const uint32_t d = (dk[k] >> shift) & 0xffu;
      buf[sm.digit_start[d] + sm.warp_hist[warp][d] + (pk[k] >> 16)] = make_uint2(dk[k], pk[k] & 0xffffu);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const uint2 v = buf[chunk0 + k * 32 + lane];
      dk[k] = v.x;
      pk[k] = v.y;
    }
    __syncthreads();
  }

  // fetch the flatten ids of the sorted order and stage {key, flatten id} in shared memory
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int slot = chunk0 + k * 32 + lane;
    if (slot < cnt) buf[slot] = make_uint2(dk[k] + kmin, (uint32_t)(src[pk[k] & 0xffffu] >> 32));
  }
  __syncthreads();
  // ties on the depth bits: order each run by flatten id (runs are rare and short; the thread that
  // owns the first element of a run insertion-sorts it)
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int slot = chunk0 + k * 32 + lane;
    if (slot < cnt - 1) {
      const uint32_t key = buf[slot].x;
      const bool run_start = (slot == 0 || buf[slot - 1].x != key) && buf[slot + 1].x == key;
      if (run_start) {
        int e = slot + 1;
        while (e < cnt && buf[e].x == key) ++e;
        for (int i = slot + 1; i < e; ++i) {
          const uint32_t v = buf[i].y;
          int j = i - 1;
          while (j >= slot && buf[j].y > v) { buf[j + 1].y = buf[j].y; --j; }
          buf[j + 1].y = v;
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < cnt; i += kTsThreads) {
    const uint2 v = buf[i];
    out_ids[i] = key_hi | (int64_t)v.x;
    out_flat[i] = (int32_t)v.y;
  }
}

// Segments of up to kTsCap entries: one sort. (Larger ones are left to tile_sort_big_kernel.)
__global__ void __launch_bounds__(kTsThreads, 2)
tile_sort_kernel(const uint64_t *__restrict__ records, const int32_t *__restrict__ offsets,
                 int64_t n_host, const int32_t *__restrict__ d_n, int64_t total_tiles, int n_tiles,
                 int tile_n_bits, int64_t *__restrict__ isect_ids, int32_t *__restrict__ flatten_ids) {
  extern __shared__ __align__(16) unsigned char ts_smem[];
  const int64_t ct = blockIdx.x;
  const int64_t n_isects = effective_n(n_host, d_n);
  const int start = offsets[ct];
  const int end = (ct == total_tiles - 1) ? (int)n_isects : offsets[ct + 1];
  const int cnt = end - start;
  if (cnt <= 0 || cnt > kTsCap) return;
  const int64_t cam = ct / n_tiles, tile = ct % n_tiles;
  const int64_t key_hi = (cam << (32 + tile_n_bits)) | (tile << 32);
  sort_segment(records + start, cnt, key_hi, carve_ts_smem(ts_smem), isect_ids + start, flatten_ids + start);
}

// Segments larger than kTsCap: the CTA first splits the segment by the top 8 varying bits of the depth
// key (a counting sort through the global `scratch` buffer), which leaves up to 256 depth-ordered
// sub-buckets; consecutive sub-buckets are then grouped up to kTsCap entries and each group is sorted in
// shared memory as above. Concatenating the groups gives the sorted segment. Only a single sub-bucket
// larger than kTsCap (thousands of splats with near-identical depth in one tile) defeats it: the
// overflow flag is raised and the caller falls back to the onesweep path.
__global__ void __launch_bounds__(kTsThreads, 2)
tile_sort_big_kernel(const uint64_t *__restrict__ records, uint64_t *__restrict__ scratch,
                     const int32_t *__restrict__ offsets, int64_t n_host, const int32_t *__restrict__ d_n,
                     int64_t total_tiles, int n_tiles, int tile_n_bits, int64_t *__restrict__ isect_ids,
                     int32_t *__restrict__ flatten_ids, int32_t *__restrict__ overflow_flag) {
  extern __shared__ __align__(16) unsigned char ts_smem[];
  __shared__ int s_hist[kTsRadix + 1];
  __shared__ int s_cursor[kTsRadix];
  __shared__ uint32_t s_mm[2];
  const int tid = threadIdx.x, lane = tid & 31;
  const int64_t ct = blockIdx.x;
  const int64_t n_isects = effective_n(n_host, d_n);
  const int start = offsets[ct];
  const int end = (ct == total_tiles - 1) ? (int)n_isects : offsets[ct + 1];
  const int cnt = end - start;
  if (cnt <= kTsCap) return;
  const int64_t cam = ct / n_tiles, tile = ct % n_tiles;
  const int64_t key_hi = (cam << (32 + tile_n_bits)) | (tile << 32);
  const uint64_t *src = records + start;
  uint64_t *tmp = scratch + start;

  if (tid == 0) { s_mm[0] = 0xffffffffu; s_mm[1] = 0u; }
  for (int i = tid; i <= kTsRadix; i += kTsThreads) s_hist[i] = 0;
  __syncthreads();
  uint32_t kmin = 0xffffffffu, kmax = 0u;
  for (int i = tid; i < cnt; i += kTsThreads) {
    const uint32_t k = (uint32_t)src[i];
    kmin = min(kmin, k);
    kmax = max(kmax, k);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
  }
  if (lane == 0) { atomicMin(&s_mm[0], kmin); atomicMax(&s_mm[1], kmax); }
  __syncthreads();
  kmin = s_mm[0];
  const uint32_t range = s_mm[1] - kmin;
  const int bits = 32 - __clz(range | 1u);
  const int shift = bits > 8 ? bits - 8 : 0;
  for (int i = tid; i < cnt; i += kTsThreads) atomicAdd(&s_hist[(((uint32_t)src[i] - kmin) >> shift) + 1], 1);
  __syncthreads();
  if (tid == 0) {
    for (int d = 0; d < kTsRadix; ++d) s_hist[d + 1] += s_hist[d];  // s_hist[d] = start of sub-bucket d
  }
  __syncthreads();
  for (int i = tid; i < kTsRadix; i += kTsThreads) s_cursor[i] = s_hist[i];
  __syncthreads();
  for (int i = tid; i < cnt; i += kTsThreads) {
    const uint64_t r = src[i];
    const int pos = atomicAdd(&s_cursor[((uint32_t)r - kmin) >> shift], 1);
    tmp[pos] = r;
  }
  __threadfence_block();
  __syncthreads();
  const TsSmem sm = carve_ts_smem(ts_smem);
  int g0 = 0;
  while (g0 < kTsRadix) {
    int g1 = g0 + 1;
    while (g1 < kTsRadix && s_hist[g1 + 1] - s_hist[g0] <= kTsCap) ++g1;
    const int gs = s_hist[g0], ge = s_hist[g1];
    if (ge - gs > kTsCap) {  // a single sub-bucket that does not fit
      if (tid == 0) atomicExch(overflow_flag, 1);
      return;
    }
    if (ge > gs)
      sort_segment(tmp + gs, ge - gs, key_hi, sm, isect_ids + start + gs, flatten_ids + start + gs);
    g0 = g1;
  }
}


}  // namespace pm

extern "C" {

int pm_bin_count(const float *means2d, const int32_t *radii, int N, int C, int tile_size,
                 int tile_width, int tile_height, int32_t *tile_counts, pm_stream_t stream) {
  if (N < 0 || C < 0 || tile_width <= 0 || tile_height <= 0) return PM_ERR_BAD_ARG;
  if (tile_size != pm::kTile) return PM_ERR_UNSUPPORTED;
  if (N == 0 || C == 0) return PM_OK;
  if (!means2d || !radii || !tile_counts) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(means2d) & 7) != 0) return PM_ERR_BAD_ARG;
  if (C > 65535 || (int64_t)N * C >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  const int nb = (N + 255) / 256;
  dim3 grid(nb, C);
  pm::bin_kernel<false><<<grid, 256, 0, pm::as_stream(stream)>>>(
      means2d, radii, nullptr, nullptr, tile_counts, N, tile_width, tile_height, pm::coprime_stride(nb), nullptr, 0,
      nullptr, nullptr, 1);
  return pm::check_launch();
}

int pm_tile_scan(const int32_t *tile_counts, int64_t T, int32_t *isect_offsets, int32_t *stats_out,
                 int32_t *tile_order, pm_stream_t stream) {
  if (T < 0 || !stats_out) return PM_ERR_BAD_ARG;
  if (T > 0 && (!tile_counts || !isect_offsets)) return PM_ERR_BAD_ARG;
  pm::tile_scan_kernel<<<1, 1024, 0, pm::as_stream(stream)>>>(tile_counts, T, isect_offsets, stats_out, tile_order);
  return pm::check_launch();
}

int pm_bin_scatter(const float *means2d, const int32_t *radii, const float *depths,
                   const int32_t *isect_offsets, int32_t *cursors, int N, int C, int tile_size,
                   int tile_width, int tile_height, uint64_t *records, int64_t capacity,
                   const float *conics, const float *opacities, int views_per_scene, pm_stream_t stream) {
  if (N < 0 || C < 0 || tile_width <= 0 || tile_height <= 0 || capacity < 0) return PM_ERR_BAD_ARG;
  if ((conics == nullptr) != (opacities == nullptr)) return PM_ERR_BAD_ARG;
  if (views_per_scene < 0 || (views_per_scene > 0 && C % views_per_scene != 0)) return PM_ERR_BAD_ARG;
  if (tile_size != pm::kTile) return PM_ERR_UNSUPPORTED;
  if (N == 0 || C == 0 || capacity == 0) return PM_OK;
  if (!means2d || !radii || !depths || !isect_offsets || !cursors || !records) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(means2d) & 7) != 0 || (reinterpret_cast<uintptr_t>(records) & 7) != 0)
    return PM_ERR_BAD_ARG;
  if (C > 65535 || (int64_t)N * C >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  const int nb = (N + 255) / 256;
  dim3 grid(nb, C);
  pm::bin_kernel<true><<<grid, 256, 0, pm::as_stream(stream)>>>(
      means2d, radii, depths, isect_offsets, cursors, N, tile_width, tile_height, pm::coprime_stride(nb), records, capacity,
      conics, opacities, views_per_scene > 0 ? views_per_scene : (C > 0 ? C : 1));
  return pm::check_launch();
}

int pm_tile_sort_capacity(void) { return pm::kTsCap; }

int pm_tile_sort(const uint64_t *records, uint64_t *scratch, const int32_t *isect_offsets, int64_t n,
                 const int32_t *d_n, int C, int tile_width, int tile_height, int max_tile_count,
                 int64_t *isect_ids, int32_t *flatten_ids, int32_t *overflow_flag, pm_stream_t stream) {
  if (n < 0 || C < 0 || tile_width <= 0 || tile_height <= 0 || !overflow_flag) return PM_ERR_BAD_ARG;
  if (n == 0 || C == 0) return PM_OK;
  if (!records || !isect_offsets || !isect_ids || !flatten_ids) return PM_ERR_BAD_ARG;
  const bool need_big = max_tile_count < 0 || max_tile_count > pm::kTsCap;
  if (need_big && !scratch) return PM_ERR_BAD_ARG;
  if (n >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  const int n_tiles = tile_width * tile_height;
  const int64_t total = (int64_t)C * n_tiles;
  if (total >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  const int tile_n_bits = pm::n_bits_for(n_tiles);
  if (tile_n_bits + pm::n_bits_for(C) > 32) return PM_ERR_RANGE;
  // per device and cheap: set on every call (a process may drive several GPUs)
  if (cudaFuncSetAttribute(pm::tile_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)pm::kTsSmemBytes) != cudaSuccess ||
      cudaFuncSetAttribute(pm::tile_sort_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)pm::kTsSmemBytes) != cudaSuccess) {
    pm::set_last_cuda_error(cudaGetLastError());
    return PM_ERR_CUDA;
  }
  cudaStream_t s = pm::as_stream(stream);
  pm::tile_sort_kernel<<<(unsigned)total, pm::kTsThreads, pm::kTsSmemBytes, s>>>(
      records, isect_offsets, n, d_n, total, n_tiles, tile_n_bits, isect_ids, flatten_ids);
  if (need_big)  // max_tile_count < 0 means "unknown" (no host sync): always launch it
    pm::tile_sort_big_kernel<<<(unsigned)total, pm::kTsThreads, pm::kTsSmemBytes, s>>>(
        records, scratch, isect_offsets, n, d_n, total, n_tiles, tile_n_bits, isect_ids, flatten_ids, overflow_flag);
  return pm::check_launch();
}

}  // extern "C"
