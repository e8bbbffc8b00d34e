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
// 47.6M plain REDs). So count and scatter privatise per CTA: a 512-slot shared-memory
// hash table keyed by tile id aggregates the CTA's ~770 intersections into ~40 global
// atomics (one per distinct tile), and CTAs are visited in a strided order so that
// concurrently resident CTAs work on far-apart Gaussians.
#include <stdlib.h>

#include "common.cuh"

namespace pm {

constexpr int kHashSlots = 512;
constexpr int kHashProbes = 16;

struct TileHash {
  int key[kHashSlots];
  int cnt[kHashSlots];
  int base[kHashSlots];
};

__device__ __forceinline__ unsigned hash_tile(int t) { return ((unsigned)t * 2654435761u) >> (32 - 9); }

// returns the slot of tile t, inserting it if needed; -1 if the probe window is full
__device__ __forceinline__ int hash_insert(TileHash &h, int t) {
  unsigned s = hash_tile(t);
  for (int p = 0; p < kHashProbes; ++p, s = (s + 1) & (kHashSlots - 1)) {
    const int prev = atomicCAS(&h.key[s], -1, t);
    if (prev == -1 || prev == t) return (int)s;
  }
  return -1;
}
__device__ __forceinline__ int hash_find(const TileHash &h, int t) {
  unsigned s = hash_tile(t);
  for (int p = 0; p < kHashProbes; ++p, s = (s + 1) & (kHashSlots - 1)) {
    const int k = h.key[s];
    if (k == t) return (int)s;
    if (k == -1) return -1;
  }
  return -1;
}

// SCATTER=false: histogram only. SCATTER=true: reserve ranges through `global` (= cursors) and write records.
template <bool SCATTER>
__global__ void __launch_bounds__(256)
bin_kernel(const float *__restrict__ means2d, const int32_t *__restrict__ radii,
           const float *__restrict__ depths, const int32_t *__restrict__ offsets,
           int32_t *__restrict__ global, int N, int tile_width, int tile_height, int chunk_stride,
           uint64_t *__restrict__ records, int64_t capacity) {
  __shared__ TileHash h;
  for (int i = threadIdx.x; i < kHashSlots; i += blockDim.x) { h.key[i] = -1; h.cnt[i] = 0; }
  __syncthreads();
  // strided CTA order: consecutive (= concurrently resident) CTAs touch far-apart chunks
  const int chunk = (int)(((int64_t)blockIdx.x * chunk_stride) % gridDim.x);
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
      if (SCATTER) rec = ((uint64_t)(uint32_t)f << 32) | (uint64_t)__float_as_uint(depths[f]);
    }
  }
  const int n_tiles = tile_width * tile_height;
  int32_t *gl = global + (size_t)cid * n_tiles;
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
  const int32_t *off = offsets + (size_t)cid * n_tiles;
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
                 int32_t *__restrict__ stats) {
  __shared__ int sm[32];
  __shared__ int sm_max[32];
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
  }
}

// ---- per-tile shared-memory sort ----
constexpr int kTsRadix = 256;
constexpr int kTsCapSmall = 8192;
constexpr int kTsCapLarge = 16384;

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

template <int THREADS, int ITEMS>
constexpr size_t ts_smem_bytes() {
  return sizeof(uint2) * THREADS * ITEMS + sizeof(uint16_t) * (THREADS / 32) * kTsRadix + sizeof(uint32_t) * (kTsRadix + 64);
}

// Sorts segments with min_cnt < cnt <= kTsThreads*ITEMS; other segments are left to the other variant.
// Registers hold, per item, the depth key and a packed (rank | local index); the flatten id stays in
// global memory (L2-resident: the segment was just read) and is fetched once at the end.
template <int THREADS, int ITEMS, bool HW_MATCH = false>
__global__ void __launch_bounds__(THREADS, (THREADS * ITEMS <= 8192 && THREADS <= 512 ? 2 : 1))
tile_sort_kernel(const uint64_t *__restrict__ records, const int32_t *__restrict__ offsets,
                 int64_t n_host, const int32_t *__restrict__ d_n, int64_t total_tiles, int n_tiles,
                 int tile_n_bits, int min_cnt, int64_t *__restrict__ isect_ids,
                 int32_t *__restrict__ flatten_ids, int32_t *__restrict__ overflow_flag) {
  constexpr int CAP = THREADS * ITEMS;
  constexpr int kTsThreads = THREADS;
  constexpr int kTsWarps = THREADS / 32;
  extern __shared__ __align__(16) unsigned char ts_smem[];
  uint2 *buf = reinterpret_cast<uint2 *>(ts_smem);                                    // [CAP] {key, local index}
  uint16_t(*warp_hist)[kTsRadix] = reinterpret_cast<uint16_t(*)[kTsRadix]>(buf + CAP);  // [warps][256]
  uint32_t *digit_start = reinterpret_cast<uint32_t *>(warp_hist + kTsWarps);         // [256]
  uint32_t *red = digit_start + kTsRadix;                                             // [64]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t ct = blockIdx.x;
  const int64_t n_isects = effective_n(n_host, d_n);
  const int start = offsets[ct];
  const int end = (ct == total_tiles - 1) ? (int)n_isects : offsets[ct + 1];
  const int cnt = end - start;
  if (cnt <= min_cnt) return;
  if (cnt > CAP) {
    if (CAP >= kTsCapLarge && tid == 0) atomicExch(overflow_flag, 1);
    return;
  }
  const int64_t cam = ct / n_tiles, tile = ct % n_tiles;
  const int64_t key_hi = (cam << (32 + tile_n_bits)) | (tile << 32);

  // load, warp-striped inside a per-warp contiguous chunk: slot(w, k, l) = w*ITEMS*32 + k*32 + l
  uint32_t dk[ITEMS], pk[ITEMS];  // pk: low 16 bits = local index in the unsorted segment, high 16 = rank scratch
  const int chunk0 = warp * (ITEMS * 32);
  uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    const int slot = chunk0 + k * 32 + lane;
    if (slot < cnt) {
      dk[k] = (uint32_t)__ldg(reinterpret_cast<const unsigned long long *>(records) + (size_t)start + slot);
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
  if (lane == 0) { red[warp] = kmin; red[32 + warp] = kmax; }
  __syncthreads();
  kmin = lane < kTsWarps ? red[lane] : 0xffffffffu;
  kmax = lane < kTsWarps ? red[32 + lane] : 0u;
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
    for (int i = tid; i < kTsWarps * kTsRadix / 2; i += kTsThreads) reinterpret_cast<uint32_t *>(warp_hist)[i] = 0;
    __syncthreads();
    // rank inside the warp's chunk (stable: item order is (k, lane))
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const uint32_t d = (dk[k] >> shift) & 0xffu;
      const unsigned peers = HW_MATCH ? __match_any_sync(0xffffffffu, d) : match_digit8(d);
      const unsigned lower = peers & ((1u << lane) - 1);
      uint32_t base = 0;
      if (lower == 0) {
        base = warp_hist[warp][d];
        warp_hist[warp][d] = (uint16_t)(base + __popc(peers));
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
        const uint32_t c = warp_hist[w][tid];
        warp_hist[w][tid] = (uint16_t)count;
        count += c;
      }
      uint32_t inc = count;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      if (lane == 31) red[warp] = inc;
      digit_start[tid] = inc - count;
    }
    __syncthreads();
    if (tid < kTsRadix) {
      uint32_t wbase = 0;
      for (int w = 0; w < warp; ++w) wbase += red[w];
      digit_start[tid] += wbase;
    }
    __syncthreads();
    // scatter to the sorted slot, then read back in chunk order
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
      const uint32_t d = (dk[k] >> shift) & 0xffu;
      buf[digit_start[d] + warp_hist[warp][d] + (pk[k] >> 16)] = make_uint2(dk[k], pk[k] & 0xffffu);
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
    if (slot < cnt) {
      const uint32_t fid = (uint32_t)(__ldg(reinterpret_cast<const unsigned long long *>(records) + (size_t)start + (pk[k] & 0xffffu)) >> 32);
      buf[slot] = make_uint2(dk[k] + kmin, fid);
    }
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
    isect_ids[(size_t)start + i] = key_hi | (int64_t)v.x;
    flatten_ids[(size_t)start + i] = (int32_t)v.y;
  }
}

static int coprime_stride(int n) {
  if (n <= 2) return 1;
  int s = (int)(n * 0.6180339887) | 1;
  auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
  while (gcd(s, n) != 1) s += 2;
  return s % n == 0 ? 1 : s;
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
      means2d, radii, nullptr, nullptr, tile_counts, N, tile_width, tile_height, pm::coprime_stride(nb), nullptr, 0);
  return pm::check_launch();
}

int pm_tile_scan(const int32_t *tile_counts, int64_t T, int32_t *isect_offsets, int32_t *stats_out,
                 pm_stream_t stream) {
  if (T < 0 || !stats_out) return PM_ERR_BAD_ARG;
  if (T > 0 && (!tile_counts || !isect_offsets)) return PM_ERR_BAD_ARG;
  pm::tile_scan_kernel<<<1, 1024, 0, pm::as_stream(stream)>>>(tile_counts, T, isect_offsets, stats_out);
  return pm::check_launch();
}

int pm_bin_scatter(const float *means2d, const int32_t *radii, const float *depths,
                   const int32_t *isect_offsets, int32_t *cursors, int N, int C, int tile_size,
                   int tile_width, int tile_height, uint64_t *records, int64_t capacity,
                   pm_stream_t stream) {
  if (N < 0 || C < 0 || tile_width <= 0 || tile_height <= 0 || capacity < 0) return PM_ERR_BAD_ARG;
  if (tile_size != pm::kTile) return PM_ERR_UNSUPPORTED;
  if (N == 0 || C == 0 || capacity == 0) return PM_OK;
  if (!means2d || !radii || !depths || !isect_offsets || !cursors || !records) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(means2d) & 7) != 0 || (reinterpret_cast<uintptr_t>(records) & 7) != 0)
    return PM_ERR_BAD_ARG;
  if (C > 65535 || (int64_t)N * C >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  const int nb = (N + 255) / 256;
  dim3 grid(nb, C);
  pm::bin_kernel<true><<<grid, 256, 0, pm::as_stream(stream)>>>(
      means2d, radii, depths, isect_offsets, cursors, N, tile_width, tile_height, pm::coprime_stride(nb), records, capacity);
  return pm::check_launch();
}

int pm_tile_sort_capacity(void) { return pm::kTsCapLarge; }

int pm_tile_sort(const uint64_t *records, const int32_t *isect_offsets, int64_t n, const int32_t *d_n,
                 int C, int tile_width, int tile_height, int max_tile_count, int64_t *isect_ids,
                 int32_t *flatten_ids, int32_t *overflow_flag, pm_stream_t stream) {
  if (n < 0 || C < 0 || tile_width <= 0 || tile_height <= 0 || !overflow_flag) return PM_ERR_BAD_ARG;
  if (n == 0 || C == 0) return PM_OK;
  if (!records || !isect_offsets || !isect_ids || !flatten_ids) return PM_ERR_BAD_ARG;
  if (n >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  const int n_tiles = tile_width * tile_height;
  const int64_t total = (int64_t)C * n_tiles;
  if (total >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  const int tile_n_bits = pm::n_bits_for(n_tiles);
  if (tile_n_bits + pm::n_bits_for(C) > 32) return PM_ERR_RANGE;
  static bool attr_set = false;
  static int small_variant = 0;
  if (!attr_set) {
    const char *e = getenv("PM_TILE_SORT_VARIANT");  // dev knob: 1 = 512 thr x 16 (default, 2 CTAs/SM), 0 = 1024x8, 2 = 1024x8 + match.any
    small_variant = e ? atoi(e) : 1;
    if (cudaFuncSetAttribute(pm::tile_sort_kernel<1024, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)pm::ts_smem_bytes<1024, 8>()) != cudaSuccess ||
        cudaFuncSetAttribute(pm::tile_sort_kernel<1024, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)pm::ts_smem_bytes<1024, 8>()) != cudaSuccess ||
        cudaFuncSetAttribute(pm::tile_sort_kernel<512, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)pm::ts_smem_bytes<512, 16>()) != cudaSuccess ||
        cudaFuncSetAttribute(pm::tile_sort_kernel<512, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)pm::ts_smem_bytes<512, 32>()) != cudaSuccess) {
      pm::set_last_cuda_error(cudaGetLastError());
      return PM_ERR_CUDA;
    }
    attr_set = true;
  }
  cudaStream_t s = pm::as_stream(stream);
  // segments of up to 8192 entries go to the small variant; the (rare) larger ones, up to 16384, to a
  // second launch. max_tile_count < 0 means "unknown" (no host sync): launch both.
  if (small_variant == 1)
    pm::tile_sort_kernel<512, 16><<<(unsigned)total, 512, pm::ts_smem_bytes<512, 16>(), s>>>(
        records, isect_offsets, n, d_n, total, n_tiles, tile_n_bits, 0, isect_ids, flatten_ids, overflow_flag);
  else if (small_variant == 2)
    pm::tile_sort_kernel<1024, 8, true><<<(unsigned)total, 1024, pm::ts_smem_bytes<1024, 8>(), s>>>(
        records, isect_offsets, n, d_n, total, n_tiles, tile_n_bits, 0, isect_ids, flatten_ids, overflow_flag);
  else
    pm::tile_sort_kernel<1024, 8><<<(unsigned)total, 1024, pm::ts_smem_bytes<1024, 8>(), s>>>(
        records, isect_offsets, n, d_n, total, n_tiles, tile_n_bits, 0, isect_ids, flatten_ids, overflow_flag);
  if (max_tile_count < 0 || max_tile_count > pm::kTsCapSmall)
    pm::tile_sort_kernel<512, 32><<<(unsigned)total, 512, pm::ts_smem_bytes<512, 32>(), s>>>(
        records, isect_offsets, n, d_n, total, n_tiles, tile_n_bits, pm::kTsCapSmall, isect_ids, flatten_ids,
        overflow_flag);
  return pm::check_launch();
}

}  // extern "C"
