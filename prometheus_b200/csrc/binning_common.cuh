// binning_common.cuh — per-CTA privatised tile histogram shared by the bin kernels (tile_sort.cu) and the
// projection kernel with the fused count (project.cu). See tile_sort.cu for the design notes.
#pragma once
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

constexpr int kWinCells = 2048;  // dense per-CTA window of tiles (covers e.g. 64 x 32 tiles)

// Visit every tile of the AABB [x0,x1) x [y0,y1). Pixel-aligned splats cover 1-2 tiles per axis (radius ~6 px on a
// 16 px grid: ~3 tiles per Gaussian), so the first 2 x 2 block is straight-line predicated code — the nested loops
// with their one- or two-trip counts were 40 % of bin_scatter's instructions, most of it loop control replayed
// under divergence (profiles/r2h_*) — and only larger splats enter the loop for the rest.
template <typename F>
__device__ __forceinline__ void for_each_tile(int x0, int x1, int y0, int y1, F fn) {
  const int nx = x1 - x0, ny = y1 - y0;
  if (nx <= 0 || ny <= 0) return;
  fn(x0, y0);
  if (nx >= 2) fn(x0 + 1, y0);
  if (ny >= 2) {
    fn(x0, y0 + 1);
    if (nx >= 2) fn(x0 + 1, y0 + 1);
  }
  if (nx > 2 || ny > 2) {
    for (int y = y0; y < y1; ++y)
      for (int x = (y < y0 + 2 ? x0 + 2 : x0); x < x1; ++x) fn(x, y);
  }
}

// Strided CTA order chunk = (block * stride) % n_blocks in 32-bit arithmetic when the product fits (the 64-bit
// remainder is ~100 instructions per thread).
__device__ __forceinline__ int strided_chunk(unsigned block, int stride, unsigned n_blocks) {
  if ((unsigned long long)n_blocks * (unsigned)stride <= 0xffffffffull) return (int)((block * (unsigned)stride) % n_blocks);
  return (int)(((unsigned long long)block * (unsigned)stride) % n_blocks);
}

// Shared-memory state of one CTA-wide histogram round.
struct TileCountSmem {
  TileHash h;
  int win[kWinCells];
  int box[4];  // min x, min y, max x, max y over the CTA's tile AABBs
};

// Add every thread's tile AABB [x0,x1) x [y0,y1) to the camera's per-tile histogram `gl` (global, int32
// [tile_height * tile_width]) with ONE global atomic per distinct tile of the CTA: dense window of tiles
// around the CTA's bounding box when it fits, 512-slot hash table otherwise. Block-collective: every thread
// of the CTA must call it (threads without a splat pass an empty box); safe to call repeatedly.
__device__ __forceinline__ void cta_tile_count(TileCountSmem &s, int x0, int x1, int y0, int y1,
                                               int32_t *__restrict__ gl, int tile_width) {
  __syncthreads();  // previous round fully flushed
  if (threadIdx.x == 0) { s.box[0] = 0x7fffffff; s.box[1] = 0x7fffffff; s.box[2] = 0; s.box[3] = 0; }
  __syncthreads();
  const bool nonempty = x1 > x0 && y1 > y0;
  {
    const int bx0 = __reduce_min_sync(0xffffffffu, nonempty ? x0 : 0x7fffffff);
    const int by0 = __reduce_min_sync(0xffffffffu, nonempty ? y0 : 0x7fffffff);
    const int bx1 = __reduce_max_sync(0xffffffffu, nonempty ? x1 : 0);
    const int by1 = __reduce_max_sync(0xffffffffu, nonempty ? y1 : 0);
    if ((threadIdx.x & 31) == 0 && bx1 > 0) {
      atomicMin(&s.box[0], bx0); atomicMin(&s.box[1], by0); atomicMax(&s.box[2], bx1); atomicMax(&s.box[3], by1);
    }
  }
  __syncthreads();
  const int wx0 = s.box[0], wy0 = s.box[1];
  const int ww = s.box[2] - wx0, wh = s.box[3] - wy0;
  if (s.box[2] == 0 || ww <= 0 || wh <= 0) return;  // nothing visible in this CTA (block-uniform)
  if (ww * wh <= kWinCells) {
    const int cells = ww * wh;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) s.win[i] = 0;
    __syncthreads();
    for_each_tile(x0, x1, y0, y1, [&](int x, int y) { atomicAdd(&s.win[(y - wy0) * ww + (x - wx0)], 1); });
    __syncthreads();
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
      const int c = s.win[i];
      if (c > 0) atomicAdd(gl + (wy0 + i / ww) * tile_width + wx0 + i % ww, c);
    }
    return;
  }
  for (int i = threadIdx.x; i < kHashSlots; i += blockDim.x) { s.h.key[i] = -1; s.h.cnt[i] = 0; }
  __syncthreads();
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      const int t = y * tile_width + x;
      const int slot = hash_insert(s.h, t);
      if (slot >= 0) atomicAdd(&s.h.cnt[slot], 1);
      else atomicAdd(gl + t, 1);  // table full (huge splats): straight to global
    }
  __syncthreads();
  for (int i = threadIdx.x; i < kHashSlots; i += blockDim.x) {
    const int t = s.h.key[i];
    if (t >= 0) atomicAdd(gl + t, s.h.cnt[i]);
  }
}

// ---- single-pass binning (module fast path) -------------------------------------------------------------------
// Count AND place in one visit: every (camera, tile) owns a fixed-capacity region of `tile_cap` records
// (records[(tile_base + t) * tile_cap + rank]), so an intersection can be written the moment its CTA has reserved a
// rank range in the tile — no offsets, hence no second pass over the pairs (pm_bin_scatter re-read 16 B per pair and
// redid the window round: 0.29 ms on C2). The per-tile counts still end up in `gl` (the exclusive scan of them gives
// the compact list positions of the sorted ids / the splat stream), and a tile that receives more than `tile_cap`
// records only loses the surplus: the caller sees max count > tile_cap and renders again with the two-pass path.
struct TileScatterSmem {
  int win[kWinCells];
  int win_base[kWinCells];
  int box[4];
};

__device__ __forceinline__ void cta_tile_count_scatter(TileScatterSmem &s, int x0, int x1, int y0, int y1,
                                                       int32_t *__restrict__ gl, int tile_width, uint64_t rec,
                                                       uint64_t *__restrict__ records, int tile_cap) {
  __syncthreads();  // previous round fully flushed
  if (threadIdx.x == 0) { s.box[0] = 0x7fffffff; s.box[1] = 0x7fffffff; s.box[2] = 0; s.box[3] = 0; }
  __syncthreads();
  const bool nonempty = x1 > x0 && y1 > y0;
  {
    const int bx0 = __reduce_min_sync(0xffffffffu, nonempty ? x0 : 0x7fffffff);
    const int by0 = __reduce_min_sync(0xffffffffu, nonempty ? y0 : 0x7fffffff);
    const int bx1 = __reduce_max_sync(0xffffffffu, nonempty ? x1 : 0);
    const int by1 = __reduce_max_sync(0xffffffffu, nonempty ? y1 : 0);
    if ((threadIdx.x & 31) == 0 && bx1 > 0) {
      atomicMin(&s.box[0], bx0); atomicMin(&s.box[1], by0); atomicMax(&s.box[2], bx1); atomicMax(&s.box[3], by1);
    }
  }
  __syncthreads();
  const int wx0 = s.box[0], wy0 = s.box[1];
  const int ww = s.box[2] - wx0, wh = s.box[3] - wy0;
  if (s.box[2] == 0 || ww <= 0 || wh <= 0) return;  // nothing visible in this CTA (block-uniform)
  if (ww * wh <= kWinCells) {
    const int cells = ww * wh;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) s.win[i] = 0;
    __syncthreads();
    for_each_tile(x0, x1, y0, y1, [&](int x, int y) { atomicAdd(&s.win[(y - wy0) * ww + (x - wx0)], 1); });
    __syncthreads();
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
      const int c = s.win[i];
      if (c > 0) {
        s.win_base[i] = atomicAdd(gl + (wy0 + i / ww) * tile_width + wx0 + i % ww, c);  // the CTA's first rank in the tile
        s.win[i] = 0;
      }
    }
    __syncthreads();
    for_each_tile(x0, x1, y0, y1, [&](int x, int y) {
      const int cell = (y - wy0) * ww + (x - wx0);
      const int rank = s.win_base[cell] + atomicAdd(&s.win[cell], 1);
      if (rank < tile_cap) records[(size_t)(y * tile_width + x) * tile_cap + rank] = rec;
    });
    return;
  }
  // scattered input on a large image: one global atomic per intersection (rare; pixel-aligned input is always dense)
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      const int t = y * tile_width + x;
      const int rank = atomicAdd(gl + t, 1);
      if (rank < tile_cap) records[(size_t)t * tile_cap + rank] = rec;
    }
}

// CTAs are visited in a strided order so that concurrently resident CTAs work on far-apart Gaussians
// (their global atomics then hit different tiles).
inline int coprime_stride(int n) {
  if (n <= 2) return 1;
  int s = (int)(n * 0.6180339887) | 1;
  auto gcd = [](int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; };
  while (gcd(s, n) != 1) s += 2;
  return s % n == 0 ? 1 : s;
}

}  // namespace pm
