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
#ifndef PM_TILE_BLOCK
#define PM_TILE_BLOCK 3
#endif
template <typename F>
__device__ __forceinline__ void for_each_tile(int x0, int x1, int y0, int y1, F fn) {
  const int nx = x1 - x0, ny = y1 - y0;
  if (nx <= 0 || ny <= 0) return;
#if PM_TILE_BLOCK == 3
  // 3 x 3 block of predicated calls (no loop control under divergence), the generic loops only beyond it
#pragma unroll
  for (int iy = 0; iy < 3; ++iy)
#pragma unroll
    for (int ix = 0; ix < 3; ++ix)
      if (ix < nx && iy < ny) fn(x0 + ix, y0 + iy);
  if (nx > 3 || ny > 3) {
    for (int y = y0; y < y1; ++y)
      for (int x = (y < y0 + 3 ? x0 + 3 : x0); x < x1; ++x) fn(x, y);
  }
#else
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
#endif
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
// (records[(camera * tiles + t) * tile_cap + rank]), so an intersection can be written the moment its CTA has
// reserved a rank range in the tile — no offsets, hence no second pass over the pairs (pm_bin_scatter re-read 16 B per
// pair and redid the window round: 0.29 ms on C2). The per-tile counts still end up in the histogram (its exclusive
// scan gives the compact list positions of the sorted ids / the splat stream), and a tile that receives more than
// `tile_cap` records only loses the surplus: the caller sees max count > tile_cap and renders again with two passes.
//
// The window round is BATCHED over up to kWinCams cameras: the projection loop leaves each thread's tile AABB and
// depth bits of every camera in shared memory and reduces the CTA's bounding box per camera on the way; then ONE
// set of barrier-separated phases (lay out the windows, zero, count, reserve, place) serves all cameras — 6 barriers
// per 8 cameras instead of 6 per camera (barrier waits were 40 % of the kernel's stall samples, profiles/r2m_*).
constexpr int kWinCams = 8;

struct TileScatterSmem {
  int win[kWinCells];          // per-cell count, then per-cell rank cursor
  int win_base[kWinCells];     // the CTA's first rank in the cell's tile
  uint32_t box[kWinCams][256]; // per (camera, thread): x0 | y0 << 10 | nx << 20 | ny << 26 (0 = no intersection)
  uint32_t dep[kWinCams][256]; // per (camera, thread): depth bits
  int bb[kWinCams][4];         // CTA bounding box per camera: min x, min y, max x, max y
  int off[kWinCams + 1];       // first window cell of each camera
  int ww[kWinCams];            // window width per camera (0 = the camera's window does not fit: direct global atomics)
};

__device__ __forceinline__ uint32_t pack_tile_box(int x0, int x1, int y0, int y1) {
  return (uint32_t)x0 | ((uint32_t)y0 << 10) | ((uint32_t)(x1 - x0) << 20) | ((uint32_t)(y1 - y0) << 26);
}
__device__ __forceinline__ bool tile_box_packable(int x0, int x1, int y0, int y1) {
  return x0 < 1024 && y0 < 1024 && x1 - x0 < 64 && y1 - y0 < 64;
}

// Before the projection loop of a camera batch (block-collective).
__device__ __forceinline__ void scatter_batch_begin(TileScatterSmem &s, int n) {
  __syncthreads();  // the previous batch is fully placed
  if (threadIdx.x < n * 4) s.bb[threadIdx.x >> 2][threadIdx.x & 3] = (threadIdx.x & 2) ? 0 : 0x7fffffff;
  __syncthreads();
}

// Inside the projection loop, camera slot k of the batch: remember the thread's AABB / depth, reduce the bounding box.
// An AABB too large for the packed form (a splat covering > 63 tiles on an axis, or a tile grid beyond 1024) is placed
// at once with one global atomic per tile.
__device__ __forceinline__ void scatter_batch_add(TileScatterSmem &s, int k, int x0, int x1, int y0, int y1, uint32_t depth_bits,
                                                  uint32_t f, int32_t *__restrict__ gl, int tile_width,
                                                  uint64_t *__restrict__ records, int tile_cap) {
  bool nonempty = x1 > x0 && y1 > y0;
  if (nonempty && !tile_box_packable(x0, x1, y0, y1)) {
    const uint64_t rec = ((uint64_t)f << 32) | depth_bits;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) {
        const int t = y * tile_width + x;
        const int rank = atomicAdd(gl + t, 1);
        if (rank < tile_cap) records[(size_t)t * tile_cap + rank] = rec;
      }
    nonempty = false;
  }
  s.box[k][threadIdx.x] = nonempty ? pack_tile_box(x0, x1, y0, y1) : 0u;
  s.dep[k][threadIdx.x] = depth_bits;
  const int bx0 = __reduce_min_sync(0xffffffffu, nonempty ? x0 : 0x7fffffff);
  const int by0 = __reduce_min_sync(0xffffffffu, nonempty ? y0 : 0x7fffffff);
  const int bx1 = __reduce_max_sync(0xffffffffu, nonempty ? x1 : 0);
  const int by1 = __reduce_max_sync(0xffffffffu, nonempty ? y1 : 0);
  if ((threadIdx.x & 31) == 0 && bx1 > 0) {
    atomicMin(&s.bb[k][0], bx0); atomicMin(&s.bb[k][1], by0); atomicMax(&s.bb[k][2], bx1); atomicMax(&s.bb[k][3], by1);
  }
}

// After the projection loop: count, reserve and place the batch's n cameras. gl0 / records0 = histogram / record
// regions of the batch's first camera (consecutive cameras follow at n_tiles / n_tiles * tile_cap); f0 = this
// thread's flatten id for that camera (+ N per camera).
__device__ __forceinline__ void scatter_batch_end(TileScatterSmem &s, int n, int32_t *__restrict__ gl0, int tile_width,
                                                  int n_tiles, uint64_t *__restrict__ records0, int tile_cap, uint32_t f0,
                                                  uint32_t N) {
  __syncthreads();  // every AABB and bounding box is in
  if (threadIdx.x == 0) {
    int o = 0;
    for (int k = 0; k < n; ++k) {
      const int w = s.bb[k][2] - s.bb[k][0], h = s.bb[k][3] - s.bb[k][1];
      const int cells = (s.bb[k][2] > 0 && w > 0 && h > 0) ? w * h : 0;
      s.off[k] = o;
      if (o + cells <= kWinCells) { s.ww[k] = cells > 0 ? w : 1; o += cells; }
      else s.ww[k] = 0;  // does not fit (scattered input on a large image): this camera goes straight to global memory
    }
    s.off[n] = o;
  }
  __syncthreads();
  const int total = s.off[n];
  for (int i = threadIdx.x; i < total; i += blockDim.x) s.win[i] = 0;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const uint32_t b = s.box[k][threadIdx.x];
    if (b == 0u) continue;
    const int x0 = b & 1023, y0 = (b >> 10) & 1023, x1 = x0 + ((b >> 20) & 63), y1 = y0 + (b >> 26);
    const int ww = s.ww[k];
    if (ww > 0) {
      const int base = s.off[k] - s.bb[k][1] * ww - s.bb[k][0];
      for_each_tile(x0, x1, y0, y1, [&](int x, int y) { atomicAdd(&s.win[base + y * ww + x], 1); });
    } else {
      const uint64_t rec = ((uint64_t)(f0 + (uint32_t)k * N) << 32) | s.dep[k][threadIdx.x];
      int32_t *gl = gl0 + (size_t)k * n_tiles;
      uint64_t *rk = records0 + (size_t)k * n_tiles * tile_cap;
      for_each_tile(x0, x1, y0, y1, [&](int x, int y) {
        const int t = y * tile_width + x;
        const int rank = atomicAdd(gl + t, 1);
        if (rank < tile_cap) rk[(size_t)t * tile_cap + rank] = rec;
      });
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int c = s.win[i];
    if (c > 0) {
      int k = 0;
      while (i >= s.off[k + 1]) ++k;
      const int ww = s.ww[k], local = i - s.off[k];
      const int t = (s.bb[k][1] + local / ww) * tile_width + s.bb[k][0] + local % ww;
      s.win_base[i] = atomicAdd(gl0 + (size_t)k * n_tiles + t, c);  // the CTA's first rank in the tile
      s.win[i] = 0;
    }
  }
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const uint32_t b = s.box[k][threadIdx.x];
    const int ww = s.ww[k];
    if (b == 0u || ww == 0) continue;
    const int x0 = b & 1023, y0 = (b >> 10) & 1023, x1 = x0 + ((b >> 20) & 63), y1 = y0 + (b >> 26);
    const int base = s.off[k] - s.bb[k][1] * ww - s.bb[k][0];
    const uint64_t rec = ((uint64_t)(f0 + (uint32_t)k * N) << 32) | s.dep[k][threadIdx.x];
    uint64_t *rk = records0 + (size_t)k * n_tiles * tile_cap;
    for_each_tile(x0, x1, y0, y1, [&](int x, int y) {
      const int cell = base + y * ww + x;
      const int rank = s.win_base[cell] + atomicAdd(&s.win[cell], 1);
      if (rank < tile_cap) rk[(unsigned)((y * tile_width + x) * tile_cap + rank)] = rec;  // < 2^31 per camera (checked at launch)
    });
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
