// binning.cu — stage 3a/3c: prefix sum of the per-pair tile counts, emission of
// the (tile | depth) sort keys, and the per-tile offsets of the sorted list.
//
// Replaces gsplat 1.4.0 isect_tiles (cumsum + emit) and isect_offset_encode,
// reached by the reference through gsplat.rasterization
// (prometheus/modules/renderers/gaussians_renderer.py:413-427).
// Integer work, HBM-bound: scan 12 B/pair, emit 20 B/pair read + 12 B/isect
// written, offsets 8 B/isect read.
#include "common.cuh"

namespace pm {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048 elements per block

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

// block-wide inclusive scan of one value per thread (blockDim.x multiple of 32, <= 1024)
__device__ __forceinline__ int block_incl_scan(int v, int *smem /*[32]*/, int &block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int inc = warp_incl_scan(v);
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int s = lane < nwarps ? smem[lane] : 0;
    s = warp_incl_scan(s);
    smem[lane] = s;
  }
  __syncthreads();
  const int base = warp > 0 ? smem[warp - 1] : 0;
  block_total = smem[nwarps - 1];
  __syncthreads();
  return inc + base;
}

// phase A: per-block totals
__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const int32_t *__restrict__ in, int64_t P, int32_t *__restrict__ block_sums) {
  __shared__ int sm[32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
    if (i < P) s += in[i];
  }
  int total;
  block_incl_scan(s, sm, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// phase B: one block turns the block totals into exclusive offsets and writes the grand total
__global__ void __launch_bounds__(1024)
scan_blocks_kernel(int32_t *__restrict__ block_sums, int64_t nb, int32_t *__restrict__ total_out) {
  __shared__ int sm[32];
  int carry = 0;
  for (int64_t base = 0; base < nb; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int v = i < nb ? block_sums[i] : 0;
    int total;
    const int inc = block_incl_scan(v, sm, total);
    if (i < nb) block_sums[i] = carry + inc - v;
    carry += total;
  }
  if (threadIdx.x == 0) *total_out = carry;
}

// phase C: per-block scan + block offset -> inclusive cumulative counts
__global__ void __launch_bounds__(kScanThreads)
scan_apply_kernel(const int32_t *__restrict__ in, int64_t P, const int32_t *__restrict__ block_sums,
                  int32_t *__restrict__ cum) {
  __shared__ int sm[32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int64_t i = base + k;
    v[k] = i < P ? in[i] : 0;
    s += v[k];
  }
  int total;
  const int inc = block_incl_scan(s, sm, total);
  int run = block_sums[blockIdx.x] + inc - s;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    run += v[k];
    const int64_t i = base + k;
    if (i < P) cum[i] = run;
  }
}

// One thread per (camera, Gaussian) pair writes its tile intersections.
__global__ void __launch_bounds__(256)
bin_emit_kernel(const float *__restrict__ means2d, const int32_t *__restrict__ radii,
                const float *__restrict__ depths, const int32_t *__restrict__ cum, int N, int64_t P,
                int tile_width, int tile_height, int tile_n_bits, int64_t *__restrict__ isect_ids,
                int32_t *__restrict__ flatten_ids, int64_t capacity) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= P) return;
  const int r = radii[f];
  if (r <= 0) return;
  const float2 m = reinterpret_cast<const float2 *>(means2d)[f];
  int x0, x1, y0, y1;
  tile_aabb(m.x, m.y, r, tile_width, tile_height, x0, x1, y0, y1);
  int64_t cur = f > 0 ? (int64_t)cum[f - 1] : 0;
  const int64_t cid = f / N;
  const int64_t cam_enc = cid << (32 + tile_n_bits);
  const int64_t depth_enc = (int64_t)__float_as_uint(depths[f]);
  for (int y = y0; y < y1; ++y)
    for (int x = x0; x < x1; ++x) {
      if (cur < capacity) {
        const int64_t tile_id = (int64_t)y * tile_width + x;
        isect_ids[cur] = cam_enc | (tile_id << 32) | depth_enc;
        flatten_ids[cur] = (int32_t)f;
      }
      ++cur;
    }
}

// isect_offset_encode: offsets[(cam,tile)] = first sorted index whose (cam,tile) >= it.
__global__ void __launch_bounds__(256)
tile_offsets_kernel(const int64_t *__restrict__ sorted_ids, int64_t n_host, const int32_t *__restrict__ d_n,
                    int n_tiles, int tile_n_bits, int64_t total_tiles, int32_t *__restrict__ offsets) {
  const int64_t n = effective_n(n_host, d_n);
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n == 0) {  // no intersections at all: every tile is empty
    for (int64_t i = idx; i < total_tiles; i += (int64_t)gridDim.x * blockDim.x) offsets[i] = 0;
    return;
  }
  if (idx >= n) return;
  const int64_t tile_mask = ((int64_t)1 << tile_n_bits) - 1;
  const int64_t hi = sorted_ids[idx] >> 32;
  const int64_t id_curr = (hi >> tile_n_bits) * n_tiles + (hi & tile_mask);
  if (idx == 0) {
    for (int64_t i = 0; i <= id_curr && i < total_tiles; ++i) offsets[i] = 0;
  } else {
    const int64_t hp = sorted_ids[idx - 1] >> 32;
    const int64_t id_prev = (hp >> tile_n_bits) * n_tiles + (hp & tile_mask);
    for (int64_t i = id_prev + 1; i <= id_curr && i < total_tiles; ++i) offsets[i] = (int32_t)idx;
  }
  if (idx == n - 1) {
    for (int64_t i = id_curr + 1; i < total_tiles; ++i) offsets[i] = (int32_t)n;
  }
}

}  // namespace pm

extern "C" {

size_t pm_scan_workspace_bytes(int64_t P) {
  if (P < 0) return 0;
  const int64_t nb = pm::ceil_div64(P, pm::kScanTile);
  return (size_t)(nb + 1) * sizeof(int32_t);
}

int pm_scan_tiles(const int32_t *tiles_per_gauss, int64_t P, int32_t *cum, int32_t *n_isects_out,
                  void *workspace, size_t workspace_bytes, pm_stream_t stream) {
  if (P < 0 || !n_isects_out) return PM_ERR_BAD_ARG;
  cudaStream_t s = pm::as_stream(stream);
  if (P == 0) {
    if (cudaMemsetAsync(n_isects_out, 0, sizeof(int32_t), s) != cudaSuccess) return PM_ERR_CUDA;
    return PM_OK;
  }
  if (!tiles_per_gauss || !cum || !workspace) return PM_ERR_BAD_ARG;
  if (P >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  if (workspace_bytes < pm_scan_workspace_bytes(P)) return PM_ERR_WORKSPACE;
  const int64_t nb = pm::ceil_div64(P, pm::kScanTile);
  int32_t *block_sums = static_cast<int32_t *>(workspace);
  pm::scan_reduce_kernel<<<(unsigned)nb, pm::kScanThreads, 0, s>>>(tiles_per_gauss, P, block_sums);
  pm::scan_blocks_kernel<<<1, 1024, 0, s>>>(block_sums, nb, n_isects_out);
  pm::scan_apply_kernel<<<(unsigned)nb, pm::kScanThreads, 0, s>>>(tiles_per_gauss, P, block_sums, cum);
  return pm::check_launch();
}

int pm_bin_emit(const float *means2d, const int32_t *radii, const float *depths, const int32_t *cum,
                int N, int C, int tile_size, int tile_width, int tile_height, int64_t *isect_ids,
                int32_t *flatten_ids, int64_t capacity, pm_stream_t stream) {
  if (N < 0 || C < 0 || tile_width <= 0 || tile_height <= 0 || capacity < 0) return PM_ERR_BAD_ARG;
  if (tile_size != pm::kTile) return PM_ERR_UNSUPPORTED;
  const int64_t P = (int64_t)N * C;
  if (P == 0 || capacity == 0) return PM_OK;
  if (!means2d || !radii || !depths || !cum || !isect_ids || !flatten_ids) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(means2d) & 7) != 0) return PM_ERR_BAD_ARG;
  const int tile_n_bits = pm::n_bits_for(tile_width * tile_height);
  const int cam_n_bits = pm::n_bits_for(C);
  if (tile_n_bits + cam_n_bits > 32) return PM_ERR_RANGE;
  pm::bin_emit_kernel<<<(unsigned)pm::ceil_div64(P, 256), 256, 0, pm::as_stream(stream)>>>(
      means2d, radii, depths, cum, N, P, tile_width, tile_height, tile_n_bits, isect_ids, flatten_ids, capacity);
  return pm::check_launch();
}

int pm_tile_offsets(const int64_t *sorted_isect_ids, int64_t n, const int32_t *d_n, int C,
                    int tile_width, int tile_height, int32_t *isect_offsets, pm_stream_t stream) {
  if (n < 0 || C < 0 || tile_width <= 0 || tile_height <= 0 || !isect_offsets) return PM_ERR_BAD_ARG;
  if (n > 0 && !sorted_isect_ids) return PM_ERR_BAD_ARG;
  const int n_tiles = tile_width * tile_height;
  const int64_t total = (int64_t)C * n_tiles;
  if (total == 0) return PM_OK;
  const int tile_n_bits = pm::n_bits_for(n_tiles);
  int64_t blocks = pm::ceil_div64(n > 0 ? n : 1, 256);
  if (n == 0) blocks = pm::ceil_div64(total, 256);
  pm::tile_offsets_kernel<<<(unsigned)blocks, 256, 0, pm::as_stream(stream)>>>(
      sorted_isect_ids, n, d_n, n_tiles, tile_n_bits, total, isect_offsets);
  return pm::check_launch();
}

}  // extern "C"
