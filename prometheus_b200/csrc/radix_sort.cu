// radix_sort.cu — stage 3b: hand-written onesweep LSD radix sort (no CUB) of
// (int64 key, int32 value) pairs on key bits [0, end_bit).
//
// Replaces cub::DeviceRadixSort::SortPairs as called by gsplat 1.4.0's
// isect_tiles (reached from prometheus/modules/renderers/gaussians_renderer.py:
// 413-427). The sort must be STABLE: Gaussians with equal (tile, depth-bits) keep
// their emission order, which is what makes the sorted flatten_ids — and hence
// the composited image — well defined and bit-comparable with the oracle.
//
// Algorithm (Adinets & Merrill's "onesweep"): one upfront kernel builds the
// digit histograms of every pass; then each 8-bit pass is ONE kernel in which a
// block (a) ranks its 4096-key tile with warp-level match + per-warp digit
// counters, (b) obtains its tile's global digit offsets through a decoupled
// look-back over the previous tiles' published counts (chained-scan), and
// (c) reorders through shared memory so each digit's run is written with
// coalesced stores. Per pass every key/value is read once and written once:
// 24 B/pair/pass + 8 B/pair for the histogram, the bandwidth floor of an LSD sort.
//
// Forward-progress note: tiles are handed out through an atomic ticket in launch
// order, so the predecessor a block spins on is always resident or finished.
// The spin is bounded; on timeout a flag is raised instead of hanging the GPU.
#include "common.cuh"

namespace pm {

constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 pairs per block
constexpr int kRadix = 256;
constexpr int kMaxPasses = 8;

constexpr uint32_t kFlagAggregate = 1u << 30;
constexpr uint32_t kFlagPrefix = 2u << 30;
constexpr uint32_t kFlagMask = 3u << 30;
constexpr uint32_t kValueMask = ~kFlagMask;
constexpr uint32_t kSpinLimit = 1u << 26;

// Workspace header (uint32 words): [kMaxPasses][256] histograms, kMaxPasses tile
// tickets, 1 error flag (+ padding), then [passes][tiles][256] look-back status.
constexpr size_t kHdrWords = kMaxPasses * kRadix + 16;

struct SortWorkspace {
  uint32_t *hist;     // [kMaxPasses][256]
  uint32_t *tickets;  // [kMaxPasses]
  uint32_t *error;    // [1]
  uint32_t *status;   // [passes][tiles][256]
};

__host__ __device__ inline SortWorkspace carve(void *ws) {
  SortWorkspace w;
  w.hist = static_cast<uint32_t *>(ws);
  w.tickets = w.hist + kMaxPasses * kRadix;
  w.error = w.tickets + kMaxPasses;
  w.status = w.hist + kHdrWords;
  return w;
}

// ---- upfront histograms of all passes -------------------------------------
__global__ void __launch_bounds__(256)
rs_histogram_kernel(const uint64_t *__restrict__ keys, int64_t n_host, const int32_t *__restrict__ d_n,
                    int num_passes, uint32_t *__restrict__ hist) {
  __shared__ uint32_t sh[kMaxPasses * kRadix];
  for (int i = threadIdx.x; i < num_passes * kRadix; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const int64_t n = effective_n(n_host, d_n);
  const int lane = threadIdx.x & 31;
  // warp-uniform trip count so the warp-synchronous match below is legal
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t first = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x - lane);
  for (int64_t base = first; base < n; base += stride) {
    const int64_t i = base + lane;
    const bool valid = i < n;
    const uint64_t key = valid ? keys[i] : 0;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      for (int p = 0; p < num_passes; ++p) {
        const uint32_t d = (uint32_t)(key >> (8 * p)) & 0xffu;
        // aggregate equal digits inside the warp: one shared atomic per distinct digit
        const unsigned peers = __match_any_sync(vmask, d);
        if ((peers & ((1u << lane) - 1)) == 0) atomicAdd(&sh[p * kRadix + d], (uint32_t)__popc(peers));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < num_passes * kRadix; i += blockDim.x) {
    const uint32_t v = sh[i];
    if (v) atomicAdd(&hist[i], v);
  }
}

// exclusive scan of each pass's 256 bins, in place (one block per pass)
__global__ void __launch_bounds__(kRadix)
rs_scan_hist_kernel(uint32_t *__restrict__ hist) {
  __shared__ uint32_t sm[kRadix];
  uint32_t *h = hist + blockIdx.x * kRadix;
  const uint32_t v = h[threadIdx.x];
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < kRadix; o <<= 1) {
    const uint32_t t = threadIdx.x >= o ? sm[threadIdx.x - o] : 0;
    __syncthreads();
    sm[threadIdx.x] += t;
    __syncthreads();
  }
  h[threadIdx.x] = sm[threadIdx.x] - v;
}

// ---- one onesweep pass -----------------------------------------------------
__global__ void __launch_bounds__(kRsThreads)
rs_onesweep_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                   uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t n_host,
                   const int32_t *__restrict__ d_n, int shift, const uint32_t *__restrict__ hist_base,
                   uint32_t *__restrict__ ticket, volatile uint32_t *__restrict__ status,
                   uint32_t *__restrict__ error_flag) {
  __shared__ uint32_t warp_hist[kRsWarps][kRadix];  // counts, then exclusive offsets across warps
  __shared__ uint32_t digit_start[kRadix];          // first slot of each digit inside the tile
  __shared__ uint32_t digit_global[kRadix];         // global index of slot 0 of each digit, minus digit_start
  __shared__ uint64_t exch[kRsTile];                // reorder buffer (keys, then values)
  __shared__ uint32_t scan_tmp[kRsWarps];
  __shared__ uint32_t tile_id_sh;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n = effective_n(n_host, d_n);
  const int64_t num_tiles = ceil_div64(n, kRsTile);
  if (tid == 0) tile_id_sh = atomicAdd(ticket, 1u);
  for (int i = tid; i < kRsWarps * kRadix; i += kRsThreads) (&warp_hist[0][0])[i] = 0;
  __syncthreads();
  const int64_t tile = tile_id_sh;
  if (tile >= num_tiles) return;
  const int64_t tile_base = tile * kRsTile;
  const int tile_count = (int)((n - tile_base) < kRsTile ? (n - tile_base) : kRsTile);

  // (a) load (warp-striped inside a per-warp contiguous chunk) and rank
  uint64_t key[kRsItems];
  uint32_t rank[kRsItems];
  const int chunk0 = warp * (kRsItems * 32);
#pragma unroll
  for (int k = 0; k < kRsItems; ++k) {
    const int local = chunk0 + k * 32 + lane;
    key[k] = local < tile_count ? keys_in[tile_base + local] : ~0ull;
  }
#pragma unroll
  for (int k = 0; k < kRsItems; ++k) {
    const int local = chunk0 + k * 32 + lane;
    const bool valid = local < tile_count;
    // invalid lanes get label 256 so they only match each other and are never counted
    const uint32_t d = valid ? ((uint32_t)(key[k] >> shift) & 0xffu) : 256u;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const unsigned lower = peers & ((1u << lane) - 1);
    uint32_t base = 0;
    if (valid && lower == 0) {
      base = warp_hist[warp][d];
      warp_hist[warp][d] = base + (uint32_t)__popc(peers);
    }
    base = __shfl_sync(0xffffffffu, base, __ffs(peers) - 1);
    rank[k] = base + (uint32_t)__popc(lower);
    __syncwarp();
  }
  __syncthreads();

  // (b) thread d owns digit d: offsets across warps, tile count, look-back
  uint32_t count = 0;
#pragma unroll
  for (int w = 0; w < kRsWarps; ++w) {
    const uint32_t c = warp_hist[w][tid];
    warp_hist[w][tid] = count;
    count += c;
  }
  volatile uint32_t *my_status = status + (size_t)tile * kRadix + tid;
  uint32_t excl = 0;
  if (tile == 0) {
    *my_status = kFlagPrefix | count;
  } else {
    *my_status = kFlagAggregate | count;
    uint32_t spins = 0;
    for (int64_t p = tile - 1; p >= 0; --p) {
      volatile uint32_t *ps = status + (size_t)p * kRadix + tid;
      uint32_t s = *ps;
      while ((s & kFlagMask) == 0) {
        if (++spins > kSpinLimit) { atomicExch(error_flag, 1u); break; }
        __nanosleep(20);
        s = *ps;
      }
      excl += s & kValueMask;
      if ((s & kFlagMask) != kFlagAggregate) break;  // inclusive prefix found (or timeout)
    }
    *my_status = kFlagPrefix | ((excl + count) & kValueMask);
  }
  // exclusive scan of `count` over the 256 digits -> digit_start
  uint32_t inc = count;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) scan_tmp[warp] = inc;
  __syncthreads();
  uint32_t wbase = 0;
#pragma unroll
  for (int w = 0; w < kRsWarps; ++w) wbase += (w < warp) ? scan_tmp[w] : 0u;
  const uint32_t dstart = wbase + inc - count;
  digit_start[tid] = dstart;
  digit_global[tid] = hist_base[tid] + excl - dstart;
  __syncthreads();

  // (c) reorder keys through shared memory, then store runs coalesced
  uint32_t pos[kRsItems];
#pragma unroll
  for (int k = 0; k < kRsItems; ++k) {
    const int local = chunk0 + k * 32 + lane;
    if (local < tile_count) {
      const uint32_t d = (uint32_t)(key[k] >> shift) & 0xffu;
      pos[k] = digit_start[d] + warp_hist[warp][d] + rank[k];
      exch[pos[k]] = key[k];
    } else {
      pos[k] = 0xffffffffu;
    }
  }
  __syncthreads();
  uint32_t dst[kRsItems];
#pragma unroll
  for (int k = 0; k < kRsItems; ++k) {
    const int slot = k * kRsThreads + tid;
    if (slot < tile_count) {
      const uint64_t kk = exch[slot];
      const uint32_t d = (uint32_t)(kk >> shift) & 0xffu;
      dst[k] = digit_global[d] + (uint32_t)slot;
      keys_out[dst[k]] = kk;
    }
  }
  __syncthreads();
  // values: same permutation, reusing the buffer as 32-bit slots
  uint32_t *exch32 = reinterpret_cast<uint32_t *>(exch);
#pragma unroll
  for (int k = 0; k < kRsItems; ++k) {
    const int local = chunk0 + k * 32 + lane;
    if (local < tile_count) exch32[pos[k]] = vals_in[tile_base + local];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kRsItems; ++k) {
    const int slot = k * kRsThreads + tid;
    if (slot < tile_count) vals_out[dst[k]] = exch32[slot];
  }
}

}  // namespace pm

extern "C" {

size_t pm_sort_workspace_bytes(int64_t n) {
  if (n < 0) return 0;
  const int64_t tiles = pm::ceil_div64(n > 0 ? n : 1, pm::kRsTile);
  return (pm::kHdrWords + (size_t)pm::kMaxPasses * (size_t)tiles * pm::kRadix) * sizeof(uint32_t);
}

int pm_sort_pairs(int64_t *keys0, int32_t *vals0, int64_t *keys1, int32_t *vals1, int64_t n,
                  const int32_t *d_n, int end_bit, void *workspace, size_t workspace_bytes,
                  int *result_buffer, pm_stream_t stream) {
  if (n < 0 || end_bit < 0 || end_bit > 64 || !result_buffer) return PM_ERR_BAD_ARG;
  const int passes = (end_bit + 7) / 8;
  *result_buffer = passes & 1;
  if (n == 0 || passes == 0) {
    *result_buffer = 0;
    return PM_OK;
  }
  if (!keys0 || !vals0 || !keys1 || !vals1 || !workspace) return PM_ERR_BAD_ARG;
  if (n >= ((int64_t)1 << 30)) return PM_ERR_RANGE;
  if (workspace_bytes < pm_sort_workspace_bytes(n)) return PM_ERR_WORKSPACE;
  cudaStream_t s = pm::as_stream(stream);
  const int64_t tiles = pm::ceil_div64(n, pm::kRsTile);
  const size_t used = (pm::kHdrWords + (size_t)passes * (size_t)tiles * pm::kRadix) * sizeof(uint32_t);
  if (cudaMemsetAsync(workspace, 0, used, s) != cudaSuccess) return PM_ERR_CUDA;
  pm::SortWorkspace w = pm::carve(workspace);

  int64_t hist_blocks = pm::ceil_div64(n, 256 * 16);
  if (hist_blocks > 148 * 8) hist_blocks = 148 * 8;
  pm::rs_histogram_kernel<<<(unsigned)hist_blocks, 256, 0, s>>>(
      reinterpret_cast<const uint64_t *>(keys0), n, d_n, passes, w.hist);
  pm::rs_scan_hist_kernel<<<passes, pm::kRadix, 0, s>>>(w.hist);

  uint64_t *kin = reinterpret_cast<uint64_t *>(keys0), *kout = reinterpret_cast<uint64_t *>(keys1);
  uint32_t *vin = reinterpret_cast<uint32_t *>(vals0), *vout = reinterpret_cast<uint32_t *>(vals1);
  for (int p = 0; p < passes; ++p) {
    pm::rs_onesweep_kernel<<<(unsigned)tiles, pm::kRsThreads, 0, s>>>(
        kin, vin, kout, vout, n, d_n, 8 * p, w.hist + p * pm::kRadix, w.tickets + p,
        w.status + (size_t)p * (size_t)tiles * pm::kRadix, w.error);
    uint64_t *tk = kin; kin = kout; kout = tk;
    uint32_t *tv = vin; vin = vout; vout = tv;
  }
  return pm::check_launch();
}

}  // extern "C"
