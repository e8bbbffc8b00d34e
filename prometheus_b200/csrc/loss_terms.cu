// loss_terms.cu — the loss-side glue of the stage-1 rendering loss in one pass each way.
//
// The reference (prometheus/systems/gs_decoder_system.py:214-239) post-processes the renderer's outputs with a
// dozen full-image ATen kernels before the actual loss networks see them:
//     mask = ones_like(depth); mask[depth <= 0.1] = 0            (:216-217)
//     disp = 1 / depth.clip(0.1, 100)                            (:218)   -> ScaleAndShiftInvariantLoss
//     F.mse_loss(images_pred, images)                            (:221)
//     (1 - masks).mean()                                         (:233)   "entropy" term
// and autograd runs the mirror image of each on the way back. Here one kernel reads the three packed renderer
// outputs (+ the target images) once and produces {sum of squared errors, sum of (1 - mask)} and the two
// per-pixel maps; one kernel turns the four cotangents into (v_images, v_depths, v_masks), the packed cotangents
// pm_composite_bwd_packed consumes. HBM-bound streaming: 8 floats read + 2 written per pixel forward.
#include "common.cuh"

namespace pm {

constexpr float kDepthLo = 0.1f, kDepthHi = 100.0f;

__global__ void __launch_bounds__(256)
loss_terms_fwd_kernel(const float *__restrict__ images, const float *__restrict__ target,
                      const float *__restrict__ depths, const float *__restrict__ masks, int64_t n_views,
                      int64_t plane, float *__restrict__ disparity, float *__restrict__ valid,
                      double *__restrict__ sums) {
  double se = 0.0, ent = 0.0;
  const int64_t n_pix = n_views * plane;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / plane, p = i - v * plane;
    const float d = __ldcs(depths + i);
    disparity[i] = 1.0f / fminf(fmaxf(d, kDepthLo), kDepthHi);
    valid[i] = d <= kDepthLo ? 0.0f : 1.0f;
    ent += (double)(1.0f - __ldcs(masks + i));
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int64_t j = (v * 3 + k) * plane + p;
      const float e = images[j] - __ldcs(target + j);
      acc = fmaf(e, e, acc);
    }
    se += (double)acc;
  }
  // block reduction in fp64 (the order differs from ATen's fp32 tree: compared with a tolerance in the tests)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    ent += __shfl_xor_sync(0xffffffffu, ent, o);
  }
  __shared__ double s_se[8], s_ent[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_se[warp] = se; s_ent[warp] = ent; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { a += s_se[w]; b += s_ent[w]; }
    atomicAdd(sums + 0, a);
    atomicAdd(sums + 1, b);
  }
}

// v_images = g_mse * 2 (x - t) / (3 n_pix)        (F.mse_loss, mean reduction)
// v_depths = -v_disp * disp^2 inside [0.1, 100], 0 outside (clamp passes the gradient on its closed interval)
// v_masks  = -g_entropy / n_pix                   ((1 - m).mean())
__global__ void __launch_bounds__(256)
loss_terms_bwd_kernel(const float *__restrict__ images, const float *__restrict__ target,
                      const float *__restrict__ depths, int64_t n_views, int64_t plane,
                      const float *__restrict__ g_mse, const float *__restrict__ g_entropy,
                      const float *__restrict__ v_disparity, float *__restrict__ v_images,
                      float *__restrict__ v_depths, float *__restrict__ v_masks) {
  const int64_t n_pix = n_views * plane;
  const float gm = g_mse ? __ldg(g_mse) * (2.0f / (float)(3 * n_pix)) : 0.0f;
  const float ge = g_entropy ? -__ldg(g_entropy) / (float)n_pix : 0.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pix; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i / plane, p = i - v * plane;
    const float d = __ldcs(depths + i);
    float vd = 0.0f;
    if (v_disparity && d >= kDepthLo && d <= kDepthHi) {
      const float r = 1.0f / d;
      vd = -__ldcs(v_disparity + i) * r * r;
    }
    v_depths[i] = vd;
    v_masks[i] = ge;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int64_t j = (v * 3 + k) * plane + p;
      v_images[j] = gm * (__ldcs(images + j) - __ldcs(target + j));
    }
  }
}

}  // namespace pm

extern "C" {

int pm_loss_terms_fwd(const float *images, const float *target, const float *depths, const float *masks,
                      int64_t n_views, int height, int width, float *disparity, float *valid, double *sums,
                      pm_stream_t stream) {
  if (!images || !target || !depths || !masks || !disparity || !valid || !sums) return PM_ERR_BAD_ARG;
  if (n_views < 0 || height <= 0 || width <= 0) return PM_ERR_BAD_ARG;
  cudaStream_t st = pm::as_stream(stream);
  if (cudaMemsetAsync(sums, 0, 2 * sizeof(double), st) != cudaSuccess) return pm::check_launch();
  const int64_t plane = (int64_t)height * width, n_pix = n_views * plane;
  if (n_pix == 0) return PM_OK;
  const int grid = (int)(n_pix / 256 + 1 < 148 * 8 ? n_pix / 256 + 1 : 148 * 8);  // persistent: 8 CTAs per SM
  pm::loss_terms_fwd_kernel<<<grid, 256, 0, st>>>(images, target, depths, masks, n_views, plane, disparity, valid, sums);
  return pm::check_launch();
}

int pm_loss_terms_bwd(const float *images, const float *target, const float *depths, int64_t n_views, int height,
                      int width, const float *g_mse, const float *g_entropy, const float *v_disparity,
                      float *v_images, float *v_depths, float *v_masks, pm_stream_t stream) {
  if (!images || !target || !depths || !v_images || !v_depths || !v_masks) return PM_ERR_BAD_ARG;
  if (n_views < 0 || height <= 0 || width <= 0) return PM_ERR_BAD_ARG;
  const int64_t plane = (int64_t)height * width, n_pix = n_views * plane;
  if (n_pix == 0) return PM_OK;
  const int grid = (int)(n_pix / 256 + 1 < 148 * 8 ? n_pix / 256 + 1 : 148 * 8);
  pm::loss_terms_bwd_kernel<<<grid, 256, 0, pm::as_stream(stream)>>>(images, target, depths, n_views, plane, g_mse,
                                                                     g_entropy, v_disparity, v_images, v_depths, v_masks);
  return pm::check_launch();
}

}  // extern "C"
