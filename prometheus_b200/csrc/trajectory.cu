// trajectory.cu — the novel-view trajectory row: pose sampling along a camera path and
// frame packing, both on the device.
//
//  * sample_cameras_kernel replaces prometheus/utils/camera.py:134-217 (sample_from_dense_cameras) together
//    with matrix_to_quaternion (prometheus/utils/math.py:46-91 — a Python loop over matrices in the
//    reference), quaternion_slerp (camera.py:81-109) and quaternion_to_matrix (math.py:18-44). The reference
//    evaluates it once per video frame from the host (export_video, prometheus/utils/io.py:167-190: 120-720
//    calls of ~60 ATen kernels each plus the per-matrix Python loop); here all frames of a trajectory are
//    sampled by one launch. Semantics are those of the reference's only well-defined call pattern (one
//    trajectory, one sample per call), INCLUDING its quirks: the strict `<` segment test sends t = 1 to
//    segment 0, and the shortest-path branch of the slerp takes |dot| but never negates the second quaternion
//    (camera.py:94-95 flips d before using it as the mask).
//  * pack_frames_u8_kernel replaces the per-frame host chain of io.py:181
//    (`.mul(1/2).add(1/2).clamp(0,1).mul(255)` -> `.astype(np.uint8)`, CHW -> HWC): 4x less device->host traffic.
//
// Compiled with -fmad=false (Makefile) so that every operation rounds once, like the fp32 ATen ops it replaces;
// what remains against the reference is the ~1 ulp difference of sinf / acosf.
#include "common.cuh"

namespace pm {

constexpr int kTrajThreads = 256;
constexpr int kTrajMaxPoses = 1024;

// math.py:46-91, wxyz
__device__ __forceinline__ float4 matrix_to_quaternion(const float *M /* 3x4 row-major pose */) {
  const float m00 = M[0], m01 = M[1], m02 = M[2], m10 = M[4], m11 = M[5], m12 = M[6], m20 = M[8], m21 = M[9], m22 = M[10];
  const float tr = ((1.0f + m00) + m11) + m22;
  float r, x, y, z;
  if (tr > 0.0f) {
    r = sqrtf(tr) / 2.0f;
    x = (m21 - m12) / (4.0f * r);
    y = (m02 - m20) / (4.0f * r);
    z = (m10 - m01) / (4.0f * r);
  } else if (m00 > m11 && m00 > m22) {
    const float S = sqrtf(((1.0f + m00) - m11) - m22) * 2.0f;
    r = (m21 - m12) / S; x = 0.25f * S; y = (m01 + m10) / S; z = (m02 + m20) / S;
  } else if (m11 > m22) {
    const float S = sqrtf(((1.0f + m11) - m00) - m22) * 2.0f;
    r = (m02 - m20) / S; x = (m01 + m10) / S; y = 0.25f * S; z = (m12 + m21) / S;
  } else {
    const float S = sqrtf(((1.0f + m22) - m00) - m11) * 2.0f;
    r = (m10 - m01) / S; x = (m02 + m20) / S; y = (m12 + m21) / S; z = 0.25f * S;
  }
  return make_float4(r, x, y, z);
}

__device__ __forceinline__ float dot4(float4 a, float4 b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }

__global__ void __launch_bounds__(kTrajThreads)
sample_cameras_kernel(const float *__restrict__ dense, const float *__restrict__ t, int N, int M, float *__restrict__ out) {
  __shared__ float4 s_q[kTrajMaxPoses];
  __shared__ float s_cum[kTrajMaxPoses];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float *cams = dense + (size_t)b * N * 18;
  for (int n = tid; n < N; n += kTrajThreads) s_q[n] = matrix_to_quaternion(cams + (size_t)n * 18);
  __syncthreads();
  // combined distance between consecutive poses: sqrt(|dp|^2 + angle^2), angle = 2 acos(<q1/|q1|, q2/|q2|>)
  for (int n = tid; n < N - 1; n += kTrajThreads) {
    const float *a = cams + (size_t)n * 18, *c = cams + (size_t)(n + 1) * 18;
    const float dx = c[3] - a[3], dy = c[7] - a[7], dz = c[11] - a[11];
    const float d_pos = sqrtf((dx * dx + dy * dy) + dz * dz);
    const float4 q1 = s_q[n], q2 = s_q[n + 1];
    const float n1 = sqrtf(dot4(q1, q1)), n2 = sqrtf(dot4(q2, q2));
    const float4 u1 = make_float4(q1.x / n1, q1.y / n1, q1.z / n1, q1.w / n1);
    const float4 u2 = make_float4(q2.x / n2, q2.y / n2, q2.z / n2, q2.w / n2);
    const float dot = fminf(fmaxf(dot4(u1, u2), -1.0f), 1.0f);
    const float angle = 2.0f * acosf(dot);
    s_cum[n + 1] = sqrtf(d_pos * d_pos + angle * angle);  // segment length for now
  }
  __syncthreads();
  if (tid == 0) {  // sequential fp32 running sum, as torch.cumsum on the CPU (N is a few hundred at most)
    float acc = 0.0f;
    s_cum[0] = 0.0f;
    for (int n = 1; n < N; ++n) { acc += s_cum[n]; s_cum[n] = acc; }
  }
  __syncthreads();
  const float total = s_cum[N - 1];
  for (int m = tid; m < M; m += kTrajThreads) {
    const float desired = t[(size_t)b * M + m] * total;
    int seg = 0;  // first segment with cum[n] <= desired < cum[n+1]; none (t = 1) -> 0, as argmax of an all-false mask
    for (int n = 0; n < N - 1; ++n)
      if (desired >= s_cum[n] && desired < s_cum[n + 1]) { seg = n; break; }
    const float lower = s_cum[seg], upper = s_cum[seg + 1];
    float frac = (desired - lower) / (upper - lower);
    frac = fminf(fmaxf(frac, 0.0f), 1.0f) + (frac != frac ? frac : 0.0f);  // clamp keeps NaN (0/0 on a repeated pose)
    const float *a = cams + (size_t)seg * 18, *c = cams + (size_t)(seg + 1) * 18;
    const float w0 = 1.0f - frac;
    const float px = w0 * a[3] + frac * c[3], py = w0 * a[7] + frac * c[7], pz = w0 * a[11] + frac * c[11];
    // slerp (camera.py:81-109)
    const float4 q0 = s_q[seg], q1 = s_q[seg + 1];
    float d = dot4(q0, q1);
    if (d < 0.0f) d = -d;  // (q1 is NOT negated: see the header comment)
    d = fminf(fmaxf(d, -1.0f), 1.0f);
    const float angle = acosf(d);
    const float isin = 1.0f / (sinf(angle) + 1e-10f);
    const float s0 = sinf((1.0f - frac) * angle), s1 = sinf(frac * angle);
    float4 q = make_float4((q0.x * s0) * isin + (q1.x * s1) * isin, (q0.y * s0) * isin + (q1.y * s1) * isin,
                           (q0.z * s0) * isin + (q1.z * s1) * isin, (q0.w * s0) * isin + (q1.w * s1) * isin);
    if (angle < 1e-5f) q = q0;
    const float qn = sqrtf(dot4(q, q));
    q = make_float4(q.x / qn, q.y / qn, q.z / qn, q.w / qn);
    // quaternion_to_matrix (math.py:18-44), r i j k = x y z w of the float4
    const float r = q.x, i = q.y, j = q.z, k = q.w;
    const float two_s = 2.0f / dot4(q, q);
    float *o = out + ((size_t)b * M + m) * 18;
    o[0] = 1.0f - two_s * (j * j + k * k); o[1] = two_s * (i * j - k * r);        o[2] = two_s * (i * k + j * r);        o[3] = px;
    o[4] = two_s * (i * j + k * r);        o[5] = 1.0f - two_s * (i * i + k * k); o[6] = two_s * (j * k - i * r);        o[7] = py;
    o[8] = two_s * (i * k - j * r);        o[9] = two_s * (j * k + i * r);        o[10] = 1.0f - two_s * (i * i + j * j); o[11] = pz;
#pragma unroll
    for (int e = 12; e < 18; ++e) o[e] = w0 * a[e] + frac * c[e];
  }
}

// images [F,3,H,W] fp32 in [-1,1] -> uint8 [F,H,W,3]; four pixels (12 bytes) per thread when W % 4 == 0
__device__ __forceinline__ uint32_t to_u8(float x) {
  const float y = __fmul_rn(fminf(fmaxf(__fadd_rn(__fmul_rn(x, 0.5f), 0.5f), 0.0f), 1.0f), 255.0f);
  return (uint32_t)y;  // truncation, as numpy's astype(uint8) on [0, 255]
}

__global__ void __launch_bounds__(256)
pack_frames_u8_kernel(const float *__restrict__ images, int64_t n_frames, int64_t plane, uint8_t *__restrict__ out, int vec4) {
  const int64_t quads = vec4 ? plane / 4 : plane;
  const int64_t total = n_frames * quads;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = idx / quads, p = idx % quads;
    const float *src = images + f * 3 * plane;
    if (vec4) {
      const float4 r = __ldcs(reinterpret_cast<const float4 *>(src) + p);
      const float4 g = __ldcs(reinterpret_cast<const float4 *>(src + plane) + p);
      const float4 b = __ldcs(reinterpret_cast<const float4 *>(src + 2 * plane) + p);
      const uint32_t by[12] = {to_u8(r.x), to_u8(g.x), to_u8(b.x), to_u8(r.y), to_u8(g.y), to_u8(b.y),
                               to_u8(r.z), to_u8(g.z), to_u8(b.z), to_u8(r.w), to_u8(g.w), to_u8(b.w)};
      uint32_t *dst = reinterpret_cast<uint32_t *>(out + (f * plane + p * 4) * 3);
#pragma unroll
      for (int w = 0; w < 3; ++w)
        dst[w] = by[4 * w] | (by[4 * w + 1] << 8) | (by[4 * w + 2] << 16) | (by[4 * w + 3] << 24);
    } else {
      uint8_t *dst = out + (f * plane + p) * 3;
      dst[0] = (uint8_t)to_u8(src[p]); dst[1] = (uint8_t)to_u8(src[plane + p]); dst[2] = (uint8_t)to_u8(src[2 * plane + p]);
    }
  }
}

}  // namespace pm

extern "C" {

int pm_sample_cameras(const float *dense_cameras, const float *t, int B, int N, int M, float *out, pm_stream_t stream) {
  if (B < 0 || N < 0 || M < 0) return PM_ERR_BAD_ARG;
  if (B == 0 || M == 0) return PM_OK;
  if (!dense_cameras || !t || !out) return PM_ERR_BAD_ARG;
  if (N < 2) return PM_ERR_BAD_ARG;  // a path needs two poses
  if (N > pm::kTrajMaxPoses) return PM_ERR_RANGE;
  pm::sample_cameras_kernel<<<B, pm::kTrajThreads, 0, pm::as_stream(stream)>>>(dense_cameras, t, N, M, out);
  return pm::check_launch();
}

int pm_pack_frames_u8(const float *images, int64_t n_frames, int height, int width, uint8_t *out, pm_stream_t stream) {
  if (n_frames < 0 || height <= 0 || width <= 0) return PM_ERR_BAD_ARG;
  if (n_frames == 0) return PM_OK;
  if (!images || !out) return PM_ERR_BAD_ARG;
  const int64_t plane = (int64_t)height * width;
  const int vec4 = (plane % 4 == 0) && (reinterpret_cast<uintptr_t>(images) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 3) == 0;
  const int64_t work = n_frames * (vec4 ? plane / 4 : plane);
  int64_t blocks = (work + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;  // grid-stride: a multiple of the SM count
  pm::pack_frames_u8_kernel<<<(unsigned)blocks, 256, 0, pm::as_stream(stream)>>>(images, n_frames, plane, out, vec4);
  return pm::check_launch();
}

}  // extern "C"

// ---- densification statistics ----------------------------------------------------
// Replaces the per-view host loop of the SDS++ refiner (prometheus/models/sdspp_refiner.py:275-280) and
// GaussianModel.add_densification_stats (prometheus/utils/gs_utils.py:247-252): for every rendered view, in
// order: vis = radii > 0; is_visible[vis] = 1; max_radii2D[vis] = max(max_radii2D[vis], radii[vis]);
// xyz_gradient_accum[vis] += |grad means2d[vis]|; denom[vis] += 1 — ~10 boolean-mask indexing kernels per view
// in the reference, one launch for all views here (one thread per Gaussian walks the views in the same order,
// so the fp32 sums round identically).
namespace pm {

__global__ void __launch_bounds__(256)
densify_stats_kernel(const int32_t *__restrict__ radii, const float *__restrict__ v_means2d, int C, int N,
                     float *__restrict__ is_visible, float *__restrict__ max_radii2d,
                     float *__restrict__ grad_accum, float *__restrict__ denom) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  float vis = is_visible[g], mr = max_radii2d[g], acc = grad_accum[g], den = denom[g];
  for (int c = 0; c < C; ++c) {
    const size_t f = (size_t)c * N + g;
    const int r = __ldg(radii + f);
    const float2 v = __ldg(reinterpret_cast<const float2 *>(v_means2d) + f);
    if (r > 0) {
      vis = 1.0f;
      mr = fmaxf(mr, (float)r);
      acc = __fadd_rn(acc, sqrtf(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y))));
      den = __fadd_rn(den, 1.0f);
    }
  }
  is_visible[g] = vis; max_radii2d[g] = mr; grad_accum[g] = acc; denom[g] = den;
}

}  // namespace pm

extern "C" int pm_densify_stats(const int32_t *radii, const float *v_means2d, int C, int N, float *is_visible,
                                float *max_radii2d, float *xyz_gradient_accum, float *denom, pm_stream_t stream) {
  if (C < 0 || N < 0) return PM_ERR_BAD_ARG;
  if (C == 0 || N == 0) return PM_OK;
  if (!radii || !v_means2d || !is_visible || !max_radii2d || !xyz_gradient_accum || !denom) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(v_means2d) & 7) != 0) return PM_ERR_BAD_ARG;
  pm::densify_stats_kernel<<<(N + 255) / 256, 256, 0, pm::as_stream(stream)>>>(radii, v_means2d, C, N, is_visible,
                                                                               max_radii2d, xyz_gradient_accum, denom);
  return pm::check_launch();
}

// ---- .splat packing ----------------------------------------------------------------
// Replaces the exporter loop of export_ply_for_gaussians (prometheus/utils/io.py:25-125): the reference inverts
// the activations (opacity -> logit, scale -> log(scale + 1e-8)), sorts the Gaussians by
// -exp(sum log-scales) / (1 + exp(-logit)) with np.argsort and then builds every 32-byte record in a Python
// loop (position 3 x f32, exp(log-scale) 3 x f32, rgba 4 x u8, normalised quaternion 4 x u8) — minutes for 2 M
// Gaussians. Here: one kernel for the sort keys (order-preserving uint32 image of the fp32 key), the
// hand-written radix sort of radix_sort.cu, one kernel that gathers and packs the records.
namespace pm {

__device__ __forceinline__ float splat_logit(float o) { return logf(o / (1.0f - o)); }   // math.py:7 inverse_sigmoid

__global__ void __launch_bounds__(256)
splat_keys_kernel(const float *__restrict__ opacity, const float *__restrict__ scales, int64_t n, int compatible,
                  int64_t *__restrict__ keys, int32_t *__restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float lg = opacity[i], l0 = scales[i * 3 + 0], l1 = scales[i * 3 + 1], l2 = scales[i * 3 + 2];
  if (compatible) { lg = splat_logit(lg); l0 = logf(l0 + 1e-8f); l1 = logf(l1 + 1e-8f); l2 = logf(l2 + 1e-8f); }
  const float key = -expf((l0 + l1) + l2) / (1.0f + expf(-lg));  // io.py:76-79, ascending
  uint32_t u = __float_as_uint(key);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // order-preserving map fp32 -> uint32
  keys[i] = (int64_t)u;
  vals[i] = (int32_t)i;
}

__global__ void __launch_bounds__(256)
splat_pack_kernel(const float *__restrict__ xyz, const float *__restrict__ rgb, const float *__restrict__ opacity,
                  const float *__restrict__ scales, const float *__restrict__ rots, const int32_t *__restrict__ order,
                  int64_t n, int compatible, uint8_t *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t g = order[i];
  float s0 = scales[g * 3 + 0], s1 = scales[g * 3 + 1], s2 = scales[g * 3 + 2], a = opacity[g];
  if (compatible) {  // the reference stores log / logit and undoes them when packing: same round trip here
    s0 = expf(logf(s0 + 1e-8f)); s1 = expf(logf(s1 + 1e-8f)); s2 = expf(logf(s2 + 1e-8f));
    a = 1.0f / (1.0f + expf(-splat_logit(a)));
  } else {           // compatible=False stores the activated values and the packer still applies exp / sigmoid
    s0 = expf(s0); s1 = expf(s1); s2 = expf(s2);
    a = 1.0f / (1.0f + expf(-a));
  }
  const float4 q = reinterpret_cast<const float4 *>(rots)[g];
  const float qn = sqrtf(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
  auto u8 = [](float v) { return (uint32_t)fminf(fmaxf(v, 0.0f), 255.0f); };  // clip(0,255).astype(uint8)
  const uint32_t col = u8(rgb[g * 3 + 0] * 255.0f) | (u8(rgb[g * 3 + 1] * 255.0f) << 8) | (u8(rgb[g * 3 + 2] * 255.0f) << 16) |
                       (u8(a * 255.0f) << 24);
  const uint32_t rot = u8((q.x / qn) * 128.0f + 128.0f) | (u8((q.y / qn) * 128.0f + 128.0f) << 8) |
                       (u8((q.z / qn) * 128.0f + 128.0f) << 16) | (u8((q.w / qn) * 128.0f + 128.0f) << 24);
  uint4 *dst = reinterpret_cast<uint4 *>(out + i * 32);
  dst[0] = make_uint4(__float_as_uint(xyz[g * 3 + 0]), __float_as_uint(xyz[g * 3 + 1]), __float_as_uint(xyz[g * 3 + 2]),
                      __float_as_uint(s0));
  dst[1] = make_uint4(__float_as_uint(s1), __float_as_uint(s2), col, rot);
}

}  // namespace pm

extern "C" {

int pm_splat_keys(const float *opacity, const float *scales, int64_t n, int compatible, int64_t *keys, int32_t *vals,
                  pm_stream_t stream) {
  if (n < 0) return PM_ERR_BAD_ARG;
  if (n == 0) return PM_OK;
  if (!opacity || !scales || !keys || !vals) return PM_ERR_BAD_ARG;
  if (n >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  pm::splat_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, pm::as_stream(stream)>>>(opacity, scales, n, compatible, keys, vals);
  return pm::check_launch();
}

int pm_splat_pack(const float *xyz, const float *rgb, const float *opacity, const float *scales, const float *rots,
                  const int32_t *order, int64_t n, int compatible, uint8_t *out, pm_stream_t stream) {
  if (n < 0) return PM_ERR_BAD_ARG;
  if (n == 0) return PM_OK;
  if (!xyz || !rgb || !opacity || !scales || !rots || !order || !out) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(rots) & 15) != 0 || (reinterpret_cast<uintptr_t>(out) & 15) != 0) return PM_ERR_BAD_ARG;
  pm::splat_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, pm::as_stream(stream)>>>(xyz, rgb, opacity, scales, rots, order, n,
                                                                                        compatible, out);
  return pm::check_launch();
}

}  // extern "C"
