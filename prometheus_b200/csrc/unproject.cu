// unproject.cu — stage 1: per-pixel 14-channel decoder maps -> pixel-aligned 3D
// Gaussians, forward and backward, one fused kernel each.
//
// Replaces the ~40 ATen launches of GaussianConverter.forward
// (reference prometheus/modules/renderers/gaussians_renderer.py:210-287) and the
// autograd graph behind them. HBM-bound by construction: 56 B read + 56 B
// written per Gaussian (fwd), every byte touched exactly once. One thread per
// pixel; the 14 channel planes are read with fully coalesced 128-B warp loads
// (channel-planar input), the AoS outputs are written as contiguous 12/16-B
// records so every 32-B sector a warp touches is fully used.
#include "common.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace pm {

// The decoder maps may arrive in half precision (the decoder runs under autocast; the reference's
// `local_gaussian_params * self.muls` promotes them to fp32, gaussians_renderer.py:219): read them in their
// storage type and widen in registers instead of paying a separate conversion pass over the 14 planes.
template <typename T> __device__ __forceinline__ float load_param(const T *p);
template <> __device__ __forceinline__ float load_param<float>(const float *p) { return __ldcs(p); }
template <> __device__ __forceinline__ float load_param<__half>(const __half *p) {
  return __half2float(__ushort_as_half(__ldcs(reinterpret_cast<const unsigned short *>(p))));
}
template <> __device__ __forceinline__ float load_param<__nv_bfloat16>(const __nv_bfloat16 *p) {
  return __uint_as_float((unsigned)__ldcs(reinterpret_cast<const unsigned short *>(p)) << 16);
}
template <typename T> __device__ __forceinline__ void store_param_grad(T *p, float v);
template <> __device__ __forceinline__ void store_param_grad<float>(float *p, float v) { __stcs(p, v); }
template <> __device__ __forceinline__ void store_param_grad<__half>(__half *p, float v) {
  __stcs(reinterpret_cast<unsigned short *>(p), __half_as_ushort(__float2half_rn(v)));
}
template <> __device__ __forceinline__ void store_param_grad<__nv_bfloat16>(__nv_bfloat16 *p, float v) {
  __stcs(reinterpret_cast<unsigned short *>(p), __bfloat16_as_ushort(__float2bfloat16_rn(v)));
}

struct UnprojCam {
  float R[9];
  float t[3];
  float fx, fy, cx, cy;
};

__device__ __forceinline__ UnprojCam load_unproj_cam(const float *__restrict__ cam, int h, int w) {
  UnprojCam c;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    c.R[r * 3 + 0] = __ldg(cam + r * 4 + 0);
    c.R[r * 3 + 1] = __ldg(cam + r * 4 + 1);
    c.R[r * 3 + 2] = __ldg(cam + r * 4 + 2);
    c.t[r] = __ldg(cam + r * 4 + 3);
  }
  const float H = __ldg(cam + 16), W = __ldg(cam + 17);
  // gaussians_renderer.py:226-227 (x scaled by h/H, y by w/W — kept as written there)
  c.fx = __ldg(cam + 12) * (float)h / H;
  c.cx = __ldg(cam + 14) * (float)h / H;
  c.fy = __ldg(cam + 13) * (float)w / W;
  c.cy = __ldg(cam + 15) * (float)w / W;
  return c;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

constexpr float kDirector3dOpacityOffset = -4.59511985013459f;  // log(0.01/0.99)
constexpr float kDirector3dScaleOffset = -4.60517018598809f;    // log(1/100)
constexpr float kDirector3dRgbDiv = 2.0f * 0.28209479177387814f;

// Where the Gaussians go: one set of the five output arrays per destination. n = 1 is the plain operator;
// n > 1 is the fused "unproject + all-gather" of the view-sharded decode (config C5): destination k is this
// rank's slot of the gather buffer on GPU k, reached with plain stores over NVLink peer mappings, so the
// exchange step needs no separate collective and no local staging copy.
constexpr int kMaxUnprojDst = 8;
struct UnprojDst {
  float *xyz, *rgb, *opacity, *scales, *quats;
};
struct UnprojDstSet {
  UnprojDst d[kMaxUnprojDst];
  int n;
  int vec16;  // every array of every destination is 16-byte aligned (set by the launcher): 16-byte stores allowed
};

template <typename TP>
__global__ void __launch_bounds__(256, 5)  // 48 registers: 5 CTAs/SM, as before the kernel was templated on the map type
unproject_fwd_kernel(const TP *__restrict__ params, const float *__restrict__ cameras,
                     const float *__restrict__ muls, int mode, float z_near, float z_far,
                     float s_min, float s_max, int h, int w, const UnprojDstSet dst) {
  const int hw = h * w;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  // Output staging (block-uniform): a full CTA whose first Gaussian is 16-byte aligned in every array hands its 256 x
  // {xyz, rgb, opacity, scales} to shared memory and writes each array's contiguous block with 16-byte stores — whole
  // 32-byte sectors per store instruction instead of 4-byte pieces at a 12-byte stride. Over NVLink (peer / multicast
  // destinations of pm_unproject_fwd_multi) a partial-sector store is a packet of its own.
  __shared__ __align__(16) float s_out[256 * 10];
  const bool staged = dst.vec16 && (blockIdx.x + 1) * 256 <= hw && (((size_t)v * hw) & 3) == 0 && blockDim.x == 256;
  if (pix >= hw) return;  // (never in a staged CTA)
  const UnprojCam c = load_unproj_cam(cameras + (size_t)v * 18, h, w);
  const TP *p = params + (size_t)v * 14 * hw + pix;
  float ch[14];
#pragma unroll
  for (int k = 0; k < 14; ++k) ch[k] = load_param<TP>(p + (size_t)k * hw) * __ldg(muls + k);

  const float i = (float)(pix % w) + 0.5f;
  const float j = (float)(pix / w) + 0.5f;
  const float u = i / c.cx + ch[3];
  const float vv = j / c.cy + ch[4];
  const float xs = (u - 1.0f) * c.cx / c.fx;   // -(u-1)*cx/fx * zs, zs = -1
  const float ys = -((vv - 1.0f) * c.cy / c.fy);
  float d0 = xs * c.R[0] + ys * c.R[1] - c.R[2];
  float d1 = xs * c.R[3] + ys * c.R[4] - c.R[5];
  float d2 = xs * c.R[6] + ys * c.R[7] - c.R[8];
  const float nrm = fmaxf(sqrtf(d0 * d0 + d1 * d1 + d2 * d2), 1e-12f);
  d0 /= nrm; d1 /= nrm; d2 /= nrm;

  float depth, op, s0, s1, s2, c0, c1, c2, q0, q1, q2, q3;
  if (mode == PM_MODE_DIRECTOR3D) {
    depth = ch[5] + 1.85f;
    c0 = ch[0] / kDirector3dRgbDiv; c1 = ch[1] / kDirector3dRgbDiv; c2 = ch[2] / kDirector3dRgbDiv;
    op = sigmoidf_(ch[6] + kDirector3dOpacityOffset);
    s0 = expf(ch[7] + kDirector3dScaleOffset);
    s1 = expf(ch[8] + kDirector3dScaleOffset);
    s2 = expf(ch[9] + kDirector3dScaleOffset);
    q0 = ch[10] + 1.0f; q1 = ch[11]; q2 = ch[12]; q3 = ch[13];
  } else if (mode == PM_MODE_GSLRM) {
    const float d_ = sigmoidf_(ch[5]);
    depth = z_near * (1.0f - d_) + z_far * d_;
    c0 = ch[0]; c1 = ch[1]; c2 = ch[2];
    op = sigmoidf_(ch[6] - 2.0f);
    s0 = fminf(fmaxf(expf(ch[7] - 2.3f), 0.0f), 0.3f);
    s1 = fminf(fmaxf(expf(ch[8] - 2.3f), 0.0f), 0.3f);
    s2 = fminf(fmaxf(expf(ch[9] - 2.3f), 0.0f), 0.3f);
    q0 = ch[10]; q1 = ch[11]; q2 = ch[12]; q3 = ch[13];
  } else {  // mvsplat
    const float d_ = sigmoidf_(ch[5]);
    depth = z_near * (1.0f - d_) + z_far * d_;
    const float a0 = sigmoidf_(ch[7]), a1 = sigmoidf_(ch[8]), a2 = sigmoidf_(ch[9]);
    s0 = s_min * (1.0f - a0) + s_max * a0;
    s1 = s_min * (1.0f - a1) + s_max * a1;
    s2 = s_min * (1.0f - a2) + s_max * a2;
    c0 = sigmoidf_(ch[0]); c1 = sigmoidf_(ch[1]); c2 = sigmoidf_(ch[2]);
    op = sigmoidf_(ch[6]);
    q0 = ch[10]; q1 = ch[11]; q2 = ch[12]; q3 = ch[13];
  }
  const float qn = fmaxf(sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3), 1e-12f);

  const size_t g = (size_t)v * hw + pix;
  const float x0 = c.t[0] + depth * d0, x1 = c.t[1] + depth * d1, x2 = c.t[2] + depth * d2;
  const float4 qv = make_float4(q0 / qn, q1 / qn, q2 / qn, q3 / qn);
  if (staged) {
    const int t = threadIdx.x;
    float *s_xyz = s_out, *s_rgb = s_out + 768, *s_sc = s_out + 1536, *s_op = s_out + 2304;
    s_xyz[t * 3 + 0] = x0; s_xyz[t * 3 + 1] = x1; s_xyz[t * 3 + 2] = x2;   // (stride 3: conflict-free)
    s_rgb[t * 3 + 0] = c0; s_rgb[t * 3 + 1] = c1; s_rgb[t * 3 + 2] = c2;
    s_sc[t * 3 + 0] = s0; s_sc[t * 3 + 1] = s1; s_sc[t * 3 + 2] = s2;
    s_op[t] = op;
    __syncthreads();
    const size_t g0 = g - t;  // the CTA's first Gaussian
    const float4 *s4 = reinterpret_cast<const float4 *>(s_out);
    for (int k = 0; k < dst.n; ++k) {
      const UnprojDst &o = dst.d[k];
      if (t < 192) {
        reinterpret_cast<float4 *>(o.xyz + g0 * 3)[t] = s4[t];
        reinterpret_cast<float4 *>(o.rgb + g0 * 3)[t] = s4[192 + t];
        reinterpret_cast<float4 *>(o.scales + g0 * 3)[t] = s4[384 + t];
      } else {
        reinterpret_cast<float4 *>(o.opacity + g0)[t - 192] = s4[576 + t - 192];
      }
      reinterpret_cast<float4 *>(o.quats)[g] = qv;
    }
    return;
  }
  for (int k = 0; k < dst.n; ++k) {
    const UnprojDst &o = dst.d[k];
    o.xyz[g * 3 + 0] = x0; o.xyz[g * 3 + 1] = x1; o.xyz[g * 3 + 2] = x2;
    o.rgb[g * 3 + 0] = c0; o.rgb[g * 3 + 1] = c1; o.rgb[g * 3 + 2] = c2;
    o.opacity[g] = op;
    o.scales[g * 3 + 0] = s0; o.scales[g * 3 + 1] = s1; o.scales[g * 3 + 2] = s2;
    reinterpret_cast<float4 *>(o.quats)[g] = qv;
  }
}

#ifndef PM_UBWD_MINBLOCKS
#define PM_UBWD_MINBLOCKS 4
#endif
template <typename TP>
__global__ void __launch_bounds__(256, PM_UBWD_MINBLOCKS)
unproject_bwd_kernel(const TP *__restrict__ params, const float *__restrict__ cameras,
                     const float *__restrict__ muls, int mode, float z_near, float z_far,
                     float s_min, float s_max, int h, int w,
                     const float *__restrict__ v_xyz, const float *__restrict__ v_rgb,
                     const float *__restrict__ v_opacity, const float *__restrict__ v_scales,
                     const float *__restrict__ v_quats, TP *__restrict__ v_params) {
  const int hw = h * w;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y;
  if (pix >= hw) return;
  const UnprojCam c = load_unproj_cam(cameras + (size_t)v * 18, h, w);
  const TP *p = params + (size_t)v * 14 * hw + pix;
  float mul[14], ch[14];
#pragma unroll
  for (int k = 0; k < 14; ++k) {
    mul[k] = __ldg(muls + k);
    ch[k] = load_param<TP>(p + (size_t)k * hw) * mul[k];
  }
  const size_t g = (size_t)v * hw + pix;
  float gx[3] = {0, 0, 0}, gc[3] = {0, 0, 0}, gs[3] = {0, 0, 0}, gq[4] = {0, 0, 0, 0}, go = 0.0f;
  if (v_xyz) { gx[0] = v_xyz[g * 3]; gx[1] = v_xyz[g * 3 + 1]; gx[2] = v_xyz[g * 3 + 2]; }
  if (v_rgb) { gc[0] = v_rgb[g * 3]; gc[1] = v_rgb[g * 3 + 1]; gc[2] = v_rgb[g * 3 + 2]; }
  if (v_scales) { gs[0] = v_scales[g * 3]; gs[1] = v_scales[g * 3 + 1]; gs[2] = v_scales[g * 3 + 2]; }
  if (v_quats) {
    const float4 t = reinterpret_cast<const float4 *>(v_quats)[g];
    gq[0] = t.x; gq[1] = t.y; gq[2] = t.z; gq[3] = t.w;
  }
  if (v_opacity) go = v_opacity[g];

  // recompute the ray
  const float i = (float)(pix % w) + 0.5f;
  const float j = (float)(pix / w) + 0.5f;
  const float u = i / c.cx + ch[3];
  const float vv = j / c.cy + ch[4];
  const float xs = (u - 1.0f) * c.cx / c.fx;
  const float ys = -((vv - 1.0f) * c.cy / c.fy);
  const float r0 = xs * c.R[0] + ys * c.R[1] - c.R[2];
  const float r1 = xs * c.R[3] + ys * c.R[4] - c.R[5];
  const float r2 = xs * c.R[6] + ys * c.R[7] - c.R[8];
  const float nrm_raw = sqrtf(r0 * r0 + r1 * r1 + r2 * r2);
  const float nrm = fmaxf(nrm_raw, 1e-12f);
  const float d0 = r0 / nrm, d1 = r1 / nrm, d2 = r2 / nrm;

  float out[14];
  float depth, ddepth_draw;
  if (mode == PM_MODE_DIRECTOR3D) {
    depth = ch[5] + 1.85f;
    ddepth_draw = 1.0f;
  } else {
    const float d_ = sigmoidf_(ch[5]);
    depth = z_near * (1.0f - d_) + z_far * d_;
    ddepth_draw = (z_far - z_near) * d_ * (1.0f - d_);
  }
  // xyz = t + depth * dn
  const float v_depth = gx[0] * d0 + gx[1] * d1 + gx[2] * d2;
  float vd0 = depth * gx[0], vd1 = depth * gx[1], vd2 = depth * gx[2];  // d/d(dn)
  if (nrm_raw > 1e-12f) {
    const float dot = vd0 * d0 + vd1 * d1 + vd2 * d2;
    vd0 = (vd0 - dot * d0) / nrm; vd1 = (vd1 - dot * d1) / nrm; vd2 = (vd2 - dot * d2) / nrm;
  } else {
    vd0 /= nrm; vd1 /= nrm; vd2 /= nrm;
  }
  const float v_xs = vd0 * c.R[0] + vd1 * c.R[3] + vd2 * c.R[6];
  const float v_ys = vd0 * c.R[1] + vd1 * c.R[4] + vd2 * c.R[7];
  out[3] = v_xs * c.cx / c.fx;
  out[4] = -v_ys * c.cy / c.fy;
  out[5] = v_depth * ddepth_draw;

  float q0, q1, q2, q3;
  if (mode == PM_MODE_DIRECTOR3D) {
    out[0] = gc[0] / kDirector3dRgbDiv; out[1] = gc[1] / kDirector3dRgbDiv; out[2] = gc[2] / kDirector3dRgbDiv;
    const float o = sigmoidf_(ch[6] + kDirector3dOpacityOffset);
    out[6] = go * o * (1.0f - o);
#pragma unroll
    for (int k = 0; k < 3; ++k) out[7 + k] = gs[k] * expf(ch[7 + k] + kDirector3dScaleOffset);
    q0 = ch[10] + 1.0f; q1 = ch[11]; q2 = ch[12]; q3 = ch[13];
  } else if (mode == PM_MODE_GSLRM) {
    out[0] = gc[0]; out[1] = gc[1]; out[2] = gc[2];
    const float o = sigmoidf_(ch[6] - 2.0f);
    out[6] = go * o * (1.0f - o);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float e = expf(ch[7 + k] - 2.3f);
      out[7 + k] = (e >= 0.0f && e <= 0.3f) ? gs[k] * e : 0.0f;  // clip passes grad inside [0, 0.3]
    }
    q0 = ch[10]; q1 = ch[11]; q2 = ch[12]; q3 = ch[13];
  } else {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float s = sigmoidf_(ch[k]);
      out[k] = gc[k] * s * (1.0f - s);
    }
    const float o = sigmoidf_(ch[6]);
    out[6] = go * o * (1.0f - o);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float a = sigmoidf_(ch[7 + k]);
      out[7 + k] = gs[k] * (s_max - s_min) * a * (1.0f - a);
    }
    q0 = ch[10]; q1 = ch[11]; q2 = ch[12]; q3 = ch[13];
  }
  const float qn_raw = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  const float qn = fmaxf(qn_raw, 1e-12f);
  if (qn_raw > 1e-12f) {
    const float n0 = q0 / qn, n1 = q1 / qn, n2 = q2 / qn, n3 = q3 / qn;
    const float dot = gq[0] * n0 + gq[1] * n1 + gq[2] * n2 + gq[3] * n3;
    out[10] = (gq[0] - dot * n0) / qn; out[11] = (gq[1] - dot * n1) / qn;
    out[12] = (gq[2] - dot * n2) / qn; out[13] = (gq[3] - dot * n3) / qn;
  } else {
    out[10] = gq[0] / qn; out[11] = gq[1] / qn; out[12] = gq[2] / qn; out[13] = gq[3] / qn;
  }
  TP *o = v_params + (size_t)v * 14 * hw + pix;
#pragma unroll
  for (int k = 0; k < 14; ++k) store_param_grad<TP>(o + (size_t)k * hw, out[k] * mul[k]);
}

// Device-side camera preparation. Reference: host round trip in
// gaussians_renderer.py:453-472 (+ MiniCam :139-169, K build :404-412).
__global__ void camera_prep_kernel(const float *__restrict__ cameras, int n, int h, int w,
                                   float *__restrict__ viewmats, float *__restrict__ Ks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *cam = cameras + (size_t)i * 18;
  // c2w with OpenGL->OpenCV flip of columns 1, 2 (:151)
  float m[9], t[3];
  for (int r = 0; r < 3; ++r) {
    m[r * 3 + 0] = cam[r * 4 + 0];
    m[r * 3 + 1] = -cam[r * 4 + 1];
    m[r * 3 + 2] = -cam[r * 4 + 2];
    t[r] = cam[r * 4 + 3];
  }
  // general 3x3 inverse (adjugate / determinant): valid for any invertible affine c2w,
  // not only rigid ones (the reference calls np.linalg.inv, :153)
  const float a00 = m[4] * m[8] - m[5] * m[7], a01 = m[2] * m[7] - m[1] * m[8], a02 = m[1] * m[5] - m[2] * m[4];
  const float a10 = m[5] * m[6] - m[3] * m[8], a11 = m[0] * m[8] - m[2] * m[6], a12 = m[2] * m[3] - m[0] * m[5];
  const float a20 = m[3] * m[7] - m[4] * m[6], a21 = m[1] * m[6] - m[0] * m[7], a22 = m[0] * m[4] - m[1] * m[3];
  const float det = m[0] * a00 + m[1] * a10 + m[2] * a20;
  const float id = 1.0f / det;
  const float inv[9] = {a00 * id, a01 * id, a02 * id, a10 * id, a11 * id, a12 * id, a20 * id, a21 * id, a22 * id};
  float *vm = viewmats + (size_t)i * 16;
  for (int r = 0; r < 3; ++r) {
    vm[r * 4 + 0] = inv[r * 3 + 0];
    vm[r * 4 + 1] = inv[r * 3 + 1];
    vm[r * 4 + 2] = inv[r * 3 + 2];
    vm[r * 4 + 3] = -(inv[r * 3 + 0] * t[0] + inv[r * 3 + 1] * t[1] + inv[r * 3 + 2] * t[2]);
  }
  vm[12] = 0.0f; vm[13] = 0.0f; vm[14] = 0.0f; vm[15] = 1.0f;
  // K: focal = size / (2 tan(fov/2)) with fov = 2 atan(0.5 size / f')  ==  f' * (w/h) resp. (h/w);
  // principal point forced to the image centre (:405-412)
  const float H = cam[16], W = cam[17];
  const float fxp = cam[12] * (float)h / H;  // :461
  const float fyp = cam[13] * (float)w / W;  // :465
  const float tan_half_fovx = 0.5f * (float)h / fxp;  // tan(atan(x)) == x
  const float tan_half_fovy = 0.5f * (float)w / fyp;
  float *K = Ks + (size_t)i * 9;
  K[0] = (float)w / (2.0f * tan_half_fovx); K[1] = 0.0f; K[2] = (float)w / 2.0f;
  K[3] = 0.0f; K[4] = (float)h / (2.0f * tan_half_fovy); K[5] = (float)h / 2.0f;
  K[6] = 0.0f; K[7] = 0.0f; K[8] = 1.0f;
}

}  // namespace pm

namespace pm {

static int unproject_check(const void *params, const float *cameras, const float *muls, int dtype, int mode,
                           int n_views, int h, int w) {
  if (!params || !cameras || !muls) return PM_ERR_BAD_ARG;
  if (n_views < 0 || h <= 0 || w <= 0) return PM_ERR_BAD_ARG;
  if (mode < PM_MODE_MVSPLAT || mode > PM_MODE_DIRECTOR3D) return PM_ERR_UNSUPPORTED;
  if (dtype < PM_DTYPE_F32 || dtype > PM_DTYPE_BF16) return PM_ERR_UNSUPPORTED;
  if ((int64_t)h * w > (int64_t)1 << 30 || n_views > 65535) return PM_ERR_RANGE;
  return PM_OK;
}

static int unproject_fwd_launch(const void *params, int dtype, const float *cameras, const float *muls, int mode,
                                float z_near, float z_far, float s_min, float s_max, int n_views, int h, int w,
                                const UnprojDstSet &dst_in, pm_stream_t stream) {
  if (n_views == 0) return PM_OK;
  const int hw = h * w;
  dim3 grid((hw + 255) / 256, n_views);
  cudaStream_t st = as_stream(stream);
  UnprojDstSet dst = dst_in;
  uintptr_t bits = 0;
  for (int k = 0; k < dst.n; ++k)
    bits |= reinterpret_cast<uintptr_t>(dst.d[k].xyz) | reinterpret_cast<uintptr_t>(dst.d[k].rgb) |
            reinterpret_cast<uintptr_t>(dst.d[k].opacity) | reinterpret_cast<uintptr_t>(dst.d[k].scales);
  dst.vec16 = dst.vec16 && (bits & 15) == 0;
  if (dtype == PM_DTYPE_F16)
    unproject_fwd_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half *>(params), cameras, muls, mode,
                                                       z_near, z_far, s_min, s_max, h, w, dst);
  else if (dtype == PM_DTYPE_BF16)
    unproject_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16 *>(params), cameras,
                                                              muls, mode, z_near, z_far, s_min, s_max, h, w, dst);
  else
    unproject_fwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float *>(params), cameras, muls, mode,
                                                      z_near, z_far, s_min, s_max, h, w, dst);
  return check_launch();
}

}  // namespace pm

extern "C" {

int pm_unproject_fwd_dt(const void *params, int params_dtype, const float *cameras, const float *muls, int mode,
                        float z_near, float z_far, float s_min, float s_max, int n_views, int h, int w,
                        float *xyz, float *rgb, float *opacity, float *scales, float *quats,
                        pm_stream_t stream) {
  if (const int rc = pm::unproject_check(params, cameras, muls, params_dtype, mode, n_views, h, w)) return rc;
  if (!xyz || !rgb || !opacity || !scales || !quats) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(quats) & 15) != 0) return PM_ERR_BAD_ARG;
  pm::UnprojDstSet dst;
  dst.n = 1;
  dst.vec16 = 0;  // local destination: L2 merges the partial sectors, staging only costs (0.058 -> 0.061 ms on C2)
  dst.d[0] = pm::UnprojDst{xyz, rgb, opacity, scales, quats};
  return pm::unproject_fwd_launch(params, params_dtype, cameras, muls, mode, z_near, z_far, s_min, s_max, n_views,
                                  h, w, dst, stream);
}

int pm_unproject_fwd(const float *params, const float *cameras, const float *muls, int mode,
                     float z_near, float z_far, float s_min, float s_max, int n_views, int h, int w,
                     float *xyz, float *rgb, float *opacity, float *scales, float *quats,
                     pm_stream_t stream) {
  return pm_unproject_fwd_dt(params, PM_DTYPE_F32, cameras, muls, mode, z_near, z_far, s_min, s_max, n_views, h, w,
                             xyz, rgb, opacity, scales, quats, stream);
}

int pm_unproject_fwd_multi(const void *params, int params_dtype, const float *cameras, const float *muls, int mode,
                           float z_near, float z_far, float s_min, float s_max, int n_views, int h, int w,
                           int n_dst, float *const *dst_ptrs, pm_stream_t stream) {
  if (const int rc = pm::unproject_check(params, cameras, muls, params_dtype, mode, n_views, h, w)) return rc;
  if (!dst_ptrs) return PM_ERR_BAD_ARG;
  if (n_dst < 1 || n_dst > pm::kMaxUnprojDst) return PM_ERR_RANGE;
  pm::UnprojDstSet dst;
  dst.n = n_dst;
  dst.vec16 = 1;  // peer / multicast destinations: whole sectors per store (cleared by the launcher if misaligned)
  for (int k = 0; k < n_dst; ++k) {
    float *const *p = dst_ptrs + 5 * k;
    if (!p[0] || !p[1] || !p[2] || !p[3] || !p[4] || (reinterpret_cast<uintptr_t>(p[4]) & 15) != 0) return PM_ERR_BAD_ARG;
    dst.d[k] = pm::UnprojDst{p[0], p[1], p[2], p[3], p[4]};
  }
  return pm::unproject_fwd_launch(params, params_dtype, cameras, muls, mode, z_near, z_far, s_min, s_max, n_views,
                                  h, w, dst, stream);
}

int pm_unproject_bwd_dt(const void *params, int params_dtype, const float *cameras, const float *muls, int mode,
                        float z_near, float z_far, float s_min, float s_max, int n_views, int h, int w,
                        const float *v_xyz, const float *v_rgb, const float *v_opacity,
                        const float *v_scales, const float *v_quats, void *v_params,
                        pm_stream_t stream) {
  if (const int rc = pm::unproject_check(params, cameras, muls, params_dtype, mode, n_views, h, w)) return rc;
  if (!v_params) return PM_ERR_BAD_ARG;
  if (v_quats && (reinterpret_cast<uintptr_t>(v_quats) & 15) != 0) return PM_ERR_BAD_ARG;
  if (n_views == 0) return PM_OK;
  const int hw = h * w;
  dim3 grid((hw + 255) / 256, n_views);
  cudaStream_t st = pm::as_stream(stream);
  if (params_dtype == PM_DTYPE_F16)
    pm::unproject_bwd_kernel<__half><<<grid, 256, 0, st>>>(
        static_cast<const __half *>(params), cameras, muls, mode, z_near, z_far, s_min, s_max, h, w, v_xyz, v_rgb,
        v_opacity, v_scales, v_quats, static_cast<__half *>(v_params));
  else if (params_dtype == PM_DTYPE_BF16)
    pm::unproject_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
        static_cast<const __nv_bfloat16 *>(params), cameras, muls, mode, z_near, z_far, s_min, s_max, h, w, v_xyz,
        v_rgb, v_opacity, v_scales, v_quats, static_cast<__nv_bfloat16 *>(v_params));
  else
    pm::unproject_bwd_kernel<float><<<grid, 256, 0, st>>>(
        static_cast<const float *>(params), cameras, muls, mode, z_near, z_far, s_min, s_max, h, w, v_xyz, v_rgb,
        v_opacity, v_scales, v_quats, static_cast<float *>(v_params));
  return pm::check_launch();
}

int pm_unproject_bwd(const float *params, const float *cameras, const float *muls, int mode,
                     float z_near, float z_far, float s_min, float s_max, int n_views, int h, int w,
                     const float *v_xyz, const float *v_rgb, const float *v_opacity,
                     const float *v_scales, const float *v_quats, float *v_params,
                     pm_stream_t stream) {
  return pm_unproject_bwd_dt(params, PM_DTYPE_F32, cameras, muls, mode, z_near, z_far, s_min, s_max, n_views, h, w,
                             v_xyz, v_rgb, v_opacity, v_scales, v_quats, v_params, stream);
}

int pm_camera_prep(const float *cameras, int n, int h, int w, float *viewmats, float *Ks,
                   pm_stream_t stream) {
  if (!cameras || !viewmats || !Ks || n < 0 || h <= 0 || w <= 0) return PM_ERR_BAD_ARG;
  if (n == 0) return PM_OK;
  pm::camera_prep_kernel<<<(n + 127) / 128, 128, 0, pm::as_stream(stream)>>>(cameras, n, h, w, viewmats, Ks);
  return pm::check_launch();
}

}  // extern "C"
