// project_bwd.cu — stage 2 backward: VJP of the EWA projection.
//
// Replaces gsplat 1.4.0 fully_fused_projection_bwd (reached through
// prometheus/modules/renderers/gaussians_renderer.py:413-427 under autograd).
// Unlike project.cu this file is compiled with the default FMA contraction: nothing here feeds an
// integer output, so only float tolerance applies, and the fp64 arithmetic is DFMA-throughput bound.
#include "common.cuh"

namespace pm {

struct ProjCam {
  float R[9];
  float t[3];
  float fx, fy, cx, cy;
};

__device__ __forceinline__ ProjCam load_proj_cam(const float *__restrict__ vm, const float *__restrict__ K) {
  ProjCam c;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    c.R[r * 3 + 0] = __ldg(vm + r * 4 + 0);
    c.R[r * 3 + 1] = __ldg(vm + r * 4 + 1);
    c.R[r * 3 + 2] = __ldg(vm + r * 4 + 2);
    c.t[r] = __ldg(vm + r * 4 + 3);
  }
  c.fx = __ldg(K + 0); c.fy = __ldg(K + 4); c.cx = __ldg(K + 2); c.cy = __ldg(K + 5);
  return c;
}

// one camera's constants in fp64, staged in shared memory once per CTA
struct BwdCam {
  double R[9];
  double t[3];
  double fx, fy;
  double lim_x_pos, lim_x_neg, lim_y_pos, lim_y_neg;
};
constexpr int kBwdMaxCams = 512;  // 72 KB of dynamic shared memory at most

__device__ __forceinline__ double sum3(double a, double b, double c) { return (a + b) + c; }
__device__ __forceinline__ void quat_to_rotmat_d(double w, double x, double y, double z, double *R) {
  const double inv_norm = 1.0 / sqrt(((x * x + y * y) + z * z) + w * w);
  x *= inv_norm; y *= inv_norm; z *= inv_norm; w *= inv_norm;
  const double x2 = x * x, y2 = y * y, z2 = z * z;
  const double xy = x * y, xz = x * z, yz = y * z;
  const double wx = w * x, wy = w * y, wz = w * z;
  R[0] = 1.0 - 2.0 * (y2 + z2); R[1] = 2.0 * (xy - wz);       R[2] = 2.0 * (xz + wy);
  R[3] = 2.0 * (xy + wz);       R[4] = 1.0 - 2.0 * (x2 + z2); R[5] = 2.0 * (yz - wx);
  R[6] = 2.0 * (xz - wy);       R[7] = 2.0 * (yz + wx);       R[8] = 1.0 - 2.0 * (x2 + y2);
}

// Backward: one thread per Gaussian loops over the cameras and accumulates in
// registers — no atomics, deterministic. Arithmetic in fp64 (the VJP chain multiplies by
// (f/z)^2 ~ 1e4 and cancels; fp64 keeps the kernel's own rounding out of the gradient error, and the
// kernel stays HBM-bound on B200). Restates fully_fused_projection_bwd
// ("VJPs of steps 5->1"); float tolerance parity only.
#ifndef PM_PBWD_MINBLOCKS
#define PM_PBWD_MINBLOCKS 6
#endif
__global__ void __launch_bounds__(128, PM_PBWD_MINBLOCKS)
project_bwd_kernel(const float *__restrict__ means, const float *__restrict__ quats,
                   const float *__restrict__ scales, const float *__restrict__ viewmats,
                   const float *__restrict__ Ks, int N, int C, int views_per_scene, int width, int height,
                   const int32_t *__restrict__ radii, const float *__restrict__ conics,
                   const float *__restrict__ v_means2d, const float *__restrict__ v_depths,
                   const float *__restrict__ v_conics, float *__restrict__ v_means,
                   float *__restrict__ v_quats, float *__restrict__ v_scales, const uint8_t *__restrict__ touched) {
  extern __shared__ __align__(16) unsigned char bwd_cam_smem[];
  BwdCam *s_cam = reinterpret_cast<BwdCam *>(bwd_cam_smem);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const ProjCam pc = load_proj_cam(viewmats + (size_t)c * 16, Ks + (size_t)c * 9);
    BwdCam bc;
#pragma unroll
    for (int i = 0; i < 9; ++i) bc.R[i] = pc.R[i];
    bc.t[0] = pc.t[0]; bc.t[1] = pc.t[1]; bc.t[2] = pc.t[2];
    bc.fx = pc.fx; bc.fy = pc.fy;
    const double cx = pc.cx, cy = pc.cy;
    const double tan_fovx = 0.5 * (double)width / bc.fx, tan_fovy = 0.5 * (double)height / bc.fy;
    bc.lim_x_pos = ((double)width - cx) / bc.fx + 0.3 * tan_fovx;
    bc.lim_x_neg = cx / bc.fx + 0.3 * tan_fovx;
    bc.lim_y_pos = ((double)height - cy) / bc.fy + 0.3 * tan_fovy;
    bc.lim_y_neg = cy / bc.fy + 0.3 * tan_fovy;
    s_cam[c] = bc;
  }
  __syncthreads();
  // scene batching (see project.cu): thread gg owns Gaussian g of scene gg / N and loops over that scene's cameras
  const int64_t gg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int n_scenes = C > 0 ? C / views_per_scene : 1;
  if (gg >= (int64_t)N * n_scenes) return;
  const int scene = (int)(gg / N);
  const int g = (int)(gg - (int64_t)scene * N);
  const double mx = means[(size_t)gg * 3 + 0], my = means[(size_t)gg * 3 + 1], mz = means[(size_t)gg * 3 + 2];
  const float4 q4 = reinterpret_cast<const float4 *>(quats)[gg];
  const double4 q = make_double4(q4.x, q4.y, q4.z, q4.w);
  const double sc[3] = {scales[(size_t)gg * 3 + 0], scales[(size_t)gg * 3 + 1], scales[(size_t)gg * 3 + 2]};
  // Sigma (world) is all the camera loop needs; Rq and M are rebuilt after it instead of being carried through
  // the loop in 36 registers. The kernel is latency-bound at low occupancy: measured on C2 with 3 / 4 / 5 / 6 / 8
  // CTAs per SM (register cap 164 / 128 / 96 / 80 / 64): 0.42 / 0.34 / 0.32 / 0.31 / 0.40 ms
  double S[9];
  {
    double Rq[9], M[9];
    quat_to_rotmat_d(q.x, q.y, q.z, q.w, Rq);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k) M[r * 3 + k] = Rq[r * 3 + k] * sc[k];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        S[i * 3 + j] = sum3(M[i * 3] * M[j * 3], M[i * 3 + 1] * M[j * 3 + 1], M[i * 3 + 2] * M[j * 3 + 2]);
  }

  double vmean[3] = {0, 0, 0};
  double vS[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // d loss / d Sigma (world), summed over cameras

  // Most (camera, Gaussian) pairs of an opaque scene arrive with all six cotangents exactly zero (never blended:
  // culled by the compositor, or behind the point where their tiles saturated) and contribute exactly nothing.
  // Pass 1 looks at 8 cameras at once — 56 independent loads in flight per thread instead of one dependent round
  // trip per camera — and keeps a bit per pair that has work; pass 2 runs the fp64 chain for those only (their
  // inputs come back from L1 / L2).
  const int cam_begin = scene * views_per_scene, cam_end = min(C, (scene + 1) * views_per_scene);
  for (int c0 = cam_begin; c0 < cam_end; c0 += 8) {
    unsigned active = 0u;
    if (touched) {
      // the compositing backward left one byte per pair that received a gradient: 1 byte per pair instead of 28
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (c0 + u < cam_end && __ldg(touched + (size_t)(c0 + u) * N + g) != 0) active |= 1u << u;
    } else {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (c0 + u < cam_end) {
        const size_t f = (size_t)(c0 + u) * N + g;
        unsigned bits = 0u;
        if (v_conics) bits |= __float_as_uint(__ldg(v_conics + f * 3 + 0)) | __float_as_uint(__ldg(v_conics + f * 3 + 1)) |
                              __float_as_uint(__ldg(v_conics + f * 3 + 2));
        if (v_means2d) bits |= __float_as_uint(__ldg(v_means2d + f * 2 + 0)) | __float_as_uint(__ldg(v_means2d + f * 2 + 1));
        if (v_depths) bits |= __float_as_uint(__ldg(v_depths + f));
        if (__ldg(radii + f) > 0 && (bits & 0x7fffffffu) != 0u) active |= 1u << u;
      }
    }
    }
    while (active) {
      const int cid = c0 + __ffs(active) - 1;
      active &= active - 1u;
      const size_t f = (size_t)cid * N + g;
      const float cq0 = __ldg(conics + f * 3 + 0), cq1 = __ldg(conics + f * 3 + 1), cq2 = __ldg(conics + f * 3 + 2);
      float gc0 = 0.0f, gc1 = 0.0f, gc2 = 0.0f, gm0 = 0.0f, gm1 = 0.0f, gd = 0.0f;
      if (v_conics) { gc0 = __ldg(v_conics + f * 3 + 0); gc1 = __ldg(v_conics + f * 3 + 1); gc2 = __ldg(v_conics + f * 3 + 2); }
      if (v_means2d) { gm0 = __ldg(v_means2d + f * 2 + 0); gm1 = __ldg(v_means2d + f * 2 + 1); }
      if (v_depths) gd = __ldg(v_depths + f);
      const BwdCam &cf = s_cam[cid];  // per-camera constants (incl. the clamp limits: 6 fp64 divisions) once per CTA
      const double *R = cf.R;
      const double *ct = cf.t;
      const double x = sum3(R[0] * mx, R[1] * my, R[2] * mz) + ct[0];
      const double y = sum3(R[3] * mx, R[4] * my, R[5] * mz) + ct[1];
      const double z = sum3(R[6] * mx, R[7] * my, R[8] * mz) + ct[2];
      // Sigma_c = R S R^T
      double Tm[9], Sc[9];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          Tm[i * 3 + j] = sum3(R[i * 3] * S[j], R[i * 3 + 1] * S[3 + j], R[i * 3 + 2] * S[6 + j]);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          Sc[i * 3 + j] = sum3(Tm[i * 3] * R[j * 3], Tm[i * 3 + 1] * R[j * 3 + 1], Tm[i * 3 + 2] * R[j * 3 + 2]);

      // conic -> cov2d : G = -Q Vc Q
      const double q00 = cq0, q01 = cq1, q11 = cq2;
      const double a00 = gc0, a01 = 0.5 * (double)gc1, a11 = gc2;
      const double t00 = q00 * a00 + q01 * a01, t01 = q00 * a01 + q01 * a11;
      const double t10 = q01 * a00 + q11 * a01, t11 = q01 * a01 + q11 * a11;
      const double G00 = -(t00 * q00 + t01 * q01);
      const double G01 = -(t00 * q01 + t01 * q11);
      const double G10 = -(t10 * q00 + t11 * q01);
      const double G11 = -(t10 * q01 + t11 * q11);

      const double fx = cf.fx, fy = cf.fy;
      const double lim_x_pos = cf.lim_x_pos, lim_x_neg = cf.lim_x_neg, lim_y_pos = cf.lim_y_pos, lim_y_neg = cf.lim_y_neg;
      const double rz = 1.0 / z, rz2 = rz * rz, rz3 = rz2 * rz;
      const double xr = x * rz, yr = y * rz;
      const bool x_in = (xr <= lim_x_pos) && (xr >= -lim_x_neg);
      const bool y_in = (yr <= lim_y_pos) && (yr >= -lim_y_neg);
      const double tx = z * fmin(lim_x_pos, fmax(-lim_x_neg, xr));
      const double ty = z * fmin(lim_y_pos, fmax(-lim_y_neg, yr));
      const double J00 = fx * rz, J02 = -fx * tx * rz2, J11 = fy * rz, J12 = -fy * ty * rz2;

      // v_Sc = J^T G J  (J = [[J00,0,J02],[0,J11,J12]])
      double vSc[9];
      vSc[0] = J00 * G00 * J00;
      vSc[1] = J00 * G01 * J11;
      vSc[2] = J00 * (G00 * J02 + G01 * J12);
      vSc[3] = J11 * G10 * J00;
      vSc[4] = J11 * G11 * J11;
      vSc[5] = J11 * (G10 * J02 + G11 * J12);
      vSc[6] = (J02 * G00 + J12 * G10) * J00;
      vSc[7] = (J02 * G01 + J12 * G11) * J11;
      vSc[8] = J02 * (G00 * J02 + G01 * J12) + J12 * (G10 * J02 + G11 * J12);

      // v_J = G J Sc^T + G^T J Sc ; only entries [0][0], [0][2], [1][1], [1][2] are used
      // (J Sc^T)[b][i] = sum_k J[b][k] Sc[i][k] ; (J Sc)[b][i] = sum_k J[b][k] Sc[k][i]
      double JSt[6], JS[6];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        JSt[0 * 3 + i] = J00 * Sc[i * 3 + 0] + J02 * Sc[i * 3 + 2];
        JSt[1 * 3 + i] = J11 * Sc[i * 3 + 1] + J12 * Sc[i * 3 + 2];
        JS[0 * 3 + i] = J00 * Sc[0 * 3 + i] + J02 * Sc[2 * 3 + i];
        JS[1 * 3 + i] = J11 * Sc[1 * 3 + i] + J12 * Sc[2 * 3 + i];
      }
      const double vJ00 = (G00 * JSt[0] + G01 * JSt[3]) + (G00 * JS[0] + G10 * JS[3]);
      const double vJ02 = (G00 * JSt[2] + G01 * JSt[5]) + (G00 * JS[2] + G10 * JS[5]);
      const double vJ11 = (G10 * JSt[1] + G11 * JSt[4]) + (G01 * JS[1] + G11 * JS[4]);
      const double vJ12 = (G10 * JSt[2] + G11 * JSt[5]) + (G01 * JS[2] + G11 * JS[5]);

      const double vm2x = gm0, vm2y = gm1;
      double vp0 = fx * rz * vm2x;
      double vp1 = fy * rz * vm2y;
      double vp2 = -(fx * x * vm2x + fy * y * vm2y) * rz2;
      if (x_in) vp0 += -fx * rz2 * vJ02; else vp2 += -fx * rz3 * vJ02 * tx;
      if (y_in) vp1 += -fy * rz2 * vJ12; else vp2 += -fy * rz3 * vJ12 * ty;
      vp2 += -fx * rz2 * vJ00 - fy * rz2 * vJ11 + 2.0 * fx * tx * rz3 * vJ02 + 2.0 * fy * ty * rz3 * vJ12;
      vp2 += (double)gd;

      // to world: v_mean += R^T vp ; vS += R^T vSc R
#pragma unroll
      for (int i = 0; i < 3; ++i) vmean[i] += R[0 * 3 + i] * vp0 + R[1 * 3 + i] * vp1 + R[2 * 3 + i] * vp2;
      double tmp[9];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          tmp[i * 3 + j] = R[0 * 3 + i] * vSc[0 * 3 + j] + R[1 * 3 + i] * vSc[1 * 3 + j] + R[2 * 3 + i] * vSc[2 * 3 + j];
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          vS[i * 3 + j] += tmp[i * 3 + 0] * R[0 * 3 + j] + tmp[i * 3 + 1] * R[1 * 3 + j] + tmp[i * 3 + 2] * R[2 * 3 + j];
    }
  }

  // Sigma = M M^T : v_M = (vS + vS^T) M ; M = Rq diag(s)
  double Rq[9], M[9];
  quat_to_rotmat_d(q.x, q.y, q.z, q.w, Rq);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int k = 0; k < 3; ++k) M[r * 3 + k] = Rq[r * 3 + k] * sc[k];
  double vM[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      vM[i * 3 + j] = (vS[i * 3 + 0] + vS[0 * 3 + i]) * M[0 * 3 + j] + (vS[i * 3 + 1] + vS[1 * 3 + i]) * M[1 * 3 + j] +
                      (vS[i * 3 + 2] + vS[2 * 3 + i]) * M[2 * 3 + j];
  double vR[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) vR[i * 3 + j] = vM[i * 3 + j] * sc[j];
#pragma unroll
  for (int j = 0; j < 3; ++j)
    v_scales[(size_t)gg * 3 + j] = (float)(Rq[0 * 3 + j] * vM[0 * 3 + j] + Rq[1 * 3 + j] * vM[1 * 3 + j] + Rq[2 * 3 + j] * vM[2 * 3 + j]);

  const double inv = 1.0 / sqrt(((q.y * q.y + q.z * q.z) + q.w * q.w) + q.x * q.x);
  const double w = q.x * inv, xq = q.y * inv, yq = q.z * inv, zq = q.w * inv;
  double vq[4];
  vq[0] = 2.0 * (-zq * vR[1] + yq * vR[2] + zq * vR[3] - xq * vR[5] - yq * vR[6] + xq * vR[7]);
  vq[1] = 2.0 * (yq * vR[1] + zq * vR[2] + yq * vR[3] - 2.0 * xq * vR[4] - w * vR[5] + zq * vR[6] + w * vR[7] - 2.0 * xq * vR[8]);
  vq[2] = 2.0 * (-2.0 * yq * vR[0] + xq * vR[1] + w * vR[2] + xq * vR[3] + zq * vR[5] - w * vR[6] + zq * vR[7] - 2.0 * yq * vR[8]);
  vq[3] = 2.0 * (-2.0 * zq * vR[0] - w * vR[1] + xq * vR[2] + w * vR[3] - 2.0 * zq * vR[4] + yq * vR[5] + xq * vR[6] + yq * vR[7]);
  const double dotp = vq[0] * w + vq[1] * xq + vq[2] * yq + vq[3] * zq;
  reinterpret_cast<float4 *>(v_quats)[gg] =
      make_float4((float)((vq[0] - dotp * w) * inv), (float)((vq[1] - dotp * xq) * inv), (float)((vq[2] - dotp * yq) * inv), (float)((vq[3] - dotp * zq) * inv));
  v_means[(size_t)gg * 3 + 0] = (float)vmean[0];
  v_means[(size_t)gg * 3 + 1] = (float)vmean[1];
  v_means[(size_t)gg * 3 + 2] = (float)vmean[2];
}

}  // namespace pm

extern "C" {

int pm_project_bwd(const float *means, const float *quats, const float *scales, const float *viewmats,
                   const float *Ks, int N, int C, int views_per_scene, int width, int height, float eps2d,
                   const int32_t *radii, const float *conics, const float *v_means2d,
                   const float *v_depths, const float *v_conics, float *v_means, float *v_quats,
                   float *v_scales, const uint8_t *touched, pm_stream_t stream) {
  (void)eps2d;  // the blur is an additive constant: no gradient term
  if (!means || !quats || !scales || !viewmats || !Ks || !radii || !conics || !v_means || !v_quats || !v_scales)
    return PM_ERR_BAD_ARG;
  if (N < 0 || C < 0 || width <= 0 || height <= 0) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(quats) & 15) != 0 || (reinterpret_cast<uintptr_t>(v_quats) & 15) != 0)
    return PM_ERR_BAD_ARG;
  if ((int64_t)N * C >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  if (N == 0) return PM_OK;
  if (C > pm::kBwdMaxCams) return PM_ERR_RANGE;
  const int Nv = views_per_scene > 0 ? views_per_scene : (C > 0 ? C : 1);
  if (C % Nv != 0) return PM_ERR_BAD_ARG;
  const int64_t total = (int64_t)N * (C > 0 ? C / Nv : 1);
  const size_t cam_bytes = sizeof(pm::BwdCam) * (size_t)(C > 0 ? C : 1);
  if (cam_bytes > 48 * 1024) {
    // per device and cheap: set on every call (a process may drive several GPUs)
    if (cudaFuncSetAttribute(pm::project_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(sizeof(pm::BwdCam) * pm::kBwdMaxCams)) != cudaSuccess) {
      pm::set_last_cuda_error(cudaGetLastError());
      return PM_ERR_CUDA;
    }
  }
  pm::project_bwd_kernel<<<(unsigned)((total + 127) / 128), 128, cam_bytes, pm::as_stream(stream)>>>(
      means, quats, scales, viewmats, Ks, N, C, Nv, width, height, radii, conics, v_means2d, v_depths,
      v_conics, v_means, v_quats, v_scales, touched);
  return pm::check_launch();
}

}  // extern "C"
