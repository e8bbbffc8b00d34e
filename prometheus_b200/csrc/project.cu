// project.cu — stage 2: EWA projection of 3D Gaussians, forward (with the fused
// per-pair tile count) and backward.
//
// Replaces gsplat 1.4.0 fully_fused_projection_{fwd,bwd} and pass 1 of
// isect_tiles, which the reference reaches through gsplat.rasterization
// (prometheus/modules/renderers/gaussians_renderer.py:413-427).
//
// THIS FILE IS COMPILED WITH -fmad=false. radii, tile counts and the depth bits
// that enter the sort keys are integer functions of the floats computed here and
// must match the oracle (oracle/gs_oracle.c, -ffp-contract=off) bit for bit, so
// every operation is one IEEE-754 rounding in the order written (three-term sums
// are (a+b)+c; nvcc's default -prec-div/-prec-sqrt give correctly rounded / and
// sqrt, and denormals are kept). The kernel is HBM-bound (40 B/Gaussian read,
// 32 B/pair written for ~250 flops), so giving up FMA contraction costs nothing.
#include "common.cuh"

namespace pm {

struct ProjCam {
  float R[9];
  float t[3];
  float fx, fy, cx, cy;
};

// clamp limits of the perspective Jacobian; per camera, so computed once per block
struct ProjLims {
  float x_pos, x_neg, y_pos, y_neg;
};
__device__ __forceinline__ ProjLims proj_lims(const ProjCam &c, int width, int height) {
  const float tan_fovx = 0.5f * (float)width / c.fx;
  const float tan_fovy = 0.5f * (float)height / c.fy;
  ProjLims l;
  l.x_pos = ((float)width - c.cx) / c.fx + 0.3f * tan_fovx;
  l.x_neg = c.cx / c.fx + 0.3f * tan_fovx;
  l.y_pos = ((float)height - c.cy) / c.fy + 0.3f * tan_fovy;
  l.y_neg = c.cy / c.fy + 0.3f * tan_fovy;
  return l;
}

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

__device__ __forceinline__ float sum3(float a, float b, float c) { return (a + b) + c; }

// quat (wxyz, un-normalised) -> rotation matrix (row-major)
__device__ __forceinline__ void quat_to_rotmat(float w, float x, float y, float z, float *R) {
  const float inv_norm = 1.0f / sqrtf(((x * x + y * y) + z * z) + w * w);
  x *= inv_norm; y *= inv_norm; z *= inv_norm; w *= inv_norm;
  const float x2 = x * x, y2 = y * y, z2 = z * z;
  const float xy = x * y, xz = x * z, yz = y * z;
  const float wx = w * x, wy = w * y, wz = w * z;
  R[0] = 1.0f - 2.0f * (y2 + z2); R[1] = 2.0f * (xy - wz);        R[2] = 2.0f * (xz + wy);
  R[3] = 2.0f * (xy + wz);        R[4] = 1.0f - 2.0f * (x2 + z2); R[5] = 2.0f * (yz - wx);
  R[6] = 2.0f * (xz - wy);        R[7] = 2.0f * (yz + wx);        R[8] = 1.0f - 2.0f * (x2 + y2);
}

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

__global__ void __launch_bounds__(256)
project_fwd_kernel(const float *__restrict__ means, const float *__restrict__ quats,
                   const float *__restrict__ scales, const float *__restrict__ viewmats,
                   const float *__restrict__ Ks, int N, int width, int height,
                   int tile_width, int tile_height, float eps2d, float near_plane, float far_plane,
                   float radius_clip, int32_t *__restrict__ radii, float *__restrict__ means2d,
                   float *__restrict__ depths, float *__restrict__ conics,
                   int32_t *__restrict__ tiles_per_gauss) {
  __shared__ ProjLims s_lims;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int cid = blockIdx.y;
  const ProjCam c = load_proj_cam(viewmats + (size_t)cid * 16, Ks + (size_t)cid * 9);
  if (threadIdx.x == 0) s_lims = proj_lims(c, width, height);  // 6 IEEE divisions, once per block
  __syncthreads();
  const bool in_range = g < N;
  const size_t f = (size_t)cid * N + (in_range ? g : 0);
  const float *R = c.R;

  int radius_i = 0;
  float m2x = 0.0f, m2y = 0.0f, depth = 0.0f, q0 = 0.0f, q1 = 0.0f, q2 = 0.0f;
  do {
    if (!in_range) break;
    const float mx = means[(size_t)g * 3 + 0], my = means[(size_t)g * 3 + 1], mz = means[(size_t)g * 3 + 2];
    const float px = sum3(R[0] * mx, R[1] * my, R[2] * mz) + c.t[0];
    const float py = sum3(R[3] * mx, R[4] * my, R[5] * mz) + c.t[1];
    const float pz = sum3(R[6] * mx, R[7] * my, R[8] * mz) + c.t[2];
    if (pz < near_plane || pz > far_plane) break;

    const float4 q = reinterpret_cast<const float4 *>(quats)[g];
    const float s0 = scales[(size_t)g * 3 + 0], s1 = scales[(size_t)g * 3 + 1], s2 = scales[(size_t)g * 3 + 2];
    float Rq[9], M[9];
    quat_to_rotmat(q.x, q.y, q.z, q.w, Rq);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      M[r * 3 + 0] = Rq[r * 3 + 0] * s0;
      M[r * 3 + 1] = Rq[r * 3 + 1] * s1;
      M[r * 3 + 2] = Rq[r * 3 + 2] * s2;
    }
    float S[9];
    S[0] = sum3(M[0] * M[0], M[1] * M[1], M[2] * M[2]);
    S[1] = sum3(M[0] * M[3], M[1] * M[4], M[2] * M[5]);
    S[2] = sum3(M[0] * M[6], M[1] * M[7], M[2] * M[8]);
    S[4] = sum3(M[3] * M[3], M[4] * M[4], M[5] * M[5]);
    S[5] = sum3(M[3] * M[6], M[4] * M[7], M[5] * M[8]);
    S[8] = sum3(M[6] * M[6], M[7] * M[7], M[8] * M[8]);
    S[3] = S[1]; S[6] = S[2]; S[7] = S[5];
    float T[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        T[r * 3 + k] = sum3(R[r * 3 + 0] * S[0 * 3 + k], R[r * 3 + 1] * S[1 * 3 + k], R[r * 3 + 2] * S[2 * 3 + k]);
    const float c00 = sum3(T[0] * R[0], T[1] * R[1], T[2] * R[2]);
    const float c01 = sum3(T[0] * R[3], T[1] * R[4], T[2] * R[5]);
    const float c02 = sum3(T[0] * R[6], T[1] * R[7], T[2] * R[8]);
    const float c11 = sum3(T[3] * R[3], T[4] * R[4], T[5] * R[5]);
    const float c12 = sum3(T[3] * R[6], T[4] * R[7], T[5] * R[8]);
    const float c22 = sum3(T[6] * R[6], T[7] * R[7], T[8] * R[8]);

    const float fx = c.fx, fy = c.fy, cx = c.cx, cy = c.cy;
    const float lim_x_pos = s_lims.x_pos, lim_x_neg = s_lims.x_neg;
    const float lim_y_pos = s_lims.y_pos, lim_y_neg = s_lims.y_neg;
    const float rz = 1.0f / pz;
    const float rz2 = rz * rz;
    const float tx = pz * fminf(lim_x_pos, fmaxf(-lim_x_neg, px * rz));
    const float ty = pz * fminf(lim_y_pos, fmaxf(-lim_y_neg, py * rz));
    const float J00 = fx * rz, J02 = -((fx * tx) * rz2);
    const float J11 = fy * rz, J12 = -((fy * ty) * rz2);
    const float A00 = J00 * c00 + J02 * c02;
    const float A01 = J00 * c01 + J02 * c12;
    const float A02 = J00 * c02 + J02 * c22;
    const float A11 = J11 * c11 + J12 * c12;
    const float A12 = J11 * c12 + J12 * c22;
    float v00 = A00 * J00 + A02 * J02;
    const float v01 = A01 * J11 + A02 * J12;
    float v11 = A11 * J11 + A12 * J12;
    const float ux = (fx * px) * rz + cx;
    const float uy = (fy * py) * rz + cy;

    v00 += eps2d;
    v11 += eps2d;
    const float det = v00 * v11 - v01 * v01;
    if (!(det > 0.0f)) break;
    const float inv_det = 1.0f / det;
    const float b = 0.5f * (v00 + v11);
    const float v1 = b + sqrtf(fmaxf(0.01f, b * b - det));
    const float radius = ceilf(3.0f * sqrtf(v1));
    if (radius <= radius_clip) break;
    if (ux + radius <= 0.0f || ux - radius >= (float)width || uy + radius <= 0.0f ||
        uy - radius >= (float)height)
      break;
    radius_i = radius > 2147483520.0f ? 2147483520 : (int)radius;
    m2x = ux; m2y = uy; depth = pz;
    q0 = v11 * inv_det; q1 = -v01 * inv_det; q2 = v00 * inv_det;
  } while (false);

  int x0 = 0, x1 = 0, y0 = 0, y1 = 0;
  if (radius_i > 0) tile_aabb(m2x, m2y, radius_i, tile_width, tile_height, x0, x1, y0, y1);
  if (in_range) {
    radii[f] = radius_i;
    reinterpret_cast<float2 *>(means2d)[f] = make_float2(m2x, m2y);
    depths[f] = depth;
    conics[f * 3 + 0] = q0; conics[f * 3 + 1] = q1; conics[f * 3 + 2] = q2;
    if (tiles_per_gauss != nullptr) tiles_per_gauss[f] = (x1 - x0) * (y1 - y0);
  }
}

// Backward: one thread per Gaussian loops over the cameras and accumulates in
// registers — no atomics, deterministic. Arithmetic in fp64 (the VJP chain multiplies by
// (f/z)^2 ~ 1e4 and cancels; fp64 keeps the kernel's own rounding out of the gradient error, and the
// kernel stays HBM-bound on B200). Restates fully_fused_projection_bwd
// ("VJPs of steps 5->1"); float tolerance parity only.
__global__ void __launch_bounds__(128)
project_bwd_kernel(const float *__restrict__ means, const float *__restrict__ quats,
                   const float *__restrict__ scales, const float *__restrict__ viewmats,
                   const float *__restrict__ Ks, int N, int C, int width, int height,
                   const int32_t *__restrict__ radii, const float *__restrict__ conics,
                   const float *__restrict__ v_means2d, const float *__restrict__ v_depths,
                   const float *__restrict__ v_conics, float *__restrict__ v_means,
                   float *__restrict__ v_quats, float *__restrict__ v_scales) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  const double mx = means[(size_t)g * 3 + 0], my = means[(size_t)g * 3 + 1], mz = means[(size_t)g * 3 + 2];
  const float4 q4 = reinterpret_cast<const float4 *>(quats)[g];
  const double4 q = make_double4(q4.x, q4.y, q4.z, q4.w);
  const double sc[3] = {scales[(size_t)g * 3 + 0], scales[(size_t)g * 3 + 1], scales[(size_t)g * 3 + 2]};
  double Rq[9], M[9], S[9];
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

  double vmean[3] = {0, 0, 0};
  double vS[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};  // d loss / d Sigma (world), summed over cameras

  for (int cid = 0; cid < C; ++cid) {
    const size_t f = (size_t)cid * N + g;
    if (radii[f] <= 0) continue;
    const ProjCam cf = load_proj_cam(viewmats + (size_t)cid * 16, Ks + (size_t)cid * 9);
    double R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = cf.R[i];
    const double ct[3] = {cf.t[0], cf.t[1], cf.t[2]};
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
    const double q00 = conics[f * 3 + 0], q01 = conics[f * 3 + 1], q11 = conics[f * 3 + 2];
    double a00 = 0.0, a01 = 0.0, a11 = 0.0;
    if (v_conics) { a00 = v_conics[f * 3 + 0]; a01 = 0.5 * v_conics[f * 3 + 1]; a11 = v_conics[f * 3 + 2]; }
    const double t00 = q00 * a00 + q01 * a01, t01 = q00 * a01 + q01 * a11;
    const double t10 = q01 * a00 + q11 * a01, t11 = q01 * a01 + q11 * a11;
    const double G00 = -(t00 * q00 + t01 * q01);
    const double G01 = -(t00 * q01 + t01 * q11);
    const double G10 = -(t10 * q00 + t11 * q01);
    const double G11 = -(t10 * q01 + t11 * q11);

    const double fx = cf.fx, fy = cf.fy, cx = cf.cx, cy = cf.cy;
    const double tan_fovx = 0.5 * (double)width / fx, tan_fovy = 0.5 * (double)height / fy;
    const double lim_x_pos = ((double)width - cx) / fx + 0.3 * tan_fovx, lim_x_neg = cx / fx + 0.3 * tan_fovx;
    const double lim_y_pos = ((double)height - cy) / fy + 0.3 * tan_fovy, lim_y_neg = cy / fy + 0.3 * tan_fovy;
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

    double vm2x = 0.0, vm2y = 0.0;
    if (v_means2d) { vm2x = v_means2d[f * 2 + 0]; vm2y = v_means2d[f * 2 + 1]; }
    double vp0 = fx * rz * vm2x;
    double vp1 = fy * rz * vm2y;
    double vp2 = -(fx * x * vm2x + fy * y * vm2y) * rz2;
    if (x_in) vp0 += -fx * rz2 * vJ02; else vp2 += -fx * rz3 * vJ02 * tx;
    if (y_in) vp1 += -fy * rz2 * vJ12; else vp2 += -fy * rz3 * vJ12 * ty;
    vp2 += -fx * rz2 * vJ00 - fy * rz2 * vJ11 + 2.0 * fx * tx * rz3 * vJ02 + 2.0 * fy * ty * rz3 * vJ12;
    if (v_depths) vp2 += v_depths[f];

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

  // Sigma = M M^T : v_M = (vS + vS^T) M ; M = Rq diag(s)
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
    v_scales[(size_t)g * 3 + j] = (float)(Rq[0 * 3 + j] * vM[0 * 3 + j] + Rq[1 * 3 + j] * vM[1 * 3 + j] + Rq[2 * 3 + j] * vM[2 * 3 + j]);

  const double inv = 1.0 / sqrt(((q.y * q.y + q.z * q.z) + q.w * q.w) + q.x * q.x);
  const double w = q.x * inv, xq = q.y * inv, yq = q.z * inv, zq = q.w * inv;
  double vq[4];
  vq[0] = 2.0 * (-zq * vR[1] + yq * vR[2] + zq * vR[3] - xq * vR[5] - yq * vR[6] + xq * vR[7]);
  vq[1] = 2.0 * (yq * vR[1] + zq * vR[2] + yq * vR[3] - 2.0 * xq * vR[4] - w * vR[5] + zq * vR[6] + w * vR[7] - 2.0 * xq * vR[8]);
  vq[2] = 2.0 * (-2.0 * yq * vR[0] + xq * vR[1] + w * vR[2] + xq * vR[3] + zq * vR[5] - w * vR[6] + zq * vR[7] - 2.0 * yq * vR[8]);
  vq[3] = 2.0 * (-2.0 * zq * vR[0] - w * vR[1] + xq * vR[2] + w * vR[3] - 2.0 * zq * vR[4] + yq * vR[5] + xq * vR[6] + yq * vR[7]);
  const double dotp = vq[0] * w + vq[1] * xq + vq[2] * yq + vq[3] * zq;
  reinterpret_cast<float4 *>(v_quats)[g] =
      make_float4((float)((vq[0] - dotp * w) * inv), (float)((vq[1] - dotp * xq) * inv), (float)((vq[2] - dotp * yq) * inv), (float)((vq[3] - dotp * zq) * inv));
  v_means[(size_t)g * 3 + 0] = (float)vmean[0];
  v_means[(size_t)g * 3 + 1] = (float)vmean[1];
  v_means[(size_t)g * 3 + 2] = (float)vmean[2];
}

}  // namespace pm

extern "C" {

int pm_project_fwd(const float *means, const float *quats, const float *scales, const float *viewmats,
                   const float *Ks, int N, int C, int width, int height, int tile_size, float eps2d,
                   float near_plane, float far_plane, float radius_clip, int32_t *radii, float *means2d,
                   float *depths, float *conics, int32_t *tiles_per_gauss, pm_stream_t stream) {
  if (!means || !quats || !scales || !viewmats || !Ks || !radii || !means2d || !depths || !conics)
    return PM_ERR_BAD_ARG;
  if (N < 0 || C < 0 || width <= 0 || height <= 0) return PM_ERR_BAD_ARG;
  if (tile_size != pm::kTile) return PM_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(quats) & 15) != 0 || (reinterpret_cast<uintptr_t>(means2d) & 7) != 0)
    return PM_ERR_BAD_ARG;
  if (C > 65535 || (int64_t)N * C >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  if (N == 0 || C == 0) return PM_OK;
  const int tw = (width + pm::kTile - 1) / pm::kTile, th = (height + pm::kTile - 1) / pm::kTile;
  dim3 grid((N + 255) / 256, C);
  pm::project_fwd_kernel<<<grid, 256, 0, pm::as_stream(stream)>>>(
      means, quats, scales, viewmats, Ks, N, width, height, tw, th, eps2d, near_plane, far_plane,
      radius_clip, radii, means2d, depths, conics, tiles_per_gauss);
  return pm::check_launch();
}

int pm_project_bwd(const float *means, const float *quats, const float *scales, const float *viewmats,
                   const float *Ks, int N, int C, int width, int height, float eps2d,
                   const int32_t *radii, const float *conics, const float *v_means2d,
                   const float *v_depths, const float *v_conics, float *v_means, float *v_quats,
                   float *v_scales, pm_stream_t stream) {
  (void)eps2d;  // the blur is an additive constant: no gradient term
  if (!means || !quats || !scales || !viewmats || !Ks || !radii || !conics || !v_means || !v_quats || !v_scales)
    return PM_ERR_BAD_ARG;
  if (N < 0 || C < 0 || width <= 0 || height <= 0) return PM_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(quats) & 15) != 0 || (reinterpret_cast<uintptr_t>(v_quats) & 15) != 0)
    return PM_ERR_BAD_ARG;
  if ((int64_t)N * C >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  if (N == 0) return PM_OK;
  pm::project_bwd_kernel<<<(N + 127) / 128, 128, 0, pm::as_stream(stream)>>>(
      means, quats, scales, viewmats, Ks, N, C, width, height, radii, conics, v_means2d, v_depths,
      v_conics, v_means, v_quats, v_scales);
  return pm::check_launch();
}

}  // extern "C"
