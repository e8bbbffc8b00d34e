// project.cu — stage 2: EWA projection of 3D Gaussians, forward (with the fused
// per-pair tile count). The backward lives in project_bwd.cu (compiled with FMA contraction).
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
#include "binning_common.cuh"

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

// One camera's constants, staged in shared memory once per CTA.
struct ProjCamS {
  ProjCam c;
  ProjLims lims;
};
constexpr int kProjCamChunk = 32;  // cameras per launch (shared-memory table); more cameras = more launches

// One thread per Gaussian, looping over the cameras of the launch: the Gaussian is read from HBM once and
// its world covariance is computed once (the per-camera grid of the first version re-read 40 B/Gaussian
// and redid the quaternion -> covariance arithmetic for every camera: 662 MB of DRAM reads for 84 MB of
// input on C2, profiles/r1e_*). COUNT additionally accumulates the per-(camera, tile) intersection
// histogram (what pm_bin_count computes from the stored radii / means2d) while the tile AABB is still in
// registers, through the per-CTA privatised histogram of binning_common.cuh.
#ifndef PM_PROJ_MINBLOCKS
#define PM_PROJ_MINBLOCKS 5  // 48 registers: 0.320 -> 0.316 ms on C2 (6: 0.318)
#endif
// (usable_tiles: common.cuh — shared with the scatter pass of the two-pass binning, tile_sort.cu)
// MODE 0: projection only; 1: + per-(camera, tile) histogram; 2: + single-pass binning (histogram and the
// intersection records, each written into its tile's fixed-capacity region — binning_common.cuh)
template <int MODE>
__global__ void __launch_bounds__(256, PM_PROJ_MINBLOCKS)
project_fwd_kernel(const float *__restrict__ means, const float *__restrict__ quats,
                   const float *__restrict__ scales, const float *__restrict__ viewmats,
                   const float *__restrict__ Ks, int N, int views_per_scene, int cam0, int n_cams, int chunk_stride,
                   int width, int height,
                   int tile_width, int tile_height, float eps2d, float near_plane, float far_plane,
                   float radius_clip, int32_t *__restrict__ radii, float *__restrict__ means2d,
                   float *__restrict__ depths, float *__restrict__ conics,
                   int32_t *__restrict__ tiles_per_gauss, int32_t *__restrict__ tile_counts,
                   uint64_t *__restrict__ records, int tile_cap, const float *__restrict__ opacities) {
  __shared__ ProjCamS s_cam[kProjCamChunk];
  __shared__ TileCountSmem s_cnt[1];     // MODE 1 (unused shared arrays are optimised away)
  __shared__ TileScatterSmem s_sct[1];   // MODE 2
  // strided CTA order (see binning_common.cuh): concurrently resident CTAs touch far-apart tiles
  const int chunk = strided_chunk(blockIdx.x, chunk_stride, gridDim.x);
  // scene batching: the Gaussian arrays hold S scenes of N Gaussians, camera c belongs to scene c / views_per_scene
  // and sees only that scene. With S > 1 the launcher guarantees N % 256 == 0, so a CTA never straddles two scenes.
  const int chunks_per_scene = (N + 255) / 256;
  const int scene = chunk / chunks_per_scene;
  const int g = (chunk - scene * chunks_per_scene) * blockDim.x + threadIdx.x;  // Gaussian index inside the scene
  const bool in_range = g < N;
  const size_t gg = (size_t)scene * N + (in_range ? g : 0);                     // index into the Gaussian arrays
  const int cam_first = scene * views_per_scene + cam0;                         // first camera of this launch, absolute
  for (int i = threadIdx.x; i < n_cams; i += blockDim.x) {
    ProjCamS pc;
    pc.c = load_proj_cam(viewmats + (size_t)(cam_first + i) * 16, Ks + (size_t)(cam_first + i) * 9);
    pc.lims = proj_lims(pc.c, width, height);
    s_cam[i] = pc;
  }
  __syncthreads();

  // opacities != NULL (binning modes only): USABLE tiles — see usable_tiles() above
  const float opac = (MODE >= 1 && opacities != nullptr && in_range) ? __ldg(opacities + gg) : 0.0f;
  float mx = 0.0f, my = 0.0f, mz = 0.0f;
  float S[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) S[i] = 0.0f;
  if (in_range) {
    mx = means[gg * 3 + 0]; my = means[gg * 3 + 1]; mz = means[gg * 3 + 2];
    const float4 q = reinterpret_cast<const float4 *>(quats)[gg];
    const float s0 = scales[gg * 3 + 0], s1 = scales[gg * 3 + 1], s2 = scales[gg * 3 + 2];
    float Rq[9], M[9];
    quat_to_rotmat(q.x, q.y, q.z, q.w, Rq);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      M[r * 3 + 0] = Rq[r * 3 + 0] * s0;
      M[r * 3 + 1] = Rq[r * 3 + 1] * s1;
      M[r * 3 + 2] = Rq[r * 3 + 2] * s2;
    }
    S[0] = sum3(M[0] * M[0], M[1] * M[1], M[2] * M[2]);
    S[1] = sum3(M[0] * M[3], M[1] * M[4], M[2] * M[5]);
    S[2] = sum3(M[0] * M[6], M[1] * M[7], M[2] * M[8]);
    S[4] = sum3(M[3] * M[3], M[4] * M[4], M[5] * M[5]);
    S[5] = sum3(M[3] * M[6], M[4] * M[7], M[5] * M[8]);
    S[8] = sum3(M[6] * M[6], M[7] * M[7], M[8] * M[8]);
    S[3] = S[1]; S[6] = S[2]; S[7] = S[5];
  }

  // MODE 2 walks the cameras in batches of kWinCams: one window round (count, reserve, place) per batch
  const int n_tiles = tile_width * tile_height;
  for (int c0 = 0; c0 < n_cams; c0 += (MODE == 2 ? kWinCams : kProjCamChunk)) {
  const int c1 = MODE == 2 ? min(n_cams, c0 + kWinCams) : n_cams;
  if (MODE == 2) scatter_batch_begin(s_sct[0], c1 - c0);
  for (int ci = c0; ci < c1; ++ci) {
    const ProjCam &c = s_cam[ci].c;
    const float *R = c.R;
    const size_t f = (size_t)(cam_first + ci) * N + (in_range ? g : 0);
    int radius_i = 0;
    float m2x = 0.0f, m2y = 0.0f, depth = 0.0f, q0 = 0.0f, q1 = 0.0f, q2 = 0.0f;
    do {
      if (!in_range) break;
      const float px = sum3(R[0] * mx, R[1] * my, R[2] * mz) + c.t[0];
      const float py = sum3(R[3] * mx, R[4] * my, R[5] * mz) + c.t[1];
      const float pz = sum3(R[6] * mx, R[7] * my, R[8] * mz) + c.t[2];
      if (pz < near_plane || pz > far_plane) break;
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
      const ProjLims &l = s_cam[ci].lims;
      const float rz = 1.0f / pz;
      const float rz2 = rz * rz;
      const float tx = pz * fminf(l.x_pos, fmaxf(-l.x_neg, px * rz));
      const float ty = pz * fminf(l.y_pos, fmaxf(-l.y_neg, py * rz));
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
    if (MODE >= 1 && opacities != nullptr && radius_i > 0) usable_tiles(m2x, m2y, q0, q1, q2, opac, x0, x1, y0, y1);
    if (in_range) {
      radii[f] = radius_i;
      reinterpret_cast<float2 *>(means2d)[f] = make_float2(m2x, m2y);
      depths[f] = depth;
      conics[f * 3 + 0] = q0; conics[f * 3 + 1] = q1; conics[f * 3 + 2] = q2;
      if (tiles_per_gauss != nullptr) tiles_per_gauss[f] = (x1 - x0) * (y1 - y0);
    }
    if (MODE == 1)
      cta_tile_count(s_cnt[0], x0, x1, y0, y1, tile_counts + (size_t)(cam_first + ci) * tile_width * tile_height, tile_width);
    if (MODE == 2) {
      const size_t tile0 = (size_t)(cam_first + ci) * n_tiles;
      scatter_batch_add(s_sct[0], ci - c0, x0, x1, y0, y1, __float_as_uint(depth), (uint32_t)f, tile_counts + tile0,
                        tile_width, records + tile0 * (size_t)tile_cap, tile_cap);
    }
  }
  if (MODE == 2) {
    const size_t tile0 = (size_t)(cam_first + c0) * n_tiles;
    scatter_batch_end(s_sct[0], c1 - c0, tile_counts + tile0, tile_width, n_tiles, records + tile0 * (size_t)tile_cap, tile_cap,
                      (uint32_t)((size_t)(cam_first + c0) * N + (in_range ? g : 0)), (uint32_t)N);
  }
  }
}

}  // namespace pm

extern "C" {

static int project_fwd_launch(const float *means, const float *quats, const float *scales, const float *viewmats,
                              const float *Ks, int N, int C, int views_per_scene, int width, int height, int tile_size,
                              float eps2d,
                              float near_plane, float far_plane, float radius_clip, int32_t *radii, float *means2d,
                              float *depths, float *conics, int32_t *tiles_per_gauss, int32_t *tile_counts,
                              uint64_t *records, int tile_cap, const float *opacities, pm_stream_t stream) {
  if (!means || !quats || !scales || !viewmats || !Ks || !radii || !means2d || !depths || !conics)
    return PM_ERR_BAD_ARG;
  if (N < 0 || C < 0 || width <= 0 || height <= 0) return PM_ERR_BAD_ARG;
  if (tile_size != pm::kTile) return PM_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(quats) & 15) != 0 || (reinterpret_cast<uintptr_t>(means2d) & 7) != 0)
    return PM_ERR_BAD_ARG;
  if (C > 65535 || (int64_t)N * C >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  if (N == 0 || C == 0) return PM_OK;
  const int Nv = views_per_scene > 0 ? views_per_scene : C;
  if (C % Nv != 0) return PM_ERR_BAD_ARG;
  const int n_scenes = C / Nv;
  if (n_scenes > 1 && N % 256 != 0) return PM_ERR_UNSUPPORTED;  // a 256-Gaussian CTA must not straddle two scenes
  const int tw = (width + pm::kTile - 1) / pm::kTile, th = (height + pm::kTile - 1) / pm::kTile;
  const int nb = n_scenes * ((N + 255) / 256);
  const int stride = pm::coprime_stride(nb);
  for (int cam0 = 0; cam0 < Nv; cam0 += pm::kProjCamChunk) {
    const int n_cams = Nv - cam0 < pm::kProjCamChunk ? Nv - cam0 : pm::kProjCamChunk;
    if (tile_counts && records)
      pm::project_fwd_kernel<2><<<nb, 256, 0, pm::as_stream(stream)>>>(
          means, quats, scales, viewmats, Ks, N, Nv, cam0, n_cams, stride, width, height, tw, th, eps2d, near_plane,
          far_plane, radius_clip, radii, means2d, depths, conics, tiles_per_gauss, tile_counts, records, tile_cap, opacities);
    else if (tile_counts)
      pm::project_fwd_kernel<1><<<nb, 256, 0, pm::as_stream(stream)>>>(
          means, quats, scales, viewmats, Ks, N, Nv, cam0, n_cams, stride, width, height, tw, th, eps2d, near_plane,
          far_plane, radius_clip, radii, means2d, depths, conics, tiles_per_gauss, tile_counts, nullptr, 0, opacities);
    else
      pm::project_fwd_kernel<0><<<nb, 256, 0, pm::as_stream(stream)>>>(
          means, quats, scales, viewmats, Ks, N, Nv, cam0, n_cams, stride, width, height, tw, th, eps2d, near_plane,
          far_plane, radius_clip, radii, means2d, depths, conics, tiles_per_gauss, nullptr, nullptr, 0, nullptr);
  }
  return pm::check_launch();
}

int pm_project_fwd(const float *means, const float *quats, const float *scales, const float *viewmats,
                   const float *Ks, int N, int C, int width, int height, int tile_size, float eps2d,
                   float near_plane, float far_plane, float radius_clip, int32_t *radii, float *means2d,
                   float *depths, float *conics, int32_t *tiles_per_gauss, pm_stream_t stream) {
  return project_fwd_launch(means, quats, scales, viewmats, Ks, N, C, C, width, height, tile_size, eps2d, near_plane,
                            far_plane, radius_clip, radii, means2d, depths, conics, tiles_per_gauss, nullptr, nullptr, 0,
                            nullptr, stream);
}

int pm_project_count_fwd(const float *means, const float *quats, const float *scales, const float *viewmats,
                         const float *Ks, int N, int C, int views_per_scene, int width, int height, int tile_size, float eps2d,
                         float near_plane, float far_plane, float radius_clip, int32_t *radii, float *means2d,
                         float *depths, float *conics, int32_t *tiles_per_gauss, int32_t *tile_counts,
                         pm_stream_t stream) {
  if (!tile_counts) return PM_ERR_BAD_ARG;
  return project_fwd_launch(means, quats, scales, viewmats, Ks, N, C, views_per_scene, width, height, tile_size, eps2d, near_plane,
                            far_plane, radius_clip, radii, means2d, depths, conics, tiles_per_gauss, tile_counts,
                            nullptr, 0, nullptr, stream);
}

int pm_project_scatter_fwd(const float *means, const float *quats, const float *scales, const float *viewmats,
                           const float *Ks, int N, int C, int views_per_scene, int width, int height, int tile_size,
                           float eps2d, float near_plane, float far_plane, float radius_clip, int32_t *radii,
                           float *means2d, float *depths, float *conics, int32_t *tile_counts, uint64_t *records,
                           int tile_cap, const float *opacities, pm_stream_t stream) {
  if (!tile_counts) return PM_ERR_BAD_ARG;
  if (records ? tile_cap <= 0 : tile_cap != 0) return PM_ERR_BAD_ARG;  // records == NULL, tile_cap == 0: histogram only
  {  // record indices inside one camera's regions are formed in 32 bits
    const int64_t tw = (width + pm::kTile - 1) / pm::kTile, th = (height + pm::kTile - 1) / pm::kTile;
    if (width > 0 && height > 0 && tw * th * (int64_t)tile_cap >= ((int64_t)1 << 31)) return PM_ERR_RANGE;
  }
  return project_fwd_launch(means, quats, scales, viewmats, Ks, N, C, views_per_scene, width, height, tile_size, eps2d,
                            near_plane, far_plane, radius_clip, radii, means2d, depths, conics, nullptr, tile_counts,
                            records, tile_cap, opacities, stream);
}

}  // extern "C"
