/*
 * gs_oracle.c — CPU oracle for the GS-decoder rendering hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under prometheus_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * PARITY UNPINNED: the arithmetic of stages 2-4 lives in the third-party wheel
 * gsplat==1.4.0 (reference README.md:57-59, requirements.txt:46), which is not
 * vendored under /root/reference, not installed here and not downloadable. The
 * reference holds no test, golden vector or fixture for this path (SURVEY.md
 * §4). Stages 2-4 below therefore restate the PUBLISHED gsplat 1.4.0 algorithm
 * (fully_fused_projection, isect_tiles, DeviceRadixSort::SortPairs,
 * isect_offset_encode, rasterize_to_pixels fwd/bwd) anchored on the reference's
 * call site prometheus/modules/renderers/gaussians_renderer.py:413-427 and on
 * the spec in SURVEY.md §8c. Stage 1 (unproject) and the camera preparation
 * restate gaussians_renderer.py:210-287, :453-472, :139-169, :404-412 and ARE
 * pinned: tests/golden/ holds outputs of the reference module itself.
 *
 * Floating-point contract (this file is the spec the CUDA kernels mirror):
 *   - compiled with -ffp-contract=off: every * + - / sqrt is one correctly
 *     rounded IEEE-754 binary32 operation, evaluated in the order written;
 *   - three-term sums are (a + b) + c;
 *   - deviations from upstream that cannot be reproduced bit-for-bit on a CPU:
 *     quaternion normalisation uses 1/sqrt (upstream: rsqrtf, <=2 ulp);
 *     compositing uses expf (upstream: __expf).
 * Integer outputs (radii, tile counts, keys, order, offsets) are exact
 * functions of those floats, so the CUDA path must match them bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

/* ------------------------------------------------------------------------- *
 * Stage 1: unproject. Restates GaussianConverter.forward,
 * gaussians_renderer.py:210-287 (modes :258-285).
 * params [V,14,h,w]; cameras [V,18]; muls [14].
 * ------------------------------------------------------------------------- */
ORC_API void orc_unproject(const float *params, const float *cameras, const float *muls,
                           int mode, float z_near, float z_far, float s_min, float s_max,
                           int V, int h, int w,
                           float *xyz, float *rgb, float *opacity, float *scales, float *quats) {
  const long hw = (long)h * w;
#pragma omp parallel for schedule(static)
  for (long idx = 0; idx < (long)V * hw; ++idx) {
    const int v = (int)(idx / hw);
    const long pix = idx % hw;
    const float *cam = cameras + (long)v * 18;
    const float *p = params + (long)v * 14 * hw + pix;
    float ch[14];
    for (int k = 0; k < 14; ++k) ch[k] = p[(long)k * hw] * muls[k]; /* :214-215 */

    /* intrinsics rescaled to the map size, :224-227 */
    const float H = cam[16], W = cam[17];
    const float fx = cam[12] * (float)h / H, cx = cam[14] * (float)h / H;
    const float fy = cam[13] * (float)w / W, cy = cam[15] * (float)w / W;

    const float i = (float)(pix % w) + 0.5f; /* :231 */
    const float j = (float)(pix / w) + 0.5f; /* :232 */
    const float u = i / cx + ch[3];          /* :234 */
    const float vv = j / cy + ch[4];         /* :235 */
    const float zs = -1.0f;
    const float xs = -(u - 1.0f) * cx / fx * zs; /* :238 */
    const float ys = (vv - 1.0f) * cy / fy * zs; /* :239 */

    /* rays_d = normalize(dir @ R^T), :243 ; R = c2w[:3,:3], row-major 3x4 */
    float d[3];
    for (int r = 0; r < 3; ++r)
      d[r] = xs * cam[r * 4 + 0] + ys * cam[r * 4 + 1] + zs * cam[r * 4 + 2];
    float nrm = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    nrm = fmaxf(nrm, 1e-12f); /* F.normalize eps */
    d[0] /= nrm; d[1] /= nrm; d[2] /= nrm;

    float depth, op, sc[3], col[3], q[4];
    if (mode == 2) { /* director3d :258-263 */
      depth = ch[5] + 1.85f;
      for (int k = 0; k < 3; ++k) col[k] = ch[k] / (2.0f * 0.28209479177387814f);
      const float opacity_offset = logf(0.01f / (1.0f - 0.01f));
      op = sigmoidf_(ch[6] + opacity_offset);
      const float scales_offset = logf(1.0f / 100.0f);
      for (int k = 0; k < 3; ++k) sc[k] = expf(ch[7 + k] + scales_offset);
      q[0] = ch[10] + 1.0f; q[1] = ch[11]; q[2] = ch[12]; q[3] = ch[13];
    } else if (mode == 1) { /* gslrm :264-273 */
      const float d_ = sigmoidf_(ch[5]);
      depth = z_near * (1.0f - d_) + z_far * d_;
      for (int k = 0; k < 3; ++k) col[k] = ch[k];
      op = sigmoidf_(ch[6] - 2.0f);
      for (int k = 0; k < 3; ++k) {
        float s = expf(ch[7 + k] - 2.3f);
        sc[k] = fminf(fmaxf(s, 0.0f), 0.3f);
      }
      for (int k = 0; k < 4; ++k) q[k] = ch[10 + k];
    } else { /* mvsplat :274-283 */
      const float d_ = sigmoidf_(ch[5]);
      depth = z_near * (1.0f - d_) + z_far * d_;
      for (int k = 0; k < 3; ++k) {
        const float s_ = sigmoidf_(ch[7 + k]);
        sc[k] = s_min * (1.0f - s_) + s_max * s_;
      }
      for (int k = 0; k < 3; ++k) col[k] = sigmoidf_(ch[k]);
      op = sigmoidf_(ch[6]);
      for (int k = 0; k < 4; ++k) q[k] = ch[10 + k];
    }
    float qn = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    qn = fmaxf(qn, 1e-12f);
    for (int k = 0; k < 4; ++k) quats[idx * 4 + k] = q[k] / qn;
    for (int k = 0; k < 3; ++k) {
      xyz[idx * 3 + k] = cam[k * 4 + 3] + depth * d[k]; /* :286 */
      rgb[idx * 3 + k] = col[k];
      scales[idx * 3 + k] = sc[k];
    }
    opacity[idx] = op;
  }
}

/* ------------------------------------------------------------------------- *
 * Stage 2: EWA projection. Restates gsplat 1.4.0 fully_fused_projection_fwd
 *  with the op order fixed below.
 * ------------------------------------------------------------------------- */
static inline float sum3(float a, float b, float c) { return (a + b) + c; }

typedef struct {
  float R[9];  /* viewmat rotation, row-major */
  float t[3];
  float fx, fy, cx, cy;
} orc_cam;

static void load_cam(const float *viewmat, const float *K, orc_cam *c) {
  for (int r = 0; r < 3; ++r) {
    for (int k = 0; k < 3; ++k) c->R[r * 3 + k] = viewmat[r * 4 + k];
    c->t[r] = viewmat[r * 4 + 3];
  }
  c->fx = K[0]; c->fy = K[4]; c->cx = K[2]; c->cy = K[5];
}

/* quat (wxyz, un-normalised) -> rotation matrix, row-major */
static void quat_to_rotmat(const float *q, float *R) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  const float inv_norm = 1.0f / sqrtf(((x * x + y * y) + z * z) + w * w);
  x *= inv_norm; y *= inv_norm; z *= inv_norm; w *= inv_norm;
  const float x2 = x * x, y2 = y * y, z2 = z * z;
  const float xy = x * y, xz = x * z, yz = y * z;
  const float wx = w * x, wy = w * y, wz = w * z;
  R[0] = 1.0f - 2.0f * (y2 + z2); R[1] = 2.0f * (xy - wz);        R[2] = 2.0f * (xz + wy);
  R[3] = 2.0f * (xy + wz);        R[4] = 1.0f - 2.0f * (x2 + z2); R[5] = 2.0f * (yz - wx);
  R[6] = 2.0f * (xz - wy);        R[7] = 2.0f * (yz + wx);        R[8] = 1.0f - 2.0f * (x2 + y2);
}

/* Sigma = (R S)(R S)^T ; 6 unique entries s00 s01 s02 s11 s12 s22 */
static void quat_scale_to_covar(const float *q, const float *s, float *S6) {
  float R[9], M[9];
  quat_to_rotmat(q, R);
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) M[r * 3 + k] = R[r * 3 + k] * s[k];
  S6[0] = sum3(M[0] * M[0], M[1] * M[1], M[2] * M[2]);
  S6[1] = sum3(M[0] * M[3], M[1] * M[4], M[2] * M[5]);
  S6[2] = sum3(M[0] * M[6], M[1] * M[7], M[2] * M[8]);
  S6[3] = sum3(M[3] * M[3], M[4] * M[4], M[5] * M[5]);
  S6[4] = sum3(M[3] * M[6], M[4] * M[7], M[5] * M[8]);
  S6[5] = sum3(M[6] * M[6], M[7] * M[7], M[8] * M[8]);
}

/* one (camera, Gaussian) pair; returns radius (0 = culled) */
static int project_one(const float *mean, const float *quat, const float *scale,
                       const orc_cam *c, int width, int height,
                       float eps2d, float near_plane, float far_plane, float radius_clip,
                       float *mean2d, float *depth, float *conic) {
  const float *R = c->R;
  /* world -> camera */
  const float px = sum3(R[0] * mean[0], R[1] * mean[1], R[2] * mean[2]) + c->t[0];
  const float py = sum3(R[3] * mean[0], R[4] * mean[1], R[5] * mean[2]) + c->t[1];
  const float pz = sum3(R[6] * mean[0], R[7] * mean[1], R[8] * mean[2]) + c->t[2];
  if (pz < near_plane || pz > far_plane) return 0;

  float S6[6];
  quat_scale_to_covar(quat, scale, S6);
  const float S[9] = {S6[0], S6[1], S6[2], S6[1], S6[3], S6[4], S6[2], S6[4], S6[5]};
  /* Sigma_c = R Sigma R^T : T = R*Sigma, then T*R^T (6 unique) */
  float T[9];
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k)
      T[r * 3 + k] = sum3(R[r * 3 + 0] * S[0 * 3 + k], R[r * 3 + 1] * S[1 * 3 + k], R[r * 3 + 2] * S[2 * 3 + k]);
  const float c00 = sum3(T[0] * R[0], T[1] * R[1], T[2] * R[2]);
  const float c01 = sum3(T[0] * R[3], T[1] * R[4], T[2] * R[5]);
  const float c02 = sum3(T[0] * R[6], T[1] * R[7], T[2] * R[8]);
  const float c11 = sum3(T[3] * R[3], T[4] * R[4], T[5] * R[5]);
  const float c12 = sum3(T[3] * R[6], T[4] * R[7], T[5] * R[8]);
  const float c22 = sum3(T[6] * R[6], T[7] * R[7], T[8] * R[8]);

  /* perspective projection (persp_proj) */
  const float fx = c->fx, fy = c->fy, cx = c->cx, cy = c->cy;
  const float tan_fovx = 0.5f * (float)width / fx;
  const float tan_fovy = 0.5f * (float)height / fy;
  const float lim_x_pos = ((float)width - cx) / fx + 0.3f * tan_fovx;
  const float lim_x_neg = cx / fx + 0.3f * tan_fovx;
  const float lim_y_pos = ((float)height - cy) / fy + 0.3f * tan_fovy;
  const float lim_y_neg = cy / fy + 0.3f * tan_fovy;
  const float rz = 1.0f / pz;
  const float rz2 = rz * rz;
  const float tx = pz * fminf(lim_x_pos, fmaxf(-lim_x_neg, px * rz));
  const float ty = pz * fminf(lim_y_pos, fmaxf(-lim_y_neg, py * rz));
  const float J00 = fx * rz, J02 = -((fx * tx) * rz2);
  const float J11 = fy * rz, J12 = -((fy * ty) * rz2);
  /* cov2d = J Sigma_c J^T with J = [[J00,0,J02],[0,J11,J12]] */
  const float A00 = J00 * c00 + J02 * c02; /* (J Sigma_c) row 0 */
  const float A01 = J00 * c01 + J02 * c12;
  const float A02 = J00 * c02 + J02 * c22;
  const float A11 = J11 * c11 + J12 * c12; /* row 1 */
  const float A12 = J11 * c12 + J12 * c22;
  float v00 = A00 * J00 + A02 * J02;
  const float v01 = A01 * J11 + A02 * J12;
  float v11 = A11 * J11 + A12 * J12;
  const float m2x = (fx * px) * rz + cx;
  const float m2y = (fy * py) * rz + cy;

  /* blur + conic (add_blur, inverse) */
  v00 += eps2d;
  v11 += eps2d;
  const float det = v00 * v11 - v01 * v01;
  if (!(det > 0.0f)) return 0;
  const float inv_det = 1.0f / det;

  /* radius */
  const float b = 0.5f * (v00 + v11);
  const float v1 = b + sqrtf(fmaxf(0.01f, b * b - det));
  const float radius = ceilf(3.0f * sqrtf(v1));
  if (radius <= radius_clip) return 0;
  if (m2x + radius <= 0.0f || m2x - radius >= (float)width ||
      m2y + radius <= 0.0f || m2y - radius >= (float)height)
    return 0;

  mean2d[0] = m2x; mean2d[1] = m2y;
  *depth = pz;
  conic[0] = v11 * inv_det;
  conic[1] = -v01 * inv_det;
  conic[2] = v00 * inv_det;
  if (radius > 2147483520.0f) return 2147483520;
  return (int)radius;
}

/* tile AABB of a projected Gaussian (isect_tiles pass 1) */
static inline int clampi_from_float(float v, int hi) {
  if (!(v > 0.0f)) return 0;
  if (v >= (float)hi) return hi;
  return (int)v;
}
static inline void tile_aabb(const float *mean2d, int radius, int tile_size, int tw, int th,
                             int *x0, int *x1, int *y0, int *y1) {
  const float ts = (float)tile_size;
  const float tr = (float)radius / ts;
  const float txc = mean2d[0] / ts, tyc = mean2d[1] / ts;
  *x0 = clampi_from_float(floorf(txc - tr), tw);
  *x1 = clampi_from_float(ceilf(txc + tr), tw);
  *y0 = clampi_from_float(floorf(tyc - tr), th);
  *y1 = clampi_from_float(ceilf(tyc + tr), th);
}

ORC_API void orc_project(const float *means, const float *quats, const float *scales,
                         const float *viewmats, const float *Ks, int N, int C,
                         int width, int height, int tile_size, float eps2d,
                         float near_plane, float far_plane, float radius_clip,
                         int32_t *radii, float *means2d, float *depths, float *conics,
                         int32_t *tiles_per_gauss) {
  const int tw = (width + tile_size - 1) / tile_size;
  const int th = (height + tile_size - 1) / tile_size;
  for (int c = 0; c < C; ++c) {
    orc_cam cam;
    load_cam(viewmats + c * 16, Ks + c * 9, &cam);
#pragma omp parallel for schedule(static)
    for (long g = 0; g < N; ++g) {
      const long f = (long)c * N + g;
      float m2[2] = {0, 0}, d = 0, con[3] = {0, 0, 0};
      const int r = project_one(means + g * 3, quats + g * 4, scales + g * 3, &cam, width, height,
                                eps2d, near_plane, far_plane, radius_clip, m2, &d, con);
      radii[f] = r;
      means2d[f * 2] = m2[0]; means2d[f * 2 + 1] = m2[1];
      depths[f] = d;
      conics[f * 3] = con[0]; conics[f * 3 + 1] = con[1]; conics[f * 3 + 2] = con[2];
      if (tiles_per_gauss) {
        int n = 0;
        if (r > 0) {
          int x0, x1, y0, y1;
          tile_aabb(m2, r, tile_size, tw, th, &x0, &x1, &y0, &y1);
          n = (x1 - x0) * (y1 - y0);
        }
        tiles_per_gauss[f] = n;
      }
    }
  }
}

/* ------------------------------------------------------------------------- *
 * Stage 3: keys, stable sort, offsets.
 * ------------------------------------------------------------------------- */
static int n_bits_for(int n) { /* floor(log2(n)) + 1 */
  int b = 0;
  while (n > 0) { ++b; n >>= 1; }
  return b;
}

ORC_API int orc_tile_n_bits(int n_tiles) { return n_bits_for(n_tiles); }
ORC_API int orc_cam_n_bits(int C) { return n_bits_for(C); }

/* returns n_isects; if isect_ids==NULL only counts */
ORC_API int64_t orc_bin_emit(const float *means2d, const int32_t *radii, const float *depths,
                             int N, int C, int tile_size, int tile_width, int tile_height,
                             int64_t *isect_ids, int32_t *flatten_ids) {
  const int tile_n_bits = n_bits_for(tile_width * tile_height);
  int64_t cur = 0;
  for (long f = 0; f < (long)C * N; ++f) {
    const int r = radii[f];
    if (r <= 0) continue;
    int x0, x1, y0, y1;
    tile_aabb(means2d + f * 2, r, tile_size, tile_width, tile_height, &x0, &x1, &y0, &y1);
    if (isect_ids) {
      const int64_t cid = f / N;
      uint32_t dbits;
      memcpy(&dbits, depths + f, 4);
      const int64_t cam_enc = cid << (32 + tile_n_bits);
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
          const int64_t tile_id = (int64_t)y * tile_width + x;
          isect_ids[cur] = cam_enc | (tile_id << 32) | (int64_t)dbits;
          flatten_ids[cur] = (int32_t)f;
          ++cur;
        }
    } else {
      cur += (int64_t)(x1 - x0) * (y1 - y0);
    }
  }
  return cur;
}

/* stable LSD radix sort on key bits [0, end_bit); result in (keys, vals) */
ORC_API void orc_sort_pairs(int64_t *keys, int32_t *vals, int64_t n, int end_bit) {
  if (n <= 1) return;
  int64_t *k2 = (int64_t *)malloc(sizeof(int64_t) * n);
  int32_t *v2 = (int32_t *)malloc(sizeof(int32_t) * n);
  int64_t *ka = keys, *kb = k2;
  int32_t *va = vals, *vb = v2;
  for (int shift = 0; shift < end_bit; shift += 8) {
    int64_t count[257];
    memset(count, 0, sizeof(count));
    for (int64_t i = 0; i < n; ++i) count[(((uint64_t)ka[i]) >> shift & 0xff) + 1]++;
    for (int d = 0; d < 256; ++d) count[d + 1] += count[d];
    for (int64_t i = 0; i < n; ++i) {
      const int d = (int)(((uint64_t)ka[i]) >> shift & 0xff);
      const int64_t dst = count[d]++;
      kb[dst] = ka[i];
      vb[dst] = va[i];
    }
    int64_t *tk = ka; ka = kb; kb = tk;
    int32_t *tv = va; va = vb; vb = tv;
  }
  if (ka != keys) {
    memcpy(keys, ka, sizeof(int64_t) * n);
    memcpy(vals, va, sizeof(int32_t) * n);
  }
  free(k2);
  free(v2);
}

/* isect_offset_encode: offsets[c,ty,tx] = first sorted index with (cam,tile) >= it */
ORC_API void orc_tile_offsets(const int64_t *sorted_ids, int64_t n, int C, int tile_width,
                              int tile_height, int32_t *offsets) {
  const int n_tiles = tile_width * tile_height;
  const int tile_n_bits = n_bits_for(n_tiles);
  const long total = (long)C * n_tiles;
  long next = 0; /* next (cam,tile) slot to fill */
  for (int64_t i = 0; i < n; ++i) {
    const int64_t hi = sorted_ids[i] >> 32;
    const long cid = (long)(hi >> tile_n_bits);
    const long tid = (long)(hi & ((1L << tile_n_bits) - 1));
    const long id = cid * n_tiles + tid;
    while (next <= id && next < total) offsets[next++] = (int32_t)i;
  }
  while (next < total) offsets[next++] = (int32_t)n;
}

/* ------------------------------------------------------------------------- *
 * Stage 4: front-to-back compositing.
 * colors is [N,D] (colors_per_camera=0) or [C,N,D]; depths [C,N] or NULL adds a
 * channel; backgrounds [C,D] or NULL. CD = D + (depths != NULL) <= 4.
 * fragile [C,H,W] uint8 (may be NULL): 1 where some decision of the pixel
 * (alpha >= 1/255, T' <= 1e-4) lies within a relative margin of its threshold,
 * i.e. where a 1-ulp difference in exp() may legitimately flip the outcome.
 * ------------------------------------------------------------------------- */
#define ALPHA_MIN (1.0f / 255.0f)
#define FRAG_ALPHA_REL 2e-5f
#define FRAG_T_REL 2e-4f

ORC_API void orc_composite_fwd(const float *means2d, const float *conics, const float *opacities,
                               const float *colors, int colors_per_camera, int D,
                               const float *depths, const float *backgrounds,
                               const int32_t *offsets, const int32_t *flatten_ids, int64_t n_isects,
                               int N, int C, int width, int height, int tile_size,
                               float *render, float *alphas, int32_t *last_ids, uint8_t *fragile,
                               float *transmittance) {
  const int tw = (width + tile_size - 1) / tile_size;
  const int th = (height + tile_size - 1) / tile_size;
  const int CD = D + (depths ? 1 : 0);
  const long n_tiles = (long)tw * th;
#pragma omp parallel for schedule(dynamic, 4)
  for (long ct = 0; ct < (long)C * n_tiles; ++ct) {
    const int c = (int)(ct / n_tiles);
    const int tile = (int)(ct % n_tiles);
    const int ty = tile / tw, txi = tile % tw;
    const int32_t start = offsets[ct];
    const int32_t end = (ct == (long)C * n_tiles - 1) ? (int32_t)n_isects : offsets[ct + 1];
    for (int iy = ty * tile_size; iy < (ty + 1) * tile_size && iy < height; ++iy)
      for (int ix = txi * tile_size; ix < (txi + 1) * tile_size && ix < width; ++ix) {
        const float px = (float)ix + 0.5f, py = (float)iy + 0.5f;
        float T = 1.0f;
        float out[4] = {0, 0, 0, 0};
        int32_t last = 0;
        uint8_t frag = 0;
        for (int32_t i = start; i < end; ++i) {
          const int32_t f = flatten_ids[i];
          const int32_t g = f % N;
          const float dx = means2d[2 * (long)f] - px;
          const float dy = means2d[2 * (long)f + 1] - py;
          const float *con = conics + 3 * (long)f;
          const float sigma = 0.5f * (con[0] * dx * dx + con[2] * dy * dy) + con[1] * dx * dy;
          const float raw = opacities[g] * expf(-sigma);
          const float alpha = fminf(0.999f, raw);
          if (fabsf(raw - ALPHA_MIN) <= FRAG_ALPHA_REL * ALPHA_MIN && sigma >= 0.0f) frag = 1;
          if (fabsf(sigma) < 1e-6f && raw >= ALPHA_MIN) frag = 1; /* sign of a rounded ~0 */
          if (sigma < 0.0f || alpha < ALPHA_MIN) continue;
          const float next_T = T * (1.0f - alpha);
          if (fabsf(next_T - 1e-4f) <= FRAG_T_REL * 1e-4f) frag = 1;
          if (next_T <= 1e-4f) break;
          const float vis = alpha * T;
          const float *col = colors + (colors_per_camera ? (long)f : (long)g) * D;
          for (int k = 0; k < D; ++k) out[k] += col[k] * vis;
          if (depths) out[D] += depths[f] * vis;
          last = i;
          T = next_T;
        }
        const long pix = ((long)c * height + iy) * width + ix;
        alphas[pix] = 1.0f - T;
        for (int k = 0; k < CD; ++k) {
          const float bg = (backgrounds && k < D) ? backgrounds[c * D + k] : 0.0f;
          render[pix * CD + k] = backgrounds ? out[k] + T * bg : out[k];
        }
        last_ids[pix] = last;
        if (fragile) fragile[pix] = frag;
        if (transmittance) transmittance[pix] = T; /* exact final T, see orc_composite_bwd */
      }
  }
}

/* Backward of the compositing, restating rasterize_to_pixels_bwd (SURVEY §8c
 * step 12). Accumulates in double so that the oracle's own summation order adds
 * no noise (tiles run in parallel, the accumulations are atomic doubles: the order
 * changes the result by ~1e-16 relative); outputs must be zero-filled by the caller.
 * transmittance == NULL: T_final = 1 - alphas[pix], the form upstream uses. That
 * subtraction rounds away up to ulp(alpha) / T_final (~3e-4 relative on an opaque
 * pixel) of the transmittance the forward actually ended with, and every term of
 * the pixel's gradient scales with it. transmittance != NULL (what the forward
 * wrote): the exact value — the form the CUDA path uses and the tests compare. */
ORC_API void orc_composite_bwd(const float *means2d, const float *conics, const float *opacities,
                               const float *colors, int colors_per_camera, int D,
                               const float *depths, const float *backgrounds,
                               const int32_t *offsets, const int32_t *flatten_ids, int64_t n_isects,
                               int N, int C, int width, int height, int tile_size,
                               const float *alphas, const int32_t *last_ids,
                               const float *v_render, const float *v_alphas,
                               double *v_means2d, double *v_conics, double *v_colors,
                               double *v_opacities, double *v_depths, const float *transmittance) {
  const int tw = (width + tile_size - 1) / tile_size;
  const int th = (height + tile_size - 1) / tile_size;
  const int CD = D + (depths ? 1 : 0);
  const long n_tiles = (long)tw * th;
  (void)n_isects;
#pragma omp parallel for schedule(dynamic, 4)
  for (long ct = 0; ct < (long)C * n_tiles; ++ct) {
    const int c = (int)(ct / n_tiles);
    const int tile = (int)(ct % n_tiles);
    const int ty = tile / tw, txi = tile % tw;
    const int32_t start = offsets[ct];
    for (int iy = ty * tile_size; iy < (ty + 1) * tile_size && iy < height; ++iy)
      for (int ix = txi * tile_size; ix < (txi + 1) * tile_size && ix < width; ++ix) {
        const long pix = ((long)c * height + iy) * width + ix;
        const float px = (float)ix + 0.5f, py = (float)iy + 0.5f;
        const float T_final = transmittance ? transmittance[pix] : 1.0f - alphas[pix];
        float T = T_final;
        float buffer[4] = {0, 0, 0, 0};
        float vr[4] = {0, 0, 0, 0};
        if (v_render)
          for (int k = 0; k < CD; ++k) vr[k] = v_render[pix * CD + k];
        const float va = v_alphas ? v_alphas[pix] : 0.0f;
        const int32_t last = last_ids[pix];
        for (int32_t i = last; i >= start; --i) {
          const int32_t f = flatten_ids[i];
          const int32_t g = f % N;
          const float dx = means2d[2 * (long)f] - px;
          const float dy = means2d[2 * (long)f + 1] - py;
          const float *con = conics + 3 * (long)f;
          const float sigma = 0.5f * (con[0] * dx * dx + con[2] * dy * dy) + con[1] * dx * dy;
          const float vis = expf(-sigma);
          const float opac = opacities[g];
          const float alpha = fminf(0.999f, opac * vis);
          if (sigma < 0.0f || alpha < ALPHA_MIN) continue;
          const float ra = 1.0f / (1.0f - alpha);
          T *= ra;
          const float fac = alpha * T;
          const long cidx = (colors_per_camera ? (long)f : (long)g) * D;
          float v_alpha = 0.0f;
          float chan[4] = {0, 0, 0, 0};
          for (int k = 0; k < D; ++k) chan[k] = colors[cidx + k];
          if (depths) chan[D] = depths[f];
          for (int k = 0; k < D; ++k) {
#pragma omp atomic
            v_colors[cidx + k] += (double)(fac * vr[k]);
          }
          if (depths) {
#pragma omp atomic
            v_depths[f] += (double)(fac * vr[D]);
          }
          for (int k = 0; k < CD; ++k) v_alpha += (chan[k] * T - buffer[k] * ra) * vr[k];
          v_alpha += T_final * ra * va;
          if (backgrounds) {
            float accum = 0.0f;
            for (int k = 0; k < D; ++k) accum += backgrounds[c * D + k] * vr[k];
            v_alpha += -T_final * ra * accum;
          }
          if (opac * vis <= 0.999f) {
            const float v_sigma = -opac * vis * v_alpha;
#pragma omp atomic
            v_conics[3 * (long)f + 0] += (double)(0.5f * v_sigma * dx * dx);
#pragma omp atomic
            v_conics[3 * (long)f + 1] += (double)(v_sigma * dx * dy);
#pragma omp atomic
            v_conics[3 * (long)f + 2] += (double)(0.5f * v_sigma * dy * dy);
#pragma omp atomic
            v_means2d[2 * (long)f + 0] += (double)(v_sigma * (con[0] * dx + con[1] * dy));
#pragma omp atomic
            v_means2d[2 * (long)f + 1] += (double)(v_sigma * (con[1] * dx + con[2] * dy));
#pragma omp atomic
            v_opacities[g] += (double)(vis * v_alpha);
          }
          for (int k = 0; k < CD; ++k) buffer[k] += chan[k] * fac;
        }
      }
  }
}

/* ------------------------------------------------------------------------- *
 * Backward of the projection (fully_fused_projection_bwd restated): VJP of
 * project_one w.r.t. mean, quat, scale. Accumulates over cameras in double.
 * ------------------------------------------------------------------------- */
static void project_bwd_one(const float *mean, const float *quat, const float *scale,
                            const orc_cam *c, int width, int height, float eps2d,
                            const float *conic, const float *v_mean2d, float v_depth,
                            const float *v_conic, double *v_mean, double *v_quat, double *v_scale) {
  (void)eps2d;
  const float *R = c->R;
  const float px = sum3(R[0] * mean[0], R[1] * mean[1], R[2] * mean[2]) + c->t[0];
  const float py = sum3(R[3] * mean[0], R[4] * mean[1], R[5] * mean[2]) + c->t[1];
  const float pz = sum3(R[6] * mean[0], R[7] * mean[1], R[8] * mean[2]) + c->t[2];
  float Rq[9], M[9];
  quat_to_rotmat(quat, Rq);
  for (int r = 0; r < 3; ++r)
    for (int k = 0; k < 3; ++k) M[r * 3 + k] = Rq[r * 3 + k] * scale[k];
  double S[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += (double)M[i * 3 + k] * M[j * 3 + k];
      S[i * 3 + j] = s;
    }
  double Sc[9], Tm[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += (double)R[i * 3 + k] * S[k * 3 + j];
      Tm[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += Tm[i * 3 + k] * (double)R[j * 3 + k];
      Sc[i * 3 + j] = s;
    }
  /* d loss / d cov2d from d loss / d conic: conic = cov2d^-1 (symmetric),
   * v_cov = -Q * Vc * Q with Vc = [[vx, vy/2],[vy/2, vz]] and Q the conic matrix */
  const double q00 = conic[0], q01 = conic[1], q11 = conic[2];
  const double a00 = v_conic[0], a01 = 0.5 * v_conic[1], a11 = v_conic[2];
  const double t00 = q00 * a00 + q01 * a01, t01 = q00 * a01 + q01 * a11;
  const double t10 = q01 * a00 + q11 * a01, t11 = q01 * a01 + q11 * a11;
  const double g00 = -(t00 * q00 + t01 * q01);
  const double g01 = -(t00 * q01 + t01 * q11);
  const double g10 = -(t10 * q00 + t11 * q01);
  const double g11 = -(t10 * q01 + t11 * q11);
  /* perspective projection VJP */
  const double fx = c->fx, fy = c->fy, cx = c->cx, cy = c->cy;
  const double tan_fovx = 0.5 * width / fx, tan_fovy = 0.5 * height / fy;
  const double lim_x_pos = (width - cx) / fx + 0.3 * tan_fovx, lim_x_neg = cx / fx + 0.3 * tan_fovx;
  const double lim_y_pos = (height - cy) / fy + 0.3 * tan_fovy, lim_y_neg = cy / fy + 0.3 * tan_fovy;
  const double x = px, y = py, z = pz;
  const double rz = 1.0 / z, rz2 = rz * rz, rz3 = rz2 * rz;
  const double xr = x * rz, yr = y * rz;
  const int x_in = (xr <= lim_x_pos && xr >= -lim_x_neg);
  const int y_in = (yr <= lim_y_pos && yr >= -lim_y_neg);
  const double tx = z * fmin(lim_x_pos, fmax(-lim_x_neg, xr));
  const double ty = z * fmin(lim_y_pos, fmax(-lim_y_neg, yr));
  const double J[6] = {fx * rz, 0, -fx * tx * rz2, 0, fy * rz, -fy * ty * rz2}; /* 2x3 */
  const double G[4] = {g00, g01, g10, g11};
  /* v_Sc = J^T G J (3x3) */
  double vSc[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) s += J[a * 3 + i] * G[a * 2 + b] * J[b * 3 + j];
      vSc[i * 3 + j] = s;
    }
  /* v_J = G J Sc^T + G^T J Sc (2x3) */
  double vJ[6];
  for (int a = 0; a < 2; ++a)
    for (int i = 0; i < 3; ++i) {
      double s = 0;
      for (int b = 0; b < 2; ++b)
        for (int k = 0; k < 3; ++k)
          s += G[a * 2 + b] * J[b * 3 + k] * Sc[i * 3 + k] + G[b * 2 + a] * J[b * 3 + k] * Sc[k * 3 + i];
      vJ[a * 3 + i] = s;
    }
  double vp[3] = {0, 0, 0};
  vp[0] += fx * rz * v_mean2d[0];
  vp[1] += fy * rz * v_mean2d[1];
  vp[2] += -(fx * x * v_mean2d[0] + fy * y * v_mean2d[1]) * rz2;
  if (x_in) vp[0] += -fx * rz2 * vJ[2];
  else vp[2] += -fx * rz3 * vJ[2] * tx;
  if (y_in) vp[1] += -fy * rz2 * vJ[5];
  else vp[2] += -fy * rz3 * vJ[5] * ty;
  vp[2] += -fx * rz2 * vJ[0] - fy * rz2 * vJ[4] + 2.0 * fx * tx * rz3 * vJ[2] + 2.0 * fy * ty * rz3 * vJ[5];
  vp[2] += v_depth;
  /* world: v_mean = R^T vp ; v_S = R^T vSc R */
  for (int i = 0; i < 3; ++i) v_mean[i] += R[0 * 3 + i] * vp[0] + R[1 * 3 + i] * vp[1] + R[2 * 3 + i] * vp[2];
  double vS[9], tmp[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += (double)R[k * 3 + i] * vSc[k * 3 + j];
      tmp[i * 3 + j] = s;
    }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += tmp[i * 3 + k] * (double)R[k * 3 + j];
      vS[i * 3 + j] = s;
    }
  /* Sigma = M M^T: v_M = (vS + vS^T) M ; M = Rq diag(s) */
  double vM[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += (vS[i * 3 + k] + vS[k * 3 + i]) * (double)M[k * 3 + j];
      vM[i * 3 + j] = s;
    }
  double vR[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) vR[i * 3 + j] = vM[i * 3 + j] * scale[j];
  for (int j = 0; j < 3; ++j)
    v_scale[j] += Rq[0 * 3 + j] * vM[0 * 3 + j] + Rq[1 * 3 + j] * vM[1 * 3 + j] + Rq[2 * 3 + j] * vM[2 * 3 + j];
  /* rotation matrix of the normalised quaternion -> un-normalised quaternion */
  const double qw0 = quat[0], qx0 = quat[1], qy0 = quat[2], qz0 = quat[3];
  const double inv = 1.0 / sqrt(qw0 * qw0 + qx0 * qx0 + qy0 * qy0 + qz0 * qz0);
  const double w = qw0 * inv, xq = qx0 * inv, yq = qy0 * inv, zq = qz0 * inv;
  /* R entries as functions of (w,x,y,z): differentiate the table in quat_to_rotmat */
  double vq[4];
  vq[0] = 2.0 * (-zq * vR[1] + yq * vR[2] + zq * vR[3] - xq * vR[5] - yq * vR[6] + xq * vR[7]);
  vq[1] = 2.0 * (yq * vR[1] + zq * vR[2] + yq * vR[3] - 2.0 * xq * vR[4] - w * vR[5] + zq * vR[6] + w * vR[7] - 2.0 * xq * vR[8]);
  vq[2] = 2.0 * (-2.0 * yq * vR[0] + xq * vR[1] + w * vR[2] + xq * vR[3] + zq * vR[5] - w * vR[6] + zq * vR[7] - 2.0 * yq * vR[8]);
  vq[3] = 2.0 * (-2.0 * zq * vR[0] - w * vR[1] + xq * vR[2] + w * vR[3] - 2.0 * zq * vR[4] + yq * vR[5] + xq * vR[6] + yq * vR[7]);
  const double qn[4] = {w, xq, yq, zq};
  const double dotp = vq[0] * qn[0] + vq[1] * qn[1] + vq[2] * qn[2] + vq[3] * qn[3];
  for (int k = 0; k < 4; ++k) v_quat[k] += (vq[k] - dotp * qn[k]) * inv;
}

ORC_API void orc_project_bwd(const float *means, const float *quats, const float *scales,
                             const float *viewmats, const float *Ks, int N, int C,
                             int width, int height, float eps2d,
                             const int32_t *radii, const float *conics,
                             const float *v_means2d, const float *v_depths, const float *v_conics,
                             double *v_means, double *v_quats, double *v_scales) {
  for (int c = 0; c < C; ++c) {
    orc_cam cam;
    load_cam(viewmats + c * 16, Ks + c * 9, &cam);
#pragma omp parallel for schedule(static)
    for (long g = 0; g < N; ++g) {
      const long f = (long)c * N + g;
      if (radii[f] <= 0) continue;
      const float zero3[3] = {0, 0, 0};
      const float *vm = v_means2d ? v_means2d + f * 2 : zero3;
      const float vd = v_depths ? v_depths[f] : 0.0f;
      const float *vc = v_conics ? v_conics + f * 3 : zero3;
      project_bwd_one(means + g * 3, quats + g * 4, scales + g * 3, &cam, width, height, eps2d,
                      conics + f * 3, vm, vd, vc, v_means + g * 3, v_quats + g * 4, v_scales + g * 3);
    }
  }
}
