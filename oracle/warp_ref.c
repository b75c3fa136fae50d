/*
 * oracle/warp_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/msda_ref.c header).
 *
 * PARITY UNPINNED: the reference's warp is kornia.warp_perspective, a third-party PyPI dependency that is not
 * vendored, not version-pinned (reference README.md:42 lists bare "kornia"; call site
 * multiview_detector/models/mvdetr.py:194-195, spelling implies kornia ~0.5.x) and not installed here. No
 * reference test or golden vector covers it. This file restates kornia's published algorithm:
 *   normal_transform_pixel(h, w)   = [[2/(w-1), 0, -1], [0, 2/(h-1), -1], [0, 0, 1]]   (eps=1e-14 if size 1)
 *   normalize_homography(M)        = Ndst @ M @ inv(Nsrc)
 *   src_norm_trans_dst_norm        = inverse(...)
 *   grid = create_meshgrid(Ho, Wo, normalized_coordinates=True)   -> linspace(-1, 1, n) per axis
 *   transform_points + convert_points_from_homogeneous:  scale = |z| > 1e-8 ? 1/(z + 1e-8) : 1
 *   F.grid_sample(src, grid, mode='bilinear', padding_mode='zeros', align_corners=False)
 * The 3x3 chain is done in double here (kornia does it in fp32 with torch.inverse; oracle/torch_port.py keeps
 * that variant, and tests bound the difference between the two).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static void normalized_inverse(const float* Mat, int Hi, int Wi, int Ho, int Wo, float* T) {
  const double eps = 1e-14;
  const double sw = Wi == 1 ? eps : Wi - 1.0, sh = Hi == 1 ? eps : Hi - 1.0;
  const double dw = Wo == 1 ? eps : Wo - 1.0, dh = Ho == 1 ? eps : Ho - 1.0;
  const double Ns_inv[9] = {sw / 2, 0, sw / 2, 0, sh / 2, sh / 2, 0, 0, 1};
  const double Nd[9] = {2 / dw, 0, -1, 0, 2 / dh, -1, 0, 0, 1};
  double A[9], N[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += (double)Mat[3 * r + k] * Ns_inv[3 * k + c];
      A[3 * r + c] = s;
    }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += Nd[3 * r + k] * A[3 * k + c];
      N[3 * r + c] = s;
    }
  /* Gauss-Jordan with partial pivoting on [N | I] */
  double aug[3][6];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 6; ++c) aug[r][c] = c < 3 ? N[3 * r + c] : (c - 3 == r ? 1.0 : 0.0);
  for (int col = 0; col < 3; ++col) {
    int piv = col;
    for (int r = col + 1; r < 3; ++r)
      if (fabs(aug[r][col]) > fabs(aug[piv][col])) piv = r;
    if (piv != col)
      for (int c = 0; c < 6; ++c) {
        double t = aug[col][c];
        aug[col][c] = aug[piv][c];
        aug[piv][c] = t;
      }
    const double d = aug[col][col];
    for (int c = 0; c < 6; ++c) aug[col][c] /= d;
    for (int r = 0; r < 3; ++r)
      if (r != col) {
        const double f = aug[r][col];
        for (int c = 0; c < 6; ++c) aug[r][c] -= f * aug[col][c];
      }
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) T[3 * r + c] = (float)aug[r][c + 3];
}

static float linspace_pm1(int i, int n) {
  if (n == 1) return -1.f;
  const float step = 2.f / (float)(n - 1);
  volatile float prod; /* keep the product rounded (no FMA contraction) */
  if (i < n / 2) {
    prod = step * (float)i;
    return -1.f + prod;
  }
  prod = step * (float)(n - 1 - i);
  return 1.f - prod;
}

/* If T_in != NULL it is used as the normalised inverse homography [BN,3,3] (lets the torch port supply kornia's
 * fp32 torch.inverse result); otherwise it is derived from Mat in double. */
void oracle_warp_fwd_f32(const float* src, const float* Mat, const float* T_in, int BN, int C, int Hi, int Wi, int Ho,
                         int Wo, float* dst) {
  const int64_t plane = (int64_t)Hi * Wi, oplane = (int64_t)Ho * Wo;
  _Pragma("omp parallel for schedule(static)") for (int n = 0; n < BN; ++n) {
    float T[9];
    if (T_in) memcpy(T, T_in + 9 * n, sizeof(T));
    else normalized_inverse(Mat + 9 * n, Hi, Wi, Ho, Wo, T);
    for (int v = 0; v < Ho; ++v)
      for (int u = 0; u < Wo; ++u) {
        const float gx = linspace_pm1(u, Wo), gy = linspace_pm1(v, Ho);
        const float X = T[0] * gx + T[1] * gy + T[2];
        const float Y = T[3] * gx + T[4] * gy + T[5];
        const float Z = T[6] * gx + T[7] * gy + T[8];
        const float sc = fabsf(Z) > 1e-8f ? 1.f / (Z + 1e-8f) : 1.f;
        const float x = X * sc, y = Y * sc;
        const float ix = ((x + 1.f) * Wi - 1.f) / 2.f, iy = ((y + 1.f) * Hi - 1.f) / 2.f;
        const float fx = floorf(ix), fy = floorf(iy);
        const float nw = (fx + 1 - ix) * (fy + 1 - iy), ne = (ix - fx) * (fy + 1 - iy);
        const float sw = (fx + 1 - ix) * (iy - fy), se = (ix - fx) * (iy - fy);
        const int ok = ix > -1.f && iy > -1.f && ix < (float)Wi && iy < (float)Hi; /* else every tap is outside */
        const int x0 = ok ? (int)fx : 0, y0 = ok ? (int)fy : 0;
        for (int c = 0; c < C; ++c) {
          const float* s = src + ((int64_t)n * C + c) * plane;
          float acc = 0.f;
          if (ok) {
            if (y0 >= 0 && x0 >= 0) acc += s[y0 * Wi + x0] * nw;
            if (y0 >= 0 && x0 + 1 <= Wi - 1) acc += s[y0 * Wi + x0 + 1] * ne;
            if (y0 + 1 <= Hi - 1 && x0 >= 0) acc += s[(y0 + 1) * Wi + x0] * sw;
            if (y0 + 1 <= Hi - 1 && x0 + 1 <= Wi - 1) acc += s[(y0 + 1) * Wi + x0 + 1] * se;
          }
          dst[((int64_t)n * C + c) * oplane + (int64_t)v * Wo + u] = acc;
        }
      }
  }
}

/* Gradient w.r.t. src (ATen grid_sampler_2d_backward, input gradient only). Sequential, deterministic. */
void oracle_warp_bwd_f32(const float* grad_dst, const float* Mat, const float* T_in, int BN, int C, int Hi, int Wi,
                         int Ho, int Wo, float* grad_src) {
  const int64_t plane = (int64_t)Hi * Wi, oplane = (int64_t)Ho * Wo;
  memset(grad_src, 0, sizeof(float) * (size_t)BN * C * plane);
  for (int n = 0; n < BN; ++n) {
    float T[9];
    if (T_in) memcpy(T, T_in + 9 * n, sizeof(T));
    else normalized_inverse(Mat + 9 * n, Hi, Wi, Ho, Wo, T);
    for (int v = 0; v < Ho; ++v)
      for (int u = 0; u < Wo; ++u) {
        const float gx = linspace_pm1(u, Wo), gy = linspace_pm1(v, Ho);
        const float X = T[0] * gx + T[1] * gy + T[2];
        const float Y = T[3] * gx + T[4] * gy + T[5];
        const float Z = T[6] * gx + T[7] * gy + T[8];
        const float sc = fabsf(Z) > 1e-8f ? 1.f / (Z + 1e-8f) : 1.f;
        const float x = X * sc, y = Y * sc;
        const float ix = ((x + 1.f) * Wi - 1.f) / 2.f, iy = ((y + 1.f) * Hi - 1.f) / 2.f;
        if (!(ix > -1.f && iy > -1.f && ix < (float)Wi && iy < (float)Hi)) continue;
        const float fx = floorf(ix), fy = floorf(iy);
        const float nw = (fx + 1 - ix) * (fy + 1 - iy), ne = (ix - fx) * (fy + 1 - iy);
        const float sw = (fx + 1 - ix) * (iy - fy), se = (ix - fx) * (iy - fy);
        const int x0 = (int)fx, y0 = (int)fy;
        for (int c = 0; c < C; ++c) {
          float* s = grad_src + ((int64_t)n * C + c) * plane;
          const float g = grad_dst[((int64_t)n * C + c) * oplane + (int64_t)v * Wo + u];
          if (y0 >= 0 && x0 >= 0) s[y0 * Wi + x0] += g * nw;
          if (y0 >= 0 && x0 + 1 <= Wi - 1) s[y0 * Wi + x0 + 1] += g * ne;
          if (y0 + 1 <= Hi - 1 && x0 >= 0) s[(y0 + 1) * Wi + x0] += g * sw;
          if (y0 + 1 <= Hi - 1 && x0 + 1 <= Wi - 1) s[(y0 + 1) * Wi + x0 + 1] += g * se;
        }
      }
  }
}
