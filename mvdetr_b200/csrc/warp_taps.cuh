// Homography + tap arithmetic shared by the warp kernels (warp.cu) and the warp->im2col kernel (im2col.cu).
// Restated kornia.warp_perspective / grid_sample semantics: see the header of warp.cu and oracle/warp_ref.c.
#pragma once

#include "common.cuh"

namespace mvd {

struct Homog {
  float t[9];
};

// T = inv(Ndst * Mat * inv(Nsrc)), all in double, rounded once to float.
__device__ inline void normalized_inverse(const float* __restrict__ Mat, int Hi, int Wi, int Ho, int Wo,
                                          float* __restrict__ T) {
  const double eps = 1e-14;  // kornia normal_transform_pixel: denominator eps when a size is 1
  const double sw = (Wi == 1) ? eps : (double)(Wi - 1), sh = (Hi == 1) ? eps : (double)(Hi - 1);
  const double dw = (Wo == 1) ? eps : (double)(Wo - 1), dh = (Ho == 1) ? eps : (double)(Ho - 1);
  double m[9];
  for (int i = 0; i < 9; ++i) m[i] = (double)Mat[i];
  // inv(Nsrc) = [[sw/2, 0, sw/2], [0, sh/2, sh/2], [0, 0, 1]]  (Nsrc = [[2/sw,0,-1],[0,2/sh,-1],[0,0,1]])
  double a[9];
  for (int r = 0; r < 3; ++r) {
    a[3 * r + 0] = m[3 * r + 0] * (sw * 0.5);
    a[3 * r + 1] = m[3 * r + 1] * (sh * 0.5);
    a[3 * r + 2] = m[3 * r + 0] * (sw * 0.5) + m[3 * r + 1] * (sh * 0.5) + m[3 * r + 2];
  }
  // Ndst * a
  double n[9];
  for (int c = 0; c < 3; ++c) {
    n[0 + c] = a[0 + c] * (2.0 / dw) - a[6 + c];
    n[3 + c] = a[3 + c] * (2.0 / dh) - a[6 + c];
    n[6 + c] = a[6 + c];
  }
  // adjugate inverse
  const double c00 = n[4] * n[8] - n[5] * n[7], c01 = n[5] * n[6] - n[3] * n[8], c02 = n[3] * n[7] - n[4] * n[6];
  const double det = n[0] * c00 + n[1] * c01 + n[2] * c02;
  const double id = 1.0 / det;
  T[0] = (float)(c00 * id);
  T[1] = (float)((n[2] * n[7] - n[1] * n[8]) * id);
  T[2] = (float)((n[1] * n[5] - n[2] * n[4]) * id);
  T[3] = (float)(c01 * id);
  T[4] = (float)((n[0] * n[8] - n[2] * n[6]) * id);
  T[5] = (float)((n[2] * n[3] - n[0] * n[5]) * id);
  T[6] = (float)(c02 * id);
  T[7] = (float)((n[1] * n[6] - n[0] * n[7]) * id);
  T[8] = (float)((n[0] * n[4] - n[1] * n[3]) * id);
}

// torch.linspace(-1, 1, n)[i] in fp32 (symmetric two-sided evaluation, as ATen does).
__device__ __forceinline__ float linspace_pm1(int i, int n) {
  if (n == 1) return -1.f;
  const float step = 2.f / (float)(n - 1);
  return (i < n / 2) ? __fadd_rn(-1.f, __fmul_rn(step, (float)i)) : __fsub_rn(1.f, __fmul_rn(step, (float)(n - 1 - i)));
}

struct Taps {
  int o00;            // offset of the north-west tap inside one source plane (may be out of range; see masks)
  int x0, y0;         // the north-west tap in source pixels (x0 in [-1, Wi-1], y0 in [-1, Hi-1] when any mask is set)
  float nw, ne, sw, se;
  bool m_nw, m_ne, m_sw, m_se;
};

__device__ __forceinline__ Taps make_taps(const float* T, int u, int v, int Hi, int Wi, int Ho, int Wo) {
  const float gx = linspace_pm1(u, Wo), gy = linspace_pm1(v, Ho);
  const float X = T[0] * gx + T[1] * gy + T[2];
  const float Y = T[3] * gx + T[4] * gy + T[5];
  const float Z = T[6] * gx + T[7] * gy + T[8];
  const float scale = (fabsf(Z) > 1e-8f) ? 1.f / (Z + 1e-8f) : 1.f;
  const float x = X * scale, y = Y * scale;
  const float ix = ((x + 1.f) * (float)Wi - 1.f) * 0.5f;
  const float iy = ((y + 1.f) * (float)Hi - 1.f) * 0.5f;
  Taps t;
  t.m_nw = t.m_ne = t.m_sw = t.m_se = false;
  t.nw = t.ne = t.sw = t.se = 0.f;
  t.o00 = 0;
  t.x0 = t.y0 = 0;
  // also false for NaN coordinates
  if (ix > -1.f && iy > -1.f && ix < (float)Wi && iy < (float)Hi) {
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float ex = (fx + 1.f) - ix, ey = (fy + 1.f) - iy;  // ix_se - ix, iy_se - iy
    const float dx = ix - fx, dy = iy - fy;
    t.nw = ex * ey;
    t.ne = dx * ey;
    t.sw = ex * dy;
    t.se = dx * dy;
    const bool lef = x0 >= 0, rig = x0 + 1 <= Wi - 1, top = y0 >= 0, bot = y0 + 1 <= Hi - 1;
    t.m_nw = top && lef;
    t.m_ne = top && rig;
    t.m_sw = bot && lef;
    t.m_se = bot && rig;
    t.o00 = y0 * Wi + x0;
    t.x0 = x0;
    t.y0 = y0;
  }
  return t;
}

}  // namespace mvd
