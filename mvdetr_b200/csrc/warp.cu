// Perspective warp of per-view feature maps onto the ground-plane grid: homography + bilinear gather.
//
// Replaces kornia.warp_perspective(src, M, dsize, align_corners=False) at the reference call site
//   ref: multiview_detector/models/mvdetr.py:194-195
// kornia is third-party and absent from the reference tree (ref: README.md:42), so the algorithm is restated
// (DESIGN.md "warp", oracle/warp_ref.c): normalise the pixel homography to [-1,1] coordinates with the
// (size-1) convention, invert, push the normalised destination meshgrid through it with the eps-guarded
// homogeneous divide, then ATen grid_sample semantics (bilinear, zeros padding, align_corners=False).
//
// One launch does all of it: the 3x3 normalise+invert runs in fp64 in thread 0 of every block (about a
// hundred flops), so there are no grid-building kernels and no [BN,Ho,Wo,2] grid tensor in HBM
// (the reference path launches ~10 tiny kernels and writes/reads a 2.4 MB grid).
// Thread mapping: x = destination pixel (consecutive u => coalesced stores and near-coalesced gathers),
// y = chunk of CCH channels (coordinates and weights computed once, reused for CCH planes), z = view.
#include "common.cuh"

namespace mvd {

constexpr int kWarpThreads = 128;

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
  }
  return t;
}

// NCHW destination: x = destination pixel, y = chunk of CCH channels, z = view.
template <int CCH>
__global__ void __launch_bounds__(kWarpThreads) warp_fwd_nchw_kernel(const float* __restrict__ src,
                                                                     const float* __restrict__ Mat, int C, int Hi,
                                                                     int Wi, int Ho, int Wo, float* __restrict__ dst) {
  __shared__ float sT[9];
  const int n = blockIdx.z;
  if (threadIdx.x == 0) normalized_inverse(Mat + 9 * n, Hi, Wi, Ho, Wo, sT);
  __syncthreads();

  const int pix = blockIdx.x * kWarpThreads + threadIdx.x;
  if (pix >= Ho * Wo) return;
  const int v = pix / Wo, u = pix - v * Wo;
  const Taps t = make_taps(sT, u, v, Hi, Wi, Ho, Wo);
  const int c0 = blockIdx.y * CCH;
  const int64_t plane = (int64_t)Hi * Wi, oplane = (int64_t)Ho * Wo;
  const float* sp = src + ((int64_t)n * C + c0) * plane + t.o00;
  float* dp = dst + ((int64_t)n * C + c0) * oplane + pix;
  const bool any = t.m_nw || t.m_ne || t.m_sw || t.m_se;
#pragma unroll 8
  for (int cc = 0; cc < CCH; ++cc) {
    if (c0 + cc >= C) break;
    float acc = 0.f;
    if (any) {
      const float* s = sp + (int64_t)cc * plane;
      if (t.m_nw) acc += __ldg(s) * t.nw;
      if (t.m_ne) acc += __ldg(s + 1) * t.ne;
      if (t.m_sw) acc += __ldg(s + Wi) * t.sw;
      if (t.m_se) acc += __ldg(s + Wi + 1) * t.se;
    }
    __stcs(dp + (int64_t)cc * oplane, acc);  // streaming store: written once, read by the next layer
  }
}

// Channels-last destination [BN, Ho, Wo, C]: a block owns 32 consecutive destination pixels of one view.
//   gather : lane = pixel (neighbouring pixels read neighbouring source texels of the same plane), the 4 warps
//            stride the channels; results go to a [channel][pixel] shared tile (pitch 33: conflict-free);
//   store  : each warp emits whole pixel rows, lane = channel, i.e. 128-byte fully coalesced stores.
constexpr int kTilePx = 32, kTileCh = 128, kTilePitch = 33;

__global__ void __launch_bounds__(kWarpThreads) warp_fwd_nhwc_kernel(const float* __restrict__ src,
                                                                     const float* __restrict__ Mat, int C, int Hi,
                                                                     int Wi, int Ho, int Wo, float* __restrict__ dst) {
  __shared__ float sT[9];
  __shared__ float tile[kTileCh * kTilePitch];
  const int n = blockIdx.y;
  if (threadIdx.x == 0) normalized_inverse(Mat + 9 * n, Hi, Wi, Ho, Wo, sT);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npix = Ho * Wo;
  const int pix0 = blockIdx.x * kTilePx;
  const int pix = pix0 + lane;
  const bool in = pix < npix;
  const int v = in ? pix / Wo : 0, u = in ? pix - v * Wo : 0;
  Taps t = make_taps(sT, u, v, Hi, Wi, Ho, Wo);
  const bool any = in && (t.m_nw || t.m_ne || t.m_sw || t.m_se);
  const int64_t plane = (int64_t)Hi * Wi;
  const float* sp = src + (int64_t)n * C * plane + t.o00;
  const int rows = min(kTilePx, npix - pix0);

  for (int c0 = 0; c0 < C; c0 += kTileCh) {
    const int nch = min(kTileCh, C - c0);
#pragma unroll 8
    for (int i = 0; i < kTileCh / 4; ++i) {
      const int c = warp + 4 * i;
      if (c >= nch) break;
      float acc = 0.f;
      if (any) {
        const float* s = sp + (int64_t)(c0 + c) * plane;
        if (t.m_nw) acc += __ldg(s) * t.nw;
        if (t.m_ne) acc += __ldg(s + 1) * t.ne;
        if (t.m_sw) acc += __ldg(s + Wi) * t.sw;
        if (t.m_se) acc += __ldg(s + Wi + 1) * t.se;
      }
      tile[c * kTilePitch + lane] = acc;
    }
    __syncthreads();
    for (int r = warp; r < rows; r += 4) {
      float* dp = dst + ((int64_t)n * npix + pix0 + r) * C + c0;
#pragma unroll
      for (int k = 0; k < kTileCh / 32; ++k) {
        const int c = lane + 32 * k;
        if (c < nch) __stcs(dp + c, tile[c * kTilePitch + r]);
      }
    }
    __syncthreads();
  }
}

template <int CCH>
__global__ void __launch_bounds__(kWarpThreads) warp_bwd_kernel(const float* __restrict__ grad_dst,
                                                                const float* __restrict__ Mat, int C, int Hi, int Wi,
                                                                int Ho, int Wo, float* __restrict__ grad_src) {
  __shared__ float sT[9];
  const int n = blockIdx.z;
  if (threadIdx.x == 0) normalized_inverse(Mat + 9 * n, Hi, Wi, Ho, Wo, sT);
  __syncthreads();

  const int pix = blockIdx.x * kWarpThreads + threadIdx.x;
  if (pix >= Ho * Wo) return;
  const int v = pix / Wo, u = pix - v * Wo;
  const Taps t = make_taps(sT, u, v, Hi, Wi, Ho, Wo);
  if (!(t.m_nw || t.m_ne || t.m_sw || t.m_se)) return;
  const int c0 = blockIdx.y * CCH;
  const int64_t plane = (int64_t)Hi * Wi, oplane = (int64_t)Ho * Wo;
  float* gp = grad_src + ((int64_t)n * C + c0) * plane + t.o00;
  const float* dp = grad_dst + ((int64_t)n * C + c0) * oplane + pix;
#pragma unroll 4
  for (int cc = 0; cc < CCH; ++cc) {
    if (c0 + cc >= C) break;
    const float g = __ldcs(dp + (int64_t)cc * oplane);
    float* s = gp + (int64_t)cc * plane;
    if (t.m_nw) atomicAdd(s, g * t.nw);
    if (t.m_ne) atomicAdd(s + 1, g * t.ne);
    if (t.m_sw) atomicAdd(s + Wi, g * t.sw);
    if (t.m_se) atomicAdd(s + Wi + 1, g * t.se);
  }
}

static int check_warp_dims(int BN, int C, int Hi, int Wi, int Ho, int Wo) {
  if (BN <= 0 || C <= 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0) return MVD_ERR_BAD_SHAPE;
  if (BN > 65535) return MVD_ERR_BAD_SHAPE;
  if ((int64_t)Hi * Wi > 0x7fffffffLL || (int64_t)Ho * Wo > 0x7fffffffLL) return MVD_ERR_BAD_SHAPE;
  return MVD_OK;
}

}  // namespace mvd

using namespace mvd;

extern "C" int mvd_warp_fwd_f32(const float* src, const float* Mat, int BN, int C, int Hi, int Wi, int Ho, int Wo,
                                float* dst, int channels_last, void* stream) {
  if (!src || !Mat || !dst) return MVD_ERR_NULL_POINTER;
  if (int e = check_warp_dims(BN, C, Hi, Wi, Ho, Wo)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (channels_last) {
    dim3 grid((unsigned)ceil_div64((int64_t)Ho * Wo, kTilePx), (unsigned)BN);
    warp_fwd_nhwc_kernel<<<grid, kWarpThreads, 0, st>>>(src, Mat, C, Hi, Wi, Ho, Wo, dst);
  } else {
    constexpr int CCH = 16;
    dim3 grid((unsigned)ceil_div64((int64_t)Ho * Wo, kWarpThreads), (unsigned)ceil_div64(C, CCH), (unsigned)BN);
    if (grid.y > 65535) return MVD_ERR_BAD_SHAPE;
    warp_fwd_nchw_kernel<CCH><<<grid, kWarpThreads, 0, st>>>(src, Mat, C, Hi, Wi, Ho, Wo, dst);
  }
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

extern "C" int mvd_warp_bwd_f32(const float* grad_dst, const float* Mat, int BN, int C, int Hi, int Wi, int Ho,
                                int Wo, float* grad_src, void* stream) {
  if (!grad_dst || !Mat || !grad_src) return MVD_ERR_NULL_POINTER;
  if (int e = check_warp_dims(BN, C, Hi, Wi, Ho, Wo)) return e;
  constexpr int CCH = 16;
  dim3 grid((unsigned)ceil_div64((int64_t)Ho * Wo, kWarpThreads), (unsigned)ceil_div64(C, CCH), (unsigned)BN);
  if (grid.y > 65535) return MVD_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  MVD_CUDA_TRY(cudaMemsetAsync(grad_src, 0, sizeof(float) * (size_t)BN * C * Hi * Wi, st));
  warp_bwd_kernel<CCH><<<grid, kWarpThreads, 0, st>>>(grad_dst, Mat, C, Hi, Wi, Ho, Wo, grad_src);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}
