// The two 3x3 convolutions around the encoder as GEMMs (callers either side of the path):
//   * downsample  nn.Conv2d(C, hidden, 3, stride 2, pad 1) + ReLU on the warped world grid
//       ref: multiview_detector/models/trans_world_feat.py:74,89
//   * upsample    nn.Upsample(Rworld_shape, bilinear, align_corners=False) -> nn.Conv2d(hidden, hidden, 3, 1, 1) + ReLU
//       ref: multiview_detector/models/trans_world_feat.py:83-84,109
// cuDNN runs both as fp32 SIMT convolutions (903 us + 471 us of a 4.0 ms frame, profiles/r01h_launches.csv). Here the
// producer of the convolution's input writes it directly in im2col order, A[token][ky][kx][c], and the convolution
// becomes one Linear GEMM (mvd_linear_f32, tensor-core fp32 emulation) whose output rows are already the token-major
// [tokens, C] layout the transformer wants -- the reference's permute-copy (trans_world_feat.py:92) disappears too.
//   warp_im2col_kernel      homography warp (same arithmetic as warp.cu) of a channels-last source, scattered into the
//                           im2col matrix of the stride-s convolution: every warped pixel is computed ONCE and stored
//                           to the <= ceil(3/s)^2 slots that use it (2.25 on average at stride 2);
//   upsample_im2col_kernel  ATen upsample_bilinear2d (align_corners=False) arithmetic on a channels-last map, scattered
//                           into the im2col matrix of the stride-1 convolution (9 slots per pixel).
// Both kernels walk the PADDED pixel domain [-1, H] x [-1, W]: padding pixels carry zeros and write them to the slots
// of border tokens, so the matrix needs no pre-zeroing.
#include "common.cuh"
#include "warp_taps.cuh"

namespace mvd {

constexpr int kImThreads = 128, kImPix = 32;

// Stores the C-vector chunk `acc` (channels c .. c+3 of padded pixel (vp, up), vp/up = pixel + 1) into every im2col
// slot that reads it. Token (oy, ox) tap (ky, kx) reads pixel (oy*s + ky - 1, ox*s + kx - 1).
// Only token rows [row0, row1) are written, at row index oy - row0 (row-sharded callers; the whole grid is row0 = 0, row1 = Ho2).
__device__ __forceinline__ void scatter_slots(float* __restrict__ A, const float4& acc, int vp, int up, int c, int C,
                                              int s, int Ho2, int Wo2, int64_t token0, int row0 = 0,
                                              int row1 = 0x7fffffff) {
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int ty = vp - ky;
    if (ty < 0 || ty % s != 0) continue;
    int oy = ty / s;
    if (oy >= Ho2 || oy < row0 || oy >= row1) continue;
    oy -= row0;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int tx = up - kx;
      if (tx < 0 || tx % s != 0) continue;
      const int ox = tx / s;
      if (ox >= Wo2) continue;
      st_stream4(A + ((token0 + (int64_t)oy * Wo2 + ox) * 9 + ky * 3 + kx) * C + c, acc);
    }
  }
}

__global__ void __launch_bounds__(kImThreads) warp_im2col_kernel(const float* __restrict__ src,
                                                                 const float* __restrict__ Mat, int C, int Hi, int Wi,
                                                                 int Ho, int Wo, int s, int Ho2, int Wo2,
                                                                 float* __restrict__ A) {
  __shared__ float sT[9];
  constexpr unsigned FULL = 0xffffffffu;
  const int n = blockIdx.y;
  if (threadIdx.x == 0) normalized_inverse(Mat + 9 * n, Hi, Wi, Ho, Wo, sT);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Wp = Wo + 2, npix = (Ho + 2) * Wp;  // padded domain
  const int pix0 = blockIdx.x * kImPix;
  const int pix = pix0 + lane;
  const int vp_own = pix / Wp, up_own = pix - vp_own * Wp;
  const bool real = pix < npix && vp_own >= 1 && vp_own <= Ho && up_own >= 1 && up_own <= Wo;
  const Taps t = make_taps(sT, real ? up_own - 1 : 0, real ? vp_own - 1 : 0, Hi, Wi, Ho, Wo);
  const unsigned mk_own = real ? ((unsigned)t.m_nw | ((unsigned)t.m_ne << 1) | ((unsigned)t.m_sw << 2) |
                                  ((unsigned)t.m_se << 3))
                               : 0u;
  const float* sbase = src + (int64_t)n * Hi * Wi * C;
  const int64_t rowC = (int64_t)Wi * C;
  const int64_t token0 = (int64_t)n * Ho2 * Wo2;

  for (int c0 = 0; c0 < C; c0 += 128) {
    const int c = c0 + lane * 4;
    const bool cok = c < C;
#pragma unroll 2
    for (int jj = 0; jj < kImPix / 4; ++jj) {
      const int j = warp * (kImPix / 4) + jj;
      if (pix0 + j >= npix) break;
      const int o = __shfl_sync(FULL, t.o00, j);
      const unsigned mk = __shfl_sync(FULL, mk_own, j);
      const float wnw = __shfl_sync(FULL, t.nw, j), wne = __shfl_sync(FULL, t.ne, j);
      const float wsw = __shfl_sync(FULL, t.sw, j), wse = __shfl_sync(FULL, t.se, j);
      const float* p = sbase + (int64_t)o * C + c;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 q0 = (cok && (mk & 1u)) ? __ldg(reinterpret_cast<const float4*>(p)) : z;
      const float4 q1 = (cok && (mk & 2u)) ? __ldg(reinterpret_cast<const float4*>(p + C)) : z;
      const float4 q2 = (cok && (mk & 4u)) ? __ldg(reinterpret_cast<const float4*>(p + rowC)) : z;
      const float4 q3 = (cok && (mk & 8u)) ? __ldg(reinterpret_cast<const float4*>(p + rowC + C)) : z;
      float4 acc;  // same operation order as warp_fwd_cl_kernel: bit-identical values
      acc.x = fmaf(q3.x, wse, fmaf(q2.x, wsw, fmaf(q1.x, wne, q0.x * wnw)));
      acc.y = fmaf(q3.y, wse, fmaf(q2.y, wsw, fmaf(q1.y, wne, q0.y * wnw)));
      acc.z = fmaf(q3.z, wse, fmaf(q2.z, wsw, fmaf(q1.z, wne, q0.z * wnw)));
      acc.w = fmaf(q3.w, wse, fmaf(q2.w, wsw, fmaf(q1.w, wne, q0.w * wnw)));
      if (cok) {
        const int pj = pix0 + j;
        const int vp = pj / Wp, up = pj - vp * Wp;
        scatter_slots(A, acc, vp, up, c, C, s, Ho2, Wo2, token0);
      }
    }
  }
}

// ATen's area_pixel_compute_source_index for align_corners = false (UpSample.cuh): scale * (dst + 0.5) - 0.5, clamped
// at 0; i1 = i0 + (i0 < in - 1); lambda = src - i0.
struct UpTap {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ UpTap up_tap(int dst, float scale, int in_size) {
  float srcf = scale * ((float)dst + 0.5f) - 0.5f;
  srcf = srcf < 0.f ? 0.f : srcf;
  UpTap t;
  t.i0 = (int)srcf;
  t.i0 = t.i0 > in_size - 1 ? in_size - 1 : t.i0;
  t.i1 = t.i0 + (t.i0 < in_size - 1 ? 1 : 0);
  t.l1 = srcf - (float)t.i0;
  t.l0 = 1.f - t.l1;
  return t;
}

// ATen: h0lambda * (w0lambda * v00 + w1lambda * v01) + h1lambda * (w0lambda * v10 + w1lambda * v11). The operation
// sequence is pinned with intrinsics so that the two kernels below give the same bits (left to the compiler, the im2col
// and the plain kernel contracted different multiply-adds: r02t).
__device__ __forceinline__ float up_blend1(const UpTap& ty, const UpTap& tx, float v00, float v01, float v10, float v11) {
  const float top = __fmaf_rn(tx.l1, v01, __fmul_rn(tx.l0, v00));
  const float bot = __fmaf_rn(tx.l1, v11, __fmul_rn(tx.l0, v10));
  return __fmaf_rn(ty.l1, bot, __fmul_rn(ty.l0, top));
}
__device__ __forceinline__ float4 up_blend(const UpTap& ty, const UpTap& tx, const float4& v00, const float4& v01,
                                           const float4& v10, const float4& v11) {
  return make_float4(up_blend1(ty, tx, v00.x, v01.x, v10.x, v11.x), up_blend1(ty, tx, v00.y, v01.y, v10.y, v11.y),
                     up_blend1(ty, tx, v00.z, v01.z, v10.z, v11.z), up_blend1(ty, tx, v00.w, v01.w, v10.w, v11.w));
}

// Rows [row0, row0 + nrows) of the upsampled grid only (nrows = Ho, row0 = 0: the whole grid): the block index walks
// the padded pixel rows [row0 - 1, row0 + nrows] that feed those tokens.
__global__ void __launch_bounds__(kImThreads) upsample_im2col_kernel(const float* __restrict__ src, int C, int Hi,
                                                                     int Wi, int Ho, int Wo, float scale_h,
                                                                     float scale_w, int row0, int nrows,
                                                                     float* __restrict__ A) {
  const int n = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Wp = Wo + 2, npix = (nrows + 2) * Wp;
  const int pix0 = blockIdx.x * kImPix;
  const float* sbase = src + (int64_t)n * Hi * Wi * C;
  const int64_t token0 = (int64_t)n * nrows * Wo;
  for (int c0 = 0; c0 < C; c0 += 128) {
    const int c = c0 + lane * 4;
    if (c >= C) continue;
#pragma unroll 2
    for (int jj = 0; jj < kImPix / 4; ++jj) {
      const int pj = pix0 + warp * (kImPix / 4) + jj;
      if (pj >= npix) break;
      const int vp = pj / Wp + row0, up = pj - (pj / Wp) * Wp;  // padded coordinates in the FULL grid
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vp >= 1 && vp <= Ho && up >= 1 && up <= Wo) {
        const UpTap ty = up_tap(vp - 1, scale_h, Hi), tx = up_tap(up - 1, scale_w, Wi);
        const float4 v00 = __ldg(reinterpret_cast<const float4*>(sbase + ((int64_t)ty.i0 * Wi + tx.i0) * C + c));
        const float4 v01 = __ldg(reinterpret_cast<const float4*>(sbase + ((int64_t)ty.i0 * Wi + tx.i1) * C + c));
        const float4 v10 = __ldg(reinterpret_cast<const float4*>(sbase + ((int64_t)ty.i1 * Wi + tx.i0) * C + c));
        const float4 v11 = __ldg(reinterpret_cast<const float4*>(sbase + ((int64_t)ty.i1 * Wi + tx.i1) * C + c));
        acc = up_blend(ty, tx, v00, v01, v10, v11);
      }
      scatter_slots(A, acc, vp, up, c, C, 1, Ho, Wo, token0, row0, row0 + nrows);
    }
  }
}

// Plain bilinear upsample (align_corners = false, ATen arithmetic as above) of a channels-last map, rows [row0, row0 +
// nrows) of the output only: the input side of the implicit-GEMM 3x3 convolution (csrc/gemm_bf16x3.cu), which fetches
// its taps by TMA instead of reading an im2col matrix. One thread = one pixel x 4 channels.
__global__ void __launch_bounds__(256) upsample_nhwc_kernel(const float* __restrict__ src, int C, int Hi, int Wi, int Ho,
                                                            int Wo, float scale_h, float scale_w, int row0, int nrows,
                                                            float* __restrict__ dst) {
  const int n = blockIdx.y;
  const int c4 = C / 4;
  const int64_t total = (int64_t)nrows * Wo * c4;
  const float* sbase = src + (int64_t)n * Hi * Wi * C;
  float* dbase = dst + (int64_t)n * Ho * Wo * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4) * 4;
    const int64_t p = i / c4;
    const int v = (int)(p / Wo) + row0, u = (int)(p % Wo);
    const UpTap ty = up_tap(v, scale_h, Hi), tx = up_tap(u, scale_w, Wi);
    const float4 v00 = __ldg(reinterpret_cast<const float4*>(sbase + ((int64_t)ty.i0 * Wi + tx.i0) * C + c));
    const float4 v01 = __ldg(reinterpret_cast<const float4*>(sbase + ((int64_t)ty.i0 * Wi + tx.i1) * C + c));
    const float4 v10 = __ldg(reinterpret_cast<const float4*>(sbase + ((int64_t)ty.i1 * Wi + tx.i0) * C + c));
    const float4 v11 = __ldg(reinterpret_cast<const float4*>(sbase + ((int64_t)ty.i1 * Wi + tx.i1) * C + c));
    const float4 acc = up_blend(ty, tx, v00, v01, v10, v11);
    *reinterpret_cast<float4*>(dbase + ((int64_t)v * Wo + u) * C + c) = acc;
  }
}

}  // namespace mvd

using namespace mvd;

extern "C" int mvd_upsample_nhwc_f32(const float* src, int BN, int C, int Hi, int Wi, int Ho, int Wo, int row0, int nrows,
                                     float* dst, void* stream) {
  if (!src || !dst) return MVD_ERR_NULL_POINTER;
  if (BN <= 0 || C <= 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0 || BN > 65535) return MVD_ERR_BAD_SHAPE;
  if (row0 < 0 || nrows <= 0 || row0 + nrows > Ho) return MVD_ERR_BAD_SHAPE;
  if (C % 4 != 0) return MVD_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) return MVD_ERR_MISALIGNED;
  const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;  // ATen area_pixel_compute_scale
  const int64_t total = (int64_t)nrows * Wo * (C / 4);
  const int64_t want = ceil_div64(total, 256);
  dim3 grid((unsigned)(want > (int64_t)kNumSMs * 16 ? (int64_t)kNumSMs * 16 : want), (unsigned)BN);
  upsample_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, C, Hi, Wi, Ho, Wo, sh, sw, row0, nrows, dst);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

extern "C" int mvd_warp_im2col_f32(const float* src, const float* Mat, int BN, int C, int Hi, int Wi, int Ho, int Wo,
                                   int stride, float* A, void* stream) {
  if (!src || !Mat || !A) return MVD_ERR_NULL_POINTER;
  if (BN <= 0 || C <= 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0 || BN > 65535) return MVD_ERR_BAD_SHAPE;
  if (stride != 1 && stride != 2) return MVD_ERR_UNSUPPORTED;
  if (C % 4 != 0) return MVD_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(A)) & 15u) return MVD_ERR_MISALIGNED;
  const int Ho2 = (Ho + 2 - 3) / stride + 1, Wo2 = (Wo + 2 - 3) / stride + 1;
  if ((int64_t)BN * Ho2 * Wo2 * 9 * C > 0x7fffffffffLL) return MVD_ERR_BAD_SHAPE;
  dim3 grid((unsigned)ceil_div64((int64_t)(Ho + 2) * (Wo + 2), kImPix), (unsigned)BN);
  warp_im2col_kernel<<<grid, kImThreads, 0, (cudaStream_t)stream>>>(src, Mat, C, Hi, Wi, Ho, Wo, stride, Ho2, Wo2, A);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

extern "C" int mvd_upsample_im2col_rows_f32(const float* src, int BN, int C, int Hi, int Wi, int Ho, int Wo, int row0,
                                            int nrows, float* A, void* stream) {
  if (!src || !A) return MVD_ERR_NULL_POINTER;
  if (BN <= 0 || C <= 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0 || BN > 65535) return MVD_ERR_BAD_SHAPE;
  if (row0 < 0 || nrows <= 0 || row0 + nrows > Ho) return MVD_ERR_BAD_SHAPE;
  if (C % 4 != 0) return MVD_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(A)) & 15u) return MVD_ERR_MISALIGNED;
  // ATen area_pixel_compute_scale<float>(in, out, align_corners=false, scale=nullopt): (float)in / out
  const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;
  dim3 grid((unsigned)ceil_div64((int64_t)(nrows + 2) * (Wo + 2), kImPix), (unsigned)BN);
  upsample_im2col_kernel<<<grid, kImThreads, 0, (cudaStream_t)stream>>>(src, C, Hi, Wi, Ho, Wo, sh, sw, row0, nrows, A);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

extern "C" int mvd_upsample_im2col_f32(const float* src, int BN, int C, int Hi, int Wi, int Ho, int Wo, float* A,
                                       void* stream) {
  return mvd_upsample_im2col_rows_f32(src, BN, C, Hi, Wi, Ho, Wo, 0, Ho, A, stream);
}
