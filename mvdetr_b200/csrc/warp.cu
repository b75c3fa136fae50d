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
#include "warp_taps.cuh"

namespace mvd {

constexpr int kWarpThreads = 128;

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


// ---------------------------------------------------------------------------------------------
// Channels-last SOURCE path ([BN, Hi, Wi, C], C % 4 == 0): every tap is C contiguous floats, so a warp fetches a tap
// with one 128-bit load per lane (512 B per request at C = 128) instead of 32 scalar loads spread over C planes.
// A block owns 32 consecutive destination pixels of one view: lane j evaluates the homography for pixel j once, the
// 4 warps take 8 pixels each and receive the tap offset / weights / mask by shuffle, lanes stride the channel quads.
//   DST_CL  : each warp stores whole pixels, 512 B coalesced (st.global.L1::no_allocate.v4)
//   !DST_CL : results go through a [channel][pixel] shared tile (column rotated by the channel quad: conflict-free
//             both ways) and leave as 128-byte rows of one channel plane (NCHW, the kornia contract).
// ---------------------------------------------------------------------------------------------
constexpr int kClPix = 32, kClChunk = 128;

template <bool DST_CL>
__global__ void __launch_bounds__(kWarpThreads) warp_fwd_cl_kernel(const float* __restrict__ src,
                                                                   const float* __restrict__ Mat, int C, int Hi,
                                                                   int Wi, int Ho, int Wo, float* __restrict__ dst) {
  __shared__ float sT[9];
  __shared__ float tile[DST_CL ? 1 : kClChunk * kClPix];
  constexpr unsigned FULL = 0xffffffffu;
  const int n = blockIdx.y;
  if (threadIdx.x == 0) normalized_inverse(Mat + 9 * n, Hi, Wi, Ho, Wo, sT);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npix = Ho * Wo;
  const int pix0 = blockIdx.x * kClPix;
  const int pix = pix0 + lane;
  const bool in = pix < npix;
  const int v = in ? pix / Wo : 0, u = in ? pix - v * Wo : 0;
  const Taps t = make_taps(sT, u, v, Hi, Wi, Ho, Wo);
  const unsigned mk_own = in ? ((unsigned)t.m_nw | ((unsigned)t.m_ne << 1) | ((unsigned)t.m_sw << 2) |
                                ((unsigned)t.m_se << 3))
                             : 0u;
  const float* sbase = src + (int64_t)n * Hi * Wi * C;
  const int64_t rowC = (int64_t)Wi * C;

  for (int c0 = 0; c0 < C; c0 += kClChunk) {
    const int c = c0 + lane * 4;
    const bool cok = c < C;
#pragma unroll 4
    for (int jj = 0; jj < kClPix / 4; ++jj) {
      const int j = warp * (kClPix / 4) + jj;
      const int o = __shfl_sync(FULL, t.o00, j);
      const unsigned mk = __shfl_sync(FULL, mk_own, j);
      const float wnw = __shfl_sync(FULL, t.nw, j), wne = __shfl_sync(FULL, t.ne, j);
      const float wsw = __shfl_sync(FULL, t.sw, j), wse = __shfl_sync(FULL, t.se, j);
      // branch-free: predicated 128-bit loads, so the unrolled pixels keep 16 loads in flight per lane
      const float* p = sbase + (int64_t)o * C + c;
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 q0 = (cok && (mk & 1u)) ? __ldg(reinterpret_cast<const float4*>(p)) : z;
      const float4 q1 = (cok && (mk & 2u)) ? __ldg(reinterpret_cast<const float4*>(p + C)) : z;
      const float4 q2 = (cok && (mk & 4u)) ? __ldg(reinterpret_cast<const float4*>(p + rowC)) : z;
      const float4 q3 = (cok && (mk & 8u)) ? __ldg(reinterpret_cast<const float4*>(p + rowC + C)) : z;
      float4 acc;
      acc.x = fmaf(q3.x, wse, fmaf(q2.x, wsw, fmaf(q1.x, wne, q0.x * wnw)));
      acc.y = fmaf(q3.y, wse, fmaf(q2.y, wsw, fmaf(q1.y, wne, q0.y * wnw)));
      acc.z = fmaf(q3.z, wse, fmaf(q2.z, wsw, fmaf(q1.z, wne, q0.z * wnw)));
      acc.w = fmaf(q3.w, wse, fmaf(q2.w, wsw, fmaf(q1.w, wne, q0.w * wnw)));
      if (DST_CL) {
        if (cok && pix0 + j < npix) st_stream4(dst + ((int64_t)n * npix + pix0 + j) * C + c, acc);
      } else {
        const int col = (j + lane) & 31;
        tile[(lane * 4 + 0) * kClPix + col] = acc.x;
        tile[(lane * 4 + 1) * kClPix + col] = acc.y;
        tile[(lane * 4 + 2) * kClPix + col] = acc.z;
        tile[(lane * 4 + 3) * kClPix + col] = acc.w;
      }
    }
    if (!DST_CL) {
      __syncthreads();
      const int nch = min(kClChunk, C - c0);
      if (in) {
        for (int cc = warp; cc < nch; cc += 4)
          __stcs(dst + ((int64_t)n * C + c0 + cc) * npix + pix, tile[cc * kClPix + ((lane + (cc >> 2)) & 31)]);
      }
      __syncthreads();
    }
  }
}

// Backward for channels-last tensors: grad_dst [BN, Ho, Wo, C] -> grad_src [BN, Hi, Wi, C] (pre-zeroed). Same
// decomposition as the forward; every tap is one vector reduction (red.global.add.v4.f32) per lane, i.e. 512 B
// contiguous per tap at C = 128, instead of C scalar atomics scattered over C planes.
__global__ void __launch_bounds__(kWarpThreads) warp_bwd_cl_kernel(const float* __restrict__ grad_dst,
                                                                   const float* __restrict__ Mat, int C, int Hi,
                                                                   int Wi, int Ho, int Wo,
                                                                   float* __restrict__ grad_src) {
  __shared__ float sT[9];
  constexpr unsigned FULL = 0xffffffffu;
  const int n = blockIdx.y;
  if (threadIdx.x == 0) normalized_inverse(Mat + 9 * n, Hi, Wi, Ho, Wo, sT);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npix = Ho * Wo;
  const int pix0 = blockIdx.x * kClPix;
  const int pix = pix0 + lane;
  const bool in = pix < npix;
  const int v = in ? pix / Wo : 0, u = in ? pix - v * Wo : 0;
  const Taps t = make_taps(sT, u, v, Hi, Wi, Ho, Wo);
  const unsigned mk_own = in ? ((unsigned)t.m_nw | ((unsigned)t.m_ne << 1) | ((unsigned)t.m_sw << 2) |
                                ((unsigned)t.m_se << 3))
                             : 0u;
  float* gbase = grad_src + (int64_t)n * Hi * Wi * C;
  const int64_t rowC = (int64_t)Wi * C;
#pragma unroll 2
  for (int jj = 0; jj < kClPix / 4; ++jj) {
    const int j = warp * (kClPix / 4) + jj;
    const int o = __shfl_sync(FULL, t.o00, j);
    const unsigned mk = __shfl_sync(FULL, mk_own, j);
    const float wnw = __shfl_sync(FULL, t.nw, j), wne = __shfl_sync(FULL, t.ne, j);
    const float wsw = __shfl_sync(FULL, t.sw, j), wse = __shfl_sync(FULL, t.se, j);
    if (!mk) continue;
    const float* gp = grad_dst + ((int64_t)n * npix + pix0 + j) * C;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 g = ld_stream4(gp + c);
      float* p = gbase + (int64_t)o * C + c;
      if (mk & 1u) red_add4(p, g.x * wnw, g.y * wnw, g.z * wnw, g.w * wnw);
      if (mk & 2u) red_add4(p + C, g.x * wne, g.y * wne, g.z * wne, g.w * wne);
      if (mk & 4u) red_add4(p + rowC, g.x * wsw, g.y * wsw, g.z * wsw, g.w * wsw);
      if (mk & 8u) red_add4(p + rowC + C, g.x * wse, g.y * wse, g.z * wse, g.w * wse);
    }
  }
}

// Batched 2-D transpose in[b][r][c] -> out[b][c][r] (fp32): NCHW <-> NHWC relayout ([BN, C, H*W] <-> [BN, H*W, C])
// for callers that hold the other layout. 32x32 tiles through shared memory, both sides 128-byte coalesced.
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int rows, int cols,
                                                        float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* ib = in + (int64_t)b * rows * cols;
  float* ob = out + (int64_t)b * rows * cols;
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + i][tx] = ld_stream(ib + (int64_t)r * cols + c);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;
    if (r < rows && c < cols) __stcs(ob + (int64_t)c * rows + r, tile[tx][ty + i]);
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
                                float* dst, int layout, void* stream) {
  if (!src || !Mat || !dst) return MVD_ERR_NULL_POINTER;
  if (int e = check_warp_dims(BN, C, Hi, Wi, Ho, Wo)) return e;
  if (layout & ~(MVD_WARP_DST_NHWC | MVD_WARP_SRC_NHWC)) return MVD_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const bool dst_cl = layout & MVD_WARP_DST_NHWC;
  if (layout & MVD_WARP_SRC_NHWC) {
    if (C % 4 != 0) return MVD_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) return MVD_ERR_MISALIGNED;
    dim3 grid((unsigned)ceil_div64((int64_t)Ho * Wo, kClPix), (unsigned)BN);
    if (dst_cl)
      warp_fwd_cl_kernel<true><<<grid, kWarpThreads, 0, st>>>(src, Mat, C, Hi, Wi, Ho, Wo, dst);
    else
      warp_fwd_cl_kernel<false><<<grid, kWarpThreads, 0, st>>>(src, Mat, C, Hi, Wi, Ho, Wo, dst);
  } else if (dst_cl) {
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

extern "C" int mvd_warp_bwd_nhwc_f32(const float* grad_dst, const float* Mat, int BN, int C, int Hi, int Wi, int Ho,
                                     int Wo, float* grad_src, void* stream) {
  if (!grad_dst || !Mat || !grad_src) return MVD_ERR_NULL_POINTER;
  if (int e = check_warp_dims(BN, C, Hi, Wi, Ho, Wo)) return e;
  if (C % 4 != 0) return MVD_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(grad_dst) | reinterpret_cast<uintptr_t>(grad_src)) & 15u) return MVD_ERR_MISALIGNED;
  cudaStream_t st = (cudaStream_t)stream;
  MVD_CUDA_TRY(cudaMemsetAsync(grad_src, 0, sizeof(float) * (size_t)BN * C * Hi * Wi, st));
  dim3 grid((unsigned)ceil_div64((int64_t)Ho * Wo, kClPix), (unsigned)BN);
  warp_bwd_cl_kernel<<<grid, kWarpThreads, 0, st>>>(grad_dst, Mat, C, Hi, Wi, Ho, Wo, grad_src);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

extern "C" int mvd_transpose_f32(const float* in, int batch, int rows, int cols, float* out, void* stream) {
  if (!in || !out) return MVD_ERR_NULL_POINTER;
  if (batch <= 0 || rows <= 0 || cols <= 0 || batch > 65535) return MVD_ERR_BAD_SHAPE;
  const int64_t gy = ceil_div64(rows, 32), gx = ceil_div64(cols, 32);
  if (gy > 65535) return MVD_ERR_BAD_SHAPE;
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)batch);
  transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, rows, cols, out);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}
