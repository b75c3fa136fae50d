// Input side of the path: camera frame (uint8, HWC as decoded) -> normalised, resized network input.
//   replaces T.Compose([T.ToTensor(), T.Normalize(mean, std), T.Resize((H*8/img_reduce, W*8/img_reduce))])
//       ref: multiview_detector/datasets/frameDataset.py:66-67   (1080x1920 -> 720x1280 per view, on the CPU in the
//       reference's data-loader workers; 7 views x 2.8 M output floats per frame)
// One kernel: thread = output pixel (all 3 channels). ToTensor/Normalize are applied per tap exactly as torch rounds
// them, ((u8 / 255) - mean) / std in fp32 (a 256-entry table per channel in shared memory), then the taps are combined
// with ATen's interpolation weights:
//   antialias != 0   torchvision >= 0.17 default for tensors = F.interpolate(mode='bilinear', antialias=True):
//                    ATen _upsample_bilinear2d_aa (UpSampleKernel.cpp, _compute_indices_min_size_weights_aa): triangle
//                    filter of support `scale` (when down-scaling), weights normalised per output coordinate, separable
//                    (rows of horizontally filtered values are then filtered vertically);
//   antialias == 0   plain bilinear, align_corners=False (ATen upsample_bilinear2d; what torchvision < 0.17 did).
// Channel order follows the input (RGB in, RGB out): out[n, c, y, x].
#include "common.cuh"

namespace mvd {
namespace {

constexpr int kPpThreads = 256;
constexpr int kPpMaxTaps = 16;  // per dimension; covers down-scaling factors up to 7.5

struct AaTaps {
  int lo, n;
  float w[kPpMaxTaps];
};

// ATen: center = scale * (i + 0.5); support = max(scale, 1) (bilinear: interp_size 2 * 0.5); weights = triangle filter
// evaluated at (j + lo - center + 0.5) * invscale, normalised by their sum.
__device__ __forceinline__ AaTaps aa_taps(int i, float scale, int in_size) {
  const float support = scale >= 1.f ? scale : 1.f;
  const float invscale = scale >= 1.f ? 1.f / scale : 1.f;
  const float center = scale * ((float)i + 0.5f);
  AaTaps t;
  t.lo = max((int)(center - support + 0.5f), 0);
  t.n = min(min((int)(center + support + 0.5f), in_size) - t.lo, kPpMaxTaps);
  float total = 0.f;
#pragma unroll
  for (int j = 0; j < kPpMaxTaps; ++j) {
    float w = 0.f;
    if (j < t.n) {
      float x = ((float)(j + t.lo) - center + 0.5f) * invscale;
      x = fabsf(x);
      w = x < 1.f ? 1.f - x : 0.f;
    }
    t.w[j] = w;
    total += w;
  }
  if (total != 0.f) {
#pragma unroll
    for (int j = 0; j < kPpMaxTaps; ++j) t.w[j] = t.w[j] / total;
  }
  return t;
}

template <bool AA>
__global__ void __launch_bounds__(kPpThreads) resize_normalize_kernel(const unsigned char* __restrict__ img, int Hi,
                                                                      int Wi, int Ho, int Wo, float sh, float sw,
                                                                      float m0, float m1, float m2, float s0, float s1,
                                                                      float s2, float* __restrict__ out) {
  __shared__ float lut[3][256];
  for (int i = threadIdx.x; i < 768; i += kPpThreads) {
    const int c = i >> 8, v = i & 255;
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    lut[c][v] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), mean), sd);  // ToTensor then Normalize, as torch rounds
  }
  __syncthreads();
  const int n = blockIdx.z;
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= Wo || y >= Ho) return;
  const unsigned char* base = img + (int64_t)n * Hi * Wi * 3;
  float r0 = 0.f, r1 = 0.f, r2 = 0.f;
  if (AA) {
    const AaTaps tx = aa_taps(x, sw, Wi), ty = aa_taps(y, sh, Hi);
    for (int j = 0; j < ty.n; ++j) {
      const unsigned char* row = base + ((int64_t)(ty.lo + j) * Wi + tx.lo) * 3;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 4
      for (int i = 0; i < tx.n; ++i) {
        const float w = tx.w[i];
        a0 = fmaf(lut[0][row[3 * i]], w, a0);
        a1 = fmaf(lut[1][row[3 * i + 1]], w, a1);
        a2 = fmaf(lut[2][row[3 * i + 2]], w, a2);
      }
      r0 = fmaf(a0, ty.w[j], r0);
      r1 = fmaf(a1, ty.w[j], r1);
      r2 = fmaf(a2, ty.w[j], r2);
    }
  } else {
    // ATen area_pixel_compute_source_index(align_corners=false): src = scale * (dst + 0.5) - 0.5, clamped at 0
    float fy = sh * ((float)y + 0.5f) - 0.5f, fx = sw * ((float)x + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = min((int)fy, Hi - 1), x0 = min((int)fx, Wi - 1);
    const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0), x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
    const unsigned char* p00 = base + ((int64_t)y0 * Wi + x0) * 3;
    const unsigned char* p01 = base + ((int64_t)y0 * Wi + x1) * 3;
    const unsigned char* p10 = base + ((int64_t)y1 * Wi + x0) * 3;
    const unsigned char* p11 = base + ((int64_t)y1 * Wi + x1) * 3;
    r0 = hy * (hx * lut[0][p00[0]] + lx * lut[0][p01[0]]) + ly * (hx * lut[0][p10[0]] + lx * lut[0][p11[0]]);
    r1 = hy * (hx * lut[1][p00[1]] + lx * lut[1][p01[1]]) + ly * (hx * lut[1][p10[1]] + lx * lut[1][p11[1]]);
    r2 = hy * (hx * lut[2][p00[2]] + lx * lut[2][p01[2]]) + ly * (hx * lut[2][p10[2]] + lx * lut[2][p11[2]]);
  }
  const int64_t plane = (int64_t)Ho * Wo;
  float* o = out + (int64_t)n * 3 * plane + (int64_t)y * Wo + x;
  __stcs(o, r0);
  __stcs(o + plane, r1);
  __stcs(o + 2 * plane, r2);
}

}  // namespace
}  // namespace mvd

using namespace mvd;

extern "C" int mvd_resize_normalize_u8(const unsigned char* img, int N, int Hi, int Wi, int Ho, int Wo,
                                       const float* mean_host, const float* std_host, int antialias, float* out,
                                       void* stream) {
  if (!img || !mean_host || !std_host || !out) return MVD_ERR_NULL_POINTER;
  if (N <= 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0 || N > 65535) return MVD_ERR_BAD_SHAPE;
  if ((int64_t)Hi * Wi > 0x2fffffffLL || (int64_t)Ho * Wo > 0x2fffffffLL) return MVD_ERR_BAD_SHAPE;
  const float sh = (float)Hi / (float)Ho, sw = (float)Wi / (float)Wo;  // ATen area_pixel_compute_scale, no scale_factor
  if (antialias && ((int)ceilf(sh >= 1.f ? sh : 1.f) * 2 + 1 > kPpMaxTaps || (int)ceilf(sw >= 1.f ? sw : 1.f) * 2 + 1 > kPpMaxTaps))
    return MVD_ERR_UNSUPPORTED;
  dim3 grid((unsigned)((Wo + 31) / 32), (unsigned)((Ho + 7) / 8), (unsigned)N);
  if (grid.y > 65535) return MVD_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  if (antialias)
    resize_normalize_kernel<true><<<grid, kPpThreads, 0, st>>>(img, Hi, Wi, Ho, Wo, sh, sw, mean_host[0], mean_host[1],
                                                                mean_host[2], std_host[0], std_host[1], std_host[2], out);
  else
    resize_normalize_kernel<false><<<grid, kPpThreads, 0, st>>>(img, Hi, Wi, Ho, Wo, sh, sw, mean_host[0], mean_host[1],
                                                                 mean_host[2], std_host[0], std_host[1], std_host[2], out);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}
