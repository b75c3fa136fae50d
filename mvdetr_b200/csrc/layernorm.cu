// Fused residual add + LayerNorm over the last dimension: out[r,:] = LN(x[r,:] + (res[r,:] + res_bias)) * gamma + beta.
// The encoder layer does this twice per layer on [N*Hd*Wd, C] tokens (eval mode, dropout = identity):
//   ref: multiview_detector/models/deformable_transformer.py:79-80 and :84-85
// torch runs it as an elementwise add (3 x 38.7 MB of traffic) plus a LayerNorm kernel with one 128-thread block
// per 128-float row (measured 15 + 114 us at Wildtrack size, profiles/r01_launches.md); here one warp owns a row,
// keeps it in registers (float4 per lane per 128 channels), and the tensor is read and written once.
#include "common.cuh"

namespace mvd {

template <int VPL>  // float4 vectors per lane: C <= 128 * VPL
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                            const float* __restrict__ res_bias,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int64_t rows, int C,
                                                            float eps, int64_t perm_inner, float* __restrict__ out,
                                                            const float* __restrict__ pos, float* __restrict__ out2) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = C >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + row * C);
  const float4* rr = res ? reinterpret_cast<const float4*>(res + row * C) : nullptr;
  float4 v[VPL];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + 32 * k;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nvec) {
      v[k] = __ldcs(xr + i);
      if (rr) {
        float4 r = __ldcs(rr + i);
        if (res_bias) {  // res is a bias-free GEMM output: (acc + bias) first, as the Linear layer rounds it
          const float4 rb = __ldg(reinterpret_cast<const float4*>(res_bias) + i);
          r.x += rb.x, r.y += rb.y, r.z += rb.z, r.w += rb.w;
        }
        v[k].x += r.x;
        v[k].y += r.y;
        v[k].z += r.z;
        v[k].w += r.w;
      }
      sum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    if (lane + 32 * k < nvec) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = 1.f / sqrtf(sq / (float)C + eps);
  // perm_inner > 0: rows are [outer][inner]; write them as [inner][outer] (view-major tokens -> cell-major, so the
  // merge convolution over all views is one GEMM over contiguous rows; ref: trans_world_feat.py:107-108)
  const int64_t orow_idx = perm_inner > 0 ? (row % perm_inner) * (rows / perm_inner) + row / perm_inner : row;
  float4* orow = reinterpret_cast<float4*>(out + orow_idx * C);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int i = lane + 32 * k;
    if (i < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i);
      float4 o;
      o.x = (v[k].x - mean) * rstd * g.x + b.x;
      o.y = (v[k].y - mean) * rstd * g.y + b.y;
      o.z = (v[k].z - mean) * rstd * g.z + b.z;
      o.w = (v[k].w - mean) * rstd * g.w + b.w;
      orow[i] = o;
      if (out2) {  // second output: the next layer's query, out + position embedding (deformable_transformer.py:71-77)
        const float4 p = __ldg(reinterpret_cast<const float4*>(pos + row * C) + i);
        reinterpret_cast<float4*>(out2 + row * C)[i] = make_float4(o.x + p.x, o.y + p.y, o.z + p.z, o.w + p.w);
      }
    }
  }
}

// In-place x[r, c] = act(x[r, c] + bias[c]) for the output of a bias-free GEMM (act = ReLU or identity). cuBLASLt
// applies an fp32 bias/ReLU epilogue of its SIMT kernels as a separate scalar pass (profiles/r01a_launches.csv:
// `cublasLt::globalKernel`, 175 us for the [75600, 512] FFN hidden at 1.8 TB/s); this is the same pass with 128-bit
// accesses.   ref: multiview_detector/models/deformable_transformer.py:82 (linear1 + ReLU), ms_deform_attn.py:96
template <bool RELU>
__global__ void __launch_bounds__(256) bias_act_kernel(float* __restrict__ x, const float* __restrict__ bias,
                                                       int64_t n4, int C4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(x)[i];
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + (int)(i % C4));
    v.x += b.x, v.y += b.y, v.z += b.z, v.w += b.w;
    if (RELU) v.x = fmaxf(v.x, 0.f), v.y = fmaxf(v.y, 0.f), v.z = fmaxf(v.z, 0.f), v.w = fmaxf(v.w, 0.f);
    reinterpret_cast<float4*>(x)[i] = v;
  }
}

}  // namespace mvd

using namespace mvd;

extern "C" int mvd_bias_act_f32(float* x, const float* bias, int64_t rows, int C, int relu, void* stream) {
  if (!x || !bias) return MVD_ERR_NULL_POINTER;
  if (rows <= 0 || C <= 0) return MVD_ERR_BAD_SHAPE;
  if (C & 3) return MVD_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(bias)) & 15u) return MVD_ERR_MISALIGNED;
  const int64_t n4 = rows * (C / 4);
  const int64_t want = ceil_div64(n4, 256);
  const int blocks = (int)(want > (int64_t)kNumSMs * 16 ? (int64_t)kNumSMs * 16 : want);
  if (relu)
    bias_act_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, bias, n4, C / 4);
  else
    bias_act_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, bias, n4, C / 4);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

extern "C" int mvd_add_layernorm_pos_f32(const float* x, const float* res, const float* res_bias, const float* gamma,
                                         const float* beta, int64_t rows, int C, float eps, int64_t perm_inner,
                                         float* out, const float* pos, float* out2, void* stream) {
  if (!x || !gamma || !beta || !out) return MVD_ERR_NULL_POINTER;
  if ((pos == nullptr) != (out2 == nullptr)) return MVD_ERR_NULL_POINTER;
  if (out2 && perm_inner > 0) return MVD_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(out2)) & 15u) return MVD_ERR_MISALIGNED;
  if (rows <= 0 || C <= 0) return MVD_ERR_BAD_SHAPE;
  if ((C & 3) || C > 1024) return MVD_ERR_UNSUPPORTED;
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                       reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
                       (res ? reinterpret_cast<uintptr_t>(res) : 0) |
                       (res_bias ? reinterpret_cast<uintptr_t>(res_bias) : 0);
  if (res_bias && !res) return MVD_ERR_NULL_POINTER;
  if (perm_inner < 0 || (perm_inner > 0 && (rows % perm_inner != 0 || out == x || out == res))) return MVD_ERR_BAD_SHAPE;
  if (al & 15u) return MVD_ERR_MISALIGNED;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t blocks = ceil_div64(rows, 8);
  if (blocks > 0x7fffffffLL) return MVD_ERR_BAD_SHAPE;
  if (C <= 128)
    add_layernorm_kernel<1><<<(int)blocks, 256, 0, st>>>(x, res, res_bias, gamma, beta, rows, C, eps, perm_inner, out, pos, out2);
  else if (C <= 256)
    add_layernorm_kernel<2><<<(int)blocks, 256, 0, st>>>(x, res, res_bias, gamma, beta, rows, C, eps, perm_inner, out, pos, out2);
  else if (C <= 512)
    add_layernorm_kernel<4><<<(int)blocks, 256, 0, st>>>(x, res, res_bias, gamma, beta, rows, C, eps, perm_inner, out, pos, out2);
  else
    add_layernorm_kernel<8><<<(int)blocks, 256, 0, st>>>(x, res, res_bias, gamma, beta, rows, C, eps, perm_inner, out, pos, out2);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

extern "C" int mvd_add_layernorm_f32(const float* x, const float* res, const float* res_bias, const float* gamma,
                                     const float* beta, int64_t rows, int C, float eps, int64_t perm_inner,
                                     float* out, void* stream) {
  return mvd_add_layernorm_pos_f32(x, res, res_bias, gamma, beta, rows, C, eps, perm_inner, out, nullptr, nullptr, stream);
}
