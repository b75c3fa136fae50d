// Multi-scale deformable attention, forward -- generic kernels (any level shapes, any Lq).
//
// Semantics follow the reference kernel ms_deformable_im2col_gpu_kernel
//   (ref: multiview_detector/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299, bilinear :33-84):
//   pixel-centre convention h_im = y*H - 0.5, strict (-1, H) validity window, per-corner zero padding,
//   value laid out [B, S, M, D] so one head-pixel is D contiguous scalars.
// The design is not the reference's (one thread per output channel, 1024-thread blocks):
//   * vec4 path (fp32, D = 4..128 power of two): D/4 lanes own one (b,q,m) pair, every corner is ONE
//     128-bit load per lane, so a D=16 head-pixel (64 B) is fetched by 4 lanes in one request;
//   * level geometry is converted once per block into shared memory (no int64 loads in the inner loop);
//   * loc / attn are read through the non-allocating path so L1 is left to the gathered value lines;
//   * optional FUSED prologue computes loc = ref + off/(W,H) and the softmax over L*P in registers
//     (ref: multiview_detector/models/ops/modules/ms_deform_attn.py:100-107), see mvd_msda_fused_fwd_f32.
//   * scalar path (fp32/fp64, any D) for odd head dims and the fp64 gradcheck contract
//     (ref: multiview_detector/models/ops/test.py:63-86).
#include "common.cuh"

namespace mvd {

// ---------------------------------------------------------------------------------------------
// scalar path: one thread per output element
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) msda_fwd_scalar_kernel(
    const T* __restrict__ value, const int64_t* __restrict__ shapes, const int64_t* __restrict__ start,
    const T* __restrict__ loc, const T* __restrict__ attn, int S, int M, int D, int L, int Lq, int P,
    int64_t n_out, T* __restrict__ out) {
  extern __shared__ Level s_lvl[];
  load_levels(s_lvl, shapes, start, L);
  __syncthreads();

  const int64_t stride_px = (int64_t)M * D;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n_out;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % D);
    const int64_t pair = idx / D;  // (b*Lq + q)*M + m
    const int m = (int)(pair % M);
    const int64_t b = pair / ((int64_t)M * Lq);
    const T* vb = value + (b * S * M + m) * (int64_t)D + c;
    const T* lp = loc + pair * L * P * 2;
    const T* ap = attn + pair * L * P;
    T acc = 0;
    for (int l = 0; l < L; ++l) {
      const Level lv = s_lvl[l];
      const T* vl = vb + (int64_t)lv.start * stride_px;
      for (int p = 0; p < P; ++p) {
        const T x = lp[0], y = lp[1], a = ap[0];
        lp += 2;
        ap += 1;
        const T h_im = y * lv.H - (T)0.5;
        const T w_im = x * lv.W - (T)0.5;
        if (h_im > (T)-1 && w_im > (T)-1 && h_im < (T)lv.H && w_im < (T)lv.W) {
          const int h0 = (int)floor(h_im), w0 = (int)floor(w_im);
          const T lh = h_im - h0, lw = w_im - w0, hh = 1 - lh, hw = 1 - lw;
          const bool top = h0 >= 0, bot = h0 + 1 <= lv.H - 1, lef = w0 >= 0, rig = w0 + 1 <= lv.W - 1;
          const T* p00 = vl + ((int64_t)h0 * lv.W + w0) * stride_px;
          const T v1 = (top && lef) ? p00[0] : (T)0;
          const T v2 = (top && rig) ? p00[stride_px] : (T)0;
          const T v3 = (bot && lef) ? p00[(int64_t)lv.W * stride_px] : (T)0;
          const T v4 = (bot && rig) ? p00[(int64_t)(lv.W + 1) * stride_px] : (T)0;
          const T val = hh * hw * v1 + hh * lw * v2 + lh * hw * v3 + lh * lw * v4;
          acc += val * a;
        }
      }
    }
    out[idx] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// vec4 path: G = D/4 lanes own one (b,q,m) pair. Work is split in two phases per chunk of G samples:
//   prep   : lane j of the group evaluates sample (s0 + j) -- loads its loc/attn (or offset/logit/ref when FUSED)
//            exactly once (no broadcast loads), and reduces it to 4 attention-scaled corner weights plus one
//            packed word (top-left pixel index + 4 corner-valid bits);
//   gather : the group walks the G samples; the 5 words come from the owning lane by width-G shuffles, every lane
//            issues one predicated 128-bit load per corner (a D=16 head-pixel = 64 B = 4 lanes) and 16 FMAs.
// Measured motivation (profiles/r01_*): the first version recomputed the sample arithmetic in all G lanes and was
// issue-bound (72 % issue-active, 182 warp-instructions per 8 samples).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

constexpr int kPxBias = 1 << 27;  // packed pixel index = px + kPxBias in the low 28 bits (host checks 2*S+2 <= 2^27)

template <int D, bool FUSED>
__global__ void __launch_bounds__(256) msda_fwd_vec4_kernel(
    const float* __restrict__ value, const int64_t* __restrict__ shapes, const int64_t* __restrict__ start,
    const float* __restrict__ loc_or_off, const float* __restrict__ attn_or_logit, const float* __restrict__ ref,
    const float* __restrict__ off_bias, const float* __restrict__ logit_bias, int S, int M, int L, int Lq, int P,
    int Lr, int64_t n_pairs, float* __restrict__ out,
    float* __restrict__ attn_out, float* __restrict__ loc_out) {
  constexpr int G = D / 4;  // lanes per pair
  constexpr int PAIRS = 256 / G;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ Level s_lvl[];
  load_levels(s_lvl, shapes, start, L);
  __syncthreads();

  const int sub = threadIdx.x % G;
  const int64_t pair_raw = (int64_t)blockIdx.x * PAIRS + threadIdx.x / G;
  const bool valid = pair_raw < n_pairs;  // tail lanes recompute the last pair (keeps shuffles full-warp)
  const int64_t pair = valid ? pair_raw : n_pairs - 1;
  const int m = (int)(pair % M);
  const int64_t bq = pair / M;
  const int64_t b = bq / Lq;
  const int q = (int)(bq % Lq);
  const int LP = L * P;
  const int64_t stride_px = (int64_t)M * D;
  const float* vb = value + (b * S * M + m) * (int64_t)D + sub * 4;
  const float2* lp = reinterpret_cast<const float2*>(loc_or_off) + pair * LP;
  const float* ap = attn_or_logit + pair * LP;

  // FUSED: softmax statistics over the L*P logits of this pair; lane `sub` owns logits sub, sub+G, ...
  float smax = 0.f, ssum = 1.f;
  const float2* rp = nullptr;
  // optional biases of the two Linear layers (per head m): raw + bias is rounded first, like Linear's acc + b
  const float* lbp = (FUSED && logit_bias) ? logit_bias + (int64_t)m * LP : nullptr;
  const float2* obp = (FUSED && off_bias) ? reinterpret_cast<const float2*>(off_bias) + (int64_t)m * LP : nullptr;
  auto logit = [&](int i) { return lbp ? __ldg(ap + i) + __ldg(lbp + i) : __ldg(ap + i); };
  if (FUSED) {
    float mx = -INFINITY;
    for (int i = sub; i < LP; i += G) mx = fmaxf(mx, logit(i));
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, o, G));
    float sum = 0.f;
    for (int i = sub; i < LP; i += G) sum += expf(logit(i) - mx);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o, G);
    smax = mx;
    ssum = sum;
    rp = reinterpret_cast<const float2*>(ref) + (int64_t)(q % Lr) * LP;
  }

  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int gl = 0, gp = 0;  // (level, point) of the sample the gather loop is at; advanced without divisions
  for (int s0 = 0; s0 < LP; s0 += G) {
    // ---- prep: this lane's sample ----
    const int s = s0 + sub;
    float w1 = 0.f, w2 = 0.f, w3 = 0.f, w4 = 0.f;
    unsigned code = 0u;
    if (s < LP) {
      const Level lv = s_lvl[s / P];
      const float fH = (float)lv.H, fW = (float)lv.W;
      float2 xy;
      float a;
      if (FUSED) {
        float2 off = ld_stream2(reinterpret_cast<const float*>(lp + s));
        if (obp) {
          const float2 ob = __ldg(obp + s);
          off.x += ob.x;
          off.y += ob.y;
        }
        const float2 r = __ldg(rp + s);
        xy.x = r.x + off.x / fW;
        xy.y = r.y + off.y / fH;
        a = expf(logit(s) - smax) / ssum;
        if (valid) {
          if (attn_out) attn_out[pair * LP + s] = a;
          if (loc_out) reinterpret_cast<float2*>(loc_out)[pair * LP + s] = xy;
        }
      } else {
        xy = ld_stream2(reinterpret_cast<const float*>(lp + s));
        a = ld_stream(ap + s);
      }
      // one fused multiply-add, as nvcc compiles the reference's `loc * spatial - 0.5` (FFMA R, R, R, -0.5 in its
      // sm_100a SASS, cuh:285-286): same cell selection as the reference build even for samples on a pixel boundary
      const float h_im = fmaf(xy.y, fH, -0.5f);
      const float w_im = fmaf(xy.x, fW, -0.5f);
      if (h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW) {
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h0 = (int)hf, w0 = (int)wf;
        const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
        const unsigned top = h0 >= 0, bot = h0 + 1 <= lv.H - 1, lef = w0 >= 0, rig = w0 + 1 <= lv.W - 1;
        const unsigned mask = (top & lef) | ((top & rig) << 1) | ((bot & lef) << 2) | ((bot & rig) << 3);
        w1 = hh * hw * a;
        w2 = hh * lw * a;
        w3 = lh * hw * a;
        w4 = lh * lw * a;
        const int px = lv.start + h0 * lv.W + w0;  // >= -W-1
        code = (unsigned)(px + kPxBias) | (mask << 28);
      }
    }
    // ---- gather: the group's samples one by one ----
    const int n = min(G, LP - s0);
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      const unsigned cj = __shfl_sync(FULL, code, j, G);
      const float a1 = __shfl_sync(FULL, w1, j, G);
      const float a2 = __shfl_sync(FULL, w2, j, G);
      const float a3 = __shfl_sync(FULL, w3, j, G);
      const float a4 = __shfl_sync(FULL, w4, j, G);
      if (cj >> 28) {
        const int px = (int)(cj & 0x0fffffffu) - kPxBias;
        const float* p00 = vb + (int64_t)px * stride_px;
        const int W = s_lvl[gl].W;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 v1 = (cj & (1u << 28)) ? ldg4(p00) : z;
        const float4 v2 = (cj & (2u << 28)) ? ldg4(p00 + stride_px) : z;
        const float4 v3 = (cj & (4u << 28)) ? ldg4(p00 + (int64_t)W * stride_px) : z;
        const float4 v4 = (cj & (8u << 28)) ? ldg4(p00 + (int64_t)(W + 1) * stride_px) : z;
        acc.x += a1 * v1.x + a2 * v2.x + a3 * v3.x + a4 * v4.x;
        acc.y += a1 * v1.y + a2 * v2.y + a3 * v3.y + a4 * v4.y;
        acc.z += a1 * v1.z + a2 * v2.z + a3 * v3.z + a4 * v4.z;
        acc.w += a1 * v1.w + a2 * v2.w + a3 * v3.w + a4 * v4.w;
      }
      if (++gp == P) {
        gp = 0;
        ++gl;
      }
    }
  }
  if (valid) *reinterpret_cast<float4*>(out + pair * D + sub * 4) = acc;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

template <typename T>
static int launch_scalar(const T* value, const int64_t* shapes, const int64_t* start, const T* loc, const T* attn,
                         int B, int S, int M, int D, int L, int Lq, int P, T* out, cudaStream_t st) {
  const int64_t n_out = (int64_t)B * Lq * M * D;
  const int64_t blocks64 = ceil_div64(n_out, 256);
  const int blocks = (int)(blocks64 > (int64_t)kNumSMs * 64 ? (int64_t)kNumSMs * 64 : blocks64);
  msda_fwd_scalar_kernel<T><<<blocks, 256, L * sizeof(Level), st>>>(value, shapes, start, loc, attn, S, M, D, L, Lq,
                                                                    P, n_out, out);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

template <int D, bool FUSED>
static int launch_vec4(const float* value, const int64_t* shapes, const int64_t* start, const float* loc,
                       const float* attn, const float* ref, const float* off_bias, const float* logit_bias, int B, int S, int M, int L, int Lq, int P, int Lr,
                       float* out, float* attn_out, float* loc_out, cudaStream_t st) {
  constexpr int PAIRS = 256 / (D / 4);
  const int64_t n_pairs = (int64_t)B * Lq * M;
  const int64_t blocks = ceil_div64(n_pairs, PAIRS);
  if (blocks > 0x7fffffffLL) return MVD_ERR_BAD_SHAPE;
  msda_fwd_vec4_kernel<D, FUSED><<<(int)blocks, 256, L * sizeof(Level), st>>>(
      value, shapes, start, loc, attn, ref, off_bias, logit_bias, S, M, L, Lq, P, Lr, n_pairs, out, attn_out, loc_out);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

template <bool FUSED>
static int dispatch_vec4(const float* value, const int64_t* shapes, const int64_t* start, const float* loc,
                         const float* attn, const float* ref, const float* off_bias, const float* logit_bias, int B, int S, int M, int D, int L, int Lq, int P,
                         int Lr, float* out, float* attn_out, float* loc_out, cudaStream_t st) {
#define MVD_CASE(DD)                                                                                              \
  case DD:                                                                                                        \
    return launch_vec4<DD, FUSED>(value, shapes, start, loc, attn, ref, off_bias, logit_bias, B, S, M, L, Lq, P,  \
                                  Lr, out, attn_out, loc_out, st)
  switch (D) {
    MVD_CASE(4);
    MVD_CASE(8);
    MVD_CASE(16);
    MVD_CASE(32);
    MVD_CASE(64);
    MVD_CASE(128);
    default:
      return MVD_ERR_UNSUPPORTED;
  }
#undef MVD_CASE
}

static bool vec4_ok(int D, int S) {
  if (2 * (int64_t)S + 2 > (int64_t)kPxBias) return false;  // packed pixel index of the vec4 kernels
  return D == 4 || D == 8 || D == 16 || D == 32 || D == 64 || D == 128;
}

static int check_dims(int B, int S, int M, int D, int L, int Lq, int P) {
  if (B <= 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0 || Lq <= 0 || P <= 0) return MVD_ERR_BAD_SHAPE;
  if (L > 4096) return MVD_ERR_BAD_SHAPE;  // level table lives in shared memory
  return MVD_OK;
}

}  // namespace mvd

using namespace mvd;

extern "C" int mvd_msda_fwd_f32(const float* value, const int64_t* shapes, const int64_t* start, const float* loc,
                                const float* attn, int B, int S, int M, int D, int L, int Lq, int P, float* out,
                                void* stream) {
  if (!value || !shapes || !start || !loc || !attn || !out) return MVD_ERR_NULL_POINTER;
  if (int e = check_dims(B, S, M, D, L, Lq, P)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec4_ok(D, S) && aligned16(value) && aligned16(out) && aligned8(loc))
    return dispatch_vec4<false>(value, shapes, start, loc, attn, nullptr, nullptr, nullptr, B, S, M, D, L, Lq, P, 1, out, nullptr,
                                nullptr, st);
  return launch_scalar<float>(value, shapes, start, loc, attn, B, S, M, D, L, Lq, P, out, st);
}

extern "C" int mvd_msda_fwd_f64(const double* value, const int64_t* shapes, const int64_t* start,
                                const double* loc, const double* attn, int B, int S, int M, int D, int L, int Lq,
                                int P, double* out, void* stream) {
  if (!value || !shapes || !start || !loc || !attn || !out) return MVD_ERR_NULL_POINTER;
  if (int e = check_dims(B, S, M, D, L, Lq, P)) return e;
  return launch_scalar<double>(value, shapes, start, loc, attn, B, S, M, D, L, Lq, P, out, (cudaStream_t)stream);
}

extern "C" int mvd_msda_fused_fwd_f32(const float* value, const int64_t* shapes, const int64_t* start,
                                      const float* offsets, const float* logits, const float* ref,
                                      const float* off_bias, const float* logit_bias, int B, int S, int M, int D,
                                      int L, int Lq, int P, int Lr, float* out, float* attn_out, float* loc_out,
                                      void* stream) {
  if (!value || !shapes || !start || !offsets || !logits || !ref || !out) return MVD_ERR_NULL_POINTER;
  if (int e = check_dims(B, S, M, D, L, Lq, P)) return e;
  if (Lr <= 0) return MVD_ERR_BAD_SHAPE;
  if (!vec4_ok(D, S)) return MVD_ERR_UNSUPPORTED;
  if (!aligned16(value) || !aligned16(out) || !aligned8(offsets) || !aligned8(ref) ||
      (loc_out && !aligned8(loc_out)) || (off_bias && !aligned8(off_bias)))
    return MVD_ERR_MISALIGNED;
  return dispatch_vec4<true>(value, shapes, start, offsets, logits, ref, off_bias, logit_bias, B, S, M, D, L, Lq, P, Lr, out, attn_out,
                             loc_out, (cudaStream_t)stream);
}
