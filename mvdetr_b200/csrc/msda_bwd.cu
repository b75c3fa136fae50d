// Multi-scale deformable attention, backward -- generic kernels.
//
// Maths follows ms_deform_attn_col2im_bilinear
//   (ref: multiview_detector/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:87-159):
//     grad_value  += w_i * grad_out * attn                      (scatter, atomics)
//     grad_attn    = sum_c grad_out_c * bilinear_c
//     grad_loc.x   = W * sum_c (d bilinear_c / d w) * grad_out_c * attn
//     grad_loc.y   = H * sum_c (d bilinear_c / d h) * grad_out_c * attn
// Design differs from the reference's 7 kernel variants (block = D threads, shared-memory reduction with
// two __syncthreads per sample: cuh:301-920; MVDeTr's D=16 gives 16-thread blocks):
//   * vec4 path (fp32, D = 4..128 power of two): D/4 lanes per (b,q,m) pair, 8 pairs per warp at D=16,
//     channel reduction by lane-group shuffles (no shared memory, no barriers), ONE 128-bit vector
//     reduction (red.global.add.v4.f32) per corner per lane instead of four scalar atomics;
//   * scalar path (fp32/fp64, any D): one warp per pair, lanes stride the channels, warp-shuffle reduce.
// grad_loc / grad_attn are written exactly once per sample (zeros for out-of-range samples), so only
// grad_value needs zeroing; the C entry points do that with cudaMemsetAsync.
#include "common.cuh"

namespace mvd {

__device__ __forceinline__ void atomic_add_t(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_t(double* p, double v) { atomicAdd(p, v); }

template <typename T>
__global__ void __launch_bounds__(256) msda_bwd_scalar_kernel(
    const T* __restrict__ grad_out, const T* __restrict__ value, const int64_t* __restrict__ shapes,
    const int64_t* __restrict__ start, const T* __restrict__ loc, const T* __restrict__ attn, int S, int M, int D,
    int L, int Lq, int P, int64_t n_pairs, T* __restrict__ grad_value, T* __restrict__ grad_loc,
    T* __restrict__ grad_attn) {
  extern __shared__ Level s_lvl[];
  load_levels(s_lvl, shapes, start, L);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t stride_px = (int64_t)M * D;
  for (int64_t pair = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); pair < n_pairs;
       pair += (int64_t)gridDim.x * warps_per_block) {
    const int m = (int)(pair % M);
    const int64_t b = pair / ((int64_t)M * Lq);
    const int64_t voff = (b * S * M + m) * (int64_t)D;
    const T* go = grad_out + pair * D;
    const T* lp = loc + pair * L * P * 2;
    const T* ap = attn + pair * L * P;
    T* glp = grad_loc + pair * L * P * 2;
    T* gap = grad_attn + pair * L * P;
    for (int l = 0; l < L; ++l) {
      const Level lv = s_lvl[l];
      const int64_t loff = voff + (int64_t)lv.start * stride_px;
      for (int p = 0; p < P; ++p) {
        const T x = lp[0], y = lp[1], a = ap[0];
        lp += 2;
        ap += 1;
        const T h_im = y * lv.H - (T)0.5;
        const T w_im = x * lv.W - (T)0.5;
        T gw = 0, gh = 0, ga = 0;
        if (h_im > (T)-1 && w_im > (T)-1 && h_im < (T)lv.H && w_im < (T)lv.W) {
          const int h0 = (int)floor(h_im), w0 = (int)floor(w_im);
          const T lh = h_im - h0, lw = w_im - w0, hh = 1 - lh, hw = 1 - lw;
          const bool top = h0 >= 0, bot = h0 + 1 <= lv.H - 1, lef = w0 >= 0, rig = w0 + 1 <= lv.W - 1;
          const int64_t o00 = loff + ((int64_t)h0 * lv.W + w0) * stride_px;
          const int64_t o01 = o00 + stride_px;
          const int64_t o10 = o00 + (int64_t)lv.W * stride_px;
          const int64_t o11 = o10 + stride_px;
          const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
          for (int c = lane; c < D; c += 32) {
            const T g = go[c];
            const T tg = g * a;
            T v1 = 0, v2 = 0, v3 = 0, v4 = 0;
            if (top && lef) {
              v1 = value[o00 + c];
              atomic_add_t(grad_value + o00 + c, w1 * tg);
            }
            if (top && rig) {
              v2 = value[o01 + c];
              atomic_add_t(grad_value + o01 + c, w2 * tg);
            }
            if (bot && lef) {
              v3 = value[o10 + c];
              atomic_add_t(grad_value + o10 + c, w3 * tg);
            }
            if (bot && rig) {
              v4 = value[o11 + c];
              atomic_add_t(grad_value + o11 + c, w4 * tg);
            }
            ga += g * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
            gw += (hh * (v2 - v1) + lh * (v4 - v3)) * tg;
            gh += (hw * (v3 - v1) + lw * (v4 - v2)) * tg;
          }
          gw *= lv.W;
          gh *= lv.H;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          gw += __shfl_xor_sync(0xffffffffu, gw, o);
          gh += __shfl_xor_sync(0xffffffffu, gh, o);
          ga += __shfl_xor_sync(0xffffffffu, ga, o);
        }
        if (lane == 0) {
          glp[0] = gw;
          glp[1] = gh;
          gap[0] = ga;
        }
        glp += 2;
        gap += 1;
      }
    }
  }
}

// vec4 path. Per chunk of C = min(G, 8) samples of one (b,q,m) pair (G = D/4 lanes):
//   prep    : lane j < C evaluates sample s0+j once (loc, attn -> fractional weights, packed pixel index + mask);
//   scatter : the group walks the C samples (values arrive by width-G shuffles); each lane loads its 4 channels of the
//             4 corners, issues one red.global.add.v4.f32 per valid corner and keeps 3 partial sums per sample;
//   reduce  : butterfly over lane strides >= C, then a transpose-reduction over the low bits, so that lane j ends up
//             with the three totals of sample s0+j and writes them (coalesced 8 B / 4 B stores, no atomics).
constexpr int kPxBiasB = 1 << 27;

template <int D>
__global__ void __launch_bounds__(256) msda_bwd_vec4_kernel(
    const float* __restrict__ grad_out, const float* __restrict__ value, const int64_t* __restrict__ shapes,
    const int64_t* __restrict__ start, const float* __restrict__ loc, const float* __restrict__ attn, int S, int M,
    int L, int Lq, int P, int64_t n_pairs, float* __restrict__ grad_value, float* __restrict__ grad_loc,
    float* __restrict__ grad_attn, int vgH, int vgW, int vgR) {
  constexpr int G = D / 4;
  constexpr int C = G < 8 ? G : 8;
  constexpr int PAIRS = 256 / G;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ Level s_lvl[];
  load_levels(s_lvl, shapes, start, L);
  __syncthreads();

  const int sub = threadIdx.x % G;
  const int64_t pair_raw = (int64_t)blockIdx.x * PAIRS + threadIdx.x / G;
  const bool valid = pair_raw < n_pairs;
  int64_t pair = valid ? pair_raw : n_pairs - 1;
  if (vgW > 0) {
    // View-grid layout (queries = vgR copies of a vgH x vgW grid): walk the pairs band by band -- (batch, row y, view r,
    // x, head) -- instead of view by view. The reds of concurrently running blocks then hit the same few rows of every
    // level and stay in L2; in query order every view's pass re-fetched all of value and grad_value from DRAM
    // (r01m ncu: 1.54 GB of DRAM traffic for 0.56 GB of algorithmic bytes).
    int64_t t = pair;
    const int mm = (int)(t % M);
    t /= M;
    const int x = (int)(t % vgW);
    t /= vgW;
    const int r = (int)(t % vgR);
    t /= vgR;
    const int y = (int)(t % vgH);
    const int64_t bb = t / vgH;
    pair = (bb * Lq + ((int64_t)r * vgH + y) * vgW + x) * M + mm;
  }
  const int m = (int)(pair % M);
  const int64_t b = pair / ((int64_t)M * Lq);
  const int LP = L * P;
  const int64_t stride_px = (int64_t)M * D;
  const int64_t voff = (b * S * M + m) * (int64_t)D + sub * 4;
  const float2* lp = reinterpret_cast<const float2*>(loc) + pair * LP;
  const float* ap = attn + pair * LP;
  float2* glp = reinterpret_cast<float2*>(grad_loc) + pair * LP;
  float* gap = grad_attn + pair * LP;
  const float4 g = __ldg(reinterpret_cast<const float4*>(grad_out + pair * D + sub * 4));

  int gl = 0, gp = 0;
  for (int s0 = 0; s0 < LP; s0 += C) {
    // ---- prep ----
    const int s = s0 + sub;
    float lh = 0.f, lw = 0.f, a = 0.f;
    unsigned code = 0u;
    if (sub < C && s < LP) {
      const Level lv = s_lvl[s / P];
      const float fH = (float)lv.H, fW = (float)lv.W;
      const float2 xy = ld_stream2(reinterpret_cast<const float*>(lp + s));
      a = ld_stream(ap + s);
      const float h_im = fmaf(xy.y, fH, -0.5f);
      const float w_im = fmaf(xy.x, fW, -0.5f);
      if (h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW) {
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h0 = (int)hf, w0 = (int)wf;
        lh = h_im - hf;
        lw = w_im - wf;
        const unsigned top = h0 >= 0, bot = h0 + 1 <= lv.H - 1, lef = w0 >= 0, rig = w0 + 1 <= lv.W - 1;
        const unsigned mask = (top & lef) | ((top & rig) << 1) | ((bot & lef) << 2) | ((bot & rig) << 3);
        code = (unsigned)(lv.start + h0 * lv.W + w0 + kPxBiasB) | (mask << 28);
      }
    }
    // ---- scatter + partial sums ----
    float pw[C], ph[C], pa[C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
      pw[j] = ph[j] = pa[j] = 0.f;
      const unsigned cj = __shfl_sync(FULL, code, j, G);
      const float jlh = __shfl_sync(FULL, lh, j, G);
      const float jlw = __shfl_sync(FULL, lw, j, G);
      const float ja = __shfl_sync(FULL, a, j, G);
      if (cj >> 28) {
        const Level lv = s_lvl[gl];
        const int64_t o00 = voff + (int64_t)((int)(cj & 0x0fffffffu) - kPxBiasB) * stride_px;
        const int64_t o01 = o00 + stride_px;
        const int64_t o10 = o00 + (int64_t)lv.W * stride_px;
        const int64_t o11 = o10 + stride_px;
        const bool c1 = cj & (1u << 28), c2 = cj & (2u << 28), c3 = cj & (4u << 28), c4 = cj & (8u << 28);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 v1 = c1 ? __ldg(reinterpret_cast<const float4*>(value + o00)) : z;
        const float4 v2 = c2 ? __ldg(reinterpret_cast<const float4*>(value + o01)) : z;
        const float4 v3 = c3 ? __ldg(reinterpret_cast<const float4*>(value + o10)) : z;
        const float4 v4 = c4 ? __ldg(reinterpret_cast<const float4*>(value + o11)) : z;
        const float hh = 1.f - jlh, hw = 1.f - jlw;
        const float w1 = hh * hw, w2 = hh * jlw, w3 = jlh * hw, w4 = jlh * jlw;
        const float4 tg = make_float4(g.x * ja, g.y * ja, g.z * ja, g.w * ja);
        if (valid) {
          if (c1) red_add4(grad_value + o00, w1 * tg.x, w1 * tg.y, w1 * tg.z, w1 * tg.w);
          if (c2) red_add4(grad_value + o01, w2 * tg.x, w2 * tg.y, w2 * tg.z, w2 * tg.w);
          if (c3) red_add4(grad_value + o10, w3 * tg.x, w3 * tg.y, w3 * tg.z, w3 * tg.w);
          if (c4) red_add4(grad_value + o11, w4 * tg.x, w4 * tg.y, w4 * tg.z, w4 * tg.w);
        }
        float sa = 0.f, sw = 0.f, sh = 0.f;
#define MVD_CH(X)                                                      \
  sa += g.X * (w1 * v1.X + w2 * v2.X + w3 * v3.X + w4 * v4.X);          \
  sw += (hh * (v2.X - v1.X) + jlh * (v4.X - v3.X)) * tg.X;              \
  sh += (hw * (v3.X - v1.X) + jlw * (v4.X - v2.X)) * tg.X;
        MVD_CH(x) MVD_CH(y) MVD_CH(z) MVD_CH(w)
#undef MVD_CH
        pa[j] = sa;
        pw[j] = sw * (float)lv.W;
        ph[j] = sh * (float)lv.H;
      }
      if (++gp == P) {
        gp = 0;
        ++gl;
      }
    }
    // ---- reduce: lanes with equal (sub mod C) first, then transpose over the low bits ----
#pragma unroll
    for (int o = G / 2; o >= C; o >>= 1) {
#pragma unroll
      for (int j = 0; j < C; ++j) {
        pw[j] += __shfl_xor_sync(FULL, pw[j], o, G);
        ph[j] += __shfl_xor_sync(FULL, ph[j], o, G);
        pa[j] += __shfl_xor_sync(FULL, pa[j], o, G);
      }
    }
#pragma unroll
    for (int o = C / 2; o > 0; o >>= 1) {
      const bool upper = sub & o;
#pragma unroll
      for (int i = 0; i < o; ++i) {
        const float sw = upper ? pw[i] : pw[i + o], kw = upper ? pw[i + o] : pw[i];
        const float sh = upper ? ph[i] : ph[i + o], kh = upper ? ph[i + o] : ph[i];
        const float sa = upper ? pa[i] : pa[i + o], ka = upper ? pa[i + o] : pa[i];
        pw[i] = kw + __shfl_xor_sync(FULL, sw, o, G);
        ph[i] = kh + __shfl_xor_sync(FULL, sh, o, G);
        pa[i] = ka + __shfl_xor_sync(FULL, sa, o, G);
      }
    }
    if (sub < C && s < LP && valid) {
      glp[s] = make_float2(pw[0], ph[0]);
      gap[s] = pa[0];
    }
  }
}

template <typename T>
static int launch_bwd_scalar(const T* grad_out, const T* value, const int64_t* shapes, const int64_t* start,
                             const T* loc, const T* attn, int B, int S, int M, int D, int L, int Lq, int P,
                             T* grad_value, T* grad_loc, T* grad_attn, cudaStream_t st) {
  const int64_t n_pairs = (int64_t)B * Lq * M;
  const int64_t want = ceil_div64(n_pairs, 8);
  const int blocks = (int)(want > (int64_t)kNumSMs * 64 ? (int64_t)kNumSMs * 64 : want);
  msda_bwd_scalar_kernel<T><<<blocks, 256, L * sizeof(Level), st>>>(grad_out, value, shapes, start, loc, attn, S, M,
                                                                    D, L, Lq, P, n_pairs, grad_value, grad_loc,
                                                                    grad_attn);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

template <int D>
static int launch_bwd_vec4(const float* grad_out, const float* value, const int64_t* shapes, const int64_t* start,
                           const float* loc, const float* attn, int B, int S, int M, int L, int Lq, int P,
                           float* grad_value, float* grad_loc, float* grad_attn, cudaStream_t st, int vgH = 0, int vgW = 0,
                           int vgR = 0) {
  constexpr int PAIRS = 256 / (D / 4);
  const int64_t n_pairs = (int64_t)B * Lq * M;
  const int64_t blocks = ceil_div64(n_pairs, PAIRS);
  if (blocks > 0x7fffffffLL) return MVD_ERR_BAD_SHAPE;
  msda_bwd_vec4_kernel<D><<<(int)blocks, 256, L * sizeof(Level), st>>>(grad_out, value, shapes, start, loc, attn, S,
                                                                       M, L, Lq, P, n_pairs, grad_value, grad_loc,
                                                                       grad_attn, vgH, vgW, vgR);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static bool al8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

}  // namespace mvd

using namespace mvd;

static int msda_bwd_f32(const float* grad_out, const float* value, const int64_t* shapes, const int64_t* start,
                        const float* loc, const float* attn, int B, int S, int M, int D, int L, int Lq, int P,
                        float* grad_value, float* grad_loc, float* grad_attn, void* stream, int vgH, int vgW, int vgR) {
  if (!grad_out || !value || !shapes || !start || !loc || !attn || !grad_value || !grad_loc || !grad_attn)
    return MVD_ERR_NULL_POINTER;
  if (B <= 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0 || Lq <= 0 || P <= 0 || L > 4096) return MVD_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  MVD_CUDA_TRY(cudaMemsetAsync(grad_value, 0, sizeof(float) * (size_t)B * S * M * D, st));
  const bool fast = al16(grad_out) && al16(value) && al16(grad_value) && al8(loc) && al8(grad_loc) &&
                    2 * (int64_t)S + 2 <= (int64_t)kPxBiasB;
  if (fast) {
#define MVD_CASE(DD)                                                                                             \
  case DD:                                                                                                       \
    return launch_bwd_vec4<DD>(grad_out, value, shapes, start, loc, attn, B, S, M, L, Lq, P, grad_value, grad_loc, \
                               grad_attn, st, vgH, vgW, vgR)
    switch (D) {
      MVD_CASE(4);
      MVD_CASE(8);
      MVD_CASE(16);
      MVD_CASE(32);
      MVD_CASE(64);
      MVD_CASE(128);
      default:
        break;
    }
#undef MVD_CASE
  }
  return launch_bwd_scalar<float>(grad_out, value, shapes, start, loc, attn, B, S, M, D, L, Lq, P, grad_value,
                                  grad_loc, grad_attn, st);
}

extern "C" int mvd_msda_bwd_f32(const float* grad_out, const float* value, const int64_t* shapes,
                                const int64_t* start, const float* loc, const float* attn, int B, int S, int M,
                                int D, int L, int Lq, int P, float* grad_value, float* grad_loc, float* grad_attn,
                                void* stream) {
  return msda_bwd_f32(grad_out, value, shapes, start, loc, attn, B, S, M, D, L, Lq, P, grad_value, grad_loc, grad_attn,
                      stream, 0, 0, 0);
}

// Same kernels and results (up to the order of the fp32 reductions, which atomics leave unspecified anyway) for the
// MVDeTr encoder layout: L levels of one H x W grid, Lq = R * H * W queries. The pairs are walked band by band so that
// the gradient reductions stay in L2 (see msda_bwd_vec4_kernel).
extern "C" int mvd_msda_bwd_banded_f32(const float* grad_out, const float* value, const int64_t* shapes,
                                       const int64_t* start, const float* loc, const float* attn, int B, int S, int M,
                                       int D, int L, int Lq, int P, int H, int W, int R, float* grad_value,
                                       float* grad_loc, float* grad_attn, void* stream) {
  if (H <= 0 || W <= 0 || R <= 0 || (int64_t)R * H * W != Lq) return MVD_ERR_BAD_SHAPE;
  return msda_bwd_f32(grad_out, value, shapes, start, loc, attn, B, S, M, D, L, Lq, P, grad_value, grad_loc, grad_attn,
                      stream, H, W, R);
}

extern "C" int mvd_msda_bwd_f64(const double* grad_out, const double* value, const int64_t* shapes,
                                const int64_t* start, const double* loc, const double* attn, int B, int S, int M,
                                int D, int L, int Lq, int P, double* grad_value, double* grad_loc,
                                double* grad_attn, void* stream) {
  if (!grad_out || !value || !shapes || !start || !loc || !attn || !grad_value || !grad_loc || !grad_attn)
    return MVD_ERR_NULL_POINTER;
  if (B <= 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0 || Lq <= 0 || P <= 0 || L > 4096) return MVD_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  MVD_CUDA_TRY(cudaMemsetAsync(grad_value, 0, sizeof(double) * (size_t)B * S * M * D, st));
  return launch_bwd_scalar<double>(grad_out, value, shapes, start, loc, attn, B, S, M, D, L, Lq, P, grad_value,
                                   grad_loc, grad_attn, st);
}
