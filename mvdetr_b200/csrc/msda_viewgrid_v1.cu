// Multi-scale deformable attention forward for the MVDeTr encoder layout ("view grid"): every level is one camera
// view of the same HxW ground-plane grid and the Lq = R*H*W queries are R copies of that grid
//   (ref: multiview_detector/models/trans_world_feat.py:92, multiview_detector/models/mvdetr.py:129-130).
// Same maths as the generic kernels of msda_fwd.cu (ref: ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299).
//
// Why a second kernel: the generic gather goes through L1 one 64-byte head-pixel (D=16) per 128-byte line, which
// ncu shows issue- and L1-bound at ~12 % of the algorithmic-HBM roofline (profiles/r01a_*). Here the R queries that sit
// on the same ground cell share their reference point, so a block that owns a TH x TW tile of cells x all R views x one
// head m touches, per level, only a (TH+2*HALO) x (TW+2*HALO) pixel window of value[:, level, :, m, :]:
//   * the window is staged in shared memory by ONE TMA tensor copy per level (5-D map d,m,x,y,level; box D x 1 x BW x
//     BH x 1). Out-of-map coordinates are zero-filled by the TMA unit, which IS the op's zero padding, so in-window
//     samples need no corner masks;
//   * windows are double buffered (level l+1 lands while level l is consumed), completion by mbarrier tx-count;
//   * one thread owns one (query, head) pair and all D channels: no cross-lane shuffles and no redundant sample
//     arithmetic. Each corner is D/4 128-bit shared loads; a per-lane rotation of the channel quads makes
//     neighbouring pixels hit disjoint bank groups, and the accumulators simply stay in that rotated order until
//     the final store;
//   * loc/attn (or offsets/logits when FUSED) are read exactly once, one 32-byte sector per (pair, level), and
//     prefetched into L2 one level ahead (no registers held across the level);
//   * sampling offsets are learned and unbounded: a sample whose 2x2 footprint is not inside the window takes the
//     masked global-memory path of the generic kernel, so the staging changes speed, never results.
#include <cuda.h>

#include "common.cuh"

namespace mvd {
namespace v1 {

constexpr int kHalo = 6;       // pixels around the tile kept in the window (default init samples +-4 px: ms_deform_attn.py:62-77)
constexpr int kMaxThreads = 448;  // 2 blocks/SM at <= 72 registers

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;  // idempotent lookup; a race only repeats it
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a TMA copy that never completes (bad descriptor, lost transaction) must surface as a launch failure,
// not as a hung GPU. 2^26 polls (each try_wait already sleeps up to a hardware time limit) is minutes, not microseconds.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, "
      "%6}], [%7];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

struct VgParams {
  const float* value;   // [B, L*H*W, M, D]  (global fallback path)
  const float* loc;     // loc [B,Lq,M,L,P,2]      or offsets when FUSED
  const float* attn;    // attn [B,Lq,M,L,P]       or logits when FUSED
  const float* ref;     // FUSED: [L, Lr, P, 2]  (level-major)
  float* out;           // [B, Lq, M*D]
  float* attn_out;      // FUSED, nullable
  float* loc_out;       // FUSED, nullable
  int H, W, M, L, R, Lr;
  int TH, TW, tiles_x;
  int BW, BH;           // window = tile + 2*halo
};

// One (pair, level) worth of sampling inputs held in registers.
template <int P>
struct LevelIn {
  float xy[2 * P];
  float a[P];
};

template <int P>
__device__ __forceinline__ void load_level(LevelIn<P>& r, const float* __restrict__ lp, const float* __restrict__ ap) {
  // loc: P*8 bytes = whole 32-byte sectors, one 256-bit request each (the L1 tag stage is the scarce resource here:
  // every lane touches its own line, so requests, not bytes, are what costs)
#pragma unroll
  for (int i = 0; i < P / 4; ++i) ld_stream8(lp + 8 * i, r.xy + 8 * i);
  if (P == 4) {
    const float4 v = ld_stream4(ap);
    r.a[0] = v.x;
    r.a[1] = v.y;
    r.a[2] = v.z;
    r.a[3] = v.w;
  } else {
#pragma unroll
    for (int i = 0; i < P / 8; ++i) ld_stream8(ap + 8 * i, r.a + 8 * i);
  }
}

// acc[k] += w1*v1 + w2*v2 + w3*v3 + w4*v4 for one channel quad
__device__ __forceinline__ void fma4(float4& acc, float w1, const float4& v1, float w2, const float4& v2, float w3,
                                     const float4& v3, float w4, const float4& v4) {
  acc.x = fmaf(w4, v4.x, fmaf(w3, v3.x, fmaf(w2, v2.x, fmaf(w1, v1.x, acc.x))));
  acc.y = fmaf(w4, v4.y, fmaf(w3, v3.y, fmaf(w2, v2.y, fmaf(w1, v1.y, acc.y))));
  acc.z = fmaf(w4, v4.z, fmaf(w3, v3.z, fmaf(w2, v2.z, fmaf(w1, v1.z, acc.z))));
  acc.w = fmaf(w4, v4.w, fmaf(w3, v3.w, fmaf(w2, v2.w, fmaf(w1, v1.w, acc.w))));
}

template <int D, int P, bool FUSED>
__global__ void __launch_bounds__(kMaxThreads, 2)
    msda_fwd_viewgrid_kernel(const __grid_constant__ CUtensorMap tmap, const VgParams prm) {
  constexpr int NQ = D / 4;             // channel quads (16-byte pieces) per head-pixel
  constexpr int PX_BYTES = D * 4;
  constexpr int PC = 4;                 // samples prepared together (P is a multiple of 4)
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long s_bar[2];

  const int H = prm.H, W = prm.W, M = prm.M, L = prm.L, R = prm.R;
  const int BW = prm.BW, BH = prm.BH;
  const int win_bytes = BW * BH * PX_BYTES;
  const uint32_t win0 = smem_u32(smem_raw);
  const uint32_t bar0 = smem_u32(&s_bar[0]);

  const int tile = blockIdx.x;
  const int ty0 = (tile / prm.tiles_x) * prm.TH, tx0 = (tile % prm.tiles_x) * prm.TW;
  const int m = blockIdx.y, b = blockIdx.z;
  const int wy0 = ty0 - kHalo, wx0 = tx0 - kHalo;  // window origin in level pixels (may be negative: TMA zero fill)

  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int nl = L < 2 ? L : 2;
    for (int l = 0; l < nl; ++l) {
      mbar_expect_tx(bar0 + 8 * l, (uint32_t)win_bytes);
      tma_load_5d(win0 + l * win_bytes, &tmap, bar0 + 8 * l, 0, m, wx0, wy0, b * L + l);
    }
  }

  // ---- this thread's (query, head) pair ----
  const int TP = prm.TH * prm.TW;
  const int r = threadIdx.x / TP, pos = threadIdx.x - r * TP;
  const int ty = pos / prm.TW, tx = pos - ty * prm.TW;
  const int y = ty0 + ty, x = tx0 + tx;
  const bool active = r < R && y < H && x < W;
  const int HW = H * W;
  const int q = r * HW + y * W + x;
  const int64_t Lq = (int64_t)R * HW;
  const int64_t pair = active ? ((int64_t)b * Lq + q) * M + m : 0;
  const int LP = L * P;
  const float* lp = prm.loc + pair * LP * 2;
  const float* ap = prm.attn + pair * LP;
  // FUSED: reference table is level-major [L, Lr, P, 2] so that neighbouring cells read neighbouring sectors
  const float* rp = FUSED ? prm.ref + (int64_t)(active ? q % prm.Lr : 0) * P * 2 : nullptr;
  const float fH = (float)H, fW = (float)W;

  // per-lane rotation of the channel quads (bank-conflict avoidance, see header): quad k of this thread's
  // accumulator holds channels 4*((k+rot)%NQ) .. +3
  const int lane = threadIdx.x & 31;
  const int rot = (NQ >= 8) ? (lane & (NQ - 1)) : (NQ == 4 ? ((lane >> 1) & 3) : ((lane >> 2) & (NQ - 1)));
  int qoff[NQ];  // byte offset of quad k inside a head-pixel
#pragma unroll
  for (int k = 0; k < NQ; ++k) qoff[k] = ((k + rot) & (NQ - 1)) * 16;

  float4 acc[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  // FUSED: online softmax over the L*P logits (ms_deform_attn.py:101-102): weights are exp(logit - running max),
  // accumulators are rescaled when the max moves, and the division by the sum happens once at the end.
  float run_max = -INFINITY, run_sum = 0.f;
  constexpr float kLog2e = 1.4426950408889634f;

  const int64_t stride_px = (int64_t)M * D;
  const float* vb = prm.value + ((int64_t)b * L * HW * M + m) * D;  // level 0, pixel 0 of this head
  const int rowb = BW * PX_BYTES;

  for (int l = 0; l < L; ++l) {
    LevelIn<P> cur, refl;
    if (active) {
      load_level<P>(cur, lp + l * P * 2, ap + l * P);
      if (FUSED) {
#pragma unroll
        for (int i = 0; i < P / 4; ++i) ldg8(rp + (int64_t)l * prm.Lr * P * 2 + 8 * i, refl.xy + 8 * i);
      }
      if (l + 1 < L) {  // next level's sectors: HBM -> L2 now, so that its loads are L2 hits
        prefetch_l2(lp + (l + 1) * P * 2);
        if (P > 4) prefetch_l2(lp + (l + 1) * P * 2 + 8);
        if (((l + 1) * P) % 8 == 0) prefetch_l2(ap + (l + 1) * P);
      }
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) cur.xy[2 * i] = cur.xy[2 * i + 1] = cur.a[i] = refl.xy[2 * i] = refl.xy[2 * i + 1] = 0.f;
    }
    if (FUSED) {
      float lm = cur.a[0];
#pragma unroll
      for (int i = 1; i < P; ++i) lm = fmaxf(lm, cur.a[i]);
      const float new_max = fmaxf(run_max, lm);
      const float sc = (run_max == -INFINITY) ? 0.f : exp2f((run_max - new_max) * kLog2e);
      run_max = new_max;
      run_sum *= sc;
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        acc[k].x *= sc;
        acc[k].y *= sc;
        acc[k].z *= sc;
        acc[k].w *= sc;
      }
#pragma unroll
      for (int i = 0; i < P; ++i) {
        cur.a[i] = (new_max == -INFINITY) ? 0.f : exp2f((cur.a[i] - new_max) * kLog2e);
        run_sum += cur.a[i];
        // same operation order as the reference module: ref + off / (W, H)
        cur.xy[2 * i] = refl.xy[2 * i] + cur.xy[2 * i] / fW;
        cur.xy[2 * i + 1] = refl.xy[2 * i + 1] + cur.xy[2 * i + 1] / fH;
      }
    }
    mbar_wait(bar0 + 8 * (l & 1), (uint32_t)((l >> 1) & 1));
    const unsigned char* win = smem_raw + (l & 1) * win_bytes;

#pragma unroll
    for (int pc = 0; pc < P; pc += PC) {
      // ---- sample arithmetic for PC samples (ref: ms_deform_im2col_cuda.cuh:285-288, :33-84) ----
      float w1[PC], w2[PC], w3[PC], w4[PC];
      int off[PC], h0[PC], w0[PC];
      unsigned valid = 0u, inwin = 0u;
#pragma unroll
      for (int i = 0; i < PC; ++i) {
        const float a = cur.a[pc + i];
        // product rounded before the subtraction, as the reference's float*int - 0.5 does
        const float h_im = __fsub_rn(__fmul_rn(cur.xy[2 * (pc + i) + 1], fH), 0.5f);
        const float w_im = __fsub_rn(__fmul_rn(cur.xy[2 * (pc + i)], fW), 0.5f);
        const bool v = h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW;  // false for NaN
        const float hf = floorf(h_im), wf = floorf(w_im);
        h0[i] = v ? (int)hf : 0;
        w0[i] = v ? (int)wf : 0;
        const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
        w1[i] = hh * hw * a;
        w2[i] = hh * lw * a;
        w3[i] = lh * hw * a;
        w4[i] = lh * lw * a;
        const int dx = w0[i] - wx0, dy = h0[i] - wy0;
        const bool in = v && (unsigned)dx < (unsigned)(BW - 1) && (unsigned)dy < (unsigned)(BH - 1);
        off[i] = in ? (dy * BW + dx) * PX_BYTES : 0;
        valid |= (unsigned)v << i;
        inwin |= (unsigned)in << i;
      }
      const bool all_in = !active || inwin == (1u << PC) - 1u;
      if (__all_sync(FULL, all_in)) {
        // ---- fast path: every sample of every lane has its 2x2 footprint in the window; no branches ----
#pragma unroll
        for (int i = 0; i < PC; ++i) {
#pragma unroll
          for (int k = 0; k < NQ; ++k) {
            const unsigned char* pk = win + off[i] + qoff[k];
            const float4 v1 = *reinterpret_cast<const float4*>(pk);
            const float4 v2 = *reinterpret_cast<const float4*>(pk + PX_BYTES);
            const float4 v3 = *reinterpret_cast<const float4*>(pk + rowb);
            const float4 v4 = *reinterpret_cast<const float4*>(pk + rowb + PX_BYTES);
            fma4(acc[k], w1[i], v1, w2[i], v2, w3[i], v3, w4[i], v4);
          }
        }
      } else if (active) {
        // ---- mixed path: per sample, window or masked global loads (generic-kernel semantics) ----
#pragma unroll
        for (int i = 0; i < PC; ++i) {
          if (!((valid >> i) & 1u)) continue;
          if ((inwin >> i) & 1u) {
#pragma unroll
            for (int k = 0; k < NQ; ++k) {
              const unsigned char* pk = win + off[i] + qoff[k];
              const float4 v1 = *reinterpret_cast<const float4*>(pk);
              const float4 v2 = *reinterpret_cast<const float4*>(pk + PX_BYTES);
              const float4 v3 = *reinterpret_cast<const float4*>(pk + rowb);
              const float4 v4 = *reinterpret_cast<const float4*>(pk + rowb + PX_BYTES);
              fma4(acc[k], w1[i], v1, w2[i], v2, w3[i], v3, w4[i], v4);
            }
          } else {
            const bool top = h0[i] >= 0, bot = h0[i] + 1 <= H - 1, lef = w0[i] >= 0, rig = w0[i] + 1 <= W - 1;
            const float* p00 = vb + ((int64_t)l * HW + (int64_t)h0[i] * W + w0[i]) * stride_px;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int k = 0; k < NQ; ++k) {
              const float* pk = p00 + (qoff[k] >> 2);
              const float4 v1 = (top && lef) ? __ldg(reinterpret_cast<const float4*>(pk)) : z;
              const float4 v2 = (top && rig) ? __ldg(reinterpret_cast<const float4*>(pk + stride_px)) : z;
              const float4 v3 = (bot && lef) ? __ldg(reinterpret_cast<const float4*>(pk + (int64_t)W * stride_px)) : z;
              const float4 v4 =
                  (bot && rig) ? __ldg(reinterpret_cast<const float4*>(pk + (int64_t)(W + 1) * stride_px)) : z;
              fma4(acc[k], w1[i], v1, w2[i], v2, w3[i], v3, w4[i], v4);
            }
          }
        }
      }
    }
    __syncthreads();  // every thread is done with this stage's window
    if (threadIdx.x == 0 && l + 2 < L) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of this stage -> async-proxy refill
      mbar_expect_tx(bar0 + 8 * (l & 1), (uint32_t)win_bytes);
      tma_load_5d(win0 + (l & 1) * win_bytes, &tmap, bar0 + 8 * (l & 1), 0, m, wx0, wy0, b * L + l + 2);
    }
  }

  if (active) {
    float* op = prm.out + pair * D;
    const float inv = FUSED ? 1.f / run_sum : 1.f;
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      float4 o = acc[k];
      if (FUSED) {
        o.x *= inv;
        o.y *= inv;
        o.z *= inv;
        o.w *= inv;
      }
      *reinterpret_cast<float4*>(op + (qoff[k] >> 2)) = o;
    }
  }
}

struct VgPlan {
  int TH, TW, BW, BH, tiles_x, tiles_y, threads;
  size_t smem;
};

// Tile = TH x TW ground cells with TH*TW*R <= kMaxThreads threads; prefers 64 cells (4x16), 32 (4x8) for many views.
static bool plan_viewgrid(int H, int W, int D, int R, VgPlan* pl) {
  int TH = 4, TW = 16;
  while (TH * TW * R > kMaxThreads && TW > 4) TW >>= 1;
  while (TH * TW * R > kMaxThreads && TH > 1) TH >>= 1;
  if (TH * TW * R > kMaxThreads) return false;
  pl->TH = TH;
  pl->TW = TW;
  pl->BW = TW + 2 * kHalo;
  pl->BH = TH + 2 * kHalo;
  pl->tiles_x = (W + TW - 1) / TW;
  pl->tiles_y = (H + TH - 1) / TH;
  pl->threads = ((TH * TW * R + 31) / 32) * 32;
  pl->smem = (size_t)2 * pl->BW * pl->BH * D * 4;
  return pl->BW <= 256 && pl->BH <= 256;
}

static int make_value_map(const float* value, int B, int H, int W, int M, int D, int L, const VgPlan& pl,
                          CUtensorMap* map) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return MVD_ERR_NO_DEVICE;
  const cuuint64_t gdim[5] = {(cuuint64_t)D, (cuuint64_t)M, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B * L};
  const cuuint64_t gstr[4] = {(cuuint64_t)D * 4, (cuuint64_t)M * D * 4, (cuuint64_t)W * M * D * 4,
                              (cuuint64_t)H * W * M * D * 4};
  const cuuint32_t box[5] = {(cuuint32_t)D, 1u, (cuuint32_t)pl.BW, (cuuint32_t)pl.BH, 1u};
  const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  const CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(value), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? MVD_OK : MVD_ERR_UNSUPPORTED;
}

template <int D, int P, bool FUSED>
static int launch_viewgrid(const CUtensorMap& map, const VgParams& prm, const VgPlan& pl, int B, cudaStream_t st) {
  auto kern = msda_fwd_viewgrid_kernel<D, P, FUSED>;
  MVD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  dim3 grid((unsigned)(pl.tiles_x * pl.tiles_y), (unsigned)prm.M, (unsigned)B);
  kern<<<grid, pl.threads, pl.smem, st>>>(map, prm);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

template <bool FUSED>
static int viewgrid_dispatch(const float* value, const float* loc, const float* attn, const float* ref, int B, int H,
                             int W, int M, int D, int L, int R, int P, int Lr, float* out, float* attn_out,
                             float* loc_out, cudaStream_t st) {
  if (B <= 0 || H <= 0 || W <= 0 || M <= 0 || D <= 0 || L <= 0 || R <= 0 || P <= 0) return MVD_ERR_BAD_SHAPE;
  if ((int64_t)B * L * H * W * M * D > 0x7fffffffLL || (int64_t)B * R * H * W * M * L * P * 2 > 0x7fffffffffLL)
    return MVD_ERR_BAD_SHAPE;
  if (M > 65535 || B > 65535) return MVD_ERR_BAD_SHAPE;
  if (!((D == 8 || D == 16 || D == 32) && (P == 4 || P == 8))) return MVD_ERR_UNSUPPORTED;
  const uintptr_t al = reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(loc) |
                       reinterpret_cast<uintptr_t>(attn) | reinterpret_cast<uintptr_t>(out) |
                       reinterpret_cast<uintptr_t>(ref) | reinterpret_cast<uintptr_t>(loc_out) |
                       reinterpret_cast<uintptr_t>(attn_out);
  if ((al & 15u) || ((reinterpret_cast<uintptr_t>(loc) | reinterpret_cast<uintptr_t>(ref)) & 31u) ||
      (P == 8 && (reinterpret_cast<uintptr_t>(attn) & 31u)))
    return MVD_ERR_MISALIGNED;
  VgPlan pl;
  if (!plan_viewgrid(H, W, D, R, &pl)) return MVD_ERR_UNSUPPORTED;
  alignas(64) CUtensorMap map;
  if (int e = make_value_map(value, B, H, W, M, D, L, pl, &map)) return e;
  VgParams prm;
  prm.value = value;
  prm.loc = loc;
  prm.attn = attn;
  prm.ref = ref;
  prm.out = out;
  prm.attn_out = attn_out;
  prm.loc_out = loc_out;
  prm.H = H;
  prm.W = W;
  prm.M = M;
  prm.L = L;
  prm.R = R;
  prm.Lr = Lr > 0 ? Lr : 1;
  prm.TH = pl.TH;
  prm.TW = pl.TW;
  prm.tiles_x = pl.tiles_x;
  prm.BW = pl.BW;
  prm.BH = pl.BH;
#define MVD_VG(DD, PP) \
  if (D == DD && P == PP) return launch_viewgrid<DD, PP, FUSED>(map, prm, pl, B, st)
  MVD_VG(8, 4);
  MVD_VG(16, 4);
  MVD_VG(32, 4);
  MVD_VG(8, 8);
  MVD_VG(16, 8);
  MVD_VG(32, 8);
#undef MVD_VG
  return MVD_ERR_UNSUPPORTED;
}

}  // namespace v1
}  // namespace mvd

namespace mvd {
using namespace v1;
// Legacy entry (A/B comparisons only; selected with MVD_VIEWGRID_IMPL=1): same contract as the public functions.
int viewgrid_v1(bool fused, const float* value, const float* loc, const float* attn, const float* ref, int B, int H,
                int W, int M, int D, int L, int R, int P, int Lr, float* out, cudaStream_t st) {
  return fused ? viewgrid_dispatch<true>(value, loc, attn, ref, B, H, W, M, D, L, R, P, Lr, out, nullptr, nullptr, st)
               : viewgrid_dispatch<false>(value, loc, attn, ref, B, H, W, M, D, L, R, P, Lr, out, nullptr, nullptr, st);
}
}  // namespace mvd
