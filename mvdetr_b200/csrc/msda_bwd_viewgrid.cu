// Multi-scale deformable attention BACKWARD for the MVDeTr encoder layout ("view grid") -- EXPERIMENTAL in round 1:
// compiled and exported (mvd_msda_bwd_viewgrid_f32) but only dispatched when MVDETR_B200_BWD_VIEWGRID=1; the default
// backward is the generic kernel of msda_bwd.cu. Not yet measured on a B200.
//
// Maths: ms_deform_attn_col2im_bilinear (ref: ops/src/cuda/ms_deform_im2col_cuda.cuh:87-159), same as msda_bwd.cu.
// Structure: the forward view-grid kernel's (msda_viewgrid.cu). A block owns TH x TW ground cells x R views x one head;
// per level the value window and the tile's loc / attn records arrive through TMA (two stages, mbarrier full/empty,
// no block barrier). One thread owns one (query, head) pair and all D channels, so
//   * the corner values come from shared memory instead of L2 (the generic kernel's 2.9 GB of L2 gather traffic at
//     Wildtrack size, profiles/r01m_ncu_full_summary.txt: lts 71 %, the kernel's bound, shared with its reductions);
//   * grad_attn / grad_loc need no cross-lane reduction at all: each thread sums its D channels and writes the
//     (pair, level) record once (32 B + 16 B vector stores);
//   * grad_value is still scattered with red.global.add.v4.f32 (one per corner and channel quad). A shared-memory
//     accumulator is not an option: fp32 atomicAdd on shared memory is an ATOMS.CAST.SPIN loop on sm_100a.
// Samples whose footprint leaves the window read their corners from global memory with the generic masks.
#include "vg_common.cuh"

namespace mvd {
namespace {

struct VgBwdParams {
  const float* value;     // [B, L*H*W, M, D]
  const float* grad_out;  // [B, Lq, M*D]
  float* grad_value;      // [B, L*H*W, M, D], zeroed by the entry point
  float* grad_loc;        // [B, Lq, M, L, P, 2]
  float* grad_attn;       // [B, Lq, M, L, P]
  int H, W, M, L, R;
  int TH, TW, tiles_x;
  int BW, BH;
  uint32_t off_a, off_b, stage_bytes, zero_off, zero_bytes, tx_bytes;
};

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

template <int D, int P>
__global__ void __launch_bounds__(kMaxThreads, (D >= 32 || P >= 8) ? 1 : 2)
    msda_vg_bwd_kernel(const __grid_constant__ CUtensorMap tm_val, const __grid_constant__ CUtensorMap tm_a,
                       const __grid_constant__ CUtensorMap tm_b, const VgBwdParams prm) {
  constexpr int NQ = D / 4;
  constexpr int PX_BYTES = D * 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long s_bar[4];  // full[0], full[1], empty[0], empty[1]

  const int H = prm.H, W = prm.W, M = prm.M, L = prm.L, R = prm.R;
  const int BW = prm.BW, BH = prm.BH;
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t bar_full = smem_u32(&s_bar[0]), bar_empty = smem_u32(&s_bar[2]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

  const int tile = blockIdx.x;
  const int ty0 = (tile / prm.tiles_x) * prm.TH, tx0 = (tile % prm.tiles_x) * prm.TW;
  const int m = blockIdx.y, b = blockIdx.z;
  const int wy0 = ty0 - kHalo, wx0 = tx0 - kHalo;

  auto issue_level = [&](int lv) {
    const uint32_t st = smem0 + (uint32_t)(lv & 1) * prm.stage_bytes, bar = bar_full + 8u * (uint32_t)(lv & 1);
    mbar_expect_tx(bar, prm.tx_bytes);
    tma_load_5d(st, &tm_val, bar, 0, m, wx0, wy0, b * L + lv);
    tma_load_5d(st + prm.off_a, &tm_a, bar, lv * 2 * P, m, tx0, ty0, b * R);
    tma_load_5d(st + prm.off_b, &tm_b, bar, lv * P, m, tx0, ty0, b * R);
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_val);
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_b);
    mbar_init(bar_full, 1);
    mbar_init(bar_full + 8, 1);
    mbar_init(bar_empty, (uint32_t)nwarps);
    mbar_init(bar_empty + 8, (uint32_t)nwarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    issue_level(0);
    if (L > 1) issue_level(1);
  }

  const int TP = prm.TH * prm.TW;
  const int r = threadIdx.x / TP, pos = threadIdx.x - r * TP;
  const int ty = pos / prm.TW, tx = pos - ty * prm.TW;
  const int y = ty0 + ty, x = tx0 + tx;
  const bool active = r < R && y < H && x < W;
  const int HW = H * W;
  const int q = r * HW + y * W + x;
  const int64_t Lq = (int64_t)R * HW;
  const int64_t pair = active ? ((int64_t)b * Lq + q) * M + m : 0;
  const float fH = (float)H, fW = (float)W;

  const int rot = (NQ >= 8) ? (lane & (NQ - 1)) : (NQ == 4 ? ((lane >> 1) & 3) : ((lane >> 2) & (NQ - 1)));
  int qoff[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) qoff[k] = ((k + rot) & (NQ - 1)) * 16;

  // this pair's output gradient, in the rotated quad order
  float4 g[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k)
    g[k] = active ? __ldg(reinterpret_cast<const float4*>(prm.grad_out + pair * D + (qoff[k] >> 2)))
                  : make_float4(0.f, 0.f, 0.f, 0.f);

  const int64_t stride_px = (int64_t)M * D;
  const int64_t head0 = ((int64_t)b * L * HW * M + m) * D;  // level 0, pixel 0 of this head
  const float* vb = prm.value + head0;
  float* gvb = prm.grad_value + head0;
  const int rowb = BW * PX_BYTES;

  for (int l = 0; l < L; ++l) {
    const int s = l & 1;
    const uint32_t ph = (uint32_t)((l >> 1) & 1);
    const unsigned char* win = smem_raw + (size_t)s * prm.stage_bytes;
    mbar_wait(bar_full + 8u * s, ph);

    float xy[2 * P], aw[P], gl[2 * P], ga[P];
    if (active) {
      read_record<2 * P>(win + prm.off_a, threadIdx.x, lane, xy);
      read_record<P>(win + prm.off_b, threadIdx.x, lane, aw);
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) xy[2 * i] = xy[2 * i + 1] = aw[i] = 0.f;
    }

#pragma unroll
    for (int i = 0; i < P; ++i) {
      const float a = aw[i];
      const float h_im = fmaf(xy[2 * i + 1], fH, -0.5f);  // one FFMA, as the reference build (cuh:285-286)
      const float w_im = fmaf(xy[2 * i], fW, -0.5f);
      const bool v = active && h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW;
      float sa = 0.f, sw = 0.f, sh = 0.f;
      if (v) {
        const float hf = floorf(h_im), wf = floorf(w_im);
        const int h0 = (int)hf, w0 = (int)wf;
        const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        const bool top = h0 >= 0, bot = h0 + 1 <= H - 1, lef = w0 >= 0, rig = w0 + 1 <= W - 1;
        const bool c1 = top && lef, c2 = top && rig, c3 = bot && lef, c4 = bot && rig;
        const int dx = w0 - wx0, dy = h0 - wy0;
        const bool in = (unsigned)dx < (unsigned)(BW - 1) && (unsigned)dy < (unsigned)(BH - 1);
        const int64_t px = ((int64_t)l * HW + (int64_t)h0 * W + w0) * stride_px;  // top-left pixel (may be outside: masks)
        const unsigned char* wp = win + (in ? (dy * BW + dx) * PX_BYTES : 0);
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
          const int qo = qoff[k] >> 2;
          float4 v1, v2, v3, v4;
          if (in) {  // window: out-of-level corners were zero-filled by the TMA unit
            v1 = *reinterpret_cast<const float4*>(wp + qoff[k]);
            v2 = *reinterpret_cast<const float4*>(wp + qoff[k] + PX_BYTES);
            v3 = *reinterpret_cast<const float4*>(wp + qoff[k] + rowb);
            v4 = *reinterpret_cast<const float4*>(wp + qoff[k] + rowb + PX_BYTES);
          } else {   // footprint outside the staged window: masked global loads
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* pk = vb + px + qo;
            v1 = c1 ? __ldg(reinterpret_cast<const float4*>(pk)) : z;
            v2 = c2 ? __ldg(reinterpret_cast<const float4*>(pk + stride_px)) : z;
            v3 = c3 ? __ldg(reinterpret_cast<const float4*>(pk + (int64_t)W * stride_px)) : z;
            v4 = c4 ? __ldg(reinterpret_cast<const float4*>(pk + (int64_t)(W + 1) * stride_px)) : z;
          }
          const float4 gk = g[k];
          const float4 tg = make_float4(gk.x * a, gk.y * a, gk.z * a, gk.w * a);
          // grad_value += w_c * grad_out * attn (cuh:125-152), one vector reduction per valid corner
          float* gp = gvb + px + qo;
          if (c1) red_add4(gp, w1 * tg.x, w1 * tg.y, w1 * tg.z, w1 * tg.w);
          if (c2) red_add4(gp + stride_px, w2 * tg.x, w2 * tg.y, w2 * tg.z, w2 * tg.w);
          if (c3) red_add4(gp + (int64_t)W * stride_px, w3 * tg.x, w3 * tg.y, w3 * tg.z, w3 * tg.w);
          if (c4) red_add4(gp + (int64_t)(W + 1) * stride_px, w4 * tg.x, w4 * tg.y, w4 * tg.z, w4 * tg.w);
          // grad_attn = sum_c g * bilinear; grad_loc = sum_c d(bilinear)/d(w,h) * g * attn (cuh:156-158)
          float4 bil, dw, dh;
          bil.x = w1 * v1.x + w2 * v2.x + w3 * v3.x + w4 * v4.x;
          bil.y = w1 * v1.y + w2 * v2.y + w3 * v3.y + w4 * v4.y;
          bil.z = w1 * v1.z + w2 * v2.z + w3 * v3.z + w4 * v4.z;
          bil.w = w1 * v1.w + w2 * v2.w + w3 * v3.w + w4 * v4.w;
          dw.x = hh * (v2.x - v1.x) + lh * (v4.x - v3.x);
          dw.y = hh * (v2.y - v1.y) + lh * (v4.y - v3.y);
          dw.z = hh * (v2.z - v1.z) + lh * (v4.z - v3.z);
          dw.w = hh * (v2.w - v1.w) + lh * (v4.w - v3.w);
          dh.x = hw * (v3.x - v1.x) + lw * (v4.x - v2.x);
          dh.y = hw * (v3.y - v1.y) + lw * (v4.y - v2.y);
          dh.z = hw * (v3.z - v1.z) + lw * (v4.z - v2.z);
          dh.w = hw * (v3.w - v1.w) + lw * (v4.w - v2.w);
          sa += dot4(gk, bil);
          sw += dot4(dw, tg);
          sh += dot4(dh, tg);
        }
      }
      ga[i] = sa;
      gl[2 * i] = sw * fW;
      gl[2 * i + 1] = sh * fH;
    }

    if (active) {  // this (pair, level)'s records: 2P and P contiguous floats
      float4* glp = reinterpret_cast<float4*>(prm.grad_loc + (pair * L + l) * 2 * P);
      float4* gap = reinterpret_cast<float4*>(prm.grad_attn + (pair * L + l) * P);
#pragma unroll
      for (int i = 0; i < 2 * P / 4; ++i) glp[i] = make_float4(gl[4 * i], gl[4 * i + 1], gl[4 * i + 2], gl[4 * i + 3]);
#pragma unroll
      for (int i = 0; i < P / 4; ++i) gap[i] = make_float4(ga[4 * i], ga[4 * i + 1], ga[4 * i + 2], ga[4 * i + 3]);
    }

    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8u * s);
    if (l + 2 < L && warp == l % nwarps && lane == 0) {
      mbar_wait(bar_empty + 8u * s, ph);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue_level(l + 2);
    }
  }
}

template <int D, int P>
int launch_vg_bwd(const CUtensorMap* maps, const VgBwdParams& prm, const VgPlan& pl, int tiles, int B, cudaStream_t st) {
  auto kern = msda_vg_bwd_kernel<D, P>;
  MVD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  dim3 grid((unsigned)tiles, (unsigned)prm.M, (unsigned)B);
  kern<<<grid, pl.threads, pl.smem, st>>>(maps[0], maps[1], maps[2], prm);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

}  // namespace
}  // namespace mvd

using namespace mvd;

extern "C" int mvd_msda_bwd_viewgrid_f32(const float* grad_out, const float* value, const float* loc, const float* attn,
                                         int B, int H, int W, int M, int D, int L, int R, int P, float* grad_value,
                                         float* grad_loc, float* grad_attn, void* stream) {
  if (!grad_out || !value || !loc || !attn || !grad_value || !grad_loc || !grad_attn) return MVD_ERR_NULL_POINTER;
  if (B <= 0 || H <= 0 || W <= 0 || M <= 0 || D <= 0 || L <= 0 || R <= 0 || P <= 0) return MVD_ERR_BAD_SHAPE;
  if ((int64_t)B * L * H * W * M * D > 0x7fffffffLL || (int64_t)B * R * H * W * M * L * P * 2 > 0x7fffffffffLL)
    return MVD_ERR_UNSUPPORTED;  // generic kernel (64-bit indexing) takes the call
  if (M > 65535 || B > 65535) return MVD_ERR_UNSUPPORTED;
  if (!((D == 8 || D == 16 || D == 32) && (P == 4 || P == 8))) return MVD_ERR_UNSUPPORTED;
  const uintptr_t al = reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(value) |
                       reinterpret_cast<uintptr_t>(loc) | reinterpret_cast<uintptr_t>(attn) |
                       reinterpret_cast<uintptr_t>(grad_value) | reinterpret_cast<uintptr_t>(grad_loc) |
                       reinterpret_cast<uintptr_t>(grad_attn);
  if (al & 15u) return MVD_ERR_MISALIGNED;
  cudaStream_t st = (cudaStream_t)stream;
  VgPlan pl;
  if (!plan_viewgrid(D, R, P, false, &pl)) return MVD_ERR_UNSUPPORTED;
  pl.tiles_x = (W + pl.TW - 1) / pl.TW;
  pl.tiles_y = (H + pl.TH - 1) / pl.TH;
  alignas(64) CUtensorMap maps[3];
  const cuuint64_t u = 1;
  {
    const cuuint64_t gdim[5] = {u * D, u * M, u * W, u * H, u * B * L};
    const cuuint64_t gstr[4] = {u * D * 4, u * M * D * 4, u * W * M * D * 4, u * H * W * M * D * 4};
    const cuuint32_t box[5] = {(cuuint32_t)D, 1u, (cuuint32_t)pl.BW, (cuuint32_t)pl.BH, 1u};
    if (int e = encode(&maps[0], value, 5, gdim, gstr, box)) return e;
  }
  {
    const cuuint64_t n = u * L * P * 2;
    const cuuint64_t gdim[5] = {n, u * M, u * W, u * H, u * B * R};
    const cuuint64_t gstr[4] = {n * 4, n * M * 4, n * M * W * 4, n * M * W * H * 4};
    const cuuint32_t box[5] = {(cuuint32_t)(2 * P), 1u, (cuuint32_t)pl.TW, (cuuint32_t)pl.TH, (cuuint32_t)R};
    if (int e = encode(&maps[1], loc, 5, gdim, gstr, box)) return e;
  }
  {
    const cuuint64_t n = u * L * P;
    const cuuint64_t gdim[5] = {n, u * M, u * W, u * H, u * B * R};
    const cuuint64_t gstr[4] = {n * 4, n * M * 4, n * M * W * 4, n * M * W * H * 4};
    const cuuint32_t box[5] = {(cuuint32_t)P, 1u, (cuuint32_t)pl.TW, (cuuint32_t)pl.TH, (cuuint32_t)R};
    if (int e = encode(&maps[2], attn, 5, gdim, gstr, box)) return e;
  }
  MVD_CUDA_TRY(cudaMemsetAsync(grad_value, 0, sizeof(float) * (size_t)B * L * H * W * M * D, st));
  VgBwdParams prm;
  prm.value = value;
  prm.grad_out = grad_out;
  prm.grad_value = grad_value;
  prm.grad_loc = grad_loc;
  prm.grad_attn = grad_attn;
  prm.H = H;
  prm.W = W;
  prm.M = M;
  prm.L = L;
  prm.R = R;
  prm.TH = pl.TH;
  prm.TW = pl.TW;
  prm.tiles_x = pl.tiles_x;
  prm.BW = pl.BW;
  prm.BH = pl.BH;
  prm.off_a = pl.off_a;
  prm.off_b = pl.off_b;
  prm.stage_bytes = pl.stage_bytes;
  prm.zero_off = pl.zero_off;
  prm.zero_bytes = pl.zero_bytes;
  prm.tx_bytes = pl.tx_bytes;
  const int tiles = pl.tiles_x * pl.tiles_y;
#define MVD_VGB(DD, PP) \
  if (D == DD && P == PP) return launch_vg_bwd<DD, PP>(maps, prm, pl, tiles, B, st)
  MVD_VGB(8, 4);
  MVD_VGB(16, 4);
  MVD_VGB(32, 4);
  MVD_VGB(8, 8);
  MVD_VGB(16, 8);
  MVD_VGB(32, 8);
#undef MVD_VGB
  return MVD_ERR_UNSUPPORTED;
}
