// Multi-scale deformable attention forward for the MVDeTr encoder layout ("view grid"): every level is one camera
// view of the same HxW ground-plane grid and the Lq = R*H*W queries are R copies of that grid
//   (ref: multiview_detector/models/trans_world_feat.py:92, multiview_detector/models/mvdetr.py:129-130).
// Same maths as the generic kernels of msda_fwd.cu (ref: ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299).
//
// A block owns a TH x TW tile of ground cells x all R views x ONE head m; one thread owns one (query, head) pair and
// all D channels (no cross-lane traffic, no redundant sample arithmetic). Per level, EVERYTHING the block reads
// arrives in shared memory through the TMA unit, two stages deep, one mbarrier transaction count per stage:
//   * the (TH+2*HALO) x (TW+2*HALO) pixel window of value[:, level, :, m, :]  (5-D map d,m,x,y,level; out-of-map
//     coordinates are zero-filled by the TMA unit, which IS the op's zero padding: in-window samples need no masks);
//   * the tile's sampling locations / offsets  [R][TH][TW][P][2]  (5-D map over loc: box 2P x 1 x TW x TH x R);
//   * the tile's attention weights / logits     [R][TH][TW][P]     (same, box P x 1 x TW x TH x R);
//   * FUSED: the tile's rows of the level-major reference table [TH][TW][P][2] (3-D map).
// r01a ncu (profiles/r01a_ncu_full_summary.txt) showed why: with per-lane global loads of loc/attn every lane touched
// its own 128-byte line (33 L1 wavefronts per request, 26 % of all data-pipe wavefronts), the loads stalled every
// level on L2 latency (long_scoreboard 22 % of samples) and the 256-bit staging registers spilled.
// Pipeline: no block-wide barrier in the level loop. Consumer warps arrive on the stage's `empty` mbarrier when done
// with a level; the refill of that stage (level l+2) is issued by one lane of warp (l mod nwarps) after it has seen
// `empty` complete, so only that warp ever waits for the slowest one.
// Corners are 128-bit shared loads with a per-lane rotation of the channel quads (neighbouring pixels hit disjoint
// bank groups; the accumulators stay in rotated order until the final store). Samples outside the level read a
// zeroed pad region with zero weights (exactly 0, no branch); samples whose 2x2 footprint leaves the window take a
// masked global-memory path, so staging changes speed, never results.
#include "vg_common.cuh"

namespace mvd {

namespace {

struct VgParams {
  const float* value;  // [B, L*H*W, M, D]  (global fallback path)
  float* out;          // [B, Lq, M*D]
  const float* off_bias;    // FUSED, nullable: [M, L, P, 2] added to the raw offsets
  const float* logit_bias;  // FUSED, nullable: [M, L, P]
  int H, W, M, L, R;
  int TH, TW, tiles_x;
  int BW, BH, halo;                           // window = tile + 2*halo
  uint32_t off_a, off_b, off_ref;             // byte offsets of the per-stage regions (window is at 0)
  uint32_t stage_bytes, zero_off, zero_bytes; // stage pitch; zero pad region (after both stages)
  uint32_t tx_bytes;                          // bytes landing per stage
};

// Blocks per SM: the D=16/P=4 stage (51 KB) fits twice; larger head dims or point counts are limited to one block by
// shared memory anyway, so they get the whole register file.
// SPLIT = 2 (D = 32): two ADJACENT lanes share a (query, head) pair and own 16 channels each. r02i: with one thread per
// pair the D=32 / P=8 stress shape ran 256 threads = 8 warps per SM with 32 accumulators each and sat at 0.06 of the
// roofline, bound by shared-memory latency; the split doubles the warps per window and halves the registers per thread.
// Both lanes of a pair repeat the sample arithmetic (identical values) and read the same records (broadcast).
constexpr int kMaxThreadsSplit = 512;
template <int D, int P, bool FUSED, int SPLIT>
__global__ void __launch_bounds__(SPLIT == 2 ? kMaxThreadsSplit : kMaxThreads, (D >= 32 || P >= 8) ? 1 : 2)
    msda_vg_kernel(const __grid_constant__ CUtensorMap tm_val, const __grid_constant__ CUtensorMap tm_a,
                   const __grid_constant__ CUtensorMap tm_b, const __grid_constant__ CUtensorMap tm_ref,
                   const VgParams prm) {
  constexpr int NQ = D / 4 / SPLIT;  // channel quads (16-byte pieces) of a head-pixel owned by this thread
  constexpr int PX_BYTES = D * 4;
  constexpr int PC = 4;  // samples prepared together (P is a multiple of 4)
  constexpr unsigned FULL = 0xffffffffu;
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
  const int wy0 = ty0 - prm.halo, wx0 = tx0 - prm.halo;  // window origin in level pixels (may be negative: TMA zero fill)

  auto issue_level = [&](int lv) {  // one lane: everything level `lv` needs -> stage lv & 1
    const uint32_t st = smem0 + (uint32_t)(lv & 1) * prm.stage_bytes, bar = bar_full + 8u * (uint32_t)(lv & 1);
    mbar_expect_tx(bar, prm.tx_bytes);
    tma_load_5d(st, &tm_val, bar, 0, m, wx0, wy0, b * L + lv);
    tma_load_5d(st + prm.off_a, &tm_a, bar, lv * 2 * P, m, tx0, ty0, b * R);
    tma_load_5d(st + prm.off_b, &tm_b, bar, lv * P, m, tx0, ty0, b * R);
    if (FUSED) tma_load_3d(st + prm.off_ref, &tm_ref, bar, tx0 * 2 * P, ty0, lv);
  };

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_val);
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_b);
    if (FUSED) prefetch_tmap(&tm_ref);
    mbar_init(bar_full, 1);
    mbar_init(bar_full + 8, 1);
    mbar_init(bar_empty, (uint32_t)nwarps);
    mbar_init(bar_empty + 8, (uint32_t)nwarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (uint32_t i = threadIdx.x * 16u; i < prm.zero_bytes; i += blockDim.x * 16u)
    *reinterpret_cast<float4*>(smem_raw + prm.zero_off + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  if (warp == 0) {
    if (elect_one()) {
      issue_level(0);
      if (L > 1) issue_level(1);
    }
    __syncwarp();
  }

  // ---- this thread's (query, head) pair ----
  const int TP = prm.TH * prm.TW;
  const int t = threadIdx.x / SPLIT, half = threadIdx.x % SPLIT;  // pair index inside the block, channel half
  const int r = t / TP, pos = t - r * TP;
  const int ty = pos / prm.TW, tx = pos - ty * prm.TW;
  const int y = ty0 + ty, x = tx0 + tx;
  const bool active = r < R && y < H && x < W;
  const int HW = H * W;
  const int q = r * HW + y * W + x;
  const int64_t Lq = (int64_t)R * HW;
  const int64_t pair = active ? ((int64_t)b * Lq + q) * M + m : 0;
  const float fH = (float)H, fW = (float)W;

  // per-lane rotation of the channel quads (bank-conflict avoidance, see header): quad k of this thread's
  // accumulator holds channels 4*((k+rot)%NQ) .. +3 of its half. SPLIT = 2: lane = 2 * pixel + half, a quarter-warp
  // covers 4 pixels x 2 halves = 8 distinct 16-byte bank groups of the 128-byte pixel pitch.
  const int rot = (NQ >= 8) ? (lane & (NQ - 1)) : (NQ == 4 ? ((lane >> 1) & 3) : ((lane >> 2) & (NQ - 1)));
  int qoff[NQ];  // byte offset of quad k inside a head-pixel
#pragma unroll
  for (int k = 0; k < NQ; ++k) qoff[k] = half * (NQ * 16) + ((k + rot) & (NQ - 1)) * 16;

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
    const int s = l & 1;
    const uint32_t ph = (uint32_t)((l >> 1) & 1);
    const unsigned char* win = smem_raw + (size_t)s * prm.stage_bytes;
    const int zoff = (int)prm.zero_off - (int)(s * prm.stage_bytes);  // zero pad relative to this stage's window
    mbar_wait(bar_full + 8u * s, ph);

    float xy[2 * P], aw[P];
    if (active) {
      read_record<2 * P>(win + prm.off_a, t, lane, xy);
      read_record<P>(win + prm.off_b, t, lane, aw);
      if (FUSED) {
        float rf[2 * P];
        read_record<2 * P>(win + prm.off_ref, pos, lane, rf);
        if (prm.off_bias) {  // block-uniform addresses: one broadcast wavefront each
          const float* ob = prm.off_bias + ((int64_t)m * L + l) * 2 * P;
#pragma unroll
          for (int i = 0; i < 2 * P; ++i) xy[i] += __ldg(ob + i);
        }
        if (prm.logit_bias) {
          const float* lb = prm.logit_bias + ((int64_t)m * L + l) * P;
#pragma unroll
          for (int i = 0; i < P; ++i) aw[i] += __ldg(lb + i);
        }
#pragma unroll
        for (int i = 0; i < P; ++i) {
          // same operation order as the reference module: ref + off / (W, H)
          xy[2 * i] = rf[2 * i] + xy[2 * i] / fW;
          xy[2 * i + 1] = rf[2 * i + 1] + xy[2 * i + 1] / fH;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) xy[2 * i] = xy[2 * i + 1] = aw[i] = 0.f;
    }
    if (FUSED) {
      float lm = aw[0];
#pragma unroll
      for (int i = 1; i < P; ++i) lm = fmaxf(lm, aw[i]);
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
        aw[i] = (new_max == -INFINITY) ? 0.f : exp2f((aw[i] - new_max) * kLog2e);
        run_sum += aw[i];
      }
    }

#pragma unroll
    for (int pc = 0; pc < P; pc += PC) {
      // ---- sample arithmetic for PC samples (ref: ms_deform_im2col_cuda.cuh:285-288, :33-84) ----
      float w1[PC], w2[PC], w3[PC], w4[PC];
      int off[PC];
      unsigned far = 0u;  // valid samples whose footprint is not inside the window (handled after the window pass)
#pragma unroll
      for (int i = 0; i < PC; ++i) {
        const float a = aw[pc + i];
        // one FFMA, exactly as nvcc compiles the reference's `loc * spatial - 0.5` (cuh:285-286)
        const float h_im = fmaf(xy[2 * (pc + i) + 1], fH, -0.5f);
        const float w_im = fmaf(xy[2 * (pc + i)], fW, -0.5f);
        // false for NaN and for threads without a query (tile overhang): those behave like outside samples
        const bool v = active && h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW;
        const float hf = floorf(h_im), wf = floorf(w_im);
        // float -> int conversion saturates (NaN -> 0): for samples outside the level h0 / w0 only pick a bank group below
        const int h0 = (int)hf, w0 = (int)wf;
        const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
        // a sample outside the level (or outside the window) contributes exactly 0 to the window pass: zero weights on
        // the zero pad (never NaN * 0)
        const int dx = w0 - wx0, dy = h0 - wy0;
        const bool in = v && (unsigned)dx < (unsigned)(BW - 1) && (unsigned)dy < (unsigned)(BH - 1);
        w1[i] = in ? hh * hw * a : 0.f;
        w2[i] = in ? hh * lw * a : 0.f;
        w3[i] = in ? lh * hw * a : 0.f;
        w4[i] = in ? lh * lw * a : 0.f;
        // zero-pad reads keep the bank group of the lane's would-be window address (its true position, offset mod 128).
        // r02i/r02o ncu source page: every corner load carried ~12 % excess wavefronts -- in the border warps the lanes
        // whose sample leaves the level all read ONE zero-pad address, which collided with a neighbour's window read.
        const unsigned nat = ((unsigned)dy * (unsigned)BW + (unsigned)dx) * (unsigned)PX_BYTES;
        off[i] = in ? (int)nat : zoff + (int)(nat & 127u);
        far |= (unsigned)(v && !in) << i;
      }
      // ---- window path, branch-free: far samples were given zero weights on the zero pad ----
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
      // ---- far samples (2x2 footprint not inside the window): masked global loads, generic-kernel semantics. Each
      //      lane walks its own far samples, so the warp iterates max-over-lanes(count) times, usually 0 or 1 ----
      while (__any_sync(FULL, far != 0u)) {
        if (far) {
          const int i = __ffs(far) - 1;
          far &= far - 1u;
          // select sample i of this chunk without dynamic register indexing
          float sx = xy[2 * pc], sy = xy[2 * pc + 1], sa = aw[pc];
#pragma unroll
          for (int j = 1; j < PC; ++j) {
            sx = (i == j) ? xy[2 * (pc + j)] : sx;
            sy = (i == j) ? xy[2 * (pc + j) + 1] : sy;
            sa = (i == j) ? aw[pc + j] : sa;
          }
          const float h_im = fmaf(sy, fH, -0.5f), w_im = fmaf(sx, fW, -0.5f);
          const float hf = floorf(h_im), wf = floorf(w_im);
          const int h0 = (int)hf, w0 = (int)wf;
          const float lh = h_im - hf, lw = w_im - wf, hh = 1.f - lh, hw = 1.f - lw;
          const float g1 = hh * hw * sa, g2 = hh * lw * sa, g3 = lh * hw * sa, g4 = lh * lw * sa;
          const bool top = h0 >= 0, bot = h0 + 1 <= H - 1, lef = w0 >= 0, rig = w0 + 1 <= W - 1;
          const float* p00 = vb + ((int64_t)l * HW + (int64_t)h0 * W + w0) * stride_px;
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int k = 0; k < NQ; ++k) {
            const float* pk = p00 + (qoff[k] >> 2);
            const float4 v1 = (top && lef) ? __ldg(reinterpret_cast<const float4*>(pk)) : z;
            const float4 v2 = (top && rig) ? __ldg(reinterpret_cast<const float4*>(pk + stride_px)) : z;
            const float4 v3 = (bot && lef) ? __ldg(reinterpret_cast<const float4*>(pk + (int64_t)W * stride_px)) : z;
            const float4 v4 = (bot && rig) ? __ldg(reinterpret_cast<const float4*>(pk + (int64_t)(W + 1) * stride_px)) : z;
            fma4(acc[k], g1, v1, g2, v2, g3, v3, g4, v4);
          }
        }
      }
    }

    // ---- release the stage; one warp (rotating) refills it with level l+2 once every warp has released it ----
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8u * s);
    if (l + 2 < L && warp == l % nwarps) {  // converged warp; the TMA instructions under elect (no uniform-register waterfall)
      mbar_wait(bar_empty + 8u * s, ph);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads of this stage -> async-proxy refill
      if (elect_one()) issue_level(l + 2);
      __syncwarp();
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

template <int D, int P, bool FUSED>
int launch_vg(const CUtensorMap* maps, const VgParams& prm, const VgPlan& pl, int tiles, int B, cudaStream_t st) {
  auto kern = msda_vg_kernel<D, P, FUSED, (D >= 32 ? 2 : 1)>;
  MVD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  dim3 grid((unsigned)tiles, (unsigned)prm.M, (unsigned)B);
  kern<<<grid, pl.threads, pl.smem, st>>>(maps[0], maps[1], maps[2], maps[3], prm);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

template <bool FUSED>
int viewgrid_dispatch(const float* value, const float* loc, const float* attn, const float* ref,
                      const float* off_bias, const float* logit_bias, int B, int H, int W, int M, int D, int L, int R,
                      int P, int Lr, float* out, cudaStream_t st, int loc_pitch = 0, int attn_pitch = 0) {
  // floats between consecutive queries in loc / attn: 0 = dense (M*L*P*2 and M*L*P); larger when both are column ranges
  // of one [queries, M*L*P*3] GEMM output (offsets and logits produced by a single launch)
  if (loc_pitch == 0) loc_pitch = M * L * P * 2;
  if (attn_pitch == 0) attn_pitch = M * L * P;
  if (loc_pitch < M * L * P * 2 || attn_pitch < M * L * P || (loc_pitch | attn_pitch) % 4 != 0) return MVD_ERR_BAD_SHAPE;
  if (B <= 0 || H <= 0 || W <= 0 || M <= 0 || D <= 0 || L <= 0 || R <= 0 || P <= 0) return MVD_ERR_BAD_SHAPE;
  // beyond the 32-bit indexing of this kernel: not an error, the generic kernel (64-bit indexing) takes the call
  if ((int64_t)B * L * H * W * M * D > 0x7fffffffLL || (int64_t)B * R * H * W * M * L * P * 2 > 0x7fffffffffLL ||
      (int64_t)B * R * H * W * (loc_pitch > attn_pitch ? loc_pitch : attn_pitch) > 0x7fffffffffLL)
    return MVD_ERR_UNSUPPORTED;
  if (M > 65535 || B > 65535) return MVD_ERR_UNSUPPORTED;
  if (!((D == 8 || D == 16 || D == 32) && (P == 4 || P == 8))) return MVD_ERR_UNSUPPORTED;
  if (FUSED && Lr != H * W) return MVD_ERR_UNSUPPORTED;  // table rows must be the grid cells (mvdetr.py:33-71)
  const uintptr_t al = reinterpret_cast<uintptr_t>(value) | reinterpret_cast<uintptr_t>(loc) |
                       reinterpret_cast<uintptr_t>(attn) | reinterpret_cast<uintptr_t>(out) |
                       reinterpret_cast<uintptr_t>(ref);
  if (al & 15u) return MVD_ERR_MISALIGNED;
  VgPlan pl;
  const int split = D >= 32 ? 2 : 1;  // must match launch_vg
  if (!plan_viewgrid(D, R, P, FUSED, &pl, viewgrid_halo(P), split, split == 2 ? kMaxThreadsSplit : kMaxThreads))
    return MVD_ERR_UNSUPPORTED;
  pl.tiles_x = (W + pl.TW - 1) / pl.TW;
  pl.tiles_y = (H + pl.TH - 1) / pl.TH;

  alignas(64) CUtensorMap maps[4];
  const cuuint64_t u = 1;
  {  // value [B*L][H][W][M][D]
    const cuuint64_t gdim[5] = {u * D, u * M, u * W, u * H, u * B * L};
    const cuuint64_t gstr[4] = {u * D * 4, u * M * D * 4, u * W * M * D * 4, u * H * W * M * D * 4};
    const cuuint32_t box[5] = {(cuuint32_t)D, 1u, (cuuint32_t)pl.BW, (cuuint32_t)pl.BH, 1u};
    if (int e = encode(&maps[0], value, 5, gdim, gstr, box)) return e;
  }
  {  // loc / offsets [B*R][H][W][M][L*P*2]
    const cuuint64_t n = u * L * P * 2, q = u * loc_pitch;
    const cuuint64_t gdim[5] = {n, u * M, u * W, u * H, u * B * R};
    const cuuint64_t gstr[4] = {n * 4, q * 4, q * W * 4, q * W * H * 4};
    const cuuint32_t box[5] = {(cuuint32_t)(2 * P), 1u, (cuuint32_t)pl.TW, (cuuint32_t)pl.TH, (cuuint32_t)R};
    if (int e = encode(&maps[1], loc, 5, gdim, gstr, box)) return e;
  }
  {  // attn / logits [B*R][H][W][M][L*P]
    const cuuint64_t n = u * L * P, q = u * attn_pitch;
    const cuuint64_t gdim[5] = {n, u * M, u * W, u * H, u * B * R};
    const cuuint64_t gstr[4] = {n * 4, q * 4, q * W * 4, q * W * H * 4};
    const cuuint32_t box[5] = {(cuuint32_t)P, 1u, (cuuint32_t)pl.TW, (cuuint32_t)pl.TH, (cuuint32_t)R};
    if (int e = encode(&maps[2], attn, 5, gdim, gstr, box)) return e;
  }
  if (FUSED) {  // reference table, level-major [L][H][W*P*2]
    const cuuint64_t n = u * W * P * 2;
    const cuuint64_t gdim[3] = {n, u * H, u * L};
    const cuuint64_t gstr[2] = {n * 4, n * H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)(pl.TW * P * 2), (cuuint32_t)pl.TH, 1u};
    if (int e = encode(&maps[3], ref, 3, gdim, gstr, box)) return e;
  } else {
    maps[3] = maps[2];
  }

  VgParams prm;
  prm.value = value;
  prm.out = out;
  prm.off_bias = off_bias;
  prm.logit_bias = logit_bias;
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
  prm.halo = pl.halo;
  prm.off_a = pl.off_a;
  prm.off_b = pl.off_b;
  prm.off_ref = pl.off_ref;
  prm.stage_bytes = pl.stage_bytes;
  prm.zero_off = pl.zero_off;
  prm.zero_bytes = pl.zero_bytes;
  prm.tx_bytes = pl.tx_bytes;
  const int tiles = pl.tiles_x * pl.tiles_y;
#define MVD_VG(DD, PP) \
  if (D == DD && P == PP) return launch_vg<DD, PP, FUSED>(maps, prm, pl, tiles, B, st)
  MVD_VG(8, 4);
  MVD_VG(16, 4);
  MVD_VG(32, 4);
  MVD_VG(8, 8);
  MVD_VG(16, 8);
  MVD_VG(32, 8);
#undef MVD_VG
  return MVD_ERR_UNSUPPORTED;
}

}  // namespace
}  // namespace mvd

using namespace mvd;

extern "C" int mvd_msda_fwd_viewgrid_f32(const float* value, const float* loc, const float* attn, int B, int H,
                                         int W, int M, int D, int L, int R, int P, float* out, void* stream) {
  if (!value || !loc || !attn || !out) return MVD_ERR_NULL_POINTER;
  return viewgrid_dispatch<false>(value, loc, attn, nullptr, nullptr, nullptr, B, H, W, M, D, L, R, P, 1, out, (cudaStream_t)stream);
}

// offsets / logits as column ranges of wider rows (one GEMM producing both): pitches in floats per query, multiples of 4
extern "C" int mvd_msda_fused_fwd_viewgrid_pitched_f32(const float* value, const float* offsets, const float* logits,
                                                       const float* ref, const float* off_bias,
                                                       const float* logit_bias, int B, int H, int W, int M, int D,
                                                       int L, int R, int P, int Lr, int off_pitch, int logit_pitch,
                                                       float* out, void* stream) {
  if (!value || !offsets || !logits || !ref || !out) return MVD_ERR_NULL_POINTER;
  if (Lr <= 0 || off_pitch <= 0 || logit_pitch <= 0) return MVD_ERR_BAD_SHAPE;
  return viewgrid_dispatch<true>(value, offsets, logits, ref, off_bias, logit_bias, B, H, W, M, D, L, R, P, Lr, out,
                                 (cudaStream_t)stream, off_pitch, logit_pitch);
}

extern "C" int mvd_msda_fused_fwd_viewgrid_f32(const float* value, const float* offsets, const float* logits,
                                               const float* ref, const float* off_bias,
                                               const float* logit_bias, int B, int H, int W, int M, int D, int L,
                                               int R, int P, int Lr, float* out, float* attn_out, float* loc_out,
                                               void* stream) {
  if (!value || !offsets || !logits || !ref || !out) return MVD_ERR_NULL_POINTER;
  if (Lr <= 0) return MVD_ERR_BAD_SHAPE;
  if (attn_out || loc_out) return MVD_ERR_UNSUPPORTED;  // normalisation is deferred here; the generic kernel writes them
  return viewgrid_dispatch<true>(value, offsets, logits, ref, off_bias, logit_bias, B, H, W, M, D, L, R, P, Lr, out,
                                 (cudaStream_t)stream);
}
