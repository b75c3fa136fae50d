// Perspective warp as ONE launch on the backbone's own layout: NCHW source tiles staged by TMA, bilinear gather from
// shared memory, destination written as NCHW (kornia contract), channels-last, or straight into the im2col matrix of
// the stride-s 3x3 convolution that consumes the warped grid.
//   replaces kornia.warp_perspective(...)                    call site ref: multiview_detector/models/mvdetr.py:194-195
//   (+ the permute-copy of ref: multiview_detector/models/trans_world_feat.py:92 in the channels-last / im2col modes)
// Same arithmetic, operation order and results as warp.cu / im2col.cu (restated kornia semantics, oracle/warp_ref.c).
//
// Why: round 1 gathered from a channels-last copy of the source, so an NCHW input (what the reference backbone always
// hands over, mvdetr.py:177-178) first went through a 51.6 MB relayout kernel (2 launches, +103 MB of HBM traffic at
// Wildtrack size). Here the relayout happens on chip:
//   * a block owns a TH x TW tile of destination pixels of one view. Every thread evaluates the homography for its
//     pixel; a block-wide min/max gives the source bounding box of the tile (a ground-plane tile is a small patch of
//     the camera image: median 36 source pixels for a 16x16 tile at Wildtrack geometry, tools: scripts/warp_bbox.py);
//   * the bounding box is fetched with TMA as boxes {bw, 4 rows, 32 channels} of the NCHW tensor map (bw = 16/32/64,
//     out-of-image coordinates zero-filled by the TMA unit = the op's zero padding), two 32 KB stages on mbarriers,
//     as many 32-channel groups per stage as fit;
//   * gather: thread = destination pixel, 4 scalar shared loads per channel (neighbouring pixels read neighbouring
//     words); results cross to channel-major through a [32][257] shared tile (conflict-free both ways) and leave as
//     128-bit stores of 4 consecutive channels, 128 contiguous bytes per pixel and 32-channel group;
//   * a tile whose bounding box does not fit a stage (strong minification) takes the same code with masked global
//     loads instead of the staged window, so staging changes speed, never results.
#include "vg_common.cuh"
#include "warp_taps.cuh"

namespace mvd {
namespace {

constexpr int kWtThreads = 256;
constexpr int kWtCC = 32;               // channels per TMA box / per transposed sub-chunk
constexpr int kWtRows = 4;              // source rows per TMA box
constexpr int kWtStageFloats = 8192;    // 32 KB per stage
constexpr int kWtPitch = kWtThreads + 1;

enum { WT_NCHW = 0, WT_NHWC = 1, WT_IM2COL = 2 };

struct WtParams {
  const float* src;  // [BN, C, Hi, Wi]
  const float* Mat;  // [BN, 3, 3]
  float* dst;
  int C, Hi, Wi, Ho, Wo;
  int tiles_x;
  int Ho2, Wo2;  // im2col: token grid
};

// MODE: destination layout. TH x TW = kWtThreads destination pixels per block. S: convolution stride (im2col only).
template <int MODE, int TH, int TW, int S>
__global__ void __launch_bounds__(kWtThreads, 2)
    warp_tma_kernel(const __grid_constant__ CUtensorMap tm16, const __grid_constant__ CUtensorMap tm32,
                    const __grid_constant__ CUtensorMap tm64, const WtParams prm) {
  static_assert(TH * TW == kWtThreads, "one thread per destination pixel");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage = reinterpret_cast<float*>(smem_raw);  // 2 x kWtStageFloats
  float* tile = stage + 2 * kWtStageFloats;           // [kWtCC][kWtPitch] (channels-last / im2col modes)
  __shared__ float sT[9];
  __shared__ int s_box[4];
  __shared__ __align__(8) unsigned long long s_bar[2];

  const int C = prm.C, Hi = prm.Hi, Wi = prm.Wi, Ho = prm.Ho, Wo = prm.Wo;
  const int n = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int PAD = (MODE == WT_IM2COL) ? 1 : 0;  // im2col walks the padded pixel domain [-1, H] x [-1, W]
  const int ty0 = (blockIdx.x / prm.tiles_x) * TH - PAD, tx0 = (blockIdx.x % prm.tiles_x) * TW - PAD;
  const uint32_t bar0 = smem_u32(&s_bar[0]);

  if (tid == 0) {
    normalized_inverse(prm.Mat + 9 * n, Hi, Wi, Ho, Wo, sT);
    s_box[0] = s_box[1] = 0x7fffffff;
    s_box[2] = s_box[3] = -0x7fffffff;
    prefetch_tmap(&tm16);
    prefetch_tmap(&tm32);
    prefetch_tmap(&tm64);
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  // ---- this thread's destination pixel and its taps (warp_taps.cuh arithmetic) ----
  const int v = ty0 + tid / TW, u = tx0 + tid % TW;
  const bool real = v >= 0 && v < Ho && u >= 0 && u < Wo;
  const Taps t = make_taps(sT, real ? u : 0, real ? v : 0, Hi, Wi, Ho, Wo);
  const bool valid = real && (t.m_nw || t.m_ne || t.m_sw || t.m_se);
  const int y0 = t.y0, x0 = t.x0;  // north-west tap in source pixels
  {
    constexpr unsigned FULL = 0xffffffffu;
    const int big = 0x7fffffff;
    int xl = valid ? x0 : big, yl = valid ? y0 : big, xh = valid ? x0 + 1 : -big, yh = valid ? y0 + 1 : -big;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      xl = min(xl, __shfl_xor_sync(FULL, xl, d));
      yl = min(yl, __shfl_xor_sync(FULL, yl, d));
      xh = max(xh, __shfl_xor_sync(FULL, xh, d));
      yh = max(yh, __shfl_xor_sync(FULL, yh, d));
    }
    if (lane == 0 && xl != big) {
      atomicMin(&s_box[0], xl);
      atomicMin(&s_box[1], yl);
      atomicMax(&s_box[2], xh);
      atomicMax(&s_box[3], yh);
    }
  }
  __syncthreads();
  // TMA fetches 16-byte granules: the innermost (x) start coordinate must be a multiple of 4 floats, otherwise the copy
  // faults as an illegal instruction (seen on B200). Round down (two's complement: also right for -1 -> -4).
  const int xmin = s_box[0] & ~3, ymin = s_box[1];
  const bool empty = s_box[2] < xmin;  // no pixel of the tile sees the source image
  const int bbw = empty ? 1 : s_box[2] - xmin + 1, bbh = empty ? 1 : s_box[3] - ymin + 1;
  const int bw = bbw <= 16 ? 16 : (bbw <= 32 ? 32 : 64);
  const int hg = (bbh + kWtRows - 1) / kWtRows;         // TMA boxes stacked vertically
  const int sub_floats = bw * kWtRows * hg * kWtCC;     // one 32-channel group of the bounding box
  const bool staged = !empty && bbw <= 64 && sub_floats <= kWtStageFloats;
  const int nsub = C / kWtCC;
  const int sps = staged ? min(nsub, kWtStageFloats / sub_floats) : 1;  // 32-channel groups per stage
  const int nchunks = (nsub + sps - 1) / sps;

  auto issue_chunk = [&](int ch) {  // one thread: 32-channel groups [ch*sps, ...) -> stage ch & 1
    const int k0 = ch * sps, k1 = min(nsub, k0 + sps);
    const uint32_t bar = bar0 + 8u * (uint32_t)(ch & 1);
    const uint32_t st = smem_u32(stage + (size_t)(ch & 1) * kWtStageFloats);
    const uint32_t box_bytes = (uint32_t)(bw * kWtRows * kWtCC * 4);
    mbar_expect_tx(bar, (uint32_t)(k1 - k0) * (uint32_t)hg * box_bytes);
    for (int k = k0; k < k1; ++k)
      for (int g = 0; g < hg; ++g) {
        const uint32_t dst = st + (uint32_t)((k - k0) * hg + g) * box_bytes;
        // the tensor map operand is a kernel parameter named in the instruction (no computed descriptor address)
        if (bw == 16) tma_load_4d(dst, &tm16, bar, xmin, ymin + g * kWtRows, k * kWtCC, n);
        else if (bw == 32) tma_load_4d(dst, &tm32, bar, xmin, ymin + g * kWtRows, k * kWtCC, n);
        else tma_load_4d(dst, &tm64, bar, xmin, ymin + g * kWtRows, k * kWtCC, n);
      }
  };
  if (staged && tid == 0) {
    issue_chunk(0);
    if (nchunks > 1) issue_chunk(1);
  }

  // staged-window addressing: element (c, y, x) of a group lives at ((g*32 + c)*4 + (y & 3)) * bw + x, g = y >> 2
  const int ry0 = y0 - ymin, ry1 = ry0 + 1, rx = x0 - xmin;
  const int o0 = valid ? (((ry0 >> 2) * kWtCC * kWtRows + (ry0 & 3)) * bw + rx) : 0;
  const int o1 = valid ? (((ry1 >> 2) * kWtCC * kWtRows + (ry1 & 3)) * bw + rx) : 0;
  const int cs = kWtRows * bw;  // channel stride inside a group
  const float wnw = valid ? t.nw : 0.f, wne = valid ? t.ne : 0.f, wsw = valid ? t.sw : 0.f, wse = valid ? t.se : 0.f;
  // global fallback addressing
  const int64_t plane = (int64_t)Hi * Wi;
  const float* gsrc = prm.src + (int64_t)n * C * plane + t.o00;

  const int64_t opix = (int64_t)v * Wo + u;  // NCHW / NHWC destination pixel
  const int64_t oplane = (int64_t)Ho * Wo;

  for (int k = 0; k < nsub; ++k) {
    const int ch = k / sps, kl = k - ch * sps;
    if (staged && kl == 0) mbar_wait(bar0 + 8u * (uint32_t)(ch & 1), (uint32_t)((ch >> 1) & 1));
    const float* sp = stage + (size_t)(ch & 1) * kWtStageFloats + (size_t)kl * sub_floats;
    const int c0 = k * kWtCC;
    // ---- gather 32 channels of this thread's pixel ----
    if (MODE == WT_NCHW) {
      if (real) {
        float* dp = prm.dst + ((int64_t)n * C + c0) * oplane + opix;
#pragma unroll 8
        for (int c = 0; c < kWtCC; ++c) {
          float acc = 0.f;
          if (staged) {
            const float* q = sp + c * cs;
            acc = fmaf(q[o1 + 1], wse, fmaf(q[o1], wsw, fmaf(q[o0 + 1], wne, q[o0] * wnw)));
            acc = valid ? acc : 0.f;
          } else if (valid) {
            const float* s = gsrc + (int64_t)(c0 + c) * plane;
            const float q0 = t.m_nw ? __ldg(s) : 0.f, q1 = t.m_ne ? __ldg(s + 1) : 0.f;
            const float q2 = t.m_sw ? __ldg(s + Wi) : 0.f, q3 = t.m_se ? __ldg(s + Wi + 1) : 0.f;
            acc = fmaf(q3, wse, fmaf(q2, wsw, fmaf(q1, wne, q0 * wnw)));
          }
          __stcs(dp + (int64_t)c * oplane, acc);
        }
      }
    } else {
      if (staged) {
#pragma unroll 8
        for (int c = 0; c < kWtCC; ++c) {
          const float* q = sp + c * cs;
          const float acc = fmaf(q[o1 + 1], wse, fmaf(q[o1], wsw, fmaf(q[o0 + 1], wne, q[o0] * wnw)));
          tile[c * kWtPitch + tid] = valid ? acc : 0.f;
        }
      } else if (!empty || k == 0) {
#pragma unroll 4
        for (int c = 0; c < kWtCC; ++c) {
          float acc = 0.f;
          if (valid) {
            const float* s = gsrc + (int64_t)(c0 + c) * plane;
            const float q0 = t.m_nw ? __ldg(s) : 0.f, q1 = t.m_ne ? __ldg(s + 1) : 0.f;
            const float q2 = t.m_sw ? __ldg(s + Wi) : 0.f, q3 = t.m_se ? __ldg(s + Wi + 1) : 0.f;
            acc = fmaf(q3, wse, fmaf(q2, wsw, fmaf(q1, wne, q0 * wnw)));
          }
          tile[c * kWtPitch + tid] = acc;
        }
      }
      __syncthreads();
      // ---- read out channel-major: 8 lanes x 4 channels per pixel, 4 pixels per warp instruction ----
      const int c4 = (lane & 7) * 4;
#pragma unroll 2
      for (int it = 0; it < kWtThreads / 32; ++it) {
        const int p = it * 32 + warp * 4 + (lane >> 3);
        const int pv = ty0 + p / TW, pu = tx0 + p % TW;
        float4 o;
        o.x = tile[(c4 + 0) * kWtPitch + p];
        o.y = tile[(c4 + 1) * kWtPitch + p];
        o.z = tile[(c4 + 2) * kWtPitch + p];
        o.w = tile[(c4 + 3) * kWtPitch + p];
        if (MODE == WT_NHWC) {
          if (pv < Ho && pu < Wo) st_stream4(prm.dst + (((int64_t)n * Ho + pv) * Wo + pu) * C + c0 + c4, o);
        } else {
          // token (oy, ox) tap (ky, kx) reads pixel (oy*S + ky - 1, ox*S + kx - 1): enumerate the taps that read (pv, pu)
          const int vp = pv + 1, up = pu + 1;  // padded coordinates, >= 0
          if (vp <= Ho + 1 && up <= Wo + 1) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const int a = vp - ky;
              if (a < 0 || (a % S) != 0) continue;
              const int oy = a / S;
              if (oy >= prm.Ho2) continue;
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const int b = up - kx;
                if (b < 0 || (b % S) != 0) continue;
                const int ox = b / S;
                if (ox >= prm.Wo2) continue;
                st_stream4(prm.dst + ((((int64_t)n * prm.Ho2 + oy) * prm.Wo2 + ox) * 9 + ky * 3 + kx) * C + c0 + c4, o);
              }
            }
          }
        }
      }
      __syncthreads();  // tile and (after the stage's last group) the stage are free again
    }
    if (staged && (kl == sps - 1 || k == nsub - 1) && ch + 2 < nchunks) {  // block-uniform
      if (MODE == WT_NCHW) __syncthreads();  // every thread is done reading the stage
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads -> async-proxy refill
        issue_chunk(ch + 2);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Channels-last / im2col destinations, second formulation (r02c profile of the kernel above: 117 us, latency-bound, 3.7 k
// instructions per warp -- a scalar shared load per tap AND per channel plus a [32][257] output transposition).
// A ground-plane tile reads a SMALL patch of the camera image (median 36 source pixels for 256 destination pixels), so
// the relayout is done on the patch, not on the output:
//   1. TMA stages the patch as [channel][row][x] (NCHW boxes, as above);
//   2. the block transposes the patch to pixel-major [patch pixel][CH + 4] (a few thousand elements);
//   3. every destination pixel then reads each tap as ONE contiguous channel vector: lanes = channel quads, 128-bit
//      shared loads, bank-conflict free, and the result leaves as 128-bit stores of consecutive channels
//      (512 contiguous bytes per pixel at CH = 128) -- no output transposition, ~4x fewer instructions.
// CH (channels per pass) is the largest of {128, 64, 32, 16, 8} whose staged patch fits; tiles whose patch does not fit
// even at CH = 8 (strong minification) gather straight from global memory with the same lane mapping.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWcRaw = 8192;    // floats per raw (TMA) stage, two stages
constexpr int kWcTr = 8704;     // floats of the transposed patch
constexpr int kWcBoxes = 6;     // TMA box widths {8, 16, 24, 32, 48, 64} floats x 4 rows x 8 channels
__device__ __constant__ int kWcBw[kWcBoxes] = {8, 16, 24, 32, 48, 64};
constexpr int kWcBoxCh = 8;     // channels per TMA box (so that every CH is a whole number of boxes)

struct WcMaps {
  CUtensorMap m[kWcBoxes];
};

template <int MODE, int S>
__global__ void __launch_bounds__(kWtThreads, 2) warp_tma_cl_kernel(const __grid_constant__ WcMaps maps, const WtParams prm) {
  constexpr int TH = 16, TW = 16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* raw = reinterpret_cast<float*>(smem_raw);  // 2 x kWcRaw
  float* tr = raw + 2 * kWcRaw;                     // kWcTr
  int* s_off = reinterpret_cast<int*>(tr + kWcTr);  // [256] element offset of the north-west tap in `tr` (pixel index)
  float* s_w = reinterpret_cast<float*>(s_off + kWtThreads);  // [256] float4 tap weights nw, ne, sw, se (0 if no valid tap)
  __shared__ float sT[9];
  __shared__ int s_box[4];
  __shared__ __align__(8) unsigned long long s_bar[2];

  const int C = prm.C, Hi = prm.Hi, Wi = prm.Wi, Ho = prm.Ho, Wo = prm.Wo;
  const int n = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int PAD = (MODE == WT_IM2COL) ? 1 : 0;
  const int ty0 = (blockIdx.x / prm.tiles_x) * TH - PAD, tx0 = (blockIdx.x % prm.tiles_x) * TW - PAD;
  const uint32_t bar0 = smem_u32(&s_bar[0]);

  if (tid == 0) {
    normalized_inverse(prm.Mat + 9 * n, Hi, Wi, Ho, Wo, sT);
    s_box[0] = s_box[1] = 0x7fffffff;
    s_box[2] = s_box[3] = -0x7fffffff;
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  const int v = ty0 + tid / TW, u = tx0 + tid % TW;
  const bool real = v >= 0 && v < Ho && u >= 0 && u < Wo;
  const Taps t = make_taps(sT, real ? u : 0, real ? v : 0, Hi, Wi, Ho, Wo);
  const bool valid = real && (t.m_nw || t.m_ne || t.m_sw || t.m_se);
  {
    constexpr unsigned FULL = 0xffffffffu;
    const int big = 0x7fffffff;
    int xl = valid ? t.x0 : big, yl = valid ? t.y0 : big, xh = valid ? t.x0 + 1 : -big, yh = valid ? t.y0 + 1 : -big;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      xl = min(xl, __shfl_xor_sync(FULL, xl, d));
      yl = min(yl, __shfl_xor_sync(FULL, yl, d));
      xh = max(xh, __shfl_xor_sync(FULL, xh, d));
      yh = max(yh, __shfl_xor_sync(FULL, yh, d));
    }
    if (lane == 0 && xl != big) {
      atomicMin(&s_box[0], xl);
      atomicMin(&s_box[1], yl);
      atomicMax(&s_box[2], xh);
      atomicMax(&s_box[3], yh);
    }
  }
  __syncthreads();
  // patch = [xmin, xmax] x [ymin, ymax]; the TMA start coordinate must be a multiple of 4 floats (16 bytes)
  const int xmin = s_box[0] & ~3, ymin = s_box[1];
  const bool empty = s_box[2] < xmin;
  const int bbw = empty ? 1 : s_box[2] - xmin + 1, bbh = empty ? 1 : s_box[3] - ymin + 1;
  int bsel = 0;
  while (bsel < kWcBoxes - 1 && kWcBw[bsel] < bbw) ++bsel;
  const int bw = kWcBw[bsel];
  const int hg = (bbh + kWtRows - 1) / kWtRows;
  // largest channel pass whose raw boxes and transposed patch both fit
  int CH = 0;
  if (!empty && bbw <= 64) {
    for (int c = 128; c >= 8; c >>= 1) {
      if (c <= C && C % c == 0 && bw * kWtRows * hg * c <= kWcRaw && bbw * bbh * (c + 4) <= kWcTr) {
        CH = c;
        break;
      }
    }
  }
  const bool staged = CH > 0;
  if (!staged) CH = (C % 128 == 0) ? 128 : 32;  // global-memory path: same lane mapping
  const int nchunks = C / CH;
  const int pitch = CH + 4;

  // per-pixel tap records for the gather (one thread = one pixel wrote them; any warp reads them)
  s_off[tid] = valid ? ((t.y0 - ymin) * bbw + (t.x0 - xmin)) : -1;
  reinterpret_cast<float4*>(s_w)[tid] = valid ? make_float4(t.nw, t.ne, t.sw, t.se) : make_float4(0.f, 0.f, 0.f, 0.f);

  // whole warp 0, converged: channels [ch*CH, (ch+1)*CH) of the patch -> raw stage ch & 1. Only the TMA instructions sit
  // under elect (a loop under `tid == 0` made the compiler wrap every UTMALDG in a uniform-register waterfall).
  auto issue_chunk = [&](int ch) {
    const uint32_t bar = bar0 + 8u * (uint32_t)(ch & 1);
    const uint32_t st = smem_u32(raw + (size_t)(ch & 1) * kWcRaw);
    const uint32_t box_bytes = (uint32_t)(bw * kWtRows * kWcBoxCh * 4);
    const int nb = CH / kWcBoxCh;
    const bool leader = elect_one();
    if (leader) mbar_expect_tx(bar, (uint32_t)nb * (uint32_t)hg * box_bytes);
    for (int g = 0; g < hg; ++g)
      for (int b = 0; b < nb; ++b)  // raw layout: [row group g][channel][4 rows][bw]
      {  // the tensor map operand is named statically (one case per box width), never a computed address
        const uint32_t dst = st + (uint32_t)(g * nb + b) * box_bytes;
        const int cy = ymin + g * kWtRows, cc = ch * CH + b * kWcBoxCh;
        if (leader) {
          switch (bsel) {
            case 0: tma_load_4d(dst, &maps.m[0], bar, xmin, cy, cc, n); break;
            case 1: tma_load_4d(dst, &maps.m[1], bar, xmin, cy, cc, n); break;
            case 2: tma_load_4d(dst, &maps.m[2], bar, xmin, cy, cc, n); break;
            case 3: tma_load_4d(dst, &maps.m[3], bar, xmin, cy, cc, n); break;
            case 4: tma_load_4d(dst, &maps.m[4], bar, xmin, cy, cc, n); break;
            default: tma_load_4d(dst, &maps.m[5], bar, xmin, cy, cc, n); break;
          }
        }
      }
    __syncwarp();
  };
  if (staged && warp == 0) {
    issue_chunk(0);
    if (nchunks > 1) issue_chunk(1);
  }
  __syncthreads();  // tap records visible

  const int QL = CH / 4;             // lanes per pixel (channel quads)
  const int PPI = 32 / (QL < 32 ? QL : 32);  // pixels per warp instruction
  const int q = lane % QL, psub = lane / QL;
  const int64_t plane = (int64_t)Hi * Wi;

  for (int ch = 0; ch < nchunks; ++ch) {
    const int c0 = ch * CH;
    if (staged) {
      mbar_wait(bar0 + 8u * (uint32_t)(ch & 1), (uint32_t)((ch >> 1) & 1));
      // ---- transpose the patch: raw [g][c][4][bw] -> tr [y*bbw + x][CH + 4] ----
      // lanes run along the patch pixels (consecutive x read consecutive words), warps stride the channels; the two
      // divisions per patch pixel are done once per thread and pixel slot, not per element
      const float* rp = raw + (size_t)(ch & 1) * kWcRaw;
      const int npx = bbw * bbh;
      const int cstride = kWtRows * bw;
      for (int px0 = 0; px0 < npx; px0 += 4 * 32) {
        int ro[4], to[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int px = px0 + j * 32 + lane;
          const int y = px / bbw, x = px - y * bbw;
          ro[j] = px < npx ? (((y >> 2) * CH) * kWtRows + (y & 3)) * bw + x : -1;
          to[j] = px * pitch;
        }
        for (int c = warp; c < CH; c += kWtThreads / 32) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (ro[j] >= 0) tr[to[j] + c] = rp[ro[j] + c * cstride];
        }
      }
      __syncthreads();
      if (warp == 0 && ch + 2 < nchunks) {  // the raw stage is free again
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue_chunk(ch + 2);
      }
    }
    // ---- gather: warp w owns tile pixels [32w, 32w + 32); PPI pixels per instruction, lane = channel quad ----
    for (int it = 0; it < 32 / PPI; ++it) {
      const int p = warp * 32 + it * PPI + psub;
      const int pv = ty0 + p / TW, pu = tx0 + p % TW;
      const int o = s_off[p];
      const float4 w4 = reinterpret_cast<const float4*>(s_w)[p];
      const float wnw = w4.x, wne = w4.y, wsw = w4.z, wse = w4.w;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o >= 0) {
        float4 q0, q1, q2, q3;
        if (staged) {
          const float* b = tr + o * pitch + 4 * q;
          q0 = *reinterpret_cast<const float4*>(b);
          q1 = *reinterpret_cast<const float4*>(b + pitch);
          q2 = *reinterpret_cast<const float4*>(b + bbw * pitch);
          q3 = *reinterpret_cast<const float4*>(b + (bbw + 1) * pitch);
        } else {  // patch too large to stage: masked global loads of the 4 channels of this lane
          const int y0 = o / bbw + ymin, x0 = o - (o / bbw) * bbw + xmin;
          const bool top = y0 >= 0, bot = y0 + 1 <= Hi - 1, lef = x0 >= 0, rig = x0 + 1 <= Wi - 1;
          const float* g = prm.src + ((int64_t)n * C + c0 + 4 * q) * plane + (int64_t)y0 * Wi + x0;
          float a0[4], a1[4], a2[4], a3[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float* gk = g + (int64_t)k * plane;
            a0[k] = (top && lef) ? __ldg(gk) : 0.f;
            a1[k] = (top && rig) ? __ldg(gk + 1) : 0.f;
            a2[k] = (bot && lef) ? __ldg(gk + Wi) : 0.f;
            a3[k] = (bot && rig) ? __ldg(gk + Wi + 1) : 0.f;
          }
          q0 = make_float4(a0[0], a0[1], a0[2], a0[3]);
          q1 = make_float4(a1[0], a1[1], a1[2], a1[3]);
          q2 = make_float4(a2[0], a2[1], a2[2], a2[3]);
          q3 = make_float4(a3[0], a3[1], a3[2], a3[3]);
        }
        acc.x = fmaf(q3.x, wse, fmaf(q2.x, wsw, fmaf(q1.x, wne, q0.x * wnw)));
        acc.y = fmaf(q3.y, wse, fmaf(q2.y, wsw, fmaf(q1.y, wne, q0.y * wnw)));
        acc.z = fmaf(q3.z, wse, fmaf(q2.z, wsw, fmaf(q1.z, wne, q0.z * wnw)));
        acc.w = fmaf(q3.w, wse, fmaf(q2.w, wsw, fmaf(q1.w, wne, q0.w * wnw)));
      }
      const int c = c0 + 4 * q;
      if (MODE == WT_NHWC) {
        if (pv < Ho && pu < Wo) st_stream4(prm.dst + (((int64_t)n * Ho + pv) * Wo + pu) * C + c, acc);
      } else {
        const int vp = pv + 1, up = pu + 1;  // padded coordinates, >= 0
        if (vp <= Ho + 1 && up <= Wo + 1) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int a = vp - ky;
            if (a < 0 || (a % S) != 0) continue;
            const int oy = a / S;
            if (oy >= prm.Ho2) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const int b = up - kx;
              if (b < 0 || (b % S) != 0) continue;
              const int ox = b / S;
              if (ox >= prm.Wo2) continue;
              st_stream4(prm.dst + ((((int64_t)n * prm.Ho2 + oy) * prm.Wo2 + ox) * 9 + ky * 3 + kx) * C + c, acc);
            }
          }
        }
      }
    }
    if (staged && ch + 1 < nchunks) __syncthreads();  // everyone is done with `tr` before the next pass overwrites it
  }
}

int encode_src_map(CUtensorMap* map, const float* src, int BN, int C, int Hi, int Wi, int bw) {
  const cuuint64_t u = 1;
  const cuuint64_t gdim[4] = {u * Wi, u * Hi, u * C, u * BN};
  const cuuint64_t gstr[3] = {u * Wi * 4, u * Hi * Wi * 4, u * C * Hi * Wi * 4};
  const cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)kWtRows, (cuuint32_t)kWtCC, 1u};
  return encode(map, src, 4, gdim, gstr, box);
}

template <int MODE, int TH, int TW, int S>
int launch_wt(const CUtensorMap* maps, WtParams prm, int BN, int dom_h, int dom_w, cudaStream_t st) {
  auto kern = warp_tma_kernel<MODE, TH, TW, S>;
  const size_t smem = (size_t)2 * kWtStageFloats * 4 + (MODE == WT_NCHW ? 0 : (size_t)kWtCC * kWtPitch * 4);
  MVD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prm.tiles_x = (dom_w + TW - 1) / TW;
  const int tiles_y = (dom_h + TH - 1) / TH;
  dim3 grid((unsigned)(prm.tiles_x * tiles_y), (unsigned)BN);
  kern<<<grid, kWtThreads, smem, st>>>(maps[0], maps[1], maps[2], prm);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

template <int MODE, int S>
int launch_wc(const float* src, int BN, int C, int Hi, int Wi, WtParams prm, int dom_h, int dom_w, cudaStream_t st) {
  WcMaps maps;
  const int bws[kWcBoxes] = {8, 16, 24, 32, 48, 64};
  for (int i = 0; i < kWcBoxes; ++i) {
    const cuuint64_t u = 1;
    const cuuint64_t gdim[4] = {u * Wi, u * Hi, u * C, u * BN};
    const cuuint64_t gstr[3] = {u * Wi * 4, u * Hi * Wi * 4, u * C * Hi * Wi * 4};
    const cuuint32_t box[4] = {(cuuint32_t)bws[i], (cuuint32_t)kWtRows, (cuuint32_t)kWcBoxCh, 1u};
    if (int e = encode(&maps.m[i], src, 4, gdim, gstr, box)) return e;
  }
  auto kern = warp_tma_cl_kernel<MODE, S>;
  const size_t smem = (size_t)(2 * kWcRaw + kWcTr) * 4 + (size_t)kWtThreads * 5 * 4;
  MVD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prm.tiles_x = (dom_w + 15) / 16;
  const int tiles_y = (dom_h + 15) / 16;
  dim3 grid((unsigned)(prm.tiles_x * tiles_y), (unsigned)BN);
  kern<<<grid, kWtThreads, smem, st>>>(maps, prm);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

}  // namespace
}  // namespace mvd

using namespace mvd;

// mode: 0 = dst [BN,C,Ho,Wo], 1 = dst [BN,Ho,Wo,C], 2 = dst = im2col matrix [BN*Ho2*Wo2, 9*C] of a 3x3 / `stride` / pad-1
// convolution (stride 1 or 2). Source is NCHW. MVD_ERR_UNSUPPORTED unless C % 32 == 0 and Wi % 4 == 0.
extern "C" int mvd_warp_tma_f32(const float* src, const float* Mat, int BN, int C, int Hi, int Wi, int Ho, int Wo,
                                float* dst, int mode, int stride, void* stream) {
  if (!src || !Mat || !dst) return MVD_ERR_NULL_POINTER;
  if (BN <= 0 || C <= 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0 || BN > 65535) return MVD_ERR_BAD_SHAPE;
  if ((int64_t)Hi * Wi > 0x3fffffffLL || (int64_t)Ho * Wo > 0x3fffffffLL) return MVD_ERR_BAD_SHAPE;
  if (mode < 0 || mode > 2) return MVD_ERR_UNSUPPORTED;
  if (mode == 2 && stride != 1 && stride != 2) return MVD_ERR_UNSUPPORTED;
  if (C % kWtCC != 0 || Wi % 4 != 0) return MVD_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) return MVD_ERR_MISALIGNED;
  cudaStream_t st = (cudaStream_t)stream;
  WtParams prm;
  prm.src = src;
  prm.Mat = Mat;
  prm.dst = dst;
  prm.C = C;
  prm.Hi = Hi;
  prm.Wi = Wi;
  prm.Ho = Ho;
  prm.Wo = Wo;
  prm.tiles_x = 0;
  prm.Ho2 = prm.Wo2 = 0;
  if (mode == 0) {
    alignas(64) CUtensorMap maps[3];
    const int bws[3] = {16, 32, 64};
    for (int i = 0; i < 3; ++i)
      if (int e = encode_src_map(&maps[i], src, BN, C, Hi, Wi, bws[i])) return e;
    return launch_wt<WT_NCHW, 8, 32, 1>(maps, prm, BN, Ho, Wo, st);
  }
  if (mode == 1) return launch_wc<WT_NHWC, 1>(src, BN, C, Hi, Wi, prm, Ho, Wo, st);
  prm.Ho2 = (Ho + 2 - 3) / stride + 1;
  prm.Wo2 = (Wo + 2 - 3) / stride + 1;
  if ((int64_t)BN * prm.Ho2 * prm.Wo2 * 9 * C > 0x7fffffffffLL) return MVD_ERR_BAD_SHAPE;
  if (stride == 2) return launch_wc<WT_IM2COL, 2>(src, BN, C, Hi, Wi, prm, Ho + 2, Wo + 2, st);
  return launch_wc<WT_IM2COL, 1>(src, BN, C, Hi, Wi, prm, Ho + 2, Wo + 2, st);
}
