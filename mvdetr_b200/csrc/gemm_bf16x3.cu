// Dense glue of the path on the 5th-generation tensor cores with fp32-level accuracy, as ONE persistent kernel per layer:
//   out[rows, N] = act(x[rows, K] @ W[N, K]^T + bias[N])
//   replaces the nn.Linear calls of an encoder layer    ref: multiview_detector/models/ops/modules/ms_deform_attn.py:96,100-101,116,
//                                                        multiview_detector/models/deformable_transformer.py:82
//   and the 3x3 convolutions (implicit GEMM: taps fetched by TMA from the channels-last map, or a materialised im2col
//   matrix) and the 1x1 merge                            ref: multiview_detector/models/trans_world_feat.py:74,82-84,107-109
//
// Why ours: round 1 used cuBLASLt's CUBLAS_COMPUTE_32F_EMULATED_16BFX9 -- 9 bf16 products, a bias pass and a separate
// Inf/NaN operand-scan kernel per call (profiles/r02c_timeline.txt: GEMM 1.34 ms + scan/patch 0.8 ms of a 3.2 ms frame).
// The layers are tall-skinny (75 600 x 128..1152 by 128..672; ~10 flop/byte).
//
// Two kernels in this file:
//  * linear_split_ts_kernel<NT> (DEFAULT, second half of the file): x terms staged in TENSOR MEMORY, NT = 2 fp16 terms /
//    3 products (default) or NT = 3 bf16 terms / 6 products; own producer warps for x and W, converged issue warps
//    under elect.sync, conv mode, optional NVLink-multicast epilogue. Its header comment has the structure.
//  * linear_bf16x3_kernel (below; MVDETR_B200_GEMM=bf16x3 with MVDETR_B200_GEMM_TS=0, mode "bf16x3ss"): the round-2
//    first version with both operands in shared memory, kept as the bit-identical comparator of the NT = 3 variant.
//
// Arithmetic of the bf16 variants: fp32 operands are split into three bf16 terms (a = a0 + a1 + a2, each the bf16 rounding
// of the remaining residual; 24 mantissa bits covered) and the SIX products with i + j <= 2 are accumulated in fp32 in
// tensor memory (scripts/emulate_bf16_split.py: the dropped three are below 2^-24 |a||b|). bf16 products are exact inside
// the tensor core's adder (16-bit significands), which a 3xTF32 split is not (22-bit products: csrc/gemm_tf32.cu measures
// 5-10x larger error).
//
// Structure of linear_bf16x3_kernel (1 CTA per SM, persistent over 128 x 128 output tiles, 320 threads, warp-specialised):
//   warp 0      TMA producer: per 32-wide K chunk the x tile [128 x 32] fp32 (SWIZZLE_128B) and the three pre-split
//               weight tiles [128 x 32] bf16 (SWIZZLE_64B; rows past `rows` / `N`, columns past K are zero-filled);
//   warps 2-5   split the x tile into three bf16 tiles in the UMMA K-major SWIZZLE_64B layout (thread = row: conflict-
//               free 16-byte reads of the 128B-swizzled source and 16-byte writes of the 64B-swizzled destinations);
//   warp 1      converged; an elected lane issues 2 k-steps x 6 tcgen05.mma.kind::f16 (M 128, N 128, K 16) per chunk,
//               commits the stages back to their producers, and after the last chunk the accumulator to the epilogue;
//   warps 6-9   epilogue: tcgen05.ld of both accumulators, sum + bias, ReLU, 32 x 32 sub-tiles transposed through shared
//               memory into full 128-byte row segments, overlapping the next tile's main loop (2 x 2 x 128 accumulator
//               columns in tensor memory).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "vg_common.cuh"

namespace mvd {
namespace {

constexpr int kBM = 128, kBN = 128, kBK = 32;
constexpr int kBThreads = 320;
constexpr int kRawBytes = kBM * kBK * 4;      // 16 KB fp32 x tile
constexpr int kHalfBytes = kBM * kBK * 2;     // 8 KB bf16 tile (x split or weight term)
// three decoupled rings: raw fp32 x tiles, weight-term triples, split x-term triples (+ the epilogue warps' stages)
constexpr int kRawStages = 4, kWStages = 3, kTStages = 2;
constexpr int kOffRaw = 0;
constexpr int kOffW = kOffRaw + kRawStages * kRawBytes;        //  64 KB
constexpr int kOffT = kOffW + kWStages * 3 * kHalfBytes;       // +72 KB
constexpr int kOffStage = kOffT + kTStages * 3 * kHalfBytes;   // +48 KB
constexpr int kStagePitch = 36;                                // floats per row of an epilogue warp's 32 x 32 stage
constexpr int kBSmemUsed = kOffStage + 4 * 32 * kStagePitch * 4;  // +18 KB = 202 KB
constexpr int kBSmem = kBSmemUsed + 1024;                      // + slack for 1024-byte alignment
constexpr uint32_t kBTmemCols = 512;          // 2 buffers x {leading product, small terms} x 128 fp32 columns
constexpr int kNumBars = 2 * kRawStages + 2 * kWStages + 2 * kTStages + 4;

struct GbParams {
  const float* bias;  // nullable
  float* out;
  int rows, K, N, relu;
  int m_blocks, n_tiles;
  int multicast;  // != 0: `out` is an NVLink multicast address (multimem.st: the NVSwitch replicates every store to all GPUs)
  int raw_stages, w_stages;  // TS kernels: depth of the fp32 x-tile ring and of the weight-term ring
  // TS kernels, conv != 0: x is a channels-last image [NB][Hi][Wi][C] and the GEMM is the 3x3 / pad 1 / stride cstride
  // convolution over it (K = 9 C, taps (ky, kx, c)): a row block is a TY x TX tile of output pixels of one image and a K
  // chunk is 32 channels of one tap, fetched by TMA straight from the image (no im2col matrix in memory)
  int conv, cHo, cWo, TX, TY, tiles_x, tiles_y, cchunks, cstride;
  // TS kernels, optional second output: out2 = out + addend (same [rows, N] layout): the first encoder layer's query
  // `src + pos` leaves the downsample convolution's epilogue instead of an element-wise kernel of its own
  const float* addend;
  float* out2;
};

__device__ __forceinline__ void tma_load_2d_b(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// UMMA shared-memory descriptor, K-major tile of 64-byte rows (SWIZZLE_64B, layout code 4): 8-row groups 512 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_k64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)(512u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

// D fp32 (c_format 1), A and B bf16 (format 1), both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t kIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBN >> 3) << 17) |
                                ((uint32_t)(kBM >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdescBf16), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {  // arrives when every MMA issued so far has retired
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// three-term bf16 split of 2 floats: packed conversions (one cvt.rn.bf16x2.f32 per term and PAIR), the bf16 values are
// widened back with a shift / mask (exact), the residuals are exact fp32 subtractions
__device__ __forceinline__ void split2(float a, float b, uint32_t& t0, uint32_t& t1, uint32_t& t2) {
  t0 = pack_bf16(a, b);  // low half = bf16(a), high half = bf16(b)
  const float ra = a - __uint_as_float(t0 << 16), rb = b - __uint_as_float(t0 & 0xffff0000u);
  t1 = pack_bf16(ra, rb);
  t2 = pack_bf16(ra - __uint_as_float(t1 << 16), rb - __uint_as_float(t1 & 0xffff0000u));
}

// 8 consecutive floats -> one 16-byte piece per term
__device__ __forceinline__ void split8(const float4& u, const float4& v, uint4& t0, uint4& t1, uint4& t2) {
  split2(u.x, u.y, t0.x, t1.x, t2.x);
  split2(u.z, u.w, t0.y, t1.y, t2.y);
  split2(v.x, v.y, t0.z, t1.z, t2.z);
  split2(v.z, v.w, t0.w, t1.w, t2.w);
}

// elect_one() (vg_common.cuh): one leader lane of a CONVERGED warp. r02p (ncu source page): with the whole issue loop under `if (lane == 0)` the
// compiler could not prove the operands of the uniform-datapath instructions (UTCHMMA, UTCBAR, UTMALDG) warp-uniform and
// wrapped every one in an ELECT / R2UR / BRA.U.ANY waterfall; the MMA warp spent two thirds of its time issuing and the
// tensor pipe idled at 36 %. The loops now run on all 32 lanes and only the asynchronous instructions sit under elect.
__global__ void __launch_bounds__(kBThreads, 1)
    linear_bf16x3_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                         const GbParams prm) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) unsigned long long s_bar[kNumBars];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[kBN];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (base - smem_u32(smem_dyn));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bars = smem_u32(&s_bar[0]);
  int nb = 0;
  auto take = [&](int n) { const uint32_t r = bars + 8u * (uint32_t)nb; nb += n; return r; };
  const uint32_t b_raw_full = take(kRawStages), b_raw_empty = take(kRawStages);   // TMA -> splitters -> TMA
  const uint32_t b_w_full = take(kWStages), b_w_empty = take(kWStages);           // TMA -> MMA -> TMA
  const uint32_t b_t_full = take(kTStages), b_t_empty = take(kTStages);           // splitters -> MMA -> splitters
  const uint32_t b_acc_full = take(2), b_acc_empty = take(2);                     // MMA -> epilogue -> MMA

  if (tid == 0) {
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_b);
    for (int i = 0; i < kRawStages; ++i) {
      mbar_init(b_raw_full + 8u * i, 1);
      mbar_init(b_raw_empty + 8u * i, 128);
    }
    for (int i = 0; i < kWStages; ++i) {
      mbar_init(b_w_full + 8u * i, 1);
      mbar_init(b_w_empty + 8u * i, 1);
    }
    for (int i = 0; i < kTStages; ++i) {
      mbar_init(b_t_full + 8u * i, 128);
      mbar_init(b_t_empty + 8u * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(b_acc_full + 8u * i, 1);
      mbar_init(b_acc_empty + 8u * i, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "r"(kBTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  const int nk = (prm.K + kBK - 1) / kBK;
  const int ntiles = prm.m_blocks * prm.n_tiles;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer (converged warp, elected lane issues) ------------
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int m0 = (tile / prm.n_tiles) * kBM, n0 = (tile % prm.n_tiles) * kBN;
      for (int kc = 0; kc < nk; ++kc, ++it) {
        const uint32_t rs = it % kRawStages, ws = it % kWStages;
        mbar_wait(b_raw_empty + 8u * rs, ((it / kRawStages) & 1u) ^ 1u);  // first lap: passes at once
        mbar_wait(b_w_empty + 8u * ws, ((it / kWStages) & 1u) ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(b_raw_full + 8u * rs, (uint32_t)kRawBytes);
          tma_load_2d_b(base + kOffRaw + rs * kRawBytes, &tm_a, b_raw_full + 8u * rs, kc * kBK, m0);
          mbar_expect_tx(b_w_full + 8u * ws, (uint32_t)(3 * kHalfBytes));
#pragma unroll
          for (int i = 0; i < 3; ++i)
            tma_load_3d(base + kOffW + (ws * 3 + i) * kHalfBytes, &tm_b, b_w_full + 8u * ws, kc * kBK, n0, i);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (converged warp, elected lane issues) --------------
    uint32_t it = 0, tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const uint32_t b = tcount & 1u, tph = (tcount >> 1) & 1u;
      mbar_wait(b_acc_empty + 8u * b, tph ^ 1u);  // the epilogue has drained this accumulator (first two tiles: at once)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // two accumulators per tile: the tensor core's fp32 accumulation truncates (~0.5 ulp of the ACCUMULATOR per
      // instruction, r02e: error grew with the number of MMAs, 6 per k-step). The leading product a0*b0 gets its own
      // accumulator (K/16 additions); the five small products (2^-8 and below) go to a second one whose rounding is
      // 2^-8 smaller. The epilogue adds the two in fp32.
      const uint32_t acc_hi = tmem + b * (uint32_t)(2 * kBN), acc_lo = acc_hi + (uint32_t)kBN;
      for (int kc = 0; kc < nk; ++kc, ++it) {
        const uint32_t ws = it % kWStages, ts = it % kTStages;
        mbar_wait(b_w_full + 8u * ws, (it / kWStages) & 1u);  // weight terms landed
        mbar_wait(b_t_full + 8u * ts, (it / kTStages) & 1u);  // x terms written by all 128 splitter threads
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a0 = base + kOffT + ts * 3 * kHalfBytes, b0 = base + kOffW + ws * 3 * kHalfBytes;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {  // UMMA K = 16 bf16 = 32 bytes inside the 64-byte swizzle row
            const uint32_t ko = 32u * (uint32_t)k;
            uint64_t da[3], db[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
              da[i] = umma_desc_k64(a0 + (uint32_t)i * kHalfBytes + ko);
              db[i] = umma_desc_k64(b0 + (uint32_t)i * kHalfBytes + ko);
            }
            // smallest terms first: (i, j) with i + j = 2, then 1; the leading product into its own accumulator
            umma_bf16(acc_lo, da[0], db[2], (kc | k) != 0);
            umma_bf16(acc_lo, da[1], db[1], 1u);
            umma_bf16(acc_lo, da[2], db[0], 1u);
            umma_bf16(acc_lo, da[0], db[1], 1u);
            umma_bf16(acc_lo, da[1], db[0], 1u);
            umma_bf16(acc_hi, da[0], db[0], (kc | k) != 0);
          }
          umma_commit(b_w_empty + 8u * ws);  // both operand stages reusable once these MMAs have read them
          umma_commit(b_t_empty + 8u * ts);
          if (kc == nk - 1) umma_commit(b_acc_full + 8u * b);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------ x-tile splitter (128 threads, thread = row) ------------------
    const int r = tid - 64;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int kc = 0; kc < nk; ++kc, ++it) {
        const uint32_t rs = it % kRawStages, ts = it % kTStages;
        mbar_wait(b_raw_full + 8u * rs, (it / kRawStages) & 1u);          // fp32 tile landed
        mbar_wait(b_t_empty + 8u * ts, ((it / kTStages) & 1u) ^ 1u);      // the MMAs that read this term stage retired
        const unsigned char* src = sm + kOffRaw + (size_t)rs * kRawBytes + (size_t)r * 128;  // 128-byte row, pieces XOR (r & 7)
        unsigned char* dst = sm + kOffT + (size_t)ts * 3 * kHalfBytes + (size_t)r * 64;       // 64-byte rows, XOR ((r >> 1) & 3)
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 8 floats -> one 16-byte bf16 piece per term
          const float4 u = *reinterpret_cast<const float4*>(src + (((2 * j) ^ (r & 7)) << 4));
          const float4 v = *reinterpret_cast<const float4*>(src + (((2 * j + 1) ^ (r & 7)) << 4));
          uint4 t0, t1, t2;
          split8(u, v, t0, t1, t2);
          const int pc = (j ^ ((r >> 1) & 3)) << 4;
          *reinterpret_cast<uint4*>(dst + pc) = t0;
          *reinterpret_cast<uint4*>(dst + kHalfBytes + pc) = t1;
          *reinterpret_cast<uint4*>(dst + 2 * kHalfBytes + pc) = t2;
        }
        mbar_arrive(b_raw_empty + 8u * rs);                            // raw tile consumed (generic reads only)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core (async proxy) reads
        mbar_arrive(b_t_full + 8u * ts);
      }
    }
  } else {
    // ------------------------------------------------ epilogue (warps 6-9: TMEM lane quarters 2, 3, 0, 1) -----------
    const int q = warp & 3;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const int m0 = (tile / prm.n_tiles) * kBM, n0 = (tile % prm.n_tiles) * kBN;
      const uint32_t b = tcount & 1u, tph = (tcount >> 1) & 1u;
      mbar_wait(b_acc_full + 8u * b, tph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // bias of this column tile -> shared memory once (every thread needs all 128 values of its row)
      asm volatile("bar.sync 1, 128;" ::: "memory");  // previous tile's readers are done with s_bias
      {
        const int et = tid - 192, n = n0 + et;
        s_bias[et] = (prm.bias && n < prm.N) ? __ldg(prm.bias + n) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const uint32_t t_hi = tmem + b * (uint32_t)(2 * kBN) + ((uint32_t)(q * 32) << 16), t_lo = t_hi + (uint32_t)kBN;
      // TMEM hands every thread one ROW (lane) of the tile; stored as such, a warp instruction would write 32 separate
      // 16-byte pieces 512 bytes apart (r02h: as multimem.st these are 16-byte NVLink packets, and the 2-GPU step got
      // slower than with ncclAllGather). The warp's 32 x 32 sub-tile is transposed through a private shared-memory
      // stage instead, so that 8 lanes write one 128-byte row segment: 4 full lines per store instruction.
      float* stg = reinterpret_cast<float*>(sm + kOffStage) + (size_t)q * 32 * kStagePitch;
      const int r0 = m0 + q * 32;
#pragma unroll 1
      for (int c = 0; c < kBN; c += 32) {
        uint32_t v[32], w[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(t_hi + (uint32_t)c));
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
              "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]),
              "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23]),
              "=r"(w[24]), "=r"(w[25]), "=r"(w[26]), "=r"(w[27]), "=r"(w[28]), "=r"(w[29]), "=r"(w[30]), "=r"(w[31])
            : "r"(t_lo + (uint32_t)c));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        __syncwarp();  // the previous chunk's readers are done with the stage
#pragma unroll
        for (int j = 0; j < 32; j += 4) {  // own row, 4 columns at a time: sum of the two accumulators + bias, ReLU
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + c + j);
          float4 o;
          o.x = (__uint_as_float(v[j]) + __uint_as_float(w[j])) + bb.x;
          o.y = (__uint_as_float(v[j + 1]) + __uint_as_float(w[j + 1])) + bb.y;
          o.z = (__uint_as_float(v[j + 2]) + __uint_as_float(w[j + 2])) + bb.z;
          o.w = (__uint_as_float(v[j + 3]) + __uint_as_float(w[j + 3])) + bb.w;
          if (prm.relu) {
            o.x = fmaxf(o.x, 0.f);
            o.y = fmaxf(o.y, 0.f);
            o.z = fmaxf(o.z, 0.f);
            o.w = fmaxf(o.w, 0.f);
          }
          *reinterpret_cast<float4*>(stg + lane * kStagePitch + j) = o;  // pitch 36 floats: conflict-free both ways
        }
        __syncwarp();
        const int piece = lane & 7, n = n0 + c + piece * 4;
        if (n < prm.N) {  // N % 4 == 0: a 4-column piece is inside or outside as a whole
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = (lane >> 3) + 4 * i;
            if (r0 + rr < prm.rows) {
              const float4 o = *reinterpret_cast<const float4*>(stg + rr * kStagePitch + piece * 4);
              float* dst = prm.out + (int64_t)(r0 + rr) * prm.N + n;
              if (prm.multicast)  // fused all-gather: one store, delivered to this row's slot on every GPU of the group
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(o.x),
                             "f"(o.y), "f"(o.z), "f"(o.w)
                             : "memory");
              else
                *reinterpret_cast<float4*>(dst) = o;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(b_acc_empty + 8u * b);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kBTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// TS variant: the x terms live in TENSOR MEMORY instead of shared memory.
// r02i: the kernel above is bound by the shared-memory pipe, not by HBM or the tensor cores -- per 32-wide K chunk it
// moves 176 KB through shared memory (TMA fill 40, split read 16 + write 24, UMMA operand reads 48 (x) + 48 (W)), 0.87 us
// measured against 0.40 us of tensor time. Here the splitters write their three bf16 terms with tcgen05.st into tensor
// memory (lane = row, one 32-bit column = two consecutive k) and the MMAs take A from there ("TS" form of tcgen05.mma):
// 104 KB per chunk (72 KB with two terms). When the K chunks of a row block fit the term stages (NT = 2: K <= 256, 8 stages
// x 32 columns; NT = 3: K <= 160) the terms stay in tensor memory and are reused by every 128-column tile of the output
// (N = 512 / 672 layers: x is read from HBM and split ONCE per row block).
//   tensor memory: columns 0-127 leading-product accumulator, 128-255 small-terms accumulator, 256-511 x-term stages
//   (one accumulator buffer: the 8 epilogue warps first drain it into registers and hand it back, then do bias / ReLU /
//   stores under the next tile's main loop)
//   warps: 0 x-tile TMA, 1 MMA, 2-5 split (TMEM lane quarter = warp & 3), 6-13 epilogue (quarter = warp & 3, column
//   half), 14 weight-term TMA
constexpr int kTsThreads = 480;  // warps: 0 x-tile TMA, 1 MMA, 2-5 split, 6-13 epilogue, 14 weight-term TMA
constexpr uint32_t kTsAccCols = 2 * kBN;      // leading + small-terms accumulators
constexpr uint32_t kTsTmemCols = 512;
// r02n: with the x and W loads issued by ONE thread the W ring (4 stages) capped the x prefetch at 4 tiles = 64 KB per SM,
// 9.5 MB in flight on the chip, and the big-K layers ran at half the HBM rate (0.68 us per chunk). The two streams now have
// their own producer warps and the x ring takes all the shared memory that is left.
template <int NT>
struct TsCfg {
  static constexpr int kRaw = NT == 2 ? 5 : 5;               // default depth of the fp32 x-tile ring (16 KB each)
  static constexpr int kW = NT == 2 ? 6 : 4;                 // default depth of the weight-term ring (NT x 8 KB each)
  static constexpr int kMaxRing = 12;
  static constexpr int kA = NT == 2 ? 8 : 5;                 // x-term stages in tensor memory
  static constexpr uint32_t kAStageCols = NT * 16;           // NT terms x 16 columns (32 k of one row per lane)
  static constexpr int kWBytes = NT * kHalfBytes;
  static constexpr int kStageBytes = 8 * 32 * kStagePitch * 4;  // epilogue transposition stages
  static constexpr int smem(int raw, int w) { return raw * kRawBytes + w * kWBytes + kStageBytes + 1024; }
  static constexpr int kNumBars = 4 * kMaxRing + 2 * kA + 2;
  static_assert(kTsAccCols + kA * kAStageCols <= kTsTmemCols, "tensor memory");
  static_assert(smem(kRaw, kW) <= 227 * 1024, "shared memory");
};

// fp16 operands (format 0) instead of bf16 (format 1): the two-term variant below
constexpr uint32_t kIdescF16 = (1u << 4) | ((uint32_t)(kBN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);

template <uint32_t IDESC>
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "n"(IDESC), "r"(accumulate)
      : "memory");
}

// Two-term fp16 split of 2 floats, second term scaled by 2^11: a = h0 + 2^-11 h1 with |a - (h0 + 2^-11 h1)| <= 2^-24 |a|
// (11 significant bits per term; a - h0 and the scaling are exact in fp32). fp16's range applies: |a| must stay below
// 65504 (larger inputs give non-finite results, never silently wrong ones); below 6e-5 the absolute error is <= 2^-25.
constexpr float kF16Scale = 2048.f;
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t& t0, uint32_t& t1) {
  const __half2 h = __floats2half2_rn(a, b);  // low half = fp16(a), high half = fp16(b)
  const float2 f = __half22float2(h);
  const __half2 g = __floats2half2_rn((a - f.x) * kF16Scale, (b - f.y) * kF16Scale);
  t0 = *reinterpret_cast<const uint32_t*>(&h);
  t1 = *reinterpret_cast<const uint32_t*>(&g);
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}

// NT = 3: three bf16 terms, six products (full fp32 range). NT = 2: two fp16 terms (second one scaled by 2^11), THREE
// products -- half the tensor work (r02m: with 6 products the big-K layers run at ~70 % of the measured bf16 tensor peak,
// the tensor cores are the bound); the small-terms accumulator is multiplied by 2^-11 in the epilogue.
template <int NT>
__global__ void __launch_bounds__(kTsThreads, 1)
    linear_split_ts_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                           const GbParams prm) {
  constexpr uint32_t IDESC = NT == 3 ? kIdescBf16 : kIdescF16;
  using Cfg = TsCfg<NT>;
  constexpr int kTsAStages = Cfg::kA, kWBytes = Cfg::kWBytes;
  constexpr uint32_t kTsAStageCols = Cfg::kAStageCols;
  const int kTsRawStages = prm.raw_stages, kTsWStages = prm.w_stages;  // ring depths chosen by the host (shared memory split)
  const int kTsOffRaw = 0, kTsOffW = kTsRawStages * kRawBytes, kTsOffStage = kTsOffW + kTsWStages * kWBytes;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) unsigned long long s_bar[TsCfg<NT>::kNumBars];
  __shared__ uint32_t s_tmem;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (base - smem_u32(smem_dyn));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bars = smem_u32(&s_bar[0]);
  int nb = 0;
  auto take = [&](int n) { const uint32_t r = bars + 8u * (uint32_t)nb; nb += n; return r; };
  const uint32_t b_raw_full = take(kTsRawStages), b_raw_empty = take(kTsRawStages);  // TMA -> splitters -> TMA
  const uint32_t b_w_full = take(kTsWStages), b_w_empty = take(kTsWStages);          // TMA -> MMA -> TMA
  const uint32_t b_a_full = take(kTsAStages), b_a_empty = take(kTsAStages);          // splitters -> MMA -> splitters
  const uint32_t b_acc_full = take(1), b_acc_empty = take(1);                        // MMA -> epilogue -> MMA

  // prologue in parallel: warp 0 allocates tensor memory while the threads of warps 1-3 initialise one barrier each
  // (one thread doing 40 mbarrier.init in a row ahead of the allocation cost ~1 us of every launch: 21 launches a frame)
  if (tid == 32) {
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_b);
  }
  if (tid >= 32 && tid < 32 + nb) {
    const int j = tid - 32, R = kTsRawStages, W = kTsWStages, A = kTsAStages;
    uint32_t count;
    if (j < R) count = 1;                      // raw_full: the TMA transaction
    else if (j < 2 * R) count = 128;           // raw_empty: every splitter thread
    else if (j < 2 * R + 2 * W) count = 1;     // w_full (TMA) / w_empty (tcgen05.commit)
    else if (j < 2 * R + 2 * W + A) count = 128;     // a_full: every splitter thread
    else if (j < 2 * R + 2 * W + 2 * A) count = 1;   // a_empty: tcgen05.commit
    else if (j == 2 * R + 2 * W + 2 * A) count = 1;  // acc_full: tcgen05.commit
    else count = 256;                          // acc_empty: every epilogue thread
    mbar_init(bars + 8u * (uint32_t)j, count);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "r"(kTsTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  const int nk = (prm.K + kBK - 1) / kBK;
  // K <= 128: the x terms of a whole row block fit the four tensor-memory stages and serve every column tile
  const bool reuse = nk <= kTsAStages;
  const int a_loads = reuse ? nk : prm.n_tiles * nk;  // x chunks loaded and split per row block

  if (warp == 0) {
    // ------------------------------------------------ TMA producer: fp32 x tiles (converged warp, elected lane) -----
    uint32_t rs = 0, rph = 0;  // ring stage, lap parity
    const uint32_t raw_tx = prm.conv ? (uint32_t)(prm.TX * prm.TY * kBK * 4) : (uint32_t)kRawBytes;
    for (int mb = blockIdx.x; mb < prm.m_blocks; mb += gridDim.x) {
      const int m0 = mb * kBM;
      int img = 0, xb = 0, yb = 0;  // conv: image, source coordinates of the tile's first output pixel for tap (0, 0)
      if (prm.conv) {
        const int tpi = prm.tiles_x * prm.tiles_y;
        img = mb / tpi;
        const int t = mb - img * tpi, ty = t / prm.tiles_x, tx = t - ty * prm.tiles_x;
        xb = tx * prm.TX * prm.cstride - 1;
        yb = ty * prm.TY * prm.cstride - 1;
      }
      for (int l = 0, kc = 0; l < a_loads; ++l) {
        mbar_wait(b_raw_empty + 8u * rs, rph ^ 1u);  // first lap: passes at once
        if (elect_one()) {
          mbar_expect_tx(b_raw_full + 8u * rs, raw_tx);
          if (prm.conv) {  // 32 channels of tap (ky, kx): rows of the stage = output pixels (y, x) of the tile, zero padding
            const int tap = kc / prm.cchunks, cc = kc - tap * prm.cchunks, ky = tap / 3, kx = tap - 3 * ky;
            tma_load_4d(base + kTsOffRaw + rs * kRawBytes, &tm_a, b_raw_full + 8u * rs, cc * kBK, xb + kx, yb + ky, img);
          } else {
            tma_load_2d_b(base + kTsOffRaw + rs * kRawBytes, &tm_a, b_raw_full + 8u * rs, kc * kBK, m0);
          }
        }
        __syncwarp();
        if (++kc == nk) kc = 0;
        if (++rs == (uint32_t)kTsRawStages) rs = 0, rph ^= 1u;
      }
    }
  } else if (warp == 14) {
    // ------------------------------------------------ TMA producer: weight terms (one chunk per column tile and k) --
    uint32_t ws = 0, wph = 0;
    for (int mb = blockIdx.x; mb < prm.m_blocks; mb += gridDim.x) {
      for (int nt = 0; nt < prm.n_tiles; ++nt) {
        const int n0 = nt * kBN;
        for (int kc = 0; kc < nk; ++kc) {
          mbar_wait(b_w_empty + 8u * ws, wph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(b_w_full + 8u * ws, (uint32_t)kWBytes);
#pragma unroll
            for (int i = 0; i < NT; ++i)
              tma_load_3d(base + kTsOffW + ws * kWBytes + i * kHalfBytes, &tm_b, b_w_full + 8u * ws, kc * kBK, n0, i);
          }
          __syncwarp();
          if (++ws == (uint32_t)kTsWStages) ws = 0, wph ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer (converged warp, one elected lane issues) ---------
    uint32_t ia_base = 0, ws = 0, wph = 0, tcount = 0;
    const uint32_t acc_hi = tmem, acc_lo = tmem + (uint32_t)kBN;
    for (int mb = blockIdx.x; mb < prm.m_blocks; mb += gridDim.x) {
      for (int nt = 0; nt < prm.n_tiles; ++nt, ++tcount) {
        mbar_wait(b_acc_empty, (tcount & 1u) ^ 1u);  // the epilogue holds the previous tile in registers (first tile: at once)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int kc = 0; kc < nk; ++kc) {
          const uint32_t a_idx = ia_base + (uint32_t)(reuse ? kc : nt * nk + kc);
          const uint32_t as = a_idx % kTsAStages;
          if (!reuse || nt == 0) mbar_wait(b_a_full + 8u * as, (a_idx / kTsAStages) & 1u);  // x terms stored by 128 threads
          mbar_wait(b_w_full + 8u * ws, wph);                                                 // weight terms landed
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a0 = tmem + kTsAccCols + as * kTsAStageCols, b0 = base + kTsOffW + ws * kWBytes;
          const bool release_a = !reuse || nt == prm.n_tiles - 1;  // last column tile: x stage reusable
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {  // UMMA K = 16 bf16: 8 tensor-memory columns of A, 32 bytes of a W row
              uint32_t ta[NT];
              uint64_t db[NT];
#pragma unroll
              for (int i = 0; i < NT; ++i) {
                ta[i] = a0 + 16u * (uint32_t)i + 8u * (uint32_t)k;
                db[i] = umma_desc_k64(b0 + (uint32_t)i * kHalfBytes + 32u * (uint32_t)k);
              }
              if constexpr (NT == 3) {
                // same order and accumulators as the shared-memory kernel above: results are bit-identical
                umma_ts<IDESC>(acc_lo, ta[0], db[2], (kc | k) != 0);
                umma_ts<IDESC>(acc_lo, ta[1], db[1], 1u);
                umma_ts<IDESC>(acc_lo, ta[2], db[0], 1u);
                umma_ts<IDESC>(acc_lo, ta[0], db[1], 1u);
                umma_ts<IDESC>(acc_lo, ta[1], db[0], 1u);
              } else {
                umma_ts<IDESC>(acc_lo, ta[0], db[1], (kc | k) != 0);  // both cross products carry the factor 2^11
                umma_ts<IDESC>(acc_lo, ta[1], db[0], 1u);
              }
              umma_ts<IDESC>(acc_hi, ta[0], db[0], (kc | k) != 0);
            }
            umma_commit(b_w_empty + 8u * ws);
            if (release_a) umma_commit(b_a_empty + 8u * as);
            if (kc == nk - 1) umma_commit(b_acc_full);  // accumulator complete
          }
          __syncwarp();
          if (++ws == (uint32_t)kTsWStages) ws = 0, wph ^= 1u;
        }
      }
      ia_base += (uint32_t)a_loads;
    }
  } else if (warp < 6) {
    // ------------------------------------------------ x-tile splitter (thread = row = tensor-memory lane) ------------
    const int r = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem + kTsAccCols + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t ia = 0, rs = 0, rph = 0;
    for (int mb = blockIdx.x; mb < prm.m_blocks; mb += gridDim.x) {
      for (int l = 0; l < a_loads; ++l, ++ia) {
        const uint32_t as = ia % kTsAStages;
        mbar_wait(b_raw_full + 8u * rs, rph);                               // fp32 tile landed
        mbar_wait(b_a_empty + 8u * as, ((ia / kTsAStages) & 1u) ^ 1u);      // the MMAs that read this stage retired
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned char* src = sm + kTsOffRaw + (size_t)rs * kRawBytes + (size_t)r * 128;  // pieces XOR (r & 7)
        uint32_t t0[16], t1[16], t2[16];  // t2: bf16 variant only
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 8 floats -> 4 packed pairs per term
          const float4 u = *reinterpret_cast<const float4*>(src + (((2 * j) ^ (r & 7)) << 4));
          const float4 v = *reinterpret_cast<const float4*>(src + (((2 * j + 1) ^ (r & 7)) << 4));
          if constexpr (NT == 3) {
            split2(u.x, u.y, t0[4 * j], t1[4 * j], t2[4 * j]);
            split2(u.z, u.w, t0[4 * j + 1], t1[4 * j + 1], t2[4 * j + 1]);
            split2(v.x, v.y, t0[4 * j + 2], t1[4 * j + 2], t2[4 * j + 2]);
            split2(v.z, v.w, t0[4 * j + 3], t1[4 * j + 3], t2[4 * j + 3]);
          } else {
            split2_f16(u.x, u.y, t0[4 * j], t1[4 * j]);
            split2_f16(u.z, u.w, t0[4 * j + 1], t1[4 * j + 1]);
            split2_f16(v.x, v.y, t0[4 * j + 2], t1[4 * j + 2]);
            split2_f16(v.z, v.w, t0[4 * j + 3], t1[4 * j + 3]);
          }
        }
        const uint32_t ta = t_lane + as * kTsAStageCols;
        tmem_st16(ta, t0);
        tmem_st16(ta + 16u, t1);
        if constexpr (NT == 3) tmem_st16(ta + 32u, t2);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        mbar_arrive(b_raw_empty + 8u * rs);  // raw tile consumed (generic reads only)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(b_a_full + 8u * as);
        if (++rs == (uint32_t)kTsRawStages) rs = 0, rph ^= 1u;
      }
    }
  } else if (warp < 14) {
    // ------------------------------------------------ epilogue (warps 6-13) ---------------------------------------
    // r02o (ncu source page): the MMA warp spent its time waiting for the accumulator to be handed back -- the epilogue
    // was the bottleneck at 3.5 us per 128 x 128 tile: a bias tile reloaded from global memory into shared memory
    // behind two 256-thread barriers per tile, and load -> add -> store chains serialised on one register quad. Now each
    // lane fetches the two 16-byte bias pieces of ITS columns before it waits for the accumulator, the raw sums go
    // through the transposition stage, and bias / ReLU are applied on the way out (8 independent loads, then 8 stores).
    const int q = warp & 3, half = (warp - 6) >> 2;  // tensor-memory lane quarter, 64-column half of the tile
    float* stg = reinterpret_cast<float*>(sm + kTsOffStage) + (size_t)(warp - 6) * 32 * kStagePitch;
    const uint32_t t_hi = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64), t_lo = t_hi + (uint32_t)kBN;
    const int piece = lane & 7;
    uint32_t tcount = 0;
    for (int mb = blockIdx.x; mb < prm.m_blocks; mb += gridDim.x) {
      // output row of each of the 8 tile rows this lane stores (row = q * 32 + (lane >> 3) + 4 i), or -1
      int64_t grow[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = q * 32 + (lane >> 3) + 4 * i;
        if (prm.conv) {
          const int tpi = prm.tiles_x * prm.tiles_y, img = mb / tpi, t = mb - img * tpi;
          const int ty = t / prm.tiles_x, tx = t - ty * prm.tiles_x;
          const int yl = r / prm.TX, xl = r - yl * prm.TX, yo = ty * prm.TY + yl;
          grow[i] = (yl < prm.TY && yo < prm.cHo) ? ((int64_t)img * prm.cHo + yo) * prm.cWo + tx * prm.TX + xl : -1;
        } else {
          const int64_t g = (int64_t)mb * kBM + r;
          grow[i] = g < prm.rows ? g : -1;
        }
      }
      for (int nt = 0; nt < prm.n_tiles; ++nt, ++tcount) {
        const int n0 = nt * kBN + half * 64 + piece * 4;  // first column of this lane's piece in column block 0
        float4 bb[2];
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
          const int n = n0 + cb * 32;
          bb[cb] = (prm.bias && n < prm.N) ? __ldg(reinterpret_cast<const float4*>(prm.bias + n))
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(b_acc_full, tcount & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float o[64];  // own row, own 64 columns: sum of the two accumulators
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          uint32_t v[16], w[16];
          tmem_ld16(t_hi + (uint32_t)c, v);
          tmem_ld16(t_lo + (uint32_t)c, w);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j)
            o[c + j] = NT == 3 ? __uint_as_float(v[j]) + __uint_as_float(w[j])
                               : fmaf(__uint_as_float(w[j]), 1.f / kF16Scale, __uint_as_float(v[j]));
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(b_acc_empty);  // accumulator handed back: the next tile's MMAs run under the rest of this epilogue
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {  // 32 x 32 sub-tiles through the warp's stage: full 128-byte row segments
          __syncwarp();  // the previous sub-tile's readers are done with the stage
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(stg + lane * kStagePitch + j) =
                make_float4(o[cb * 32 + j], o[cb * 32 + j + 1], o[cb * 32 + j + 2], o[cb * 32 + j + 3]);
          __syncwarp();
          const int n = n0 + cb * 32;
          if (n < prm.N) {  // N % 4 == 0: a 4-column piece is inside or outside as a whole
            float4 xs[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              xs[i] = *reinterpret_cast<const float4*>(stg + ((lane >> 3) + 4 * i) * kStagePitch + piece * 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float4 x = xs[i];
              x.x += bb[cb].x;
              x.y += bb[cb].y;
              x.z += bb[cb].z;
              x.w += bb[cb].w;
              if (prm.relu) {
                x.x = fmaxf(x.x, 0.f);
                x.y = fmaxf(x.y, 0.f);
                x.z = fmaxf(x.z, 0.f);
                x.w = fmaxf(x.w, 0.f);
              }
              if (grow[i] >= 0) {
                float* dst = prm.out + grow[i] * prm.N + n;
                if (prm.multicast)
                  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(x.x),
                               "f"(x.y), "f"(x.z), "f"(x.w)
                               : "memory");
                else
                  *reinterpret_cast<float4*>(dst) = x;
                if (prm.out2) {
                  const float4 a = __ldg(reinterpret_cast<const float4*>(prm.addend + grow[i] * prm.N + n));
                  *reinterpret_cast<float4*>(prm.out2 + grow[i] * prm.N + n) =
                      make_float4(x.x + a.x, x.y + a.y, x.z + a.z, x.w + a.w);
                }
              }
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTsTmemCols) : "memory");
  }
}

// W [n] fp32 -> terms [3][n] bf16: t0 = bf16(w), t1 = bf16(w - t0), t2 = bf16(w - t0 - t1)
__global__ void bf16_split3_kernel(const float* __restrict__ w, int64_t n, __nv_bfloat16* __restrict__ t) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = w[i];
  const __nv_bfloat16 h = __float2bfloat16_rn(a);
  const float r1 = a - __bfloat162float(h);
  const __nv_bfloat16 m = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(m);
  t[i] = h;
  t[n + i] = m;
  t[2 * n + i] = __float2bfloat16_rn(r2);
}

// W [n] fp32 -> terms [2][n] fp16: t0 = fp16(w), t1 = fp16((w - t0) * 2^11)
__global__ void f16_split2_kernel(const float* __restrict__ w, int64_t n, __half* __restrict__ t) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = w[i];
  const __half h = __float2half_rn(a);
  t[i] = h;
  t[n + i] = __float2half_rn((a - __half2float(h)) * kF16Scale);
}

}  // namespace
}  // namespace mvd

using namespace mvd;

extern "C" int mvd_f16_split2_f32(const float* w, int64_t n, void* terms, void* stream) {
  if (!w || !terms) return MVD_ERR_NULL_POINTER;
  if (n <= 0) return MVD_ERR_BAD_SHAPE;
  f16_split2_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(w, n, reinterpret_cast<__half*>(terms));
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

extern "C" int mvd_bf16_split3_f32(const float* w, int64_t n, void* terms, void* stream) {
  if (!w || !terms) return MVD_ERR_NULL_POINTER;
  if (n <= 0) return MVD_ERR_BAD_SHAPE;
  bf16_split3_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(w, n,
                                                                                      reinterpret_cast<__nv_bfloat16*>(terms));
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

static int encode_w_terms(CUtensorMap* tm_b, const void* w_terms, int K, int N, int terms) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return MVD_ERR_NO_DEVICE;
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)terms};
  const cuuint64_t gstr[2] = {(cuuint64_t)K * 2, (cuuint64_t)N * K * 2};
  const cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)kBN, 1u};
  if (enc(tm_b, terms == 3 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
          const_cast<void*>(w_terms), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return MVD_ERR_UNSUPPORTED;
  return MVD_OK;
}

// persistent over row blocks; the column tiles of a row block run back to back on one SM
static int launch_ts(const CUtensorMap& tm_a, const CUtensorMap& tm_b, GbParams& prm, int terms, void* stream) {
  auto kern = terms == 3 ? linear_split_ts_kernel<3> : linear_split_ts_kernel<2>;
  // ring depths: defaults of TsCfg, or MVD_GEMM_RINGS="raw,w" (tuning aid; must fit 227 KB and 12 stages each)
  static int env_raw = -1, env_w = -1;
  if (env_raw < 0) {
    int r = 0, w = 0;
    const char* e = getenv("MVD_GEMM_RINGS");
    if (!e || sscanf(e, "%d,%d", &r, &w) != 2) r = w = 0;
    env_w = w;
    env_raw = r;
  }
  prm.raw_stages = env_raw > 0 ? env_raw : (terms == 3 ? TsCfg<3>::kRaw : TsCfg<2>::kRaw);
  prm.w_stages = env_w > 0 ? env_w : (terms == 3 ? TsCfg<3>::kW : TsCfg<2>::kW);
  const int smem = terms == 3 ? TsCfg<3>::smem(prm.raw_stages, prm.w_stages) : TsCfg<2>::smem(prm.raw_stages, prm.w_stages);
  if (prm.raw_stages < 2 || prm.w_stages < 2 || prm.raw_stages > 12 || prm.w_stages > 12 || smem > 227 * 1024)
    return MVD_ERR_UNSUPPORTED;
  MVD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const unsigned grid = (unsigned)(prm.m_blocks < kNumSMs ? prm.m_blocks : kNumSMs);
  kern<<<grid, kTsThreads, smem, (cudaStream_t)stream>>>(tm_a, tm_b, prm);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

static int linear_bf16x3(const float* x, const void* w_terms, const float* bias, int64_t rows, int K, int N, int relu,
                         float* out, int multicast, int ts, void* stream, int terms = 3) {
  if (!x || !w_terms || !out) return MVD_ERR_NULL_POINTER;
  if (rows <= 0 || K <= 0 || N <= 0 || rows > 0x7fffffffLL) return MVD_ERR_BAD_SHAPE;
  if (K % 8 != 0 || N % 4 != 0) return MVD_ERR_UNSUPPORTED;  // 16-byte row pitch of the bf16 terms, 16-byte output pieces
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_terms) |
                       reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias);
  if (al & 15u) return MVD_ERR_MISALIGNED;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return MVD_ERR_NO_DEVICE;
  alignas(64) CUtensorMap tm_a, tm_b;
  {
    const cuuint32_t estr[2] = {1u, 1u};
    const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kBM};
    if (enc(&tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MVD_ERR_UNSUPPORTED;
  }
  if (int e = encode_w_terms(&tm_b, w_terms, K, N, terms)) return e;
  GbParams prm = {};
  prm.bias = bias;
  prm.out = out;
  prm.rows = (int)rows;
  prm.K = K;
  prm.N = N;
  prm.relu = relu;
  prm.m_blocks = (int)ceil_div64(rows, kBM);
  prm.n_tiles = (int)ceil_div64(N, kBN);
  prm.multicast = multicast;
  if (ts) return launch_ts(tm_a, tm_b, prm, terms, stream);
  MVD_CUDA_TRY(cudaFuncSetAttribute(linear_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBSmem));
  const int64_t tiles = (int64_t)prm.m_blocks * prm.n_tiles;
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  linear_bf16x3_kernel<<<grid, kBThreads, kBSmem, (cudaStream_t)stream>>>(tm_a, tm_b, prm);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

// 3x3 / pad 1 / stride {1, 2} convolution of a channels-last image as an implicit GEMM on the TS kernel: no im2col
// matrix. Output [NB * Ho * Wo, N] (channels-last). Bit-identical to mvd_linear_* over the (ky, kx, c)-ordered im2col
// matrix of the same image: the K chunks, their order and the accumulators are the same.
static int conv3x3_nhwc(const float* src, const void* w_terms, const float* bias, int NB, int Hi, int Wi, int C,
                        int stride, int N, int relu, int terms, float* out, const float* addend, float* out2,
                        int multicast, void* stream) {
  if (!src || !w_terms || !out || ((addend == nullptr) != (out2 == nullptr))) return MVD_ERR_NULL_POINTER;
  if ((reinterpret_cast<uintptr_t>(addend) | reinterpret_cast<uintptr_t>(out2)) & 15u) return MVD_ERR_MISALIGNED;
  if (NB <= 0 || Hi <= 0 || Wi <= 0 || C <= 0 || N <= 0) return MVD_ERR_BAD_SHAPE;
  if ((stride != 1 && stride != 2) || (terms != 2 && terms != 3)) return MVD_ERR_BAD_SHAPE;
  if (C % kBK != 0 || N % 4 != 0) return MVD_ERR_UNSUPPORTED;  // a K chunk is 32 channels of one tap
  const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(w_terms) |
                       reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(bias);
  if (al & 15u) return MVD_ERR_MISALIGNED;
  const int Ho = (Hi + 2 - 3) / stride + 1, Wo = (Wi + 2 - 3) / stride + 1;
  if ((int64_t)NB * Ho * Wo > 0x7fffffffLL) return MVD_ERR_BAD_SHAPE;
  // tile = TY x TX output pixels, TX the largest divisor of Wo that fits 128 rows (a TMA box cannot wrap image rows)
  int TX = 0;
  for (int d = (Wo < kBM ? Wo : kBM); d >= 1; --d)
    if (Wo % d == 0) {
      TX = d;
      break;
    }
  if (TX < 16) return MVD_ERR_UNSUPPORTED;  // e.g. a prime width above 128: the caller keeps the im2col route
  int TY = kBM / TX;
  if (TY > Ho) TY = Ho;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return MVD_ERR_NO_DEVICE;
  alignas(64) CUtensorMap tm_a, tm_b;
  {  // to load n elements with traversal stride s the box spans n * s (cuTensorMapEncodeTiled: ceil(box / stride) loaded)
    const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)NB};
    const cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)Wi * C * 4, (cuuint64_t)Hi * Wi * C * 4};
    const cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)(TX * stride), (cuuint32_t)(TY * stride), 1u};
    const cuuint32_t estr[4] = {1u, (cuuint32_t)stride, (cuuint32_t)stride, 1u};
    if (box[1] > 256u || box[2] > 256u) return MVD_ERR_UNSUPPORTED;
    if (enc(&tm_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(src), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MVD_ERR_UNSUPPORTED;
  }
  const int K = 9 * C;
  if (int e = encode_w_terms(&tm_b, w_terms, K, N, terms)) return e;
  GbParams prm = {};
  prm.bias = bias;
  prm.out = out;
  prm.rows = NB * Ho * Wo;
  prm.K = K;
  prm.N = N;
  prm.relu = relu;
  prm.conv = 1;
  prm.cHo = Ho;
  prm.cWo = Wo;
  prm.TX = TX;
  prm.TY = TY;
  prm.tiles_x = Wo / TX;
  prm.tiles_y = (Ho + TY - 1) / TY;
  prm.cchunks = C / kBK;
  prm.cstride = stride;
  prm.m_blocks = NB * prm.tiles_x * prm.tiles_y;
  prm.n_tiles = (int)ceil_div64(N, kBN);
  prm.multicast = multicast;
  prm.addend = addend;
  prm.out2 = out2;
  return launch_ts(tm_a, tm_b, prm, terms, stream);
}

extern "C" int mvd_linear_bf16x3_f32(const float* x, const void* w_terms, const float* bias, int64_t rows, int K, int N,
                                     int relu, float* out, void* stream) {
  return linear_bf16x3(x, w_terms, bias, rows, K, N, relu, out, 0, 0, stream);
}

// Same arithmetic (bit-identical results), x terms staged in tensor memory: see linear_split_ts_kernel<3>.
extern "C" int mvd_linear_bf16x3_ts_f32(const float* x, const void* w_terms, const float* bias, int64_t rows, int K, int N,
                                        int relu, float* out, void* stream) {
  return linear_bf16x3(x, w_terms, bias, rows, K, N, relu, out, 0, 1, stream);
}

// Same GEMM whose epilogue IS the all-gather: `out_mc` is the NVLink multicast mapping of a symmetric buffer (every GPU
// of the group maps the same physical layout); each 16-byte piece of the result leaves as one multimem.st that the
// NVSwitch replicates into every GPU's copy, tile by tile while the main loop of the next tile runs. The caller
// separates producers from consumers with a cross-GPU barrier (mvdetr_b200/sharded.py).
extern "C" int mvd_linear_bf16x3_multicast_f32(const float* x, const void* w_terms, const float* bias, int64_t rows, int K,
                                               int N, int relu, float* out_mc, void* stream) {
  return linear_bf16x3(x, w_terms, bias, rows, K, N, relu, out_mc, 1, 0, stream);
}

// Two fp16 terms, three products (see linear_split_ts_kernel<2>): w_terms [2][N][K] fp16 from mvd_f16_split2_f32.
extern "C" int mvd_linear_f16x2_f32(const float* x, const void* w_terms, const float* bias, int64_t rows, int K, int N,
                                    int relu, float* out, void* stream) {
  return linear_bf16x3(x, w_terms, bias, rows, K, N, relu, out, 0, 1, stream, 2);
}

extern "C" int mvd_linear_f16x2_multicast_f32(const float* x, const void* w_terms, const float* bias, int64_t rows, int K,
                                              int N, int relu, float* out_mc, void* stream) {
  return linear_bf16x3(x, w_terms, bias, rows, K, N, relu, out_mc, 1, 1, stream, 2);
}

// terms = 2: w_terms from mvd_f16_split2_f32 (default); terms = 3: from mvd_bf16_split3_f32. The weight matrix is
// [N][9 C] with columns ordered (ky, kx, c).
extern "C" int mvd_conv3x3_nhwc_f32(const float* src, const void* w_terms, const float* bias, int NB, int Hi, int Wi,
                                    int C, int stride, int N, int relu, int terms, float* out, const float* addend,
                                    float* out2, void* stream) {
  return conv3x3_nhwc(src, w_terms, bias, NB, Hi, Wi, C, stride, N, relu, terms, out, addend, out2, 0, stream);
}

extern "C" int mvd_linear_bf16x3_ts_multicast_f32(const float* x, const void* w_terms, const float* bias, int64_t rows,
                                                  int K, int N, int relu, float* out_mc, void* stream) {
  return linear_bf16x3(x, w_terms, bias, rows, K, N, relu, out_mc, 1, 1, stream);
}

namespace mvd {
namespace {
// dst_mc row of local row r: r itself (inner == 0) or, with rows laid out [outer][inner] locally, the transposed position
// (r % inner) * outer_total + outer0 + r / inner  (view-major tokens of the local views -> cell-major rows of ALL views)
__global__ void multicast_copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst_mc, int64_t rows, int c4,
                                      int64_t inner, int64_t outer_total, int64_t outer0) {
  const int64_t n4 = rows * c4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4, c = i - r * c4;
    const int64_t orow = inner > 0 ? (r % inner) * outer_total + outer0 + r / inner : r;
    const float4 v = src[i];
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst_mc + orow * c4 + c), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
  }
}
}  // namespace
}  // namespace mvd

// All-gather of a tensor whose producer is not one of the kernels above: src [rows, C] (C % 4 == 0, 16-byte aligned) is
// written through the multicast address `dst_mc` into every GPU's copy of a symmetric buffer.
//   inner == 0: row r -> row r of dst_mc (dst_mc already points at this rank's slot);
//   inner  > 0: local rows are [outer_local][inner]; row r -> row (r % inner) * outer_total + outer0 + r / inner of
//               dst_mc (base of the whole buffer): view-major tokens leave cell-major, so the merge convolution over all
//               views (ref: mvd/models/trans_world_feat.py:107-108) reads contiguous rows on every GPU without a
//               permute-copy.
extern "C" int mvd_multicast_copy_f32(const float* src, float* dst_mc, int64_t rows, int C, int64_t inner,
                                      int64_t outer_total, int64_t outer0, void* stream) {
  if (!src || !dst_mc) return MVD_ERR_NULL_POINTER;
  if (rows <= 0 || C <= 0 || (C & 3) || inner < 0 || (inner > 0 && (rows % inner != 0 || outer_total <= 0 || outer0 < 0)))
    return MVD_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst_mc)) & 15u) return MVD_ERR_MISALIGNED;
  const int64_t n4 = rows * (C / 4);
  const int64_t want = ceil_div64(n4, 256);
  const int blocks = (int)(want > (int64_t)kNumSMs * 8 ? (int64_t)kNumSMs * 8 : want);
  multicast_copy_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(src),
                                                                  reinterpret_cast<float4*>(dst_mc), rows, C / 4, inner,
                                                                  outer_total, outer0);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}
