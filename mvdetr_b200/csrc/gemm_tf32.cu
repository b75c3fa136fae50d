// Dense glue of the encoder layer on the 5th-generation tensor cores, with fp32-level accuracy:
//   out[rows, N] = act(x[rows, K] @ W[N, K]^T + bias[N])
//   replaces the six nn.Linear calls per encoder layer  ref: multiview_detector/models/ops/modules/ms_deform_attn.py:96,100-101,116,
//                                                        multiview_detector/models/deformable_transformer.py:82
//   and (through the im2col matrices of im2col.cu / warp_tma.cu) the convolutions of multiview_detector/models/trans_world_feat.py:74,82-84
//
// Round 1 ran these through cuBLASLt's CUBLAS_COMPUTE_32F_EMULATED_16BFX9: nine bf16 products per GEMM plus a separate
// Inf/NaN operand scan kernel per call (24.5 % of the frame, no switch in cuBLAS 12.9), 68 % of the frame in library
// time. The layers are tall-skinny (75 600 x 128 by 128 x {128..512}): at ~10 flop/byte they are HBM-bound on B200,
// so the kernel's job is to stream A once and keep everything else on chip.
//
// 3xTF32 split, evaluated by tcgen05.mma.kind::tf32 with fp32 accumulation in tensor memory:
//   a = a_hi + a_lo with a_hi = RN_tf32(a) (11 significant bits) and a_lo = a - a_hi exactly (13 bits, of which the
//   tensor core reads the top 11); likewise the weights, split once on the device (mvd_tf32_split_f32).
//   x @ W^T ~= a_hi b_hi + a_hi b_lo + a_lo b_hi; the dropped terms are <= 2^-22 |a||b| each (scripts/emulate_tf32_split.py:
//   max error vs fp64 as cuBLASLt's BF16x9, below a native fp32 GEMM).
// Structure (one CTA = 128 rows x 128 output columns, 128 threads, up to 3 CTAs per SM):
//   per 32-wide K chunk: one thread issues TMA loads of the A tile [128 x 32] fp32 and of the B_hi / B_lo tiles
//   [128 x 32] (SWIZZLE_128B, K-major: each row of the chunk is one 128-byte swizzle row; rows past `rows` / `N` are
//   zero-filled by the TMA unit); all threads split the A tile in place into hi / lo (elementwise on the swizzled
//   bytes, so the split never needs to know the swizzle); one thread issues 4 k-steps x 3 tcgen05.mma (M 128, N 128,
//   K 8) and commits them to an mbarrier; the stage is reused when that barrier flips. Co-resident CTAs overlap each
//   other's load / split / MMA / epilogue phases.
//   epilogue: each warp reads its 32 accumulator rows from TMEM (tcgen05.ld 32x32b), adds the bias, applies ReLU and
//   writes 16-byte pieces of its own rows.
#include <cuda.h>

#include "vg_common.cuh"

namespace mvd {
namespace {

constexpr int kGtM = 128, kGtN = 128, kGtK = 32;  // CTA tile; K chunk = one 128-byte swizzle row of fp32
constexpr int kGtThreads = 128;
constexpr int kGtTileBytes = kGtM * kGtK * 4;      // 16 KB: A, A_lo, B_hi, B_lo tiles all have this size
constexpr int kGtSmem = 4 * kGtTileBytes + 1024;   // + slack for the 1024-byte alignment of the swizzle atoms
constexpr uint32_t kGtTmemCols = 128;

struct GtParams {
  const float* bias;  // nullable
  float* out;
  int rows, K, N, relu;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// UMMA shared-memory matrix descriptor: K-major tile whose rows are 128-byte swizzle rows (SWIZZLE_128B), 8-row groups
// 1024 bytes apart (SBO), one swizzle atom along K (LBO unused), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);   // start address, 16-byte units
  d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                        // version
  d |= (uint64_t)2 << 61;                        // layout type: SWIZZLE_128B
  return d;
}

// Instruction descriptor: D fp32 (c_format 1), A and B tf32 (format 2), both K-major, N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kGtN >> 3) << 17) |
                                ((uint32_t)(kGtM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdescTf32), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kGtThreads, 3)
    linear_tf32x3_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_bhi,
                         const __grid_constant__ CUtensorMap tm_blo, const GtParams prm) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) unsigned long long s_bar[2];  // [0] TMA landed, [1] MMAs retired
  __shared__ uint32_t s_tmem;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;  // swizzle atoms are 1024-byte aligned
  unsigned char* sm = smem_dyn + (base - smem_u32(smem_dyn));
  const uint32_t a_hi = base, a_lo = base + kGtTileBytes, b_hi = base + 2 * kGtTileBytes, b_lo = base + 3 * kGtTileBytes;
  const uint32_t bar_full = smem_u32(&s_bar[0]), bar_mma = smem_u32(&s_bar[1]);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * kGtM, n0 = blockIdx.y * kGtN;

  if (tid == 0) {
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_bhi);
    prefetch_tmap(&tm_blo);
    mbar_init(bar_full, 1);
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {  // one warp allocates the accumulator columns and gives the permit back
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                 "r"(kGtTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = s_tmem;

  const int nchunks = (prm.K + kGtK - 1) / kGtK;
  for (int kc = 0; kc < nchunks; ++kc) {
    const uint32_t ph = (uint32_t)(kc & 1);
    if (tid == 0) {
      mbar_expect_tx(bar_full, 3u * kGtTileBytes);
      tma_load_2d(a_hi, &tm_a, bar_full, kc * kGtK, m0);
      tma_load_2d(b_hi, &tm_bhi, bar_full, kc * kGtK, n0);
      tma_load_2d(b_lo, &tm_blo, bar_full, kc * kGtK, n0);
    }
    mbar_wait(bar_full, ph);
    // split the A tile in place: hi = RN_tf32(a), lo = a - hi (same swizzled byte position in the lo tile)
    float4* ph4 = reinterpret_cast<float4*>(sm);
    float4* pl4 = reinterpret_cast<float4*>(sm + kGtTileBytes);
#pragma unroll
    for (int i = 0; i < kGtTileBytes / 16 / kGtThreads; ++i) {
      const int idx = i * kGtThreads + tid;
      const float4 a = ph4[idx];
      float4 h, l;
      uint32_t t;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(a.x));
      h.x = __uint_as_float(t);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(a.y));
      h.y = __uint_as_float(t);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(a.z));
      h.z = __uint_as_float(t);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(a.w));
      h.w = __uint_as_float(t);
      l.x = a.x - h.x;
      l.y = a.y - h.y;
      l.z = a.z - h.z;
      l.w = a.w - h.w;
      ph4[idx] = h;
      pl4[idx] = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> tensor-core (async proxy) reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int k = 0; k < kGtK / 8; ++k) {  // UMMA K = 8 tf32 = 32 bytes: advance the start address inside the swizzle row
        const uint64_t dah = umma_desc_k128(a_hi + 32u * k), dal = umma_desc_k128(a_lo + 32u * k);
        const uint64_t dbh = umma_desc_k128(b_hi + 32u * k), dbl = umma_desc_k128(b_lo + 32u * k);
        umma_tf32(tmem_d, dal, dbh, (kc | k) != 0);  // small terms first
        umma_tf32(tmem_d, dah, dbl, 1u);
        umma_tf32(tmem_d, dah, dbh, 1u);
      }
      // arrives on the mbarrier when every MMA issued so far has retired (implies tcgen05.fence::before_thread_sync)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma)
                   : "memory");
    }
    mbar_wait(bar_mma, ph);  // operands consumed: the stage may be overwritten, the accumulator read
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  // ---- epilogue: warp w owns accumulator lanes (rows) 32w .. 32w+31; thread = one row ----
  const int row = m0 + tid;
  float* orow = prm.out + (int64_t)row * prm.N + n0;
#pragma unroll 1
  for (int c = 0; c < kGtN; c += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row < prm.rows) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const int n = n0 + c + j;
        if (n < prm.N) {  // N % 4 == 0: a 4-column piece is inside or outside as a whole
          float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                 __uint_as_float(v[j + 3]));
          if (prm.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(prm.bias + n));
            o.x += b.x;
            o.y += b.y;
            o.z += b.z;
            o.w += b.w;
          }
          if (prm.relu) {
            o.x = fmaxf(o.x, 0.f);
            o.y = fmaxf(o.y, 0.f);
            o.z = fmaxf(o.z, 0.f);
            o.w = fmaxf(o.w, 0.f);
          }
          *reinterpret_cast<float4*>(orow + c + j) = o;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(kGtTmemCols) : "memory");
  }
}

__global__ void tf32_split_kernel(const float* __restrict__ w, int64_t n, float* __restrict__ hi, float* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = w[i];
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(a));
  const float h = __uint_as_float(t);
  hi[i] = h;
  lo[i] = a - h;
}

int encode_2d_sw128(CUtensorMap* map, const float* base, int64_t rows, int K) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return MVD_ERR_NO_DEVICE;
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)K * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kGtK, (cuuint32_t)kGtM};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? MVD_OK : MVD_ERR_UNSUPPORTED;
}

}  // namespace
}  // namespace mvd

using namespace mvd;

extern "C" int mvd_tf32_split_f32(const float* w, int64_t n, float* hi, float* lo, void* stream) {
  if (!w || !hi || !lo) return MVD_ERR_NULL_POINTER;
  if (n <= 0) return MVD_ERR_BAD_SHAPE;
  tf32_split_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(w, n, hi, lo);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

extern "C" int mvd_linear_tf32x3_f32(const float* x, const float* w_hi, const float* w_lo, const float* bias,
                                     int64_t rows, int K, int N, int relu, float* out, void* stream) {
  if (!x || !w_hi || !w_lo || !out) return MVD_ERR_NULL_POINTER;
  if (rows <= 0 || K <= 0 || N <= 0 || rows > 0x7fffffffLL) return MVD_ERR_BAD_SHAPE;
  if (K % 4 != 0 || N % 4 != 0) return MVD_ERR_UNSUPPORTED;  // 16-byte row pitch (TMA) and 16-byte output pieces
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w_hi) |
                       reinterpret_cast<uintptr_t>(w_lo) | reinterpret_cast<uintptr_t>(out) |
                       reinterpret_cast<uintptr_t>(bias);
  if (al & 15u) return MVD_ERR_MISALIGNED;
  alignas(64) CUtensorMap maps[3];
  if (int e = encode_2d_sw128(&maps[0], x, rows, K)) return e;
  if (int e = encode_2d_sw128(&maps[1], w_hi, N, K)) return e;
  if (int e = encode_2d_sw128(&maps[2], w_lo, N, K)) return e;
  MVD_CUDA_TRY(cudaFuncSetAttribute(linear_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGtSmem));
  GtParams prm;
  prm.bias = bias;
  prm.out = out;
  prm.rows = (int)rows;
  prm.K = K;
  prm.N = N;
  prm.relu = relu;
  dim3 grid((unsigned)ceil_div64(rows, kGtM), (unsigned)ceil_div64(N, kGtN));
  if (grid.y > 65535) return MVD_ERR_BAD_SHAPE;
  linear_tf32x3_kernel<<<grid, kGtThreads, kGtSmem, (cudaStream_t)stream>>>(maps[0], maps[1], maps[2], prm);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}
