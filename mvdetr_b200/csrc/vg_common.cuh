// Shared by the view-grid kernels (msda_viewgrid.cu forward, msda_bwd_viewgrid.cu backward): TMA / mbarrier wrappers,
// shared-memory record readers, the tile plan and the tensor-map encoder. Everything lives in an anonymous namespace
// (one private copy per translation unit).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace mvd {
namespace {

constexpr int kHalo = 6;          // pixels around the tile kept in the window (default init samples +-P px, P = 4: ms_deform_attn.py:62-77)
// P > 4: the default pattern reaches +-P px, whose 2x2 footprints need P + 1 (r02i: the 8-point stress shape ran at 0.06 of
// the roofline with a quarter of its samples on the masked global path)
inline int viewgrid_halo(int P) { return P <= 4 ? kHalo : P + 1; }
constexpr int kMaxThreads = 448;  // 2 blocks/SM at <= 72 registers

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], "
      "[%6];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], "
      "[%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
// One leader lane of a CONVERGED warp: uniform-datapath instructions (UTMALDG, UTCHMMA, UTCBAR) issued under
// `if (lane == 0)` get wrapped in an ELECT / R2UR / BRA.U.ANY waterfall by the compiler; under elect.sync they do not.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// acc[k] += w1*v1 + w2*v2 + w3*v3 + w4*v4 for one channel quad
__device__ __forceinline__ void fma4(float4& acc, float w1, const float4& v1, float w2, const float4& v2, float w3,
                                     const float4& v3, float w4, const float4& v4) {
  acc.x = fmaf(w4, v4.x, fmaf(w3, v3.x, fmaf(w2, v2.x, fmaf(w1, v1.x, acc.x))));
  acc.y = fmaf(w4, v4.y, fmaf(w3, v3.y, fmaf(w2, v2.y, fmaf(w1, v1.y, acc.y))));
  acc.z = fmaf(w4, v4.z, fmaf(w3, v3.z, fmaf(w2, v2.z, fmaf(w1, v1.z, acc.z))));
  acc.w = fmaf(w4, v4.w, fmaf(w3, v3.w, fmaf(w2, v2.w, fmaf(w1, v1.w, acc.w))));
}

// Reads N floats (N % 4 == 0) of this thread's record from a [thread][N] shared array with 128-bit loads. For
// N == 8 the two halves are fetched in an order that depends on bit 2 of the lane, so the 8 lanes of a quarter-warp
// phase cover all 32 banks (records are 32 bytes apart: lanes j and j+4 would otherwise collide).
template <int N>
__device__ __forceinline__ void read_record(const unsigned char* base, int idx, int lane, float* dst) {
  const float4* p = reinterpret_cast<const float4*>(base + (size_t)idx * N * 4);
  if (N == 8) {
    const int h = (lane >> 2) & 1;
    const float4 a = p[h], b = p[h ^ 1];
    const float4 lo = h ? b : a, hi = h ? a : b;
    dst[0] = lo.x, dst[1] = lo.y, dst[2] = lo.z, dst[3] = lo.w;
    dst[4] = hi.x, dst[5] = hi.y, dst[6] = hi.z, dst[7] = hi.w;
  } else {
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
      const float4 a = p[i];
      dst[4 * i] = a.x, dst[4 * i + 1] = a.y, dst[4 * i + 2] = a.z, dst[4 * i + 3] = a.w;
    }
  }
}

struct VgPlan {
  int TH, TW, BW, BH, halo, tiles_x, tiles_y, threads;
  uint32_t off_a, off_b, off_ref, stage_bytes, zero_off, zero_bytes, tx_bytes;
  size_t smem;
};

inline uint32_t up128(size_t v) { return (uint32_t)((v + 127) & ~(size_t)127); }

// Tile = TH x TW ground cells with TH*TW*R <= kMaxThreads threads; prefers 64 cells (4x16), 32 (4x8) for many views; the
// tile is narrowed further while two stages of window + records do not fit the shared memory of one block.
// `split` threads share one (query, head) pair (each owns D / split channels); `max_threads` is the block-size limit.
bool plan_viewgrid(int D, int R, int P, bool fused, VgPlan* pl, int halo = kHalo, int split = 1,
                   int max_threads = kMaxThreads) {
  int TH = 4, TW = 16;
  while (TH * TW * R * split > max_threads && TW > 4) TW >>= 1;
  while (TH * TW * R * split > max_threads && TH > 1) TH >>= 1;
  if (TH * TW * R * split > max_threads || R > 256) return false;
  // few views per block (the view-sharded multi-GPU path: R = 1 per rank): 64 threads per window left the SM at 6 warps
  // (r02l timeline: 57 us per launch for a seventh of the queries). A taller tile doubles the threads per window.
  if (TH == 4 && 2 * TH * TW * R * split <= max_threads / 2) TH = 8;
  for (;; TW >>= 1) {
    pl->TH = TH;
    pl->TW = TW;
    pl->halo = halo;
    pl->BW = TW + 2 * halo;
    pl->BH = TH + 2 * halo;
    pl->threads = ((TH * TW * R * split + 31) / 32) * 32;
    const size_t nbox = (size_t)TH * TW * R;
    const size_t win = (size_t)pl->BW * pl->BH * D * 4, a = nbox * 2 * P * 4, b = nbox * P * 4,
                 rf = fused ? (size_t)TH * TW * 2 * P * 4 : 0;
    pl->off_a = up128(win);
    pl->off_b = pl->off_a + up128(a);
    pl->off_ref = pl->off_b + up128(b);
    pl->stage_bytes = pl->off_ref + up128(rf);
    pl->tx_bytes = (uint32_t)(win + a + b + rf);
    pl->zero_off = 2 * pl->stage_bytes;
    pl->zero_bytes = up128((size_t)pl->BW * D * 4 + 2 * D * 4) + 128;  // reach of a 2x2 footprint from its top-left pixel (+ bank-preserving start offset < 128)
    pl->smem = (size_t)pl->zero_off + pl->zero_bytes;
    if (pl->smem <= 227 * 1024 && 2 * P * TW <= 256) return true;
    if (TW <= 4) return false;
  }
}

int encode(CUtensorMap* map, const float* base, int rank, const cuuint64_t* gdim, const cuuint64_t* gstr,
           const cuuint32_t* box) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return MVD_ERR_NO_DEVICE;
  const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
  for (int i = 0; i + 1 < rank; ++i)
    if (gstr[i] % 16 != 0) return MVD_ERR_UNSUPPORTED;
  const CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), gdim, gstr,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return rc == CUDA_SUCCESS ? MVD_OK : MVD_ERR_UNSUPPORTED;
}

}  // namespace
}  // namespace mvd
