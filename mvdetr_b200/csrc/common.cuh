// Shared device/host helpers for libmvdetr_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mvdetr_b200.h"

#define MVD_CUDA_TRY(expr)                       \
  do {                                           \
    cudaError_t e__ = (expr);                    \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)

// Launch-error check that never synchronises: picks up configuration errors of the launch just made.
#define MVD_LAUNCH_CHECK()                       \
  do {                                           \
    cudaError_t e__ = cudaGetLastError();        \
    if (e__ != cudaSuccess) return (int)e__;     \
  } while (0)

namespace mvd {

constexpr int kNumSMs = 148;  // B200

// Per-level geometry as the kernels consume it (converted once per block from the int64 device arrays the
// reference API hands over: mvd/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:274-277 re-reads them per thread).
struct Level {
  int H, W, start, pad;
};

__device__ __forceinline__ void load_levels(Level* s_lvl, const int64_t* __restrict__ shapes,
                                            const int64_t* __restrict__ start, int L) {
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    Level lv;
    lv.H = (int)shapes[2 * l];
    lv.W = (int)shapes[2 * l + 1];
    lv.start = (int)start[l];
    lv.pad = 0;
    s_lvl[l] = lv;
  }
}

template <typename T>
__device__ __forceinline__ T ldg(const T* p) {
  return __ldg(p);
}

// Streaming (read-once) loads: keep them out of L1 so the gathered `value` lines stay resident.
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ld_stream2(const float* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// 256-bit global load (sm_100+): one request per lane for a whole 32-byte sector. p must be 32-byte aligned.
__device__ __forceinline__ void ld_stream8(const float* p, float* v) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void ldg8(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// Vector float atomic add without return (sm_90+): one L2 reduction op for 4 channels.
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__host__ __device__ __forceinline__ int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace mvd
