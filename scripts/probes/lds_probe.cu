// Shared-memory gather micro-benchmark (B200): cycles per LDS.128 warp instruction for the address patterns the
// view-grid MSDA kernel could use. Answers: does a 128-bit shared load whose lanes share addresses cost fewer
// wavefronts (so de-duplicating corners across lanes would pay), and what do 64-byte-pitch pixel gathers cost.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/lds_probe scripts/probes/lds_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

template <int WIDTH>
__global__ void probe(const int* __restrict__ offs, long long* cyc, float* sink, int smem_floats) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < smem_floats; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  int o = offs[threadIdx.x & 31];  // byte offset for this lane
  float acc = 0.f;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 8
  for (int it = 0; it < kIters; ++it) {
    const unsigned a = base + (unsigned)o;  // asm volatile: the loads are neither hoisted nor merged
    if (WIDTH == 16) {
      float4 v;
      asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
      acc += v.x + v.y + v.z + v.w;
    } else if (WIDTH == 8) {
      float2 v;
      asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
      acc += v.x + v.y;
    } else {
      float v;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
      acc += v;
    }
    o ^= (it & 1) ? 2048 : 4096;  // same pattern, different base
  }
  __syncthreads();  // the whole block: warp 0 alone would only time its own (prioritised) instruction stream
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
  const int threads = 1024, blocks = 148;
  int h[32];
  int* d_off;
  long long* d_cyc;
  float* d_sink;
  cudaMalloc(&d_off, 32 * 4);
  cudaMalloc(&d_cyc, blocks * 8);
  cudaMalloc(&d_sink, blocks * threads * 4);
  const int smem_floats = 16384;  // 64 KB
  cudaFuncSetAttribute(probe<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_floats * 4);
  cudaFuncSetAttribute(probe<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_floats * 4);
  cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_floats * 4);
  struct Pat { const char* name; int width; int (*f)(int); };
  Pat pats[] = {
      {"v4 distinct consecutive (512 B)", 16, [](int l) { return l * 16; }},
      {"v4 pairs share (lane>>1)*16 (256 B)", 16, [](int l) { return (l >> 1) * 16; }},
      {"v4 quads share (lane>>2)*16 (128 B)", 16, [](int l) { return (l >> 2) * 16; }},
      {"v4 lane and lane+16 share (256 B)", 16, [](int l) { return (l & 15) * 16; }},
      {"v4 lane and lane+8 share (128 B)", 16, [](int l) { return (l & 7) * 16; }},
      {"v4 all same (broadcast)", 16, [](int l) { return 0; }},
      {"v4 64B-pitch pixels, same quad (4-way)", 16, [](int l) { return l * 64; }},
      {"v4 64B-pitch pixels, quad rot (lane>>1)&3", 16, [](int l) { return l * 64 + ((l >> 1) & 3) * 16; }},
      {"v4 64B-pitch, pairs of lanes same pixel", 16, [](int l) { return (l >> 1) * 64 + ((l >> 2) & 3) * 16; }},
      {"v4 128B-pitch pixels (D=32), rot lane&7", 16, [](int l) { return l * 128 + (l & 7) * 16; }},
      {"v2 distinct consecutive (256 B)", 8, [](int l) { return l * 8; }},
      {"v2 64B-pitch pixels, rot", 8, [](int l) { return l * 64 + (l & 7) * 8; }},
      {"v1 distinct consecutive (128 B)", 4, [](int l) { return l * 4; }},
      {"v1 all same", 4, [](int l) { return 0; }},
      {"v1 2-way conflict", 4, [](int l) { return l * 8; }},
  };
  for (const Pat& p : pats) {
    for (int l = 0; l < 32; ++l) h[l] = p.f(l);
    cudaMemcpy(d_off, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) {
      if (p.width == 16) probe<16><<<blocks, threads, smem_floats * 4>>>(d_off, d_cyc, d_sink, smem_floats);
      else if (p.width == 8) probe<8><<<blocks, threads, smem_floats * 4>>>(d_off, d_cyc, d_sink, smem_floats);
      else probe<4><<<blocks, threads, smem_floats * 4>>>(d_off, d_cyc, d_sink, smem_floats);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long c[148];
    cudaMemcpy(c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)c[i];
    avg /= blocks;
    // 32 warps per block (1 block per SM at 1024 threads): cycles per warp-instruction at the SM level
    printf("%-46s %7.2f cyc / warp-LDS (SM-level)\n", p.name, avg / (double)kIters / 32.0);
  }
  return 0;
}
