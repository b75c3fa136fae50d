// C-ABI glue that is not a kernel: version, error strings and the host-buffer convenience entry points
// declared in include/mvdetr_b200.h.
#include <stdio.h>

#include "common.cuh"

extern "C" int mvd_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char* mvd_error_string(int code) {
  switch (code) {
    case MVD_OK:
      return "ok";
    case MVD_ERR_NULL_POINTER:
      return "mvdetr_b200: a required pointer argument is NULL";
    case MVD_ERR_BAD_SHAPE:
      return "mvdetr_b200: a dimension is non-positive or too large";
    case MVD_ERR_UNSUPPORTED:
      return "mvdetr_b200: argument combination not supported by this entry point";
    case MVD_ERR_MISALIGNED:
      return "mvdetr_b200: pointer alignment requirement not met";
    case MVD_ERR_NO_DEVICE:
      return "mvdetr_b200: no CUDA device or driver entry point unavailable";
    default:
      break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "mvdetr_b200: unknown error code";
}

namespace {

// Stream-ordered scratch that frees itself on every return path.
struct DevBuf {
  void* p = nullptr;
  cudaStream_t st;
  explicit DevBuf(cudaStream_t s) : st(s) {}
  cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes, st); }
  ~DevBuf() {
    if (p) cudaFreeAsync(p, st);
  }
};

}  // namespace

extern "C" int mvd_msda_fwd_f32_host(const float* value_host, const int64_t* shapes_host,
                                     const int64_t* start_host, const float* loc_host, const float* attn_host,
                                     int B, int S, int M, int D, int L, int Lq, int P, float* out_host,
                                     void* stream) {
  if (!value_host || !shapes_host || !start_host || !loc_host || !attn_host || !out_host)
    return MVD_ERR_NULL_POINTER;
  if (B <= 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0 || Lq <= 0 || P <= 0) return MVD_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t nv = (size_t)B * S * M * D, na = (size_t)B * Lq * M * L * P, no = (size_t)B * Lq * M * D;
  DevBuf dv(st), ds(st), dl(st), da(st), dout(st);
  MVD_CUDA_TRY(dv.alloc(nv * 4));
  MVD_CUDA_TRY(ds.alloc((size_t)L * 3 * 8));
  MVD_CUDA_TRY(dl.alloc(na * 8));
  MVD_CUDA_TRY(da.alloc(na * 4));
  MVD_CUDA_TRY(dout.alloc(no * 4));
  int64_t* dshapes = (int64_t*)ds.p;
  int64_t* dstart = dshapes + 2 * L;
  MVD_CUDA_TRY(cudaMemcpyAsync(dv.p, value_host, nv * 4, cudaMemcpyHostToDevice, st));
  MVD_CUDA_TRY(cudaMemcpyAsync(dshapes, shapes_host, (size_t)L * 16, cudaMemcpyHostToDevice, st));
  MVD_CUDA_TRY(cudaMemcpyAsync(dstart, start_host, (size_t)L * 8, cudaMemcpyHostToDevice, st));
  MVD_CUDA_TRY(cudaMemcpyAsync(dl.p, loc_host, na * 8, cudaMemcpyHostToDevice, st));
  MVD_CUDA_TRY(cudaMemcpyAsync(da.p, attn_host, na * 4, cudaMemcpyHostToDevice, st));
  int rc = mvd_msda_fwd_f32((const float*)dv.p, dshapes, dstart, (const float*)dl.p, (const float*)da.p, B, S, M,
                            D, L, Lq, P, (float*)dout.p, stream);
  if (rc != MVD_OK) return rc;
  MVD_CUDA_TRY(cudaMemcpyAsync(out_host, dout.p, no * 4, cudaMemcpyDeviceToHost, st));
  MVD_CUDA_TRY(cudaStreamSynchronize(st));
  return MVD_OK;
}

extern "C" int mvd_warp_fwd_f32_host(const float* src_host, const float* Mat_host, int BN, int C, int Hi, int Wi,
                                     int Ho, int Wo, float* dst_host, void* stream) {
  if (!src_host || !Mat_host || !dst_host) return MVD_ERR_NULL_POINTER;
  if (BN <= 0 || C <= 0 || Hi <= 0 || Wi <= 0 || Ho <= 0 || Wo <= 0) return MVD_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t ns = (size_t)BN * C * Hi * Wi, nd = (size_t)BN * C * Ho * Wo;
  DevBuf dsrc(st), dm(st), ddst(st);
  MVD_CUDA_TRY(dsrc.alloc(ns * 4));
  MVD_CUDA_TRY(dm.alloc((size_t)BN * 36));
  MVD_CUDA_TRY(ddst.alloc(nd * 4));
  MVD_CUDA_TRY(cudaMemcpyAsync(dsrc.p, src_host, ns * 4, cudaMemcpyHostToDevice, st));
  MVD_CUDA_TRY(cudaMemcpyAsync(dm.p, Mat_host, (size_t)BN * 36, cudaMemcpyHostToDevice, st));
  int rc = mvd_warp_fwd_f32((const float*)dsrc.p, (const float*)dm.p, BN, C, Hi, Wi, Ho, Wo, (float*)ddst.p, 0,
                            stream);
  if (rc != MVD_OK) return rc;
  MVD_CUDA_TRY(cudaMemcpyAsync(dst_host, ddst.p, nd * 4, cudaMemcpyDeviceToHost, st));
  MVD_CUDA_TRY(cudaStreamSynchronize(st));
  return MVD_OK;
}
