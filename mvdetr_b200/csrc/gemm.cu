// Dense glue of the encoder layer: out = act(x @ W^T + bias) for the six Linear layers per layer
//   ref: multiview_detector/models/ops/modules/ms_deform_attn.py:96,100-101,116 (value_proj, sampling_offsets,
//        attention_weights, output_proj), multiview_detector/models/deformable_transformer.py:82 (linear1/2)
// These are plain library GEMMs (M = 75 600 tokens, K, N <= 512) and stay in cuBLASLt. What this file adds is HOW the
// library is driven on B200: torch's bundled cuBLAS 12.8 runs fp32 GEMMs on the SIMT pipe (cutlass simt sgemm,
// 40-45 TFLOP/s measured, profiles/r01g_launches.csv: 2.5 ms of a 5.0 ms frame). cuBLASLt 12.9 (shipped with the CUDA
// toolkit of this image) has CUBLAS_COMPUTE_32F_EMULATED_16BFX9: every fp32 operand is split into three bf16 terms
// (8+8+8 mantissa bits) and the 9 partial products run on the tcgen05 tensor cores with fp32 accumulation --
// fp32-level accuracy (checked against fp64 in tests/test_gemm_gpu.py) at tensor-core speed.
// The toolkit's libcublasLt is loaded by absolute path at first use (dlopen; it becomes a second, private copy next
// to the one torch links -- the two never share handles), so libmvdetr_b200.so keeps having no link-time dependency.
// `precision` 0 asks the same library for native fp32 (CUBLAS_COMPUTE_32F).
#include <cublasLt.h>
#include <dlfcn.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace mvd {
namespace {

struct LtApi {
  void* so = nullptr;
  cublasStatus_t (*Create)(cublasLtHandle_t*) = nullptr;
  cublasStatus_t (*DescCreate)(cublasLtMatmulDesc_t*, cublasComputeType_t, cudaDataType_t) = nullptr;
  cublasStatus_t (*DescSet)(cublasLtMatmulDesc_t, cublasLtMatmulDescAttributes_t, const void*, size_t) = nullptr;
  cublasStatus_t (*DescDestroy)(cublasLtMatmulDesc_t) = nullptr;
  cublasStatus_t (*LayoutCreate)(cublasLtMatrixLayout_t*, cudaDataType, uint64_t, uint64_t, int64_t) = nullptr;
  cublasStatus_t (*LayoutDestroy)(cublasLtMatrixLayout_t) = nullptr;
  cublasStatus_t (*PrefCreate)(cublasLtMatmulPreference_t*) = nullptr;
  cublasStatus_t (*PrefSet)(cublasLtMatmulPreference_t, cublasLtMatmulPreferenceAttributes_t, const void*,
                            size_t) = nullptr;
  cublasStatus_t (*PrefDestroy)(cublasLtMatmulPreference_t) = nullptr;
  cublasStatus_t (*Heuristic)(cublasLtHandle_t, cublasLtMatmulDesc_t, cublasLtMatrixLayout_t, cublasLtMatrixLayout_t,
                              cublasLtMatrixLayout_t, cublasLtMatrixLayout_t, cublasLtMatmulPreference_t, int,
                              cublasLtMatmulHeuristicResult_t*, int*) = nullptr;
  cublasStatus_t (*Matmul)(cublasLtHandle_t, cublasLtMatmulDesc_t, const void*, const void*, cublasLtMatrixLayout_t,
                           const void*, cublasLtMatrixLayout_t, const void*, const void*, cublasLtMatrixLayout_t,
                           void*, cublasLtMatrixLayout_t, const cublasLtMatmulAlgo_t*, void*, size_t,
                           cudaStream_t) = nullptr;
  size_t (*GetVersion)(void) = nullptr;
  cublasStatus_t (*Destroy)(cublasLtHandle_t) = nullptr;
  cublasLtHandle_t handle = nullptr;  // handle of the device the CURRENT call runs on (see handle_for)
  bool ok = false;
};

struct Plan {
  cublasLtMatmulDesc_t desc = nullptr;
  cublasLtMatrixLayout_t a = nullptr, b = nullptr, c = nullptr;
  cublasLtMatmulAlgo_t algo;
  size_t workspace = 0;
  bool valid = false;
};

std::mutex g_mu;
LtApi g_lt;
bool g_lt_tried = false;
// key: rows, K, N, epilogue flags (1 = bias, 2 = relu), precision, device
std::map<std::tuple<int64_t, int, int, int, int, int>, Plan> g_plans;  // invalid plans are cached too (negative result)
std::map<int, cublasLtHandle_t> g_handles;                             // one cuBLASLt handle per device (context)

// Lock held, library loaded. A cuBLASLt handle belongs to the context it was created in: one per device.
cublasLtHandle_t handle_for(int dev) {
  auto it = g_handles.find(dev);
  if (it != g_handles.end()) return it->second;
  cublasLtHandle_t h = nullptr;
  if (g_lt.Create(&h) != CUBLAS_STATUS_SUCCESS) h = nullptr;
  g_handles[dev] = h;
  return h;
}

template <typename F>
bool sym(void* so, const char* name, F* out) {
  *out = reinterpret_cast<F>(dlsym(so, name));
  return *out != nullptr;
}

// Lock held. Loads <toolkit>/lib64/libcublasLt.so.12 by absolute path ($MVD_CUBLASLT overrides).
bool load_lt() {
  if (g_lt_tried) return g_lt.ok;
  g_lt_tried = true;
  const char* env = getenv("MVD_CUBLASLT");
  const char* cands[] = {env, "/usr/local/cuda/lib64/libcublasLt.so.12", "/usr/local/cuda/targets/x86_64-linux/lib/libcublasLt.so.12"};
  for (const char* path : cands) {
    if (!path || !path[0]) continue;
    void* so = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!so) continue;
    LtApi api;
    api.so = so;
    const bool all = sym(so, "cublasLtCreate", &api.Create) && sym(so, "cublasLtMatmulDescCreate", &api.DescCreate) &&
                     sym(so, "cublasLtMatmulDescSetAttribute", &api.DescSet) &&
                     sym(so, "cublasLtMatmulDescDestroy", &api.DescDestroy) &&
                     sym(so, "cublasLtMatrixLayoutCreate", &api.LayoutCreate) &&
                     sym(so, "cublasLtMatrixLayoutDestroy", &api.LayoutDestroy) &&
                     sym(so, "cublasLtMatmulPreferenceCreate", &api.PrefCreate) &&
                     sym(so, "cublasLtMatmulPreferenceSetAttribute", &api.PrefSet) &&
                     sym(so, "cublasLtMatmulPreferenceDestroy", &api.PrefDestroy) &&
                     sym(so, "cublasLtMatmulAlgoGetHeuristic", &api.Heuristic) &&
                     sym(so, "cublasLtMatmul", &api.Matmul) && sym(so, "cublasLtGetVersion", &api.GetVersion);
    if (!all || api.GetVersion() < 120900) {
      dlclose(so);
      continue;
    }
    api.ok = true;
    g_lt = api;
    return true;
  }
  return false;
}

// Lock held. Row-major out[rows, N] = x[rows, K] @ W[N, K]^T is, in cuBLAS' column-major terms,
// C'(N x rows, ld N) = op_T(A = W as K x N, ld K) * (B = x as K x rows, ld K).
int make_plan(int64_t rows, int K, int N, int epi, int precision, const float* bias, size_t ws_avail, Plan* out) {
  Plan p;
  const cublasComputeType_t ct = precision ? CUBLAS_COMPUTE_32F_EMULATED_16BFX9 : CUBLAS_COMPUTE_32F;
  if (g_lt.DescCreate(&p.desc, ct, CUDA_R_32F) != CUBLAS_STATUS_SUCCESS) return MVD_ERR_UNSUPPORTED;
  const cublasOperation_t tr = CUBLAS_OP_T, no = CUBLAS_OP_N;
  g_lt.DescSet(p.desc, CUBLASLT_MATMUL_DESC_TRANSA, &tr, sizeof(tr));
  g_lt.DescSet(p.desc, CUBLASLT_MATMUL_DESC_TRANSB, &no, sizeof(no));
  cublasLtEpilogue_t e = CUBLASLT_EPILOGUE_DEFAULT;
  if ((epi & 1) && (epi & 2)) e = CUBLASLT_EPILOGUE_RELU_BIAS;
  else if (epi & 1) e = CUBLASLT_EPILOGUE_BIAS;
  else if (epi & 2) e = CUBLASLT_EPILOGUE_RELU;
  bool okc = g_lt.DescSet(p.desc, CUBLASLT_MATMUL_DESC_EPILOGUE, &e, sizeof(e)) == CUBLAS_STATUS_SUCCESS;
  if (epi & 1) okc = okc && g_lt.DescSet(p.desc, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias)) == CUBLAS_STATUS_SUCCESS;
  okc = okc && g_lt.LayoutCreate(&p.a, CUDA_R_32F, (uint64_t)K, (uint64_t)N, K) == CUBLAS_STATUS_SUCCESS;
  okc = okc && g_lt.LayoutCreate(&p.b, CUDA_R_32F, (uint64_t)K, (uint64_t)rows, K) == CUBLAS_STATUS_SUCCESS;
  okc = okc && g_lt.LayoutCreate(&p.c, CUDA_R_32F, (uint64_t)N, (uint64_t)rows, N) == CUBLAS_STATUS_SUCCESS;
  cublasLtMatmulPreference_t pref = nullptr;
  okc = okc && g_lt.PrefCreate(&pref) == CUBLAS_STATUS_SUCCESS;
  int found = 0;
  cublasLtMatmulHeuristicResult_t res[16];
  if (okc) {
    g_lt.PrefSet(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &ws_avail, sizeof(ws_avail));
    if (g_lt.Heuristic(g_lt.handle, p.desc, p.a, p.b, p.c, p.c, pref, 4, res, &found) !=
        CUBLAS_STATUS_SUCCESS)
      found = 0;
  }
  if (pref) g_lt.PrefDestroy(pref);
  int pick = -1;
  for (int i = 0; i < found && pick < 0; ++i)
    if (res[i].state == CUBLAS_STATUS_SUCCESS && res[i].workspaceSize <= ws_avail) pick = i;
  if (pick < 0) {
    if (p.a) g_lt.LayoutDestroy(p.a);
    if (p.b) g_lt.LayoutDestroy(p.b);
    if (p.c) g_lt.LayoutDestroy(p.c);
    if (p.desc) g_lt.DescDestroy(p.desc);
    return MVD_ERR_UNSUPPORTED;
  }
  p.algo = res[pick].algo;
  p.workspace = res[pick].workspaceSize;
  p.valid = true;
  *out = p;
  return MVD_OK;
}

}  // namespace
}  // namespace mvd

using namespace mvd;

extern "C" int mvd_linear_available(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  return load_lt() ? (int)g_lt.GetVersion() : 0;
}

extern "C" int mvd_linear_f32(const float* x, const float* W, const float* bias, int64_t rows, int K, int N, int relu,
                              int precision, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !W || !out) return MVD_ERR_NULL_POINTER;
  if (rows <= 0 || K <= 0 || N <= 0 || rows > 0x7fffffffLL) return MVD_ERR_BAD_SHAPE;
  if (precision != 0 && precision != 1) return MVD_ERR_UNSUPPORTED;
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(out) |
                       reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(workspace);
  if (al & 15u) return MVD_ERR_MISALIGNED;
  int dev = 0;
  MVD_CUDA_TRY(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu);
  if (!load_lt()) return MVD_ERR_NO_DEVICE;
  g_lt.handle = handle_for(dev);
  if (!g_lt.handle) return MVD_ERR_NO_DEVICE;
  const int epi = (bias ? 1 : 0) | (relu ? 2 : 0);
  const auto key = std::make_tuple(rows, K, N, epi, precision, dev);
  auto it = g_plans.find(key);
  if (it == g_plans.end()) {
    Plan p;
    const int e = make_plan(rows, K, N, epi, precision, bias, workspace ? workspace_bytes : 0, &p);
    if (e != MVD_OK && e != MVD_ERR_UNSUPPORTED) return e;
    it = g_plans.emplace(key, p).first;  // p.valid == false records "no algorithm": the query is not repeated
  }
  Plan& p = it->second;
  if (!p.valid) return MVD_ERR_UNSUPPORTED;
  if (p.workspace > (workspace ? workspace_bytes : 0)) return MVD_ERR_UNSUPPORTED;
  if (epi & 1) g_lt.DescSet(p.desc, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bias, sizeof(bias));
  const float alpha = 1.f, beta = 0.f;
  const cublasStatus_t st = g_lt.Matmul(g_lt.handle, p.desc, &alpha, W, p.a, x, p.b, &beta, out, p.c, out, p.c, &p.algo,
                                        workspace, workspace ? workspace_bytes : 0, (cudaStream_t)stream);
  if (st != CUBLAS_STATUS_SUCCESS) return MVD_ERR_UNSUPPORTED;
  return MVD_OK;
}
