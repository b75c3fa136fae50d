// Output side of the path: ground-plane heatmap -> detections, on the GPU.
//   mvd_decode_candidates_f32   sigmoid + offset decode + score threshold + compaction
//       replaces mvdet_decode(torch.sigmoid(world_heatmap.cpu()), world_offset.cpu(), reduce)  ref: multiview_detector/utils/decode.py:80-93
//       and the per-frame threshold `scores > cls_thres`, `positions[b, ids]`                   ref: multiview_detector/trainer.py:121-133
//   mvd_distance_nms_f32        greedy distance-based non-maximum suppression
//       replaces nms(pos, s, 20, np.inf)                                                        ref: multiview_detector/utils/nms.py:7-44, trainer.py:134
// The reference moves both maps to the host and runs a Python while-loop with one torch.norm per kept detection; here
// only the kept detections leave the GPU.
//
// Semantics kept (checked bit-for-bit on indices against the reference's functions, tests/test_decode_gpu.py):
//   * position of cell (y, x): ((x + off_x) * reduce, (y + off_y) * reduce), two separately rounded fp32 operations
//     (decode.py:84-90: `xy = xy + offset; xy *= reduce`); swapped to (row, col) order for 'ij' datasets (trainer.py:127-130);
//   * candidates are the cells with sigmoid(h) > cls_thres, numbered in row-major cell order (boolean-mask order);
//   * NMS visits candidates by descending score (equal scores: larger candidate number first, which is what the
//     reference's stable ascending sort read from the back does), keeps one, and drops every remaining candidate whose
//     Euclidean distance to it is NOT > dist_thres; only the top_k best candidates take part (nms.py:26-27).
#include "common.cuh"

namespace mvd {
namespace {

constexpr int kDecThreads = 256;
constexpr int kNmsThreads = 1024;
constexpr int kNmsSmemKeys = 4096;  // candidates sorted in shared memory up to this many (32 KB of keys)

__global__ void __launch_bounds__(kDecThreads) decode_candidates_kernel(
    const float* __restrict__ heat, const float* __restrict__ offset, int H, int W, float reduce, float thres,
    int swap_xy, int cap, int* __restrict__ count, int* __restrict__ cell, float* __restrict__ pos,
    float* __restrict__ score) {
  const int b = blockIdx.y;
  const int idx = blockIdx.x * kDecThreads + threadIdx.x;
  if (idx >= H * W) return;
  const float h = __ldg(heat + (int64_t)b * H * W + idx);
  const float s = 1.f / (1.f + expf(-h));  // torch.sigmoid (full-precision division and exp)
  if (!(s > thres)) return;
  const int y = idx / W, x = idx - y * W;
  float px = (float)x, py = (float)y;
  if (offset) {
    px = __fadd_rn(px, __ldg(offset + ((int64_t)b * 2 + 0) * H * W + idx));
    py = __fadd_rn(py, __ldg(offset + ((int64_t)b * 2 + 1) * H * W + idx));
  } else {
    px = __fadd_rn(px, 0.5f);
    py = __fadd_rn(py, 0.5f);
  }
  px = __fmul_rn(px, reduce);
  py = __fmul_rn(py, reduce);
  const int slot = atomicAdd(count + b, 1);  // arrival order; mvd_distance_nms_f32 restores row-major order
  if (slot < cap) {
    const int64_t o = (int64_t)b * cap + slot;
    cell[o] = idx;
    score[o] = s;
    pos[2 * o] = swap_xy ? py : px;
    pos[2 * o + 1] = swap_xy ? px : py;
  }
}

// In-place ascending bitonic sort of n_pad (power of two) 64-bit keys by the whole block.
__device__ void bitonic_sort(unsigned long long* keys, int n_pad) {
  for (int k = 2; k <= n_pad; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const unsigned long long a = keys[i], c = keys[p];
          const bool up = (i & k) == 0;
          if ((a > c) == up) {
            keys[i] = c;
            keys[p] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ unsigned orderable(float f) {  // monotone map float -> unsigned
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// One block per batch element. In: the unordered candidates of decode_candidates_kernel. Out: the candidates in
// row-major cell order (o_*: the reference's `pos` / `s` arrays after the boolean mask) and the kept candidate numbers.
__global__ void __launch_bounds__(kNmsThreads) distance_nms_kernel(
    const int* __restrict__ count, const int* __restrict__ cell, const float* __restrict__ pos,
    const float* __restrict__ score, int cap, int cap_pad, float dist_thres, int top_k,
    unsigned long long* __restrict__ ws_keys, unsigned char* __restrict__ ws_supp, int* __restrict__ o_cell,
    float* __restrict__ o_pos, float* __restrict__ o_score, int* __restrict__ keep, int* __restrict__ keep_count) {
  __shared__ unsigned long long s_keys[kNmsSmemKeys];
  __shared__ int s_next, s_kept;
  const int b = blockIdx.x;
  const int n = min(count[b], cap);
  int n_pad = 1;
  while (n_pad < n) n_pad <<= 1;
  unsigned long long* keys = n_pad <= kNmsSmemKeys ? s_keys : ws_keys + (int64_t)b * cap_pad;
  unsigned char* supp = ws_supp + (int64_t)b * cap_pad;
  const int64_t base = (int64_t)b * cap;
  if (n == 0) {
    if (threadIdx.x == 0) keep_count[b] = 0;
    return;
  }
  // ---- 1. row-major order: sort (cell, slot) ----
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x)
    keys[i] = i < n ? (((unsigned long long)(unsigned)cell[base + i] << 32) | (unsigned)i) : ~0ull;
  __syncthreads();
  bitonic_sort(keys, n_pad);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int slot = (int)(keys[i] & 0xffffffffu);
    o_cell[base + i] = cell[base + slot];
    o_score[base + i] = score[base + slot];
    o_pos[2 * (base + i)] = pos[2 * (base + slot)];
    o_pos[2 * (base + i) + 1] = pos[2 * (base + slot) + 1];
  }
  __syncthreads();
  // ---- 2. visiting order: descending score, ties -> larger candidate number first ----
  for (int i = threadIdx.x; i < n_pad; i += blockDim.x)
    keys[i] = i < n ? (((unsigned long long)(~orderable(o_score[base + i])) << 32) | (unsigned)(0xffffffffu - (unsigned)i))
                    : ~0ull;
  for (int i = threadIdx.x; i < n; i += blockDim.x) supp[i] = 0;
  __syncthreads();
  bitonic_sort(keys, n_pad);
  // ---- 3. greedy suppression ----
  const int limit = (top_k > 0 && top_k < n) ? top_k : n;
  if (threadIdx.x == 0) s_kept = 0;
  int start = 0;  // used by thread 0 only
  while (true) {
    if (threadIdx.x == 0) {
      int k = start;
      while (k < limit && supp[k]) ++k;
      s_next = k;
      if (k < limit) keep[base + s_kept++] = (int)(0xffffffffu - (unsigned)(keys[k] & 0xffffffffu));
    }
    __syncthreads();
    const int k = s_next;
    if (k >= limit) break;
    const int i = (int)(0xffffffffu - (unsigned)(keys[k] & 0xffffffffu));
    const float tx = o_pos[2 * (base + i)], ty = o_pos[2 * (base + i) + 1];
    for (int j = k + 1 + threadIdx.x; j < limit; j += blockDim.x) {
      if (supp[j]) continue;
      const int c = (int)(0xffffffffu - (unsigned)(keys[j] & 0xffffffffu));
      const float dx = __fsub_rn(tx, o_pos[2 * (base + c)]), dy = __fsub_rn(ty, o_pos[2 * (base + c) + 1]);
      const float d = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      if (!(d > dist_thres)) supp[j] = 1;  // nms.py:43 keeps `dists > dist_thres`
    }
    __syncthreads();  // marks visible to thread 0's scan; everyone has read s_next
    start = k + 1;
  }
  if (threadIdx.x == 0) keep_count[b] = s_kept;
}

int pad_pow2(int n) {
  int p = 1;
  while (p < n) p <<= 1;
  return p;
}

}  // namespace
}  // namespace mvd

using namespace mvd;

extern "C" size_t mvd_distance_nms_workspace_bytes(int B, int cap) {
  if (B <= 0 || cap <= 0) return 0;
  return (size_t)B * pad_pow2(cap) * (sizeof(unsigned long long) + 1) + 16;
}

extern "C" int mvd_decode_candidates_f32(const float* heatmap, const float* offset, int B, int H, int W, float reduce,
                                         float cls_thres, int swap_xy, int cap, int* cand_count, int* cand_cell,
                                         float* cand_pos, float* cand_score, void* stream) {
  if (!heatmap || !cand_count || !cand_cell || !cand_pos || !cand_score) return MVD_ERR_NULL_POINTER;
  if (B <= 0 || H <= 0 || W <= 0 || cap <= 0 || B > 65535 || (int64_t)H * W > 0x3fffffffLL) return MVD_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  MVD_CUDA_TRY(cudaMemsetAsync(cand_count, 0, sizeof(int) * (size_t)B, st));
  dim3 grid((unsigned)ceil_div64((int64_t)H * W, kDecThreads), (unsigned)B);
  decode_candidates_kernel<<<grid, kDecThreads, 0, st>>>(heatmap, offset, H, W, reduce, cls_thres, swap_xy, cap,
                                                         cand_count, cand_cell, cand_pos, cand_score);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}

extern "C" int mvd_distance_nms_f32(const int* cand_count, const int* cand_cell, const float* cand_pos,
                                    const float* cand_score, int B, int cap, float dist_thres, int top_k,
                                    void* workspace, size_t workspace_bytes, int* out_cell, float* out_pos,
                                    float* out_score, int* keep, int* keep_count, void* stream) {
  if (!cand_count || !cand_cell || !cand_pos || !cand_score || !workspace || !out_cell || !out_pos || !out_score ||
      !keep || !keep_count)
    return MVD_ERR_NULL_POINTER;
  if (B <= 0 || cap <= 0 || B > 65535) return MVD_ERR_BAD_SHAPE;
  if (workspace_bytes < mvd_distance_nms_workspace_bytes(B, cap)) return MVD_ERR_BAD_SHAPE;
  if (reinterpret_cast<uintptr_t>(workspace) & 7u) return MVD_ERR_MISALIGNED;
  const int cap_pad = pad_pow2(cap);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(workspace);
  unsigned char* supp = reinterpret_cast<unsigned char*>(keys + (size_t)B * cap_pad);
  distance_nms_kernel<<<B, kNmsThreads, 0, (cudaStream_t)stream>>>(cand_count, cand_cell, cand_pos, cand_score, cap,
                                                                  cap_pad, dist_thres, top_k, keys, supp, out_cell,
                                                                  out_pos, out_score, keep, keep_count);
  MVD_LAUNCH_CHECK();
  return MVD_OK;
}
