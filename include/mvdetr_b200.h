/*
 * mvdetr_b200.h -- C ABI of libmvdetr_b200.so: the B200 (sm_100a) implementation of MVDeTr's
 * per-frame multiview fusion hot path (perspective warp + multi-scale deformable attention).
 *
 * Every entry point is what the reference's FFI for this path would bind. "ref:" paths are
 * relative to the reference tree (hou-yz/MVDeTr @ 66cae15), mvd/ = multiview_detector/.
 *
 * Conventions (all functions):
 *   - plain pointers + sizes, no torch / ATen types;
 *   - device pointers unless the argument name ends in `_host`;
 *   - all tensors dense, row-major, layouts given per function;
 *   - asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream),
 *     never synchronise, never allocate device memory, re-entrant, no global mutable state
 *     (exception: the `*_host` convenience entry points, documented below);
 *   - return 0 (MVD_OK) or a negative MVD_ERR_* for argument errors, or a positive cudaError_t
 *     if a CUDA runtime call / launch failed. Never throw. `mvd_error_string` decodes both.
 *     (ref prints launch errors and carries on: mvd/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:948-952;
 *      ours reports them.)
 */
#ifndef MVDETR_B200_H_
#define MVDETR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MVD_API __attribute__((visibility("default")))
#else
#define MVD_API
#endif

#define MVD_OK 0
#define MVD_ERR_NULL_POINTER (-1)   /* a required pointer is NULL                                  */
#define MVD_ERR_BAD_SHAPE (-2)      /* a dimension is <= 0 or the element count overflows int64    */
#define MVD_ERR_UNSUPPORTED (-3)    /* combination not supported by this entry point               */
#define MVD_ERR_MISALIGNED (-4)     /* pointer alignment requirement of a fast path not met        */
#define MVD_ERR_NO_DEVICE (-5)      /* no CUDA device / driver entry point unavailable             */

/* Library version: major*10000 + minor*100 + patch. */
MVD_API int mvd_version(void);

/* Human-readable text for a return code of any function below (static storage). */
MVD_API const char* mvd_error_string(int code);

/* ------------------------------------------------------------------------------------------
 * Multi-scale deformable attention, forward.
 *   replaces ms_deform_attn_cuda_forward      ref: mvd/models/ops/src/cuda/ms_deform_attn_cuda.cu:20-80
 *   (kernel ms_deformable_im2col_gpu_kernel   ref: mvd/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299)
 *   bound by pybind `ms_deform_attn_forward`  ref: mvd/models/ops/src/vision.cpp:13
 *
 *   value  [B, S, M, D]          S = sum_l H_l*W_l, level l occupies rows [start[l], start[l]+H_l*W_l)
 *   shapes [L, 2] int64 (H_l, W_l), start [L] int64        -- device memory, as the reference passes them
 *   loc    [B, Lq, M, L, P, 2]   (x, y) normalised to [0,1] over each level (outside allowed: zero padding)
 *   attn   [B, Lq, M, L, P]
 *   out    [B, Lq, M*D]          fully overwritten (no pre-zeroing needed)
 *
 *   out[b,q,m,:] = sum_l sum_p attn[b,q,m,l,p] * bilinear(value_l[b,:,m,:], x*W_l-0.5, y*H_l-0.5)
 *
 * The whole batch is processed in one launch; the reference's `im2col_step` batch chunking
 * (ms_deform_attn_cuda.cu:50-75) does not change results and is validated by the Python mirror.
 * ------------------------------------------------------------------------------------------ */
MVD_API int mvd_msda_fwd_f32(const float* value, const int64_t* shapes, const int64_t* start,
                     const float* loc, const float* attn,
                     int B, int S, int M, int D, int L, int Lq, int P,
                     float* out, void* stream);
MVD_API int mvd_msda_fwd_f64(const double* value, const int64_t* shapes, const int64_t* start,
                     const double* loc, const double* attn,
                     int B, int S, int M, int D, int L, int Lq, int P,
                     double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Multi-scale deformable attention, backward.
 *   replaces ms_deform_attn_cuda_backward     ref: mvd/models/ops/src/cuda/ms_deform_attn_cuda.cu:83-153
 *   (kernels ms_deformable_col2im_gpu_kernel_* ref: mvd/models/ops/src/cuda/ms_deform_im2col_cuda.cuh:301-920)
 *   bound by pybind `ms_deform_attn_backward` ref: mvd/models/ops/src/vision.cpp:14
 *
 *   grad_out   [B, Lq, M*D]
 *   grad_value [B, S, M, D]        zeroed by this call (cudaMemsetAsync on `stream`), then accumulated
 *                                  with floating-point atomics => summation order is not deterministic,
 *                                  exactly as in the reference (cuh:125-152)
 *   grad_loc   [B, Lq, M, L, P, 2] fully overwritten
 *   grad_attn  [B, Lq, M, L, P]    fully overwritten
 * ------------------------------------------------------------------------------------------ */
MVD_API int mvd_msda_bwd_f32(const float* grad_out, const float* value, const int64_t* shapes, const int64_t* start,
                     const float* loc, const float* attn,
                     int B, int S, int M, int D, int L, int Lq, int P,
                     float* grad_value, float* grad_loc, float* grad_attn, void* stream);
/* Same kernels for the MVDeTr encoder layout (L levels of one H x W grid, Lq = R * H * W queries; shapes / start still
 * describe it): the (query, head) pairs are walked band by band -- (batch, row, view, x, head) -- so the gradient
 * reductions of concurrently running blocks hit the same few rows of every level and stay in L2 instead of re-fetching
 * value / grad_value from DRAM once per query view. Results equal mvd_msda_bwd_f32's up to the (unspecified) order of
 * the fp32 atomic reductions. ref: ms_deform_im2col_cuda.cuh:301-920 (col2im), call site ms_deform_attn_func.py:35. */
MVD_API int mvd_msda_bwd_banded_f32(const float* grad_out, const float* value, const int64_t* shapes,
                            const int64_t* start, const float* loc, const float* attn, int B, int S, int M, int D, int L,
                            int Lq, int P, int H, int W, int R, float* grad_value, float* grad_loc, float* grad_attn,
                            void* stream);
MVD_API int mvd_msda_bwd_f64(const double* grad_out, const double* value, const int64_t* shapes, const int64_t* start,
                     const double* loc, const double* attn,
                     int B, int S, int M, int D, int L, int Lq, int P,
                     double* grad_value, double* grad_loc, double* grad_attn, void* stream);

/* ------------------------------------------------------------------------------------------
 * Forward for the MVDeTr encoder layout ("view grid"): every level is one camera view of the
 * same HxW ground-plane grid, S = L*H*W, and the Lq queries are laid out [R, H, W] (R replicas of
 * the grid, R = L in MVDeTr: mvd/models/trans_world_feat.py:92, mvd/models/mvdetr.py:129-130).
 * Same maths and same result as mvd_msda_fwd_f32; the grid knowledge (passed from the HOST, so no
 * device->host read of `shapes` is needed) lets the kernel stage per-(level, head) value windows
 * in shared memory with TMA. Sampling locations are still arbitrary: samples that fall outside the
 * staged window take a global-memory path, so the hint can only change speed, never results.
 *   value [B, L*H*W, M, D], loc [B, R*H*W, M, L, P, 2], attn [B, R*H*W, M, L, P], out [B, R*H*W, M*D]
 * Returns MVD_ERR_UNSUPPORTED when (D, P, R) has no tiled instantiation (D in {8,16,32}, P in {4,8}); callers then
 * use mvd_msda_fwd_f32. MVD_ERR_NO_DEVICE when the driver's cuTensorMapEncodeTiled cannot be resolved.
 * ------------------------------------------------------------------------------------------ */
MVD_API int mvd_msda_fwd_viewgrid_f32(const float* value, const float* loc, const float* attn,
                              int B, int H, int W, int M, int D, int L, int R, int P,
                              float* out, void* stream);

/* EXPERIMENTAL (round 1: compiled and exported, not dispatched by default, not yet measured): backward for the MVDeTr
 * encoder layout with the forward view-grid kernel's structure -- value windows and loc/attn records staged by TMA,
 * one thread per (query, head) pair, no cross-lane reductions. Same contract and results as mvd_msda_bwd_f32 with
 * S = L*H*W and Lq = R*H*W; grad_value is zeroed by the call. MVD_ERR_UNSUPPORTED for other (D, P) than
 * D in {8,16,32}, P in {4,8}.   replaces ms_deform_attn_cuda_backward  ref: mvd/models/ops/src/cuda/ms_deform_attn_cuda.cu:83-153 */
MVD_API int mvd_msda_bwd_viewgrid_f32(const float* grad_out, const float* value, const float* loc, const float* attn,
                              int B, int H, int W, int M, int D, int L, int R, int P,
                              float* grad_value, float* grad_loc, float* grad_attn, void* stream);

/* View-grid variant of mvd_msda_fused_fwd_f32 below (grid given as host ints; S = L*H*W, Lq = R*H*W).
 * `ref_lm` is the reference table in LEVEL-MAJOR order [L, Lr, P, 2] (query q reads row q % Lr of every level),
 * (Lr must equal H*W: one table row per ground cell). All pointers 16-byte aligned. The softmax is evaluated online (running max, one division at the
 * end), so results agree with mvd_msda_fused_fwd_f32 to rounding; attn_out / loc_out must be NULL here
 * (MVD_ERR_UNSUPPORTED otherwise: callers that need them use mvd_msda_fused_fwd_f32). */
MVD_API int mvd_msda_fused_fwd_viewgrid_f32(const float* value, const float* offsets, const float* logits,
                                    const float* ref_lm, const float* off_bias, const float* logit_bias,
                                    int B, int H, int W, int M, int D, int L, int R, int P, int Lr,
                                    float* out, float* attn_out, float* loc_out, void* stream);
/* Same call with offsets / logits given as COLUMN RANGES of wider rows: off_pitch / logit_pitch = floats between
 * consecutive queries (multiples of 4, >= M*L*P*2 and M*L*P). Lets ONE GEMM over the concatenated weights of
 * sampling_offsets and attention_weights (ref: ms_deform_attn.py:100-101) feed the kernel: the query rows are read and
 * split once, one launch less per layer. */
MVD_API int mvd_msda_fused_fwd_viewgrid_pitched_f32(const float* value, const float* offsets, const float* logits,
                                            const float* ref, const float* off_bias, const float* logit_bias, int B,
                                            int H, int W, int M, int D, int L, int R, int P, int Lr, int off_pitch,
                                            int logit_pitch, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused sampling-location + softmax + deformable attention, forward ("next" row).
 *   replaces, in one launch, the elementwise tail of MSDeformAttn.forward
 *       ref: mvd/models/ops/modules/ms_deform_attn.py:100-107   (softmax over L*P; loc = ref + off/(W_l,H_l))
 *   followed by MSDeformAttnFunction.forward   ref: mvd/models/ops/functions/ms_deform_attn_func.py:22-28
 *
 *   value   [B, S, M, D]
 *   offsets [B, Lq, M, L, P, 2]  raw output of the `sampling_offsets` Linear (pixels of level l)
 *   logits  [B, Lq, M, L*P]      raw output of the `attention_weights` Linear (pre-softmax)
 *   ref     [Lr, L, P, 2]        normalised reference points, shared by the batch; query q uses row q % Lr
 *                                (MVDeTr: Lr = H*W, the table of mvd/models/mvdetr.py:33-71 before its
 *                                 `.repeat([num_cam,1,1,1])` at :130)
 *   off_bias   (nullable) [M, L, P, 2]  added to `offsets` first: lets the caller run the `sampling_offsets` Linear as a
 *   logit_bias (nullable) [M, L, P]     bias-free GEMM (`attention_weights` likewise); same rounding as Linear's acc + b
 *   out     [B, Lq, M*D]
 *   attn_out (nullable) [B, Lq, M, L, P]  softmax-ed weights, written when non-NULL (needed by backward)
 *   loc_out  (nullable) [B, Lq, M, L, P, 2] sampling locations, written when non-NULL
 * ------------------------------------------------------------------------------------------ */
MVD_API int mvd_msda_fused_fwd_f32(const float* value, const int64_t* shapes, const int64_t* start,
                           const float* offsets, const float* logits, const float* ref,
                           const float* off_bias, const float* logit_bias,
                           int B, int S, int M, int D, int L, int Lq, int P, int Lr,
                           float* out, float* attn_out, float* loc_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Perspective warp (homography + bilinear gather), forward.
 *   replaces kornia.warp_perspective(src, M, dsize, mode='bilinear', padding_mode='zeros',
 *                                    align_corners=False)      call site ref: mvd/models/mvdetr.py:194-195
 *   (kornia is a third-party dependency of the reference, un-vendored and unpinned -- ref: README.md:42;
 *    the restated algorithm is in DESIGN.md and oracle/warp_ref.c)
 *
 *   src [BN, C, Hi, Wi] (NCHW)     Mat [BN, 3, 3] row-major, maps SOURCE pixel (x,y,1) -> DEST pixel
 *   dst [BN, C, Ho, Wo] (NCHW)     fully overwritten
 *
 *   Per view: A = Ndst * Mat * inv(Nsrc), T = inv(A) (normalised coords, computed in-kernel in fp64,
 *   rounded to fp32), then per dst pixel (u,v): g = (linspace(-1,1,Wo)[u], linspace(-1,1,Ho)[v]),
 *   (X,Y,Z) = T*(g,1), (x,y) = (X,Y) * (|Z|>1e-8 ? 1/(Z+1e-8) : 1),
 *   ix = ((x+1)*Wi-1)/2, iy = ((y+1)*Hi-1)/2, 4-tap bilinear with zero padding.
 *
 *   `layout` is a bit set of MVD_WARP_DST_NHWC (dst written as [BN, Ho, Wo, C]: lets the caller skip the
 *   permute-copy of ref: mvd/models/trans_world_feat.py:92) and MVD_WARP_SRC_NHWC (src read as [BN, Hi, Wi, C],
 *   C % 4 == 0, 16-byte aligned: every tap is one contiguous C-vector -- the fast path; a torch channels_last
 *   tensor has this layout). 0 = NCHW in, NCHW out (the kornia contract). Same values in every layout.
 * ------------------------------------------------------------------------------------------ */
#define MVD_WARP_DST_NHWC 1
#define MVD_WARP_SRC_NHWC 2
MVD_API int mvd_warp_fwd_f32(const float* src, const float* Mat,
                     int BN, int C, int Hi, int Wi, int Ho, int Wo,
                     float* dst, int layout, void* stream);

/* The same warp as ONE launch on an NCHW source (the backbone's layout, ref: mvd/models/mvdetr.py:177-178): per destination
 * tile the source bounding box is staged in shared memory by TMA (boxes of the 4-D NCHW tensor map, out-of-image
 * coordinates zero-filled = the op's zero padding) and gathered from there; tiles whose bounding box does not fit take
 * masked global loads in the same kernel. Bit-identical results to mvd_warp_fwd_f32 / mvd_warp_im2col_f32.
 *   mode 0: dst [BN, C, Ho, Wo]   (kornia contract)           replaces kornia.warp_perspective, mvd/models/mvdetr.py:194-195
 *   mode 1: dst [BN, Ho, Wo, C]   (skips the permute-copy of   ref: mvd/models/trans_world_feat.py:92)
 *   mode 2: dst = im2col matrix [BN*Ho2*Wo2, 9*C] of the 3x3 / `stride` / pad-1 convolution over the warped grid
 *           (stride in {1,2}; layout as mvd_warp_im2col_f32)    ref: mvd/models/trans_world_feat.py:74,89
 * MVD_ERR_UNSUPPORTED unless C % 32 == 0 and Wi % 4 == 0 (callers then use mvd_warp_fwd_f32 / mvd_warp_im2col_f32);
 * src and dst 16-byte aligned. */
MVD_API int mvd_warp_tma_f32(const float* src, const float* Mat,
                     int BN, int C, int Hi, int Wi, int Ho, int Wo,
                     float* dst, int mode, int stride, void* stream);

/* Backward of the warp w.r.t. `src` (the backbone trains through it, ref: mvd/models/mvdetr.py:177-195;
 * `Mat` carries no gradient in MVDeTr).  Replaces ATen grid_sampler_2d_backward as reached from kornia.
 *   grad_dst [BN, C, Ho, Wo]   grad_src [BN, C, Hi, Wi] zeroed by this call, then atomically accumulated. */
MVD_API int mvd_warp_bwd_f32(const float* grad_dst, const float* Mat,
                     int BN, int C, int Hi, int Wi, int Ho, int Wo,
                     float* grad_src, void* stream);
/* Same backward for channels-last tensors: grad_dst [BN, Ho, Wo, C] -> grad_src [BN, Hi, Wi, C] (C % 4 == 0),
 * one 128-bit vector reduction per tap per lane instead of C scalar atomics. */
MVD_API int mvd_warp_bwd_nhwc_f32(const float* grad_dst, const float* Mat,
                          int BN, int C, int Hi, int Wi, int Ho, int Wo,
                          float* grad_src, void* stream);

/* Batched 2-D transpose in[batch][rows][cols] -> out[batch][cols][rows] (fp32, out must not alias in): the
 * NCHW <-> NHWC relayout ([BN, C, H*W] <-> [BN, H*W, C]) for callers of the warp that hold the other layout
 * (replaces the permute-copy of ref: mvd/models/trans_world_feat.py:92 when it cannot be avoided). */
MVD_API int mvd_transpose_f32(const float* in, int batch, int rows, int cols, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused residual add + LayerNorm over the last dimension (caller-side glue of the encoder layer, eval mode):
 *   out[r,:] = LayerNorm(x[r,:] + (res[r,:] + res_bias); eps) * gamma + beta     res may be NULL (plain LayerNorm);
 *   res_bias [C] may be NULL: it is the bias of the Linear layer that produced `res` when that GEMM ran bias-free
 *   replaces `src = self.norm1(src + self.dropout1(src2))` / `self.norm2(src + self.dropout3(src2))`
 *       ref: mvd/models/deformable_transformer.py:79-80,84-85
 *   x, res, out [rows, C] (out may alias x or res); gamma, beta [C]; C % 4 == 0, C <= 1024, 16-byte aligned.
 *   perm_inner > 0 (must divide rows; out must not alias): input row o*perm_inner + i is written to output row
 *   i*(rows/perm_inner) + o, i.e. view-major tokens [N][cells] leave cell-major [cells][N], which turns the merge
 *   convolution over all views (ref: mvd/models/trans_world_feat.py:107-108) into one GEMM over contiguous rows.
 * ------------------------------------------------------------------------------------------ */
MVD_API int mvd_add_layernorm_f32(const float* x, const float* res, const float* res_bias, const float* gamma, const float* beta,
                                  int64_t rows, int C, float eps, int64_t perm_inner, float* out, void* stream);
/* Same, with a second output out2[r,:] = out[r,:] + pos[r,:] (pos, out2 [rows, C]; both or neither NULL; perm_inner must
 * be 0 with them): the next encoder layer's query `with_pos_embed(src, pos)` written by the LayerNorm that produces src,
 * replacing a separate elementwise add   ref: mvd/models/deformable_transformer.py:71-77. */
MVD_API int mvd_add_layernorm_pos_f32(const float* x, const float* res, const float* res_bias, const float* gamma,
                                      const float* beta, int64_t rows, int C, float eps, int64_t perm_inner, float* out,
                                      const float* pos, float* out2, void* stream);

/* In-place x[r, c] = act(x[r, c] + bias[c]) over [rows, C] (C % 4 == 0, 16-byte aligned), act = ReLU when `relu` != 0:
 * the bias (+ReLU) of a Linear layer whose GEMM ran bias-free.
 *   replaces the bias/activation of `self.linear1` + F.relu  ref: mvd/models/deformable_transformer.py:82
 *   and of `value_proj`                                       ref: mvd/models/ops/modules/ms_deform_attn.py:96 */
MVD_API int mvd_bias_act_f32(float* x, const float* bias, int64_t rows, int C, int relu, void* stream);

/* ------------------------------------------------------------------------------------------
 * The two 3x3 convolutions around the encoder as GEMMs: their input is produced directly in im2col order
 *   A[token][ky][kx][c]   (token = (view, oy, ox) row-major; tap (ky,kx) reads input pixel (oy*s+ky-1, ox*s+kx-1), zeros outside)
 * so that conv + bias + ReLU is one mvd_linear_f32 call with the weight reshaped to [C_out, 3*3*C_in] (ky, kx, c_in order)
 * and the GEMM's output rows are the token-major [tokens, C_out] layout the transformer consumes.
 *   mvd_warp_im2col_f32     perspective warp (arithmetic of mvd_warp_fwd_f32) of a channels-last source [BN,Hi,Wi,C]
 *                           onto the Ho x Wo grid, written as the im2col matrix of a 3x3 / stride s / pad 1 convolution:
 *                           A [BN*Ho2*Wo2, 9*C], Ho2 = (Ho-1)/s + 1.   s in {1,2}, C % 4 == 0.
 *       replaces kornia.warp_perspective + the input side of `self.downsample`
 *                           ref: mvd/models/mvdetr.py:194-195, mvd/models/trans_world_feat.py:74,89,92
 *   mvd_upsample_im2col_f32 bilinear upsample (ATen upsample_bilinear2d, align_corners=False) of a channels-last map
 *                           [BN,Hi,Wi,C] to Ho x Wo, written as the im2col matrix of a 3x3 / stride 1 / pad 1
 *                           convolution: A [BN*Ho*Wo, 9*C].
 *       replaces nn.Upsample + the input side of the 3x3 conv   ref: mvd/models/trans_world_feat.py:83-84,109
 * A is fully overwritten (padding slots included).
 * ------------------------------------------------------------------------------------------ */
MVD_API int mvd_warp_im2col_f32(const float* src, const float* Mat, int BN, int C, int Hi, int Wi, int Ho, int Wo,
                        int stride, float* A, void* stream);
MVD_API int mvd_upsample_im2col_f32(const float* src, int BN, int C, int Hi, int Wi, int Ho, int Wo, float* A, void* stream);
/* Same, rows [row0, row0 + nrows) of the upsampled grid only: A [BN*nrows*Wo, 9*C] (row-sharded tail of the multi-GPU
 * path: every rank convolves its own band of ground-plane rows; the band's halo rows are computed, not exchanged). */
MVD_API int mvd_upsample_im2col_rows_f32(const float* src, int BN, int C, int Hi, int Wi, int Ho, int Wo, int row0,
                                 int nrows, float* A, void* stream);

/* ------------------------------------------------------------------------------------------
 * Input side of the path: decoded camera frames -> normalised, resized network input in one kernel.
 *   replaces T.Compose([T.ToTensor(), T.Normalize(mean, std), T.Resize((Ho, Wo))])   ref: mvd/datasets/frameDataset.py:66-67
 *   img [N, Hi, Wi, 3] uint8 (HWC, as decoded)  ->  out [N, 3, Ho, Wo] fp32, out[n,c] = resize(((img/255) - mean[c]) / std[c])
 *   mean_host / std_host: 3 floats each in HOST memory (read during the call).
 *   antialias != 0: torchvision >= 0.17's behaviour on tensors, F.interpolate(bilinear, antialias=True) = ATen
 *   _upsample_bilinear2d_aa (triangle filter of support Hi/Ho when down-scaling, normalised weights, separable);
 *   antialias == 0: plain bilinear, align_corners=False. MVD_ERR_UNSUPPORTED for antialiased down-scaling beyond 7.5x.
 * ------------------------------------------------------------------------------------------ */
MVD_API int mvd_resize_normalize_u8(const unsigned char* img, int N, int Hi, int Wi, int Ho, int Wo,
                            const float* mean_host, const float* std_host, int antialias, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Output side of the path: ground-plane heatmap -> detections on the GPU, so that only the kept
 * detections cross to the host (the reference copies both maps to the CPU and loops in Python).
 *   mvd_decode_candidates_f32: for every cell with sigmoid(heatmap) > cls_thres writes (in arrival order)
 *       cand_cell  [B, cap]     row-major cell number y*W + x
 *       cand_pos   [B, cap, 2]  ((x + off_x) * reduce, (y + off_y) * reduce), swapped to (row, col) when swap_xy != 0
 *                               (offset NULL: + 0.5 instead); each operation rounded separately, as torch does
 *       cand_score [B, cap]     sigmoid(heatmap)
 *       cand_count [B]          number of such cells (zeroed by the call; may exceed cap: the excess is dropped,
 *                               cap = H*W can never overflow)
 *     heatmap [B, 1, H, W] logits, offset [B, 2, H, W] (x, y) or NULL.
 *     replaces mvdet_decode(torch.sigmoid(world_heatmap), world_offset, reduce)  ref: mvd/utils/decode.py:80-93
 *              + `ids = scores > cls_thres; pos, s = positions[b, ids], scores[b, ids, 0]`  ref: mvd/trainer.py:121-133
 *   mvd_distance_nms_f32: per batch element, reorders the candidates into row-major cell order (out_cell / out_pos /
 *     out_score [B, cap(,2)]: the arrays the reference's boolean mask produces) and runs greedy distance NMS over
 *     them: visit by descending score (ties: larger candidate number first), keep, drop every other candidate whose
 *     distance to it is not > dist_thres; only the top_k best take part (top_k <= 0: all).
 *       keep [B, cap] candidate numbers in the order kept, keep_count [B]
 *     workspace: device scratch of mvd_distance_nms_workspace_bytes(B, cap) bytes, 8-byte aligned.
 *     replaces nms(pos, s, 20, np.inf)   ref: mvd/utils/nms.py:7-44, call site mvd/trainer.py:134
 * ------------------------------------------------------------------------------------------ */
MVD_API size_t mvd_distance_nms_workspace_bytes(int B, int cap);
MVD_API int mvd_decode_candidates_f32(const float* heatmap, const float* offset, int B, int H, int W, float reduce,
                              float cls_thres, int swap_xy, int cap, int* cand_count, int* cand_cell,
                              float* cand_pos, float* cand_score, void* stream);
MVD_API int mvd_distance_nms_f32(const int* cand_count, const int* cand_cell, const float* cand_pos,
                         const float* cand_score, int B, int cap, float dist_thres, int top_k,
                         void* workspace, size_t workspace_bytes, int* out_cell, float* out_pos,
                         float* out_score, int* keep, int* keep_count, void* stream);

/* ------------------------------------------------------------------------------------------
 * Linear layer out[rows, N] = act(x[rows, K] @ W[N, K]^T + bias[N]) through the CUDA toolkit's cuBLASLt (>= 12.9, loaded
 * by absolute path at first use: /usr/local/cuda/lib64/libcublasLt.so.12 or $MVD_CUBLASLT). Library GEMM, not a kernel
 * of ours; what it buys on B200 is `precision` = 1: CUBLAS_COMPUTE_32F_EMULATED_16BFX9, fp32 operands split into three
 * bf16 terms, nine tensor-core products, fp32 accumulation -- fp32-level accuracy at tensor-core speed. `precision` = 0
 * is native fp32 (CUBLAS_COMPUTE_32F) from the same library.
 *   replaces the six nn.Linear calls per encoder layer  ref: mvd/models/ops/modules/ms_deform_attn.py:96,100-101,116,
 *                                                        mvd/models/deformable_transformer.py:82
 *   bias nullable; relu != 0 applies ReLU in the GEMM epilogue; all pointers 16-byte aligned; `workspace` (device,
 *   nullable) is scratch the caller owns for the duration of the call on `stream`.
 * Exception to the "no global state" rule: a process-wide cuBLASLt handle and a plan cache (mutex protected).
 * Returns MVD_ERR_NO_DEVICE when no suitable cuBLASLt can be loaded, MVD_ERR_UNSUPPORTED when the library has no
 * algorithm for the request (callers then use their own GEMM path).
 * mvd_linear_available(): the loaded cuBLASLt version (e.g. 120901), 0 if none.
 * ------------------------------------------------------------------------------------------ */
MVD_API int mvd_linear_available(void);
MVD_API int mvd_linear_f32(const float* x, const float* W, const float* bias, int64_t rows, int K, int N, int relu,
                   int precision, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * The same Linear layer as OUR kernel on the 5th-generation tensor cores (tcgen05.mma.kind::tf32, accumulators in tensor
 * memory, operands staged by TMA): out[rows, N] = act(x[rows, K] @ W[N, K]^T + bias[N]) with fp32-level accuracy through a
 * 3xTF32 split -- x = x_hi + x_lo is split inside the kernel, the weight once by mvd_tf32_split_f32 (hi = RN_tf32(w),
 * lo = w - hi), out = x_hi w_hi + x_hi w_lo + x_lo w_hi accumulated in fp32 (max error vs fp64 below a native fp32 GEMM).
 * One launch per layer: no operand scan, no separate bias pass, no workspace.
 *   replaces the nn.Linear calls of  ref: mvd/models/ops/modules/ms_deform_attn.py:96,100-101,116,
 *                                         mvd/models/deformable_transformer.py:82
 *   K % 4 == 0, N % 4 == 0, all pointers 16-byte aligned; bias nullable; relu != 0 applies ReLU.
 * ------------------------------------------------------------------------------------------ */
MVD_API int mvd_tf32_split_f32(const float* w, int64_t n, float* hi, float* lo, void* stream);
/* Preferred variant: three-term bf16 split (a = a0 + a1 + a2), the six products with i + j <= 2 through
 * tcgen05.mma.kind::f16 -- bf16 products are exact in the tensor core's adder, so the error is at cuBLASLt BF16x9's level
 * (below a native fp32 GEMM) with 6 instead of 9 products. Persistent, warp-specialised kernel: TMA producer warp, MMA
 * warp, 4 splitter warps, 4 epilogue warps; 3-stage operand ring; two accumulators in tensor memory so a tile's
 * epilogue overlaps the next tile's main loop.
 *   mvd_bf16_split3_f32: w [n] fp32 -> terms [3][n] bf16 (device, 16-byte aligned), once per weight.
 *   mvd_linear_bf16x3_f32: x [rows, K] fp32, w_terms [3][N][K] bf16; K % 8 == 0, N % 4 == 0. */
MVD_API int mvd_bf16_split3_f32(const float* w, int64_t n, void* terms, void* stream);
MVD_API int mvd_linear_bf16x3_f32(const float* x, const void* w_terms, const float* bias, int64_t rows, int K, int N,
                          int relu, float* out, void* stream);
/* Multi-GPU: the same GEMM with the all-gather fused into its epilogue. `out_mc` is the NVLink MULTICAST
 * address of a symmetric buffer slot; results leave as multimem.st, replicated by the NVSwitch into every GPU's copy,
 * tile by tile under the next tile's main loop (replaces GEMM -> ncclAllGather of the per-layer `value` rows,
 * ref for the data flow: mvd/models/ops/modules/ms_deform_attn.py:96). mvd_multicast_copy_f32 does the same for a tensor
 * [rows, C] produced elsewhere; with inner > 0 it also transposes [outer][inner] rows into [inner][outer_total] rows
 * (view-major tokens -> cell-major rows for the merge conv, ref: mvd/models/trans_world_feat.py:107-108).
 * Consumers on other GPUs must be separated from the producers by a cross-GPU barrier. */
MVD_API int mvd_linear_bf16x3_multicast_f32(const float* x, const void* w_terms, const float* bias, int64_t rows, int K,
                                    int N, int relu, float* out_mc, void* stream);
MVD_API int mvd_multicast_copy_f32(const float* src, float* dst_mc, int64_t rows, int C, int64_t inner,
                           int64_t outer_total, int64_t outer0, void* stream);
/* 3x3 / pad 1 / stride {1, 2} convolution of a channels-last fp32 image src [NB][Hi][Wi][C] (C % 32 == 0) as an
 * IMPLICIT GEMM on the same tcgen05 kernel: a row block is a tile of output pixels of one image, a K chunk is 32
 * channels of one tap fetched by TMA straight from the image (4-D tensor map with traversal stride = conv stride,
 * out-of-image coordinates zero-filled = the padding). No im2col matrix exists in memory. w_terms: the split of the
 * [N][9 C] weight matrix with columns ordered (ky, kx, c) (terms = 2: mvd_f16_split2_f32, 3: mvd_bf16_split3_f32).
 * out [NB * Ho * Wo][N] channels-last, Ho = (Hi - 1) / stride + 1. Bit-identical to mvd_linear_* over the im2col matrix
 * of mvd_warp_im2col_f32 / mvd_upsample_im2col_f32. Replaces self.downsample and the upsample conv
 * (ref: multiview_detector/models/trans_world_feat.py:74,82-84,89,109). MVD_ERR_UNSUPPORTED when the output width has
 * no divisor in [16, 128] (callers keep the im2col route).
 * mvd_upsample_nhwc_f32: bilinear upsample (align_corners = false, ATen arithmetic) of a channels-last map, output rows
 * [row0, row0 + nrows) only: the input of that convolution (ref: trans_world_feat.py:83 nn.Upsample).
 * addend / out2 (both or neither, [NB * Ho * Wo][N]): out2 = out + addend from the same epilogue -- the first encoder
 * layer's query `src + pos` (ref: multiview_detector/models/deformable_transformer.py:71-77) without a kernel of its own. */
MVD_API int mvd_conv3x3_nhwc_f32(const float* src, const void* w_terms, const float* bias, int NB, int Hi, int Wi, int C,
                         int stride, int N, int relu, int terms, float* out, const float* addend, float* out2,
                         void* stream);
MVD_API int mvd_upsample_nhwc_f32(const float* src, int BN, int C, int Hi, int Wi, int Ho, int Wo, int row0, int nrows,
                          float* dst, void* stream);
/* Same two GEMMs (same arguments, bit-identical results) with the three bf16 terms of x staged in TENSOR MEMORY by the
 * splitter warps (tcgen05.st) and consumed by tcgen05.mma's A-from-TMEM form: 40 % less shared-memory traffic per K
 * chunk, and for K <= 128 the terms of a 128-row block are reused by every 128-column tile of the output (x is read and
 * split once per row block for the N = 224 / 448 / 512 layers, ref: ms_deform_attn.py:100-101, deformable_transformer.py:82). */
MVD_API int mvd_linear_bf16x3_ts_f32(const float* x, const void* w_terms, const float* bias, int64_t rows, int K, int N,
                             int relu, float* out, void* stream);
MVD_API int mvd_linear_bf16x3_ts_multicast_f32(const float* x, const void* w_terms, const float* bias, int64_t rows, int K,
                                       int N, int relu, float* out_mc, void* stream);
/* Half the tensor work: TWO fp16 terms per operand (a = h0 + 2^-11 h1, |a - h0 - 2^-11 h1| <= 2^-24 |a|) and the THREE
 * products h0*g0 (own accumulator), h0*g1 + h1*g0 (second accumulator, scaled by 2^-11 in the epilogue); fp16 products
 * are exact in the tensor core's fp32 adder. Same kernel structure as the _ts_ variant (terms of x in tensor memory).
 * Range: |x| and |w| below 65504 (fp16); beyond that the result is non-finite, never silently wrong. Same layers as
 * above (ref: ms_deform_attn.py:96,100-101,116; deformable_transformer.py:82; trans_world_feat.py:74,82-84).
 *   mvd_f16_split2_f32: w [n] fp32 -> terms [2][n] fp16 (device, 16-byte aligned), once per weight.
 *   mvd_linear_f16x2_f32: x [rows, K] fp32, w_terms [2][N][K] fp16; K % 8 == 0, N % 4 == 0. */
MVD_API int mvd_f16_split2_f32(const float* w, int64_t n, void* terms, void* stream);
MVD_API int mvd_linear_f16x2_f32(const float* x, const void* w_terms, const float* bias, int64_t rows, int K, int N,
                         int relu, float* out, void* stream);
MVD_API int mvd_linear_f16x2_multicast_f32(const float* x, const void* w_terms, const float* bias, int64_t rows, int K,
                                   int N, int relu, float* out_mc, void* stream);
MVD_API int mvd_linear_tf32x3_f32(const float* x, const float* w_hi, const float* w_lo, const float* bias,
                          int64_t rows, int K, int N, int relu, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Host-buffer convenience entry points (used for end-to-end timing and by non-PyTorch callers):
 * same arguments, but every pointer is HOST memory (pinned or pageable). They allocate device
 * scratch, copy in, run the kernel above, copy the result back and synchronise `stream` before
 * returning. `shapes`/`start` are host int64 arrays here.
 * ------------------------------------------------------------------------------------------ */
MVD_API int mvd_msda_fwd_f32_host(const float* value_host, const int64_t* shapes_host, const int64_t* start_host,
                          const float* loc_host, const float* attn_host,
                          int B, int S, int M, int D, int L, int Lq, int P,
                          float* out_host, void* stream);
MVD_API int mvd_warp_fwd_f32_host(const float* src_host, const float* Mat_host,
                          int BN, int C, int Hi, int Wi, int Ho, int Wo,
                          float* dst_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MVDETR_B200_H_ */
