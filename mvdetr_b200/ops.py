"""Host-side mirror of the reference's operator interface for the hot path, over the C ABI.

  ms_deform_attn_forward / ms_deform_attn_backward
        same names, argument order and error behaviour as the pybind module `MultiScaleDeformableAttention`
        (ref: multiview_detector/models/ops/src/vision.cpp:13-15, ms_deform_attn.h:20-61,
         cuda/ms_deform_attn_cuda.cu:20-153)
  MSDeformAttnFunction
        same autograd.Function as ref: multiview_detector/models/ops/functions/ms_deform_attn_func.py:21-38
  msda_fused_forward
        extra entry point: loc/softmax arithmetic of ms_deform_attn.py:100-107 inside the kernel
  warp_perspective
        kornia.warp_perspective signature as used at ref: multiview_detector/models/mvdetr.py:194-195

PyTorch is used for device memory and streams only; all arithmetic happens in libmvdetr_b200.so.
"""
import os
import weakref

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _C

_VIEWGRID = os.environ.get("MVDETR_B200_VIEWGRID", "1") != "0"
_BWD_VIEWGRID = os.environ.get("MVDETR_B200_BWD_VIEWGRID", "0") == "1"  # experimental TMA-staged backward (opt-in)
_BWD_BANDED = os.environ.get("MVDETR_B200_BWD_BANDED", "1") == "1"  # encoder layout: band-by-band pair order in the backward
_WARP_CL = os.environ.get("MVDETR_B200_WARP_CL", "1") != "0"  # 0: always the scalar NCHW-source warp kernels (A/B switch)
_WARP_TMA = os.environ.get("MVDETR_B200_WARP_TMA", "1") != "0"  # 0: relayout + channels-last gather (round-1 path)


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class _on_device:
    """Cheap device guard: only switches when the tensor lives on a non-current device."""

    def __init__(self, t):
        self.idx = t.device.index
        self.prev = None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if self.idx is not None and self.idx != cur:
            self.prev = cur
            torch.cuda.set_device(self.idx)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)
        return False


def _check_msda_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step,
                       grad_output=None):
    """The reference's preconditions, same order and wording (ms_deform_attn.h:29-38, ms_deform_attn_cuda.cu:28-52)."""
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    named = [("value", value), ("spatial_shapes", spatial_shapes), ("level_start_index", level_start_index),
             ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)]
    if grad_output is not None:
        named.append(("grad_output", grad_output))
    for name, t in named:
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")
    for name, t in named:
        if not t.is_cuda:
            raise RuntimeError(f"{name} must be a CUDA tensor")
    if value.dtype not in (torch.float32, torch.float64):
        raise RuntimeError(f'"ms_deform_attn_forward_cuda" not implemented for \'{value.dtype}\'')
    for name, t in named[3:]:
        if t.dtype != value.dtype:
            raise RuntimeError(f"{name} must have the dtype of value ({value.dtype}), got {t.dtype}")
    if spatial_shapes.dtype != torch.int64 or level_start_index.dtype != torch.int64:
        raise RuntimeError("spatial_shapes and level_start_index must be int64 tensors")
    if value.dim() != 4 or sampling_loc.dim() != 6 or attn_weight.dim() != 5:
        raise RuntimeError("expected value [B,S,M,D], sampling_loc [B,Lq,M,L,P,2], attn_weight [B,Lq,M,L,P]")
    B, S, M, D = value.shape
    L = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    if tuple(sampling_loc.shape) != (B, Lq, M, L, P, 2) or tuple(attn_weight.shape) != (B, Lq, M, L, P):
        raise RuntimeError(f"inconsistent shapes: value {tuple(value.shape)}, sampling_loc "
                           f"{tuple(sampling_loc.shape)}, attn_weight {tuple(attn_weight.shape)}, levels {L}")
    if level_start_index.numel() != L:
        raise RuntimeError("level_start_index must have one entry per level")
    step = min(B, int(im2col_step))
    if step <= 0 or B % step != 0:
        raise RuntimeError(f"batch({B}) must divide im2col_step({step})")
    if grad_output is not None and grad_output.numel() != B * Lq * M * D:
        raise RuntimeError("grad_output must have B*Lq*M*D elements")
    return B, S, M, D, L, Lq, P


class _TensorCache:
    """Derived data cached per tensor OBJECT (weak reference + version), never per address: the caching allocator hands
    the same data_ptr to a new tensor as soon as the old one dies, so an address key would serve stale entries."""

    def __init__(self):
        self._d = {}

    def get(self, t):
        ent = self._d.get(id(t))
        if ent is not None and ent[0]() is t and ent[1] == t._version:
            return ent[2]
        return None

    def put(self, t, payload):
        key = id(t)
        self._d[key] = (weakref.ref(t, lambda _r, k=key, d=self._d: d.pop(k, None)), t._version, payload)
        return payload


_shape_cache = _TensorCache()


def _host_shapes(spatial_shapes):
    """Host copy of the [L,2] int64 device tensor, cached on the tensor object (and its version): one device->host read
    the first time a shapes tensor is seen, none afterwards (keeps the autograd path graph-capturable once warm)."""
    got = _shape_cache.get(spatial_shapes)
    if got is None:
        got = _shape_cache.put(spatial_shapes, spatial_shapes.tolist())
    return got


def _viewgrid_geometry(value, spatial_shapes, S, L, Lq):
    """Returns (H, W, R) when every level is the same HxW grid and the queries are R copies of that grid
    (MVDeTr's encoder layout), else None. Reads `spatial_shapes` back to the host (one small sync; the reference's
    own caller already synchronises on the same tensor at ms_deform_attn.py:94)."""
    if not _VIEWGRID or value.dtype != torch.float32 or S % L != 0:
        return None
    hw = S // L
    if Lq % hw != 0:
        return None
    shapes = _host_shapes(spatial_shapes)
    H, W = shapes[0]
    if H * W != hw or any(s != [H, W] for s in shapes):
        return None
    return H, W, Lq // hw


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    """-> output [B, Lq, M*D]. Drop-in for MSDA.ms_deform_attn_forward (ms_deform_attn_func.py:25)."""
    B, S, M, D, L, Lq, P = _check_msda_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                              im2col_step)
    out = torch.empty((B, Lq, M * D), dtype=value.dtype, device=value.device)
    if Lq == 0:  # no queries: the reference returns its (empty) zero-initialised output (ms_deform_attn_cuda.cu:54)
        return out
    if S == 0:   # no keys: every sample is outside every level
        return out.zero_()
    with _on_device(value):
        if value.dtype == torch.float32:
            geo = _viewgrid_geometry(value, spatial_shapes, S, L, Lq)
            if geo is not None:
                H, W, R = geo
                rc = _C.lib.mvd_msda_fwd_viewgrid_f32(value.data_ptr(), sampling_loc.data_ptr(),
                                                      attn_weight.data_ptr(), B, H, W, M, D, L, R, P,
                                                      out.data_ptr(), _stream(value))
                if rc == 0:
                    return out
                if rc != -3:  # MVD_ERR_UNSUPPORTED -> generic kernel
                    _C.check(rc, "mvd_msda_fwd_viewgrid_f32")
            fn = _C.lib.mvd_msda_fwd_f32
        else:
            fn = _C.lib.mvd_msda_fwd_f64
        rc = fn(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(), sampling_loc.data_ptr(),
                attn_weight.data_ptr(), B, S, M, D, L, Lq, P, out.data_ptr(), _stream(value))
    _C.check(rc, "mvd_msda_fwd")
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output,
                            im2col_step):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight]. Drop-in for MSDA.ms_deform_attn_backward."""
    B, S, M, D, L, Lq, P = _check_msda_inputs(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                                              im2col_step, grad_output)
    grad_value = torch.empty_like(value)  # zeroed inside the C call
    grad_loc = torch.empty_like(sampling_loc)
    grad_attn = torch.empty_like(attn_weight)
    if Lq == 0 or S == 0:  # nothing sampled: all gradients are zero (ms_deform_attn_cuda.cu:121-123 zero-fills them)
        return [grad_value.zero_(), grad_loc.zero_(), grad_attn.zero_()]
    if _BWD_VIEWGRID and value.dtype == torch.float32:  # experimental opt-in (round 1: not the default path)
        geo = _viewgrid_geometry(value, spatial_shapes, S, L, Lq)
        if geo is not None:
            H, W, R = geo
            with _on_device(value):
                rc = _C.lib.mvd_msda_bwd_viewgrid_f32(grad_output.data_ptr(), value.data_ptr(), sampling_loc.data_ptr(),
                                                      attn_weight.data_ptr(), B, H, W, M, D, L, R, P,
                                                      grad_value.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(),
                                                      _stream(value))
            if rc == 0:
                return [grad_value, grad_loc, grad_attn]
            if rc != -3:
                _C.check(rc, "mvd_msda_bwd_viewgrid_f32")
    if _BWD_BANDED and value.dtype == torch.float32:  # encoder layout: same kernels, pairs walked band by band (L2-resident reds)
        geo = _viewgrid_geometry(value, spatial_shapes, S, L, Lq)
        if geo is not None:
            H, W, R = geo
            with _on_device(value):
                rc = _C.lib.mvd_msda_bwd_banded_f32(grad_output.data_ptr(), value.data_ptr(), spatial_shapes.data_ptr(),
                                                    level_start_index.data_ptr(), sampling_loc.data_ptr(),
                                                    attn_weight.data_ptr(), B, S, M, D, L, Lq, P, H, W, R,
                                                    grad_value.data_ptr(), grad_loc.data_ptr(), grad_attn.data_ptr(),
                                                    _stream(value))
            _C.check(rc, "mvd_msda_bwd_banded_f32")
            return [grad_value, grad_loc, grad_attn]
    fn = _C.lib.mvd_msda_bwd_f32 if value.dtype == torch.float32 else _C.lib.mvd_msda_bwd_f64
    with _on_device(value):
        rc = fn(grad_output.data_ptr(), value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                sampling_loc.data_ptr(), attn_weight.data_ptr(), B, S, M, D, L, Lq, P, grad_value.data_ptr(),
                grad_loc.data_ptr(), grad_attn.data_ptr(), _stream(value))
    _C.check(rc, "mvd_msda_bwd")
    return [grad_value, grad_loc, grad_attn]


class MSDeformAttnFunction(Function):
    """Same contract as ref ms_deform_attn_func.py:21-38: apply(value, value_spatial_shapes,
    value_level_start_index, sampling_locations, attention_weights, im2col_step)."""

    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step):
        ctx.im2col_step = im2col_step
        output = ms_deform_attn_forward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                        attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, start, loc, attn = ctx.saved_tensors
        grad_value, grad_loc, grad_attn = ms_deform_attn_backward(value, shapes, start, loc, attn,
                                                                  grad_output.contiguous(), ctx.im2col_step)
        return grad_value, None, None, grad_loc, grad_attn, None


def msda_viewgrid_forward(value, sampling_loc, attn_weight, H, W):
    """Forward for the MVDeTr encoder layout with the grid given as host ints (no device->host read):
    value [B, L*H*W, M, D], sampling_loc [B, R*H*W, M, L, P, 2], attn_weight [B, R*H*W, M, L, P]."""
    B, S, M, D = value.shape
    Lq, L, P = sampling_loc.shape[1], sampling_loc.shape[3], sampling_loc.shape[4]
    if S != L * H * W or Lq % (H * W) != 0:
        raise RuntimeError("msda_viewgrid_forward: value/queries are not L resp. R copies of an HxW grid")
    for name, t in (("value", value), ("sampling_loc", sampling_loc), ("attn_weight", attn_weight)):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError(f"{name} must be a contiguous fp32 CUDA tensor")
    out = torch.empty((B, Lq, M * D), dtype=value.dtype, device=value.device)
    with _on_device(value):
        rc = _C.lib.mvd_msda_fwd_viewgrid_f32(value.data_ptr(), sampling_loc.data_ptr(), attn_weight.data_ptr(), B, H,
                                              W, M, D, L, Lq // (H * W), P, out.data_ptr(), _stream(value))
    _C.check(rc, "mvd_msda_fwd_viewgrid_f32")
    return out


def msda_fused_forward(value, spatial_shapes, level_start_index, offsets, logits, ref_table, want_aux=False,
                       grid_hw=None, ref_table_lm=None, off_bias=None, logit_bias=None):
    """out = MSDA(value, loc = ref + offsets/(W_l,H_l), attn = softmax(logits)) in one kernel (inference path).

    value [B,S,M,D] fp32; offsets [B,Lq,M,L,P,2] and logits [B,Lq,M,L*P]: raw Linear outputs; ref_table [Lr,L,P,2]
    (query q reads row q % Lr). Returns out [B,Lq,M*D], plus (attn, loc) when want_aux.
    grid_hw=(H, W) (host ints) asserts the MVDeTr encoder layout -- every level an HxW grid, S = L*H*W, Lq = R*H*W --
    and selects the TMA-staged view-grid kernel; the generic kernel is used when that layout has no instantiation.
    ref_table_lm: optional precomputed level-major copy of ref_table, [L,Lr,P,2] (made on the fly when None).
    off_bias [M*L*P*2] / logit_bias [M*L*P]: optional biases of the two Linear layers, added in the kernel so the
    caller can run both as bias-free GEMMs (cuBLASLt applies an fp32 bias in a separate pass over the output).
    offsets / logits may also be the two column ranges of ONE contiguous [B*Lq, M*L*P*3] GEMM output (views with that row
    pitch): the view-grid kernel reads them in place through strided tensor maps."""
    B, S, M, D = value.shape
    _, Lq, _, L, P, _ = offsets.shape
    for name, t in (("value", value), ("ref_table", ref_table)):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError(f"{name} must be a contiguous fp32 CUDA tensor")
    off_pitch, log_pitch = M * L * P * 2, M * L * P
    pitched = False
    if not (offsets.is_contiguous() and logits.is_contiguous()):
        # accepted: rows [B*Lq] with a common pitch, dense inside a row
        o2, l2 = offsets.reshape(B * Lq, off_pitch) if offsets.dim() == 6 else offsets, logits.reshape(B * Lq, log_pitch)
        if not (o2.stride(1) == 1 and l2.stride(1) == 1 and o2.stride(0) == l2.stride(0) and o2.stride(0) % 4 == 0 and
                o2.data_ptr() % 16 == 0 and l2.data_ptr() % 16 == 0 and o2.data_ptr() == offsets.data_ptr()):
            raise RuntimeError("msda_fused_forward: offsets / logits must be contiguous or column ranges of one row-major buffer")
        off_pitch = log_pitch = o2.stride(0)
        pitched = True
        if want_aux or grid_hw is None or not _VIEWGRID:  # only the view-grid kernel reads strided rows
            offsets, logits, pitched = offsets.contiguous(), logits.contiguous(), False
            off_pitch, log_pitch = M * L * P * 2, M * L * P
    for name, t in (("offsets", offsets), ("logits", logits)):
        if not (t.is_cuda and t.dtype == torch.float32):
            raise RuntimeError(f"{name} must be a fp32 CUDA tensor")
    if logits.numel() != B * Lq * M * L * P or tuple(ref_table.shape[1:]) != (L, P, 2):
        raise RuntimeError("msda_fused_forward: inconsistent shapes")
    for name, t, n in (("off_bias", off_bias, M * L * P * 2), ("logit_bias", logit_bias, M * L * P)):
        if t is not None and not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32 and t.numel() == n):
            raise RuntimeError(f"msda_fused_forward: {name} must be a contiguous fp32 CUDA tensor of {n} elements")
    ob = off_bias.data_ptr() if off_bias is not None else None
    lb = logit_bias.data_ptr() if logit_bias is not None else None
    out = torch.empty((B, Lq, M * D), dtype=value.dtype, device=value.device)
    attn = torch.empty((B, Lq, M, L, P), dtype=value.dtype, device=value.device) if want_aux else None
    loc = torch.empty_like(offsets) if want_aux else None
    aux = (attn.data_ptr() if want_aux else None, loc.data_ptr() if want_aux else None)
    with _on_device(value):
        if grid_hw is not None and _VIEWGRID and not want_aux:
            H, W = int(grid_hw[0]), int(grid_hw[1])
            if S != L * H * W or Lq % (H * W) != 0:
                raise RuntimeError("msda_fused_forward: grid_hw does not match value/queries")
            if ref_table_lm is None:
                ref_table_lm = ref_table.permute(1, 0, 2, 3).contiguous()
            elif tuple(ref_table_lm.shape) != (L, ref_table.shape[0], P, 2) or not ref_table_lm.is_contiguous():
                raise RuntimeError("msda_fused_forward: ref_table_lm must be contiguous [L, Lr, P, 2]")
            if pitched:
                rc = _C.lib.mvd_msda_fused_fwd_viewgrid_pitched_f32(value.data_ptr(), offsets.data_ptr(),
                                                                    logits.data_ptr(), ref_table_lm.data_ptr(), ob, lb, B,
                                                                    H, W, M, D, L, Lq // (H * W), P, ref_table.shape[0],
                                                                    off_pitch, log_pitch, out.data_ptr(), _stream(value))
            else:
                rc = _C.lib.mvd_msda_fused_fwd_viewgrid_f32(value.data_ptr(), offsets.data_ptr(), logits.data_ptr(),
                                                            ref_table_lm.data_ptr(), ob, lb, B, H, W, M, D, L,
                                                            Lq // (H * W), P,
                                                            ref_table.shape[0], out.data_ptr(), None, None,
                                                            _stream(value))
            if rc == 0:
                return (out, attn, loc) if want_aux else out
            if rc != -3:  # MVD_ERR_UNSUPPORTED -> generic kernel
                _C.check(rc, "mvd_msda_fused_fwd_viewgrid_f32")
        if pitched:  # the generic kernel reads dense tensors
            offsets, logits = offsets.contiguous(), logits.contiguous()
        rc = _C.lib.mvd_msda_fused_fwd_f32(value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
                                           offsets.data_ptr(), logits.data_ptr(), ref_table.data_ptr(), ob, lb, B, S,
                                           M, D, L, Lq, P, ref_table.shape[0], out.data_ptr(), *aux, _stream(value))
    _C.check(rc, "mvd_msda_fused_fwd_f32")
    return (out, attn, loc) if want_aux else out


def add_layer_norm(x, res, weight, bias, eps=1e-5, res_bias=None, perm_inner=0, pos=None):
    """LayerNorm(x + (res + res_bias)) over the last dim in one kernel (inference glue of the encoder layer,
    ref: multiview_detector/models/deformable_transformer.py:79-80,84-85). res may be None; res_bias [C] is the bias
    of the Linear that produced `res` when that GEMM was run bias-free. perm_inner > 0: rows [outer][inner] are written
    as [inner][outer] (same shape returned; see mvd_add_layernorm_f32). pos (same shape as x): additionally returns
    out + pos, the next layer's query -> (out, out + pos)."""
    C = x.shape[-1]
    for name, t in (("x", x), ("res", res), ("weight", weight), ("bias", bias), ("res_bias", res_bias)):
        if t is not None and not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError(f"add_layer_norm: {name} must be a contiguous fp32 CUDA tensor")
    if res is not None and res.shape != x.shape:
        raise RuntimeError("add_layer_norm: x and res must have the same shape")
    out = torch.empty_like(x)
    out2 = None
    if pos is not None:
        if not (pos.is_cuda and pos.is_contiguous() and pos.dtype == torch.float32 and pos.numel() == x.numel()):
            raise RuntimeError("add_layer_norm: pos must be a contiguous fp32 CUDA tensor with x's element count")
        out2 = torch.empty_like(x)
    with _on_device(x):
        rc = _C.lib.mvd_add_layernorm_pos_f32(x.data_ptr(), res.data_ptr() if res is not None else None,
                                              res_bias.data_ptr() if res_bias is not None else None, weight.data_ptr(),
                                              bias.data_ptr(), x.numel() // C, C, float(eps), int(perm_inner),
                                              out.data_ptr(), pos.data_ptr() if pos is not None else None,
                                              out2.data_ptr() if out2 is not None else None, _stream(x))
    _C.check(rc, "mvd_add_layernorm_pos_f32")
    return out if pos is None else (out, out2)


_DST_NHWC, _SRC_NHWC = 1, 2  # MVD_WARP_* layout bits of include/mvdetr_b200.h


def _tma_warp_ok(src):
    """The one-launch TMA warp takes a plain NCHW source with C % 32 == 0 and Wi % 4 == 0 (16-byte row pitch)."""
    return (_WARP_TMA and src.is_contiguous() and src.shape[1] % 32 == 0 and src.shape[3] % 4 == 0 and
            src.data_ptr() % 16 == 0)


def _warp_tma(src, mat, Ho, Wo, dst, mode, stride=1):
    BN, C, Hi, Wi = src.shape
    with _on_device(src):
        rc = _C.lib.mvd_warp_tma_f32(src.data_ptr(), mat.data_ptr(), BN, C, Hi, Wi, Ho, Wo, dst.data_ptr(), mode,
                                     int(stride), _stream(src))
    if rc == -3:
        return False
    _C.check(rc, "mvd_warp_tma_f32")
    return True


def warp_launch_names(src, channels_last=False, im2col=False):
    """Kernels one warp call launches for this source (bench.py reports it next to the timing)."""
    if _tma_warp_ok(src):
        return "warp_tma_kernel<%s> (1 launch)" % ("IM2COL" if im2col else "NHWC" if channels_last else "NCHW")
    cl_src = src.is_contiguous(memory_format=torch.channels_last) and not src.is_contiguous()
    pre = "" if cl_src else "mvd_transpose_f32 + "
    if im2col:
        return pre + "warp_im2col_kernel"
    return pre + ("warp_fwd_cl_kernel<NHWC dst>" if channels_last else "warp_fwd_cl_kernel<NCHW dst>")


def warp_launch_count(im2col=True):
    """Launches of the warp stage per frame on an NCHW source (1 with the TMA kernel, else relayout + gather)."""
    return 1 if _WARP_TMA else 2


def warp_im2col(src, M, dsize, stride=2):
    """Perspective warp of src [BN,C,Hi,Wi] (as warp_perspective) written as the im2col matrix of a 3x3 / stride /
    pad-1 convolution over the warped grid: returns A [BN*Ho2*Wo2, 9*C] (token-major, taps (ky,kx,c)) and (Ho2, Wo2).
    Inference only (no autograd)."""
    BN, C, Hi, Wi = src.shape
    Ho, Wo = int(dsize[0]), int(dsize[1])
    if not (src.is_cuda and src.dtype == torch.float32 and C % 4 == 0):
        raise RuntimeError("warp_im2col: fp32 CUDA source with C % 4 == 0 required")
    mat = M.detach().to(device=src.device, dtype=torch.float32).contiguous()
    Ho2, Wo2 = (Ho - 1) // stride + 1, (Wo - 1) // stride + 1
    A = torch.empty((BN * Ho2 * Wo2, 9 * C), dtype=src.dtype, device=src.device)
    if _tma_warp_ok(src) and _warp_tma(src, mat, Ho, Wo, A, 2, stride):  # NCHW source: one launch, TMA-staged
        return A, (Ho2, Wo2)
    src_cl, _ = _as_nhwc(src)
    with _on_device(src):
        rc = _C.lib.mvd_warp_im2col_f32(src_cl.data_ptr(), mat.data_ptr(), BN, C, Hi, Wi, Ho, Wo, int(stride),
                                        A.data_ptr(), _stream(src))
    _C.check(rc, "mvd_warp_im2col_f32")
    return A, (Ho2, Wo2)


def upsample_im2col(x_cl, dsize, rows=None):
    """x_cl [BN,Hi,Wi,C] channels-last fp32 CUDA -> im2col matrix [BN*Ho*Wo, 9*C] of a 3x3 / stride-1 / pad-1 convolution
    over its bilinear upsample (align_corners=False) to dsize. rows=(row0, nrows): only that band of output rows
    ([BN*nrows*Wo, 9*C]; the view-sharded tail)."""
    BN, Hi, Wi, C = x_cl.shape
    Ho, Wo = int(dsize[0]), int(dsize[1])
    if not (x_cl.is_cuda and x_cl.is_contiguous() and x_cl.dtype == torch.float32 and C % 4 == 0):
        raise RuntimeError("upsample_im2col: contiguous fp32 CUDA [BN,Hi,Wi,C] with C % 4 == 0 required")
    row0, nrows = (0, Ho) if rows is None else (int(rows[0]), int(rows[1]))
    A = torch.empty((BN * nrows * Wo, 9 * C), dtype=x_cl.dtype, device=x_cl.device)
    with _on_device(x_cl):
        rc = _C.lib.mvd_upsample_im2col_rows_f32(x_cl.data_ptr(), BN, C, Hi, Wi, Ho, Wo, row0, nrows, A.data_ptr(),
                                                 _stream(x_cl))
    _C.check(rc, "mvd_upsample_im2col_rows_f32")
    return A


def transpose_last2(x):
    """[batch, rows, cols] fp32 CUDA contiguous -> [batch, cols, rows] contiguous (mvd_transpose_f32): the NCHW <-> NHWC
    relayout kernel."""
    batch, rows, cols = x.shape
    out = torch.empty((batch, cols, rows), dtype=x.dtype, device=x.device)
    with _on_device(x):
        rc = _C.lib.mvd_transpose_f32(x.data_ptr(), batch, rows, cols, out.data_ptr(), _stream(x))
    _C.check(rc, "mvd_transpose_f32")
    return out


def _as_nhwc(x):
    """NCHW-shaped fp32 CUDA tensor -> (tensor whose storage is [N,H,W,C] contiguous, made_copy)."""
    N, C, H, W = x.shape
    if x.is_contiguous(memory_format=torch.channels_last) and not (C == 1 or H * W == 1):
        return x.permute(0, 2, 3, 1), False
    x = x.contiguous()
    return transpose_last2(x.view(N, C, H * W)).view(N, H, W, C), True


# "bf16x3" / "tf32x3": OUR tcgen05 kernels (split operands, fp32-level accuracy) | "bf16x9" / "fp32": cuBLASLt 12.9 | "torch"
_GEMM_MODE = os.environ.get("MVDETR_B200_GEMM", "f16x2")
# bf16x3 only: x terms staged in tensor memory (tcgen05.mma with A from TMEM) instead of shared memory; same results
_GEMM_TS = os.environ.get("MVDETR_B200_GEMM_TS", "1") == "1"
_gemm_ws = {}
_tf32_split_cache = _TensorCache()
_bf16_split_cache = _TensorCache()
_f16_split_cache = _TensorCache()


def _bf16_split3(weight):
    """[3, N, K] bf16 terms of an fp32 weight (t0 = bf16(w), t1 = bf16(w - t0), t2 = bf16(w - t0 - t1)), computed on the
    device once per weight object and version."""
    got = _bf16_split_cache.get(weight)
    if got is None:
        w = weight.detach()
        terms = torch.empty((3, *w.shape), dtype=torch.bfloat16, device=w.device)
        with _on_device(w):
            rc = _C.lib.mvd_bf16_split3_f32(w.data_ptr(), w.numel(), terms.data_ptr(), _stream(w))
        _C.check(rc, "mvd_bf16_split3_f32")
        got = _bf16_split_cache.put(weight, terms)
    return got


def _f16_split2(weight):
    """[2, N, K] fp16 terms of an fp32 weight (t0 = fp16(w), t1 = fp16((w - t0) * 2^11)), once per weight object and version."""
    got = _f16_split_cache.get(weight)
    if got is None:
        w = weight.detach()
        terms = torch.empty((2, *w.shape), dtype=torch.float16, device=w.device)
        with _on_device(w):
            rc = _C.lib.mvd_f16_split2_f32(w.data_ptr(), w.numel(), terms.data_ptr(), _stream(w))
        _C.check(rc, "mvd_f16_split2_f32")
        got = _f16_split_cache.put(weight, terms)
    return got


def _tf32_split(weight):
    """(hi, lo) fp32 tensors with hi = RN_tf32(weight), lo = weight - hi; computed on the device once per weight object
    and version (a few hundred KB each). Pass the SAME tensor object every call (a Parameter, a cached derived weight):
    a fresh view per call is split again each time."""
    got = _tf32_split_cache.get(weight)
    if got is None:
        w = weight.detach()
        hi, lo = torch.empty_like(w), torch.empty_like(w)
        with _on_device(w):
            rc = _C.lib.mvd_tf32_split_f32(w.data_ptr(), w.numel(), hi.data_ptr(), lo.data_ptr(), _stream(w))
        _C.check(rc, "mvd_tf32_split_f32")
        got = _tf32_split_cache.put(weight, (hi, lo))
    return got


def linear_available():
    """cuBLASLt version behind ops.linear (>= 120900), or 0 when the toolkit library cannot be loaded."""
    return int(_C.lib.mvd_linear_available())


def linear(x, weight, bias=None, relu=False, mode=None, out=None):
    """act(x @ weight.T + bias) for x [rows, K], weight [N, K] (nn.Linear layout), fp32 CUDA, through mvd_linear_f32.
    mode "bf16x9": fp32 emulated on the tensor cores (cuBLASLt 12.9 CUBLAS_COMPUTE_32F_EMULATED_16BFX9, fp32-level
    accuracy); "fp32": the same library's native fp32; "torch": torch.mm + our bias kernel (explicit opt-in with
    MVDETR_B200_GEMM=torch or mode="torch"). There is NO silent fallback: when the toolkit's cuBLASLt cannot be loaded
    or has no algorithm for the request, this raises and names the switch."""
    mode = mode or _GEMM_MODE
    rows, K = x.shape
    N = weight.shape[0]
    x = x.contiguous()
    if mode in ("tf32x3", "bf16x3", "bf16x3ts", "bf16x3ss", "f16x2"):
        for name, t in (("x", x), ("weight", weight), ("bias", bias)):
            if t is not None and not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise RuntimeError(f"linear: {name} must be a contiguous fp32 CUDA tensor")
        if out is None:
            out = torch.empty((rows, N), dtype=x.dtype, device=x.device)
        elif not (out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (rows, N)):
            raise RuntimeError("linear: out must be a contiguous fp32 CUDA tensor [rows, N]")
        bp = bias.data_ptr() if bias is not None else None
        if mode == "f16x2" and K % 8 == 0:
            terms = _f16_split2(weight)
            with _on_device(x):
                rc = _C.lib.mvd_linear_f16x2_f32(x.data_ptr(), terms.data_ptr(), bp, rows, K, N, 1 if relu else 0,
                                                 out.data_ptr(), _stream(x))
            _C.check(rc, "mvd_linear_f16x2_f32")
            return out
        if mode != "tf32x3" and K % 8 == 0:
            terms = _bf16_split3(weight)
            ts = mode == "bf16x3ts" or (mode == "bf16x3" and _GEMM_TS)   # "bf16x3ss": shared-memory operands explicitly
            fn = _C.lib.mvd_linear_bf16x3_ts_f32 if ts else _C.lib.mvd_linear_bf16x3_f32
            with _on_device(x):
                rc = fn(x.data_ptr(), terms.data_ptr(), bp, rows, K, N, 1 if relu else 0, out.data_ptr(), _stream(x))
            _C.check(rc, "mvd_linear_bf16x3_ts_f32" if ts else "mvd_linear_bf16x3_f32")
            return out
        w_hi, w_lo = _tf32_split(weight)
        with _on_device(x):
            rc = _C.lib.mvd_linear_tf32x3_f32(x.data_ptr(), w_hi.data_ptr(), w_lo.data_ptr(), bp, rows, K, N,
                                              1 if relu else 0, out.data_ptr(), _stream(x))
        _C.check(rc, "mvd_linear_tf32x3_f32")
        return out
    if mode != "torch":
        for name, t in (("x", x), ("weight", weight), ("bias", bias)):
            if t is not None and not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise RuntimeError(f"linear: {name} must be a contiguous fp32 CUDA tensor")
        # scratch per (device, stream): two streams running GEMMs concurrently must not share it
        ws_key = (x.device, _stream(x))
        ws = _gemm_ws.get(ws_key)
        if ws is None:
            ws = _gemm_ws[ws_key] = torch.empty(64 << 20, dtype=torch.uint8, device=x.device)
        if out is None:
            out = torch.empty((rows, N), dtype=x.dtype, device=x.device)
        elif not (out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (rows, N)):
            raise RuntimeError("linear: out must be a contiguous fp32 CUDA tensor [rows, N]")
        with _on_device(x):
            rc = _C.lib.mvd_linear_f32(x.data_ptr(), weight.data_ptr(), bias.data_ptr() if bias is not None else None,
                                       rows, K, N, 1 if relu else 0, 1 if mode == "bf16x9" else 0, out.data_ptr(),
                                       ws.data_ptr(), ws.numel(), _stream(x))
        if rc == 0:
            return out
        if rc in (-3, -5):
            raise RuntimeError(
                f"mvdetr_b200.linear: cuBLASLt >= 12.9 path unavailable ({_C.error_string(rc)}; rows={rows}, K={K}, "
                f"N={N}, mode={mode}). Set MVDETR_B200_GEMM=torch to run the dense layers through torch.mm instead "
                "(about 1.5x slower frames); it is never selected silently.")
        _C.check(rc, "mvd_linear_f32")
    res = torch.mm(x, weight.t(), out=out) if out is not None else torch.mm(x, weight.t())
    if bias is not None and N % 4 == 0:
        return bias_act_(res, bias, relu=relu)
    if bias is not None:
        res.add_(bias)
    return torch.relu_(res) if relu else res


# "implicit": the 3x3 convolutions of the fast path fetch their taps by TMA from the channels-last map (no im2col matrix);
# "im2col": round-1/2 route (warp / upsample kernels write the im2col matrix, one Linear GEMM reads it). Same results.
_CONV_MODE = os.environ.get("MVDETR_B200_CONV3X3", "implicit")


def conv3x3_implicit_ok(c_in, n_out):
    """True when the fast path should try conv3x3_nhwc: own split-operand kernel with the x terms in tensor memory selected,
    32-channel chunks. (An output width without a divisor in [16, 128] is reported by the library: conv3x3_nhwc -> None.)"""
    return (_CONV_MODE == "implicit" and (_GEMM_MODE == "f16x2" or (_GEMM_MODE == "bf16x3" and _GEMM_TS)) and
            c_in % 32 == 0 and n_out % 4 == 0)


def conv3x3_nhwc(x_cl, w2d, bias=None, stride=1, relu=False, add=None):
    """3x3 / pad 1 / stride convolution of x_cl [NB,Hi,Wi,C] (channels-last storage, fp32 CUDA) with the weight matrix
    w2d [N, 9*C] whose columns are ordered (ky, kx, c) -> [NB*Ho*Wo, N] (channels-last rows), as an implicit GEMM on
    our tcgen05 kernel (mvd_conv3x3_nhwc_f32). Bit-identical to linear(im2col(x), w2d). Returns None when the library
    reports the shape unsupported (callers keep the im2col route). add [NB*Ho*Wo, N] (or broadcastable view of that
    element count): returns (out, out + add), the sum written by the same epilogue."""
    NB, Hi, Wi, C = x_cl.shape
    N = w2d.shape[0]
    if w2d.shape[1] != 9 * C:
        raise RuntimeError("conv3x3_nhwc: weight matrix must be [N, 9*C] with (ky, kx, c) columns")
    for name, t in (("x", x_cl), ("weight", w2d), ("bias", bias)):
        if t is not None and not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError(f"conv3x3_nhwc: {name} must be a contiguous fp32 CUDA tensor")
    Ho, Wo = (Hi - 1) // stride + 1, (Wi - 1) // stride + 1
    if _GEMM_MODE == "f16x2":
        terms, nt = _f16_split2(w2d), 2
    else:
        terms, nt = _bf16_split3(w2d), 3
    out = torch.empty((NB * Ho * Wo, N), dtype=x_cl.dtype, device=x_cl.device)
    out2 = None
    if add is not None:
        if not (add.is_cuda and add.is_contiguous() and add.dtype == torch.float32 and add.numel() == out.numel()):
            raise RuntimeError("conv3x3_nhwc: add must be a contiguous fp32 CUDA tensor with the output's element count")
        out2 = torch.empty_like(out)
    with _on_device(x_cl):
        rc = _C.lib.mvd_conv3x3_nhwc_f32(x_cl.data_ptr(), terms.data_ptr(), bias.data_ptr() if bias is not None else None,
                                         NB, Hi, Wi, C, int(stride), N, 1 if relu else 0, nt, out.data_ptr(),
                                         add.data_ptr() if add is not None else None,
                                         out2.data_ptr() if out2 is not None else None, _stream(x_cl))
    if rc == -3:
        return None
    _C.check(rc, "mvd_conv3x3_nhwc_f32")
    return out if add is None else (out, out2)


def upsample_nhwc(x_cl, dsize, rows=None, out=None):
    """Bilinear upsample (align_corners=False, ATen arithmetic) of x_cl [BN,Hi,Wi,C] channels-last fp32 CUDA to
    [BN,Ho,Wo,C]; rows=(row0, nrows) writes only that band of output rows into `out` (the rest is left untouched)."""
    BN, Hi, Wi, C = x_cl.shape
    Ho, Wo = int(dsize[0]), int(dsize[1])
    if not (x_cl.is_cuda and x_cl.is_contiguous() and x_cl.dtype == torch.float32 and C % 4 == 0):
        raise RuntimeError("upsample_nhwc: contiguous fp32 CUDA [BN,Hi,Wi,C] with C % 4 == 0 required")
    row0, nrows = (0, Ho) if rows is None else (int(rows[0]), int(rows[1]))
    if out is None:
        out = torch.empty((BN, Ho, Wo, C), dtype=x_cl.dtype, device=x_cl.device)
    elif not (out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (BN, Ho, Wo, C)):
        raise RuntimeError("upsample_nhwc: out must be a contiguous fp32 CUDA [BN,Ho,Wo,C]")
    with _on_device(x_cl):
        rc = _C.lib.mvd_upsample_nhwc_f32(x_cl.data_ptr(), BN, C, Hi, Wi, Ho, Wo, row0, nrows, out.data_ptr(),
                                          _stream(x_cl))
    _C.check(rc, "mvd_upsample_nhwc_f32")
    return out


def linear_multicast(x, weight, bias, mc_ptr, relu=False):
    """act(x @ weight.T + bias) written through the NVLink MULTICAST address `mc_ptr` (int; the slot of a symmetric buffer
    as mapped by torch.distributed._symmetric_memory): the GEMM's epilogue is the all-gather -- every 16-byte piece
    leaves as one multimem.st that the NVSwitch replicates into every GPU's copy of the buffer (mvdetr_b200/sharded.py
    separates producers and consumers with a cross-GPU barrier). bf16x3 kernel only (K % 8 == 0, N % 4 == 0)."""
    rows, K = x.shape
    N = weight.shape[0]
    x = x.contiguous()
    for name, t in (("x", x), ("weight", weight), ("bias", bias)):
        if t is not None and not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError(f"linear_multicast: {name} must be a contiguous fp32 CUDA tensor")
    if _GEMM_MODE == "f16x2":
        terms, fn = _f16_split2(weight), _C.lib.mvd_linear_f16x2_multicast_f32
    else:
        terms = _bf16_split3(weight)
        fn = _C.lib.mvd_linear_bf16x3_ts_multicast_f32 if _GEMM_TS else _C.lib.mvd_linear_bf16x3_multicast_f32
    with _on_device(x):
        rc = fn(x.data_ptr(), terms.data_ptr(), bias.data_ptr() if bias is not None else None, rows, K, N,
                1 if relu else 0, int(mc_ptr), _stream(x))
    _C.check(rc, "mvd_linear_bf16x3_multicast_f32")


def multicast_copy(src, mc_ptr, inner=0, outer_total=0, outer0=0):
    """src [rows, C] (contiguous fp32 CUDA, C % 4 == 0) -> the symmetric buffer at multicast address `mc_ptr` on every
    GPU. inner > 0: local rows [outer_local][inner] land transposed at row (r % inner) * outer_total + outer0 + r // inner
    (view-major tokens -> cell-major rows of all views)."""
    if not (src.is_cuda and src.is_contiguous() and src.dtype == torch.float32 and src.dim() == 2 and src.shape[1] % 4 == 0):
        raise RuntimeError("multicast_copy: contiguous fp32 CUDA [rows, C] with C % 4 == 0 required")
    if src.numel() == 0:
        return
    with _on_device(src):
        rc = _C.lib.mvd_multicast_copy_f32(src.data_ptr(), int(mc_ptr), src.shape[0], src.shape[1], int(inner),
                                           int(outer_total), int(outer0), _stream(src))
    _C.check(rc, "mvd_multicast_copy_f32")


def gemm_mode_text():
    """One line describing how ops.linear runs (bench.py's config.gemm)."""
    lt = linear_available()
    if _GEMM_MODE == "torch":
        return "torch.mm fp32 (cuBLAS SIMT), explicit MVDETR_B200_GEMM=torch"
    if _GEMM_MODE == "bf16x3":
        return ("bf16x3: own persistent tcgen05.mma.kind::f16 kernel (3-term bf16 split in-kernel, 6 products, fp32 "
                "accumulate in TMEM, bias/ReLU epilogue, TMA rings; x terms in "
                + ("tensor memory (TS-form MMA)" if _GEMM_TS else "shared memory") + "), fp32-level accuracy")
    if _GEMM_MODE == "f16x2":
        return ("f16x2: own persistent tcgen05.mma.kind::f16 kernel (2-term fp16 split in-kernel, second term scaled by "
                "2^11, 3 products, fp32 accumulate in TMEM, x terms in tensor memory), fp32-level accuracy for |x|,|w| < 65504")
    if _GEMM_MODE == "tf32x3":
        return ("tf32x3: own tcgen05.mma.kind::tf32 kernel (3xTF32 split in-kernel, fp32 accumulate in TMEM, bias/ReLU "
                "epilogue), fp32-level accuracy")
    if not lt:
        return "UNAVAILABLE: cuBLASLt >= 12.9 not loadable (ops.linear raises)"
    return (f"{_GEMM_MODE} via cuBLASLt {lt} (fp32 in/out; bf16x9 = CUBLAS_COMPUTE_32F_EMULATED_16BFX9, fp32-accurate "
            "tensor-core emulation)")


def pos_add_launches(layers):
    """Launches of the query = src + pos add per frame that are OURS: layers 2.. get their query from the previous
    layer's LayerNorm kernel (no extra launch); the first layer's add is a torch elementwise kernel (not counted)."""
    return 0


def msda_bwd_kernel_name(value, hw, Lq):
    """Which backward kernel ms_deform_attn_backward dispatches for this layout (bench.py reports it)."""
    B, S, M, D = value.shape
    uniform = all(s == hw[0] for s in hw)
    if _BWD_VIEWGRID and uniform and value.dtype == torch.float32 and D in (8, 16, 32):
        return f"msda_vg_bwd_kernel<{D}> (TMA-staged view grid)"
    if D % 4 != 0:
        return "msda_bwd_scalar_kernel"
    banded = _BWD_BANDED and uniform and value.dtype == torch.float32 and Lq % (hw[0][0] * hw[0][1]) == 0
    return f"msda_bwd_vec4_kernel<{D}> (" + ("pairs walked band by band" if banded else "generic, query order") + ")"


def bias_act_(x, bias, relu=False):
    """In place: x = act(x + bias) over the last dim (mvd_bias_act_f32); x is the output of a bias-free GEMM."""
    C = x.shape[-1]
    for name, t in (("x", x), ("bias", bias)):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError(f"bias_act_: {name} must be a contiguous fp32 CUDA tensor")
    if bias.numel() != C:
        raise RuntimeError("bias_act_: bias must have one entry per column")
    with _on_device(x):
        rc = _C.lib.mvd_bias_act_f32(x.data_ptr(), bias.data_ptr(), x.numel() // C, C, 1 if relu else 0, _stream(x))
    _C.check(rc, "mvd_bias_act_f32")
    return x


class _WarpPerspective(Function):
    """Layout policy: with C % 4 == 0 the gather runs on a channels-last source (every tap one contiguous C-vector).
    A torch channels_last input is used in place; a plain NCHW input is relaid out once by mvd_transpose_f32
    (51.6 MB read + write at Wildtrack size, cheaper than gathering scalars from C planes). Other C take the
    scalar NCHW kernels."""

    @staticmethod
    def forward(ctx, src, mat, Ho, Wo, channels_last):
        BN, C, Hi, Wi = src.shape
        shape = (BN, Ho, Wo, C) if channels_last else (BN, C, Ho, Wo)
        dst = torch.empty(shape, dtype=src.dtype, device=src.device)
        vec = C % 4 == 0 and _WARP_CL
        if vec and _tma_warp_ok(src) and _warp_tma(src, mat, Ho, Wo, dst, 1 if channels_last else 0):
            ctx.save_for_backward(mat)
            ctx.geom = (BN, C, Hi, Wi, Ho, Wo, channels_last, vec)
            return dst
        layout = _DST_NHWC if channels_last else 0
        if vec:
            src, _ = _as_nhwc(src)
            layout |= _SRC_NHWC
        else:
            src = src.contiguous()
        with _on_device(src):
            rc = _C.lib.mvd_warp_fwd_f32(src.data_ptr(), mat.data_ptr(), BN, C, Hi, Wi, Ho, Wo, dst.data_ptr(),
                                         layout, _stream(src))
        _C.check(rc, "mvd_warp_fwd_f32")
        ctx.save_for_backward(mat)
        ctx.geom = (BN, C, Hi, Wi, Ho, Wo, channels_last, vec)
        return dst

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_dst):
        (mat,) = ctx.saved_tensors
        BN, C, Hi, Wi, Ho, Wo, channels_last, vec = ctx.geom
        if vec:
            g_cl = grad_dst.contiguous() if channels_last else _as_nhwc(grad_dst)[0]
            grad_src = torch.empty((BN, Hi, Wi, C), dtype=grad_dst.dtype, device=grad_dst.device)
            with _on_device(grad_dst):
                rc = _C.lib.mvd_warp_bwd_nhwc_f32(g_cl.data_ptr(), mat.data_ptr(), BN, C, Hi, Wi, Ho, Wo,
                                                  grad_src.data_ptr(), _stream(grad_dst))
            _C.check(rc, "mvd_warp_bwd_nhwc_f32")
            return grad_src.permute(0, 3, 1, 2), None, None, None, None  # NCHW-shaped view, channels_last strides
        if channels_last:
            grad_dst = grad_dst.permute(0, 3, 1, 2)
        grad_dst = grad_dst.contiguous()
        grad_src = torch.empty((BN, C, Hi, Wi), dtype=grad_dst.dtype, device=grad_dst.device)
        with _on_device(grad_dst):
            rc = _C.lib.mvd_warp_bwd_f32(grad_dst.data_ptr(), mat.data_ptr(), BN, C, Hi, Wi, Ho, Wo,
                                         grad_src.data_ptr(), _stream(grad_dst))
        _C.check(rc, "mvd_warp_bwd_f32")
        return grad_src, None, None, None, None


def warp_perspective(src, M, dsize, mode="bilinear", padding_mode="zeros", align_corners=None, channels_last=False):
    """kornia.warp_perspective(src [BN,C,H,W], M [BN,3,3] (src pixel -> dst pixel), dsize=(Ho,Wo)) -> [BN,C,Ho,Wo].

    Only the combination MVDeTr's hot path uses is implemented in CUDA (mvdetr.py:194-195):
    mode='bilinear', padding_mode='zeros', align_corners=False; anything else raises NotImplementedError.
    Differentiable w.r.t. `src`. `channels_last=True` (extension) returns [BN,Ho,Wo,C]."""
    if mode != "bilinear" or padding_mode != "zeros" or align_corners is not False:
        raise NotImplementedError(
            "mvdetr_b200.warp_perspective implements mode='bilinear', padding_mode='zeros', align_corners=False "
            f"(got mode={mode!r}, padding_mode={padding_mode!r}, align_corners={align_corners!r})")
    if not torch.is_tensor(src) or not torch.is_tensor(M):
        raise TypeError("Input type is not a torch.Tensor")
    if src.dim() != 4:
        raise ValueError(f"Input src must be a BxCxHxW tensor. Got {tuple(src.shape)}")
    if M.dim() != 3 or tuple(M.shape[-2:]) != (3, 3) or M.shape[0] != src.shape[0]:
        raise ValueError(f"Input M must be a Bx3x3 tensor. Got {tuple(M.shape)}")
    if not src.is_cuda:
        raise RuntimeError("mvdetr_b200.warp_perspective: src must be a CUDA tensor (no CPU fallback)")
    if src.dtype != torch.float32:
        raise RuntimeError(f"mvdetr_b200.warp_perspective: fp32 only, got {src.dtype}")
    Ho, Wo = int(dsize[0]), int(dsize[1])
    mat = M.detach().to(device=src.device, dtype=torch.float32).contiguous()
    return _WarpPerspective.apply(src, mat, Ho, Wo, bool(channels_last))
