"""ctypes binding of lib/libmvdetr_b200.so (the C ABI declared in include/mvdetr_b200.h).

There is no CPU fallback: if the library is missing, importing this module raises ImportError with the build
command, and every op in this package fails with it.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmvdetr_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"mvdetr_b200: CUDA library not built ({LIB_PATH} missing). Build it with `python mvdetr_b200/build.py` "
        "(needs nvcc; there is deliberately no CPU or PyTorch fallback).")

lib = ctypes.CDLL(LIB_PATH)

_p, _i = ctypes.c_void_p, ctypes.c_int

# name -> argtypes; must list every function declared in include/mvdetr_b200.h (tests/test_abi.py checks it).
SIGNATURES = {
    "mvd_version": [],
    "mvd_error_string": [_i],
    "mvd_msda_fwd_f32": [_p] * 5 + [_i] * 7 + [_p, _p],
    "mvd_msda_fwd_f64": [_p] * 5 + [_i] * 7 + [_p, _p],
    "mvd_msda_bwd_f32": [_p] * 6 + [_i] * 7 + [_p] * 4,
    "mvd_msda_bwd_banded_f32": [_p] * 6 + [_i] * 10 + [_p] * 4,
    "mvd_msda_bwd_f64": [_p] * 6 + [_i] * 7 + [_p] * 4,
    "mvd_msda_fwd_viewgrid_f32": [_p] * 3 + [_i] * 8 + [_p, _p],
    "mvd_msda_bwd_viewgrid_f32": [_p] * 4 + [_i] * 8 + [_p] * 4,
    "mvd_msda_fused_fwd_f32": [_p] * 8 + [_i] * 8 + [_p] * 4,
    "mvd_msda_fused_fwd_viewgrid_f32": [_p] * 6 + [_i] * 9 + [_p] * 4,
    "mvd_msda_fused_fwd_viewgrid_pitched_f32": [_p] * 6 + [_i] * 11 + [_p, _p],
    "mvd_add_layernorm_f32": [_p] * 5 + [ctypes.c_int64, _i, ctypes.c_float, ctypes.c_int64, _p, _p],
    "mvd_add_layernorm_pos_f32": [_p] * 5 + [ctypes.c_int64, _i, ctypes.c_float, ctypes.c_int64, _p, _p, _p, _p],
    "mvd_warp_im2col_f32": [_p, _p] + [_i] * 7 + [_p, _p],
    "mvd_upsample_im2col_f32": [_p] + [_i] * 6 + [_p, _p],
    "mvd_upsample_im2col_rows_f32": [_p] + [_i] * 8 + [_p, _p],
    "mvd_resize_normalize_u8": [_p, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p],
    "mvd_distance_nms_workspace_bytes": [_i, _i],
    "mvd_decode_candidates_f32": [_p, _p, _i, _i, _i, ctypes.c_float, ctypes.c_float, _i, _i, _p, _p, _p, _p, _p],
    "mvd_distance_nms_f32": [_p, _p, _p, _p, _i, _i, ctypes.c_float, _i, _p, ctypes.c_size_t, _p, _p, _p, _p, _p, _p],
    "mvd_linear_available": [],
    "mvd_linear_f32": [_p, _p, _p, ctypes.c_int64, _i, _i, _i, _i, _p, _p, ctypes.c_size_t, _p],
    "mvd_tf32_split_f32": [_p, ctypes.c_int64, _p, _p, _p],
    "mvd_linear_tf32x3_f32": [_p, _p, _p, _p, ctypes.c_int64, _i, _i, _i, _p, _p],
    "mvd_bf16_split3_f32": [_p, ctypes.c_int64, _p, _p],
    "mvd_linear_bf16x3_f32": [_p, _p, _p, ctypes.c_int64, _i, _i, _i, _p, _p],
    "mvd_linear_bf16x3_multicast_f32": [_p, _p, _p, ctypes.c_int64, _i, _i, _i, _p, _p],
    "mvd_conv3x3_nhwc_f32": [_p, _p, _p] + [_i] * 8 + [_p, _p, _p, _p],
    "mvd_upsample_nhwc_f32": [_p] + [_i] * 8 + [_p, _p],
    "mvd_f16_split2_f32": [_p, ctypes.c_int64, _p, _p],
    "mvd_linear_f16x2_f32": [_p, _p, _p, ctypes.c_int64, _i, _i, _i, _p, _p],
    "mvd_linear_f16x2_multicast_f32": [_p, _p, _p, ctypes.c_int64, _i, _i, _i, _p, _p],
    "mvd_linear_bf16x3_ts_f32": [_p, _p, _p, ctypes.c_int64, _i, _i, _i, _p, _p],
    "mvd_linear_bf16x3_ts_multicast_f32": [_p, _p, _p, ctypes.c_int64, _i, _i, _i, _p, _p],
    "mvd_multicast_copy_f32": [_p, _p, ctypes.c_int64, _i, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _p],
    "mvd_bias_act_f32": [_p, _p, ctypes.c_int64, _i, _i, _p],
    "mvd_warp_fwd_f32": [_p, _p] + [_i] * 6 + [_p, _i, _p],
    "mvd_warp_tma_f32": [_p, _p] + [_i] * 6 + [_p, _i, _i, _p],
    "mvd_warp_bwd_f32": [_p, _p] + [_i] * 6 + [_p, _p],
    "mvd_warp_bwd_nhwc_f32": [_p, _p] + [_i] * 6 + [_p, _p],
    "mvd_transpose_f32": [_p, _i, _i, _i, _p, _p],
    "mvd_msda_fwd_f32_host": [_p] * 5 + [_i] * 7 + [_p, _p],
    "mvd_warp_fwd_f32_host": [_p, _p] + [_i] * 6 + [_p, _p],
}

for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = (ctypes.c_char_p if _name == "mvd_error_string" else
                   ctypes.c_size_t if _name.endswith("_bytes") else ctypes.c_int)


def error_string(code):
    return lib.mvd_error_string(int(code)).decode()


def check(code, what):
    """Turns a non-zero C-ABI return code into a RuntimeError (the reference raises RuntimeError through pybind)."""
    if code != 0:
        raise RuntimeError(f"{what}: {error_string(code)} (code {code})")
