"""ctypes binding of oracle/liboracle.so (plain-C restatement; see msda_ref.c / warp_ref.c headers).

TEST INFRASTRUCTURE. numpy arrays in, numpy arrays out; shapes as in include/mvdetr_b200.h.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("msda_ref.c", "warp_ref.c", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ints(*xs):
    return [ctypes.c_int(int(x)) for x in xs]


def _dims(value, shapes, loc):
    B, S, M, D = value.shape
    L = shapes.shape[0]
    _, Lq, _, _, P, _ = loc.shape
    assert loc.shape == (B, Lq, M, L, P, 2), loc.shape
    return _ints(B, S, M, D, L, Lq, P)


def level_start_index(shapes):
    shapes = np.asarray(shapes, dtype=np.int64).reshape(-1, 2)
    return np.concatenate([[0], np.cumsum(shapes[:, 0] * shapes[:, 1])[:-1]]).astype(np.int64)


def msda_forward(value, shapes, start, loc, attn):
    """value [B,S,M,D], shapes [L,2] int64, start [L] int64, loc [B,Lq,M,L,P,2], attn [B,Lq,M,L,P] -> [B,Lq,M*D]."""
    dt = np.float64 if value.dtype == np.float64 else np.float32
    fn = lib().oracle_msda_fwd_f64 if dt == np.float64 else lib().oracle_msda_fwd_f32
    value, loc, attn = _c(value, dt), _c(loc, dt), _c(attn, dt)
    shapes, start = _c(shapes, np.int64), _c(start, np.int64)
    dims = _dims(value, shapes, loc)
    B, _, M, D = value.shape
    Lq = loc.shape[1]
    out = np.empty((B, Lq, M * D), dtype=dt)
    fn(_ptr(value), _ptr(shapes), _ptr(start), _ptr(loc), _ptr(attn), *dims, _ptr(out))
    return out


def msda_backward(grad_out, value, shapes, start, loc, attn):
    """-> (grad_value, grad_loc, grad_attn) with the shapes of value, loc, attn."""
    dt = np.float64 if value.dtype == np.float64 else np.float32
    fn = lib().oracle_msda_bwd_f64 if dt == np.float64 else lib().oracle_msda_bwd_f32
    grad_out, value, loc, attn = _c(grad_out, dt), _c(value, dt), _c(loc, dt), _c(attn, dt)
    shapes, start = _c(shapes, np.int64), _c(start, np.int64)
    dims = _dims(value, shapes, loc)
    gv, gl, ga = np.empty_like(value), np.empty_like(loc), np.empty_like(attn)
    fn(_ptr(grad_out), _ptr(value), _ptr(shapes), _ptr(start), _ptr(loc), _ptr(attn), *dims, _ptr(gv), _ptr(gl),
       _ptr(ga))
    return gv, gl, ga


def msda_prep(offsets, logits, ref, shapes):
    """offsets [B,Lq,M,L,P,2], logits [B,Lq,M,L*P], ref [Lr,L,P,2] -> (loc, attn) fp32."""
    offsets, logits, ref = _c(offsets, np.float32), _c(logits, np.float32), _c(ref, np.float32)
    shapes = _c(shapes, np.int64)
    B, Lq, M, L, P, _ = offsets.shape
    Lr = ref.shape[0]
    loc = np.empty_like(offsets)
    attn = np.empty((B, Lq, M, L, P), dtype=np.float32)
    lib().oracle_msda_prep_f32(_ptr(offsets), _ptr(logits), _ptr(ref), _ptr(shapes), *_ints(B, M, L, Lq, P, Lr),
                               _ptr(loc), _ptr(attn))
    return loc, attn


def warp_forward(src, mat, dsize, T=None):
    """src [BN,C,Hi,Wi], mat [BN,3,3] (src pixel -> dst pixel), dsize (Ho,Wo) -> [BN,C,Ho,Wo].
    T: optional precomputed normalised inverse homographies [BN,3,3] fp32 (else derived from mat in double)."""
    src, mat = _c(src, np.float32), _c(mat, np.float32)
    BN, C, Hi, Wi = src.shape
    Ho, Wo = dsize
    dst = np.empty((BN, C, Ho, Wo), dtype=np.float32)
    Tc = None if T is None else _c(T, np.float32)
    lib().oracle_warp_fwd_f32(_ptr(src), _ptr(mat), None if Tc is None else _ptr(Tc),
                              *_ints(BN, C, Hi, Wi, Ho, Wo), _ptr(dst))
    return dst


def warp_backward(grad_dst, mat, src_hw, T=None):
    """grad_dst [BN,C,Ho,Wo] -> grad_src [BN,C,Hi,Wi]."""
    grad_dst, mat = _c(grad_dst, np.float32), _c(mat, np.float32)
    BN, C, Ho, Wo = grad_dst.shape
    Hi, Wi = src_hw
    gs = np.empty((BN, C, Hi, Wi), dtype=np.float32)
    Tc = None if T is None else _c(T, np.float32)
    lib().oracle_warp_bwd_f32(_ptr(grad_dst), _ptr(mat), None if Tc is None else _ptr(Tc),
                              *_ints(BN, C, Hi, Wi, Ho, Wo), _ptr(gs))
    return gs
