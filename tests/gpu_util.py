"""Helpers shared by the -m gpu tests (all arithmetic under test goes through the C ABI via mvdetr_b200.ops)."""
import importlib.util
import os

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REF_SO = os.path.join(REPO, "oracle", "_ref", "MultiScaleDeformableAttention.so")
_ref_ext = None


def ref_cuda_ext():
    """The reference's own CUDA op built for sm_100a by oracle/build_ref.sh, or None when it was not built."""
    global _ref_ext
    if _ref_ext is None and os.path.exists(_REF_SO):
        spec = importlib.util.spec_from_file_location("MultiScaleDeformableAttention", _REF_SO)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_ext = mod
    return _ref_ext


def dev(a, device, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return t.to(dtype) if dtype is not None else t


def start_index(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def make_problem(B, shapes_list, M, D, Lq, P, seed, dtype=torch.float32, spread=1.3, device="cpu"):
    """Seeded random MSDA inputs (CPU generator => identical on every box)."""
    g = torch.Generator().manual_seed(seed)
    shapes = torch.as_tensor(shapes_list, dtype=torch.long)
    S = int(shapes.prod(1).sum())
    L = shapes.shape[0]
    value = torch.randn(B, S, M, D, generator=g).to(dtype)
    loc = (torch.rand(B, Lq, M, L, P, 2, generator=g) * spread - (spread - 1) / 2).to(dtype)
    attn = torch.softmax(torch.randn(B, Lq, M, L * P, generator=g), -1).view(B, Lq, M, L, P).to(dtype)
    grad_out = torch.randn(B, Lq, M * D, generator=g).to(dtype)
    start = start_index(shapes)
    return tuple(t.to(device) for t in (value, shapes, start, loc, attn, grad_out))


def viewgrid_problem(L, H, W, M, D, P, seed, R=None, offset_px=3.0, device="cpu"):
    """MVDeTr encoder layout: L views of an HxW grid, queries = R copies of the grid, locations = identity reference
    points (mvdetr.py:33-71 with zs=0) + head/point-dependent offsets of a few pixels + per-query jitter."""
    g = torch.Generator().manual_seed(seed)
    R = L if R is None else R
    S, Lq = L * H * W, R * H * W
    ys, xs = torch.meshgrid(torch.linspace(0.5, H - 0.5, H), torch.linspace(0.5, W - 0.5, W), indexing="ij")
    ref = torch.stack((xs / W, ys / H), -1).reshape(1, H * W, 1, 1, 1, 2).repeat(1, R, 1, 1, 1, 1)
    theta = torch.arange(M, dtype=torch.float32) * (2 * np.pi / M)
    d = torch.stack([theta.cos(), theta.sin()], -1)
    d = (d / d.abs().max(-1, keepdim=True)[0]).view(1, 1, M, 1, 1, 2) * torch.arange(1, P + 1).view(1, 1, 1, 1, P, 1)
    jitter = torch.randn(1, Lq, M, L, P, 2, generator=g) * offset_px / 3
    loc = ref + (d + jitter) / torch.tensor([W, H], dtype=torch.float32)
    value = torch.randn(1, S, M, D, generator=g)
    attn = torch.softmax(torch.randn(1, Lq, M, L * P, generator=g), -1).view(1, Lq, M, L, P)
    shapes = torch.as_tensor([[H, W]] * L, dtype=torch.long)
    grad_out = torch.randn(1, Lq, M * D, generator=g)
    return tuple(t.contiguous().to(device) for t in (value, shapes, start_index(shapes), loc, attn, grad_out))
