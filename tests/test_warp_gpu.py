"""GPU parity tests for the perspective warp kernel (through the C ABI) vs the oracle.
PARITY UNPINNED against the reference itself: kornia is third-party and absent; the oracle restates it."""
import numpy as np
import pytest
import torch

from mvdetr_b200 import ops
from oracle import cpu_oracle as co
from oracle import torch_port as tp
from tests.gpu_util import dev

pytestmark = pytest.mark.gpu
ATOL = 1e-4


def test_golden_small_forward_backward(golden, cuda):
    g = golden("warp_small.npz")
    dsize = tuple(int(v) for v in g["dsize"])
    src = dev(g["src"], cuda).requires_grad_(True)
    out = ops.warp_perspective(src, dev(g["mats"], cuda), dsize, align_corners=False)
    assert np.abs(out.detach().cpu().numpy() - g["out"]).max() <= ATOL
    out.backward(dev(g["grad_out"], cuda))
    assert np.abs(src.grad.cpu().numpy() - g["grad_src"]).max() <= ATOL * max(1.0, np.abs(g["grad_src"]).max())


def test_world_feat_mini_projection_chain(golden, cuda):
    """Projection matrices built exactly as MVDeTr does (mvdetr.py:82-95,155-161; golden holds them)."""
    g = golden("world_feat_mini.npz")
    out = ops.warp_perspective(dev(g["imgs_feat"], cuda), dev(g["proj_mats"], cuda), tuple(g["Rworld"]),
                               align_corners=False)
    assert np.abs(out.cpu().numpy() - g["world_in"]).max() <= ATOL
    cl = ops.warp_perspective(dev(g["imgs_feat"], cuda), dev(g["proj_mats"], cuda), tuple(g["Rworld"]),
                              align_corners=False, channels_last=True)
    assert torch.equal(cl.permute(0, 3, 1, 2), out)


def _ring_homographies(n, Hi, Wi, Ho, Wo, seed):
    """Random but well-conditioned src->dst pixel homographies whose images cover part of the destination."""
    rng = np.random.RandomState(seed)
    mats = []
    for _ in range(n):
        sx, sy = Wo / Wi * rng.uniform(0.7, 1.6), Ho / Hi * rng.uniform(0.7, 1.6)
        th = rng.uniform(-0.5, 0.5)
        A = np.array([[sx * np.cos(th), -sy * np.sin(th), rng.uniform(-0.2, 0.2) * Wo],
                      [sx * np.sin(th), sy * np.cos(th), rng.uniform(-0.2, 0.2) * Ho],
                      [rng.uniform(-1, 1) * 1e-3, rng.uniform(-1, 1) * 1e-3, 1.0]])
        mats.append(A * rng.uniform(0.001, 2.0))  # homographies are scale-free; MVDeTr's have ~1e-3 scale
    return np.stack(mats).astype(np.float32)


@pytest.mark.parametrize("C,Hi,Wi,Ho,Wo", [(1, 5, 7, 6, 9), (3, 1, 1, 4, 4), (20, 9, 16, 1, 1), (17, 12, 20, 15, 33)])
def test_odd_shapes_vs_c_oracle(cuda, C, Hi, Wi, Ho, Wo):
    rng = np.random.RandomState(C)
    src = rng.randn(2, C, Hi, Wi).astype(np.float32)
    mats = _ring_homographies(2, max(Hi, 2), max(Wi, 2), max(Ho, 2), max(Wo, 2), seed=C)
    out = ops.warp_perspective(dev(src, cuda), dev(mats, cuda), (Ho, Wo), align_corners=False)
    ref = co.warp_forward(src, mats, (Ho, Wo))
    assert np.abs(out.cpu().numpy() - ref).max() <= ATOL


def test_full_wildtrack_size(cuda):
    """[7,128,90,160] -> [7,128,120,360] (BASELINE config 2): vs the C oracle (OpenMP, seconds), plus linearity and
    <warp(x), g> = <x, warp^T(g)> (the backward is the exact adjoint of the forward)."""
    BN, C, Hi, Wi, Ho, Wo = 7, 128, 90, 160, 120, 360
    g = torch.Generator().manual_seed(0)
    src = torch.randn(BN, C, Hi, Wi, generator=g)
    mats = torch.from_numpy(_ring_homographies(BN, Hi, Wi, Ho, Wo, seed=0))
    d_src = src.to(cuda).requires_grad_(True)
    out = ops.warp_perspective(d_src, mats.to(cuda), (Ho, Wo), align_corners=False)
    ref = co.warp_forward(src.numpy(), mats.numpy(), (Ho, Wo))
    assert (ref != 0).mean() > 0.2  # the synthetic cameras do see the plane
    assert np.abs(out.detach().cpu().numpy() - ref).max() <= ATOL
    # kornia-restatement (fp32 torch.inverse) agrees with the in-kernel fp64 inverse to the same tolerance
    assert (tp.warp_perspective(src, mats, (Ho, Wo)) - out.detach().cpu()).abs().max().item() <= 5e-4
    gout = torch.randn(BN, C, Ho, Wo, generator=g).to(cuda)
    out.backward(gout)
    lhs = (out.detach() * gout).sum().item()
    rhs = (d_src.detach() * d_src.grad).sum().item()
    assert abs(lhs - rhs) <= 1e-3 * max(1.0, abs(lhs))
    src2 = torch.randn(BN, C, Hi, Wi, generator=g).to(cuda)
    o2 = ops.warp_perspective(src2, mats.to(cuda), (Ho, Wo), align_corners=False)
    o12 = ops.warp_perspective(d_src.detach() * 2 - src2, mats.to(cuda), (Ho, Wo), align_corners=False)
    assert (o12 - (out.detach() * 2 - o2)).abs().max().item() <= 1e-4


def test_host_buffer_entry_point(cuda):
    from mvdetr_b200 import _C
    src = torch.randn(2, 8, 9, 16)
    mats = torch.from_numpy(_ring_homographies(2, 9, 16, 12, 20, seed=4))
    out_host = torch.empty(2, 8, 12, 20)
    rc = _C.lib.mvd_warp_fwd_f32_host(src.data_ptr(), mats.data_ptr(), 2, 8, 9, 16, 12, 20, out_host.data_ptr(), None)
    assert rc == 0, _C.error_string(rc)
    out = ops.warp_perspective(src.to(cuda), mats.to(cuda), (12, 20), align_corners=False)
    assert torch.equal(out.cpu(), out_host)


@pytest.mark.parametrize("C,Hi,Wi,Ho,Wo", [(128, 18, 32, 24, 72), (8, 9, 16, 15, 33), (132, 7, 5, 3, 11), (256, 12, 20, 10, 30)])
def test_every_layout_gives_the_same_values(cuda, C, Hi, Wi, Ho, Wo):
    """NCHW / channels_last source x NCHW / channels-last destination: the vector (channels-last source) kernels, the
    scalar NCHW kernels and the C oracle agree; a torch channels_last input is consumed in place."""
    rng = np.random.RandomState(C + Ho)
    src = rng.randn(3, C, Hi, Wi).astype(np.float32)
    mats = _ring_homographies(3, Hi, Wi, Ho, Wo, seed=C)
    ref = co.warp_forward(src, mats, (Ho, Wo))
    d_src, d_mats = dev(src, cuda), dev(mats, cuda)
    outs = {}
    for name, s in (("nchw", d_src), ("cl", d_src.contiguous(memory_format=torch.channels_last))):
        for cl_out in (False, True):
            o = ops.warp_perspective(s, d_mats, (Ho, Wo), align_corners=False, channels_last=cl_out)
            outs[(name, cl_out)] = o.permute(0, 3, 1, 2) if cl_out else o
    old = ops._WARP_CL
    try:
        ops._WARP_CL = False  # scalar kernels reading the NCHW source directly
        outs[("scalar", False)] = ops.warp_perspective(d_src, d_mats, (Ho, Wo), align_corners=False)
        outs[("scalar", True)] = ops.warp_perspective(d_src, d_mats, (Ho, Wo), align_corners=False,
                                                      channels_last=True).permute(0, 3, 1, 2)
    finally:
        ops._WARP_CL = old
    base = outs[("nchw", False)]
    assert np.abs(base.cpu().numpy() - ref).max() <= ATOL
    for key, o in outs.items():
        assert o.shape == base.shape, key
        assert (o - base).abs().max().item() <= 1e-6, key


def test_channels_last_backward_is_the_adjoint(cuda):
    """Vector backward (red.global.add.v4) against the scalar-atomics backward and the adjoint identity, for both
    gradient layouts autograd can hand over."""
    BN, C, Hi, Wi, Ho, Wo = 2, 64, 14, 22, 19, 37
    g = torch.Generator().manual_seed(5)
    src = torch.randn(BN, C, Hi, Wi, generator=g).to(cuda)
    mats = torch.from_numpy(_ring_homographies(BN, Hi, Wi, Ho, Wo, seed=9)).to(cuda)
    gout = torch.randn(BN, C, Ho, Wo, generator=g).to(cuda)
    grads = []
    for flag, go in ((True, gout), (True, gout.contiguous(memory_format=torch.channels_last)), (False, gout)):
        old = ops._WARP_CL
        try:
            ops._WARP_CL = flag
            s = src.clone().requires_grad_(True)
            out = ops.warp_perspective(s, mats, (Ho, Wo), align_corners=False)
            out.backward(go)
        finally:
            ops._WARP_CL = old
        grads.append(s.grad)
        lhs, rhs = (out.detach() * gout).sum().item(), (src * s.grad).sum().item()
        assert abs(lhs - rhs) <= 1e-3 * max(1.0, abs(lhs))
    assert grads[0].shape == src.shape
    for gr in grads[:2]:
        assert (gr - grads[2]).abs().max().item() <= 1e-4 * max(1.0, grads[2].abs().max().item())
    # channels-last OUTPUT: gradient arrives as [BN, Ho, Wo, C]
    s = src.clone().requires_grad_(True)
    ops.warp_perspective(s, mats, (Ho, Wo), align_corners=False, channels_last=True).backward(
        gout.permute(0, 2, 3, 1).contiguous())
    assert (s.grad - grads[2]).abs().max().item() <= 1e-4 * max(1.0, grads[2].abs().max().item())


@pytest.mark.parametrize("batch,rows,cols", [(1, 1, 1), (3, 37, 65), (2, 128, 14400), (1, 33, 32)])
def test_transpose_kernel(cuda, batch, rows, cols):
    x = torch.randn(batch, rows, cols, generator=torch.Generator().manual_seed(rows)).to(cuda)
    assert torch.equal(ops.transpose_last2(x), x.transpose(1, 2).contiguous())


def _im2col_reference(x_nchw, stride):
    """[BN,C,H,W] -> [BN*Ho2*Wo2, 9*C] with (ky, kx, c) column order, via F.unfold (3x3, pad 1)."""
    BN, C = x_nchw.shape[:2]
    cols = torch.nn.functional.unfold(x_nchw, 3, padding=1, stride=stride)          # [BN, C*9, L], rows (c, ky, kx)
    L = cols.shape[-1]
    return cols.view(BN, C, 9, L).permute(0, 3, 2, 1).reshape(BN * L, 9 * C)


@pytest.mark.parametrize("C,Hi,Wi,Ho,Wo,stride", [(128, 18, 32, 24, 72, 2), (8, 9, 16, 15, 33, 2), (132, 7, 5, 3, 11, 1),
                                                   (16, 12, 20, 1, 1, 2), (32, 10, 10, 2, 2, 2)])
def test_warp_im2col_matches_warp_then_unfold(cuda, C, Hi, Wi, Ho, Wo, stride):
    rng = np.random.RandomState(C + Ho)
    src = dev(rng.randn(3, C, Hi, Wi).astype(np.float32), cuda)
    mats = dev(_ring_homographies(3, max(Hi, 2), max(Wi, 2), max(Ho, 2), max(Wo, 2), seed=C), cuda)
    world = ops.warp_perspective(src, mats, (Ho, Wo), align_corners=False)
    A, (Ho2, Wo2) = ops.warp_im2col(src, mats, (Ho, Wo), stride=stride)
    want = _im2col_reference(world, stride)
    assert A.shape == want.shape and Ho2 * Wo2 * 3 == want.shape[0]
    assert torch.equal(A, want)  # same arithmetic as the warp kernel, zeros in the padding slots


@pytest.mark.parametrize("C,Hi,Wi,Ho,Wo", [(128, 15, 45, 30, 90), (8, 7, 9, 15, 20), (36, 5, 4, 11, 6), (4, 1, 1, 3, 2)])
def test_upsample_im2col_matches_interpolate_then_unfold(cuda, C, Hi, Wi, Ho, Wo):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(2, C, Hi, Wi, generator=g).to(cuda)
    up = torch.nn.functional.interpolate(x, size=(Ho, Wo), mode="bilinear", align_corners=False)
    want = _im2col_reference(up, 1)
    A = ops.upsample_im2col(x.permute(0, 2, 3, 1).contiguous(), (Ho, Wo))
    assert A.shape == want.shape
    assert (A - want).abs().max().item() <= 1e-6


def _tma_cases():
    """(C, Hi, Wi, Ho, Wo, homography kind): geometries that exercise every branch of the TMA kernel."""
    return [(32, 90, 160, 120, 360, "ring"),       # Wildtrack geometry, one 32-channel group
            (128, 24, 32, 40, 72, "ring"),         # several groups per stage
            (64, 128, 128, 16, 16, "shrink"),      # strong minification: bounding box > 64 px -> global-load path
            (64, 16, 16, 70, 90, "zoom"),          # strong magnification: 1-2 source pixels per tile
            (96, 20, 28, 33, 47, "outside"),       # image lands mostly / entirely outside: empty tiles write zeros
            (32, 12, 16, 5, 7, "ring")]            # destination smaller than one tile


def _homographies(kind, n, Hi, Wi, Ho, Wo, seed):
    if kind == "ring":
        return _ring_homographies(n, Hi, Wi, Ho, Wo, seed)
    rng = np.random.RandomState(seed)
    mats = []
    for i in range(n):
        if kind == "shrink":   # whole source onto the small destination
            A = np.array([[Wo / Wi, 0.02, 0.3], [-0.01, Ho / Hi, 0.2], [1e-5, 2e-5, 1.0]])
        elif kind == "zoom":   # a few source pixels cover the destination
            A = np.array([[Wo / 3.0, 0.0, -Wo * (1 + i)], [0.0, Ho / 2.5, -Ho * 2.0], [0.0, 0.0, 1.0]])
        else:                  # "outside": view 0 far away from the destination, others straddle its border
            sh = [5.0, 0.6, -0.7][i % 3]
            A = np.array([[Wo / Wi, 0.0, sh * Wo], [0.0, Ho / Hi, sh * Ho], [0.0, 0.0, 1.0]])
        mats.append(A * rng.uniform(0.01, 2.0))
    return np.stack(mats).astype(np.float32)


@pytest.mark.parametrize("C,Hi,Wi,Ho,Wo,kind", _tma_cases())
def test_tma_warp_kernel_every_mode(cuda, C, Hi, Wi, Ho, Wo, kind):
    """The one-launch TMA-staged warp (NCHW source) in its three destination modes: bit-identical to the round-1
    kernels (relayout + channels-last gather) and within 1e-4 of the C oracle."""
    rng = np.random.RandomState(C + Ho)
    src = rng.randn(3, C, Hi, Wi).astype(np.float32)
    mats = _homographies(kind, 3, Hi, Wi, Ho, Wo, seed=C + Wo)
    ref = co.warp_forward(src, mats, (Ho, Wo))
    d_src, d_mats = dev(src, cuda), dev(mats, cuda)
    assert ops._tma_warp_ok(d_src)
    res = {}
    for tma in (True, False):
        old = ops._WARP_TMA
        try:
            ops._WARP_TMA = tma
            res[tma] = (ops.warp_perspective(d_src, d_mats, (Ho, Wo), align_corners=False),
                        ops.warp_perspective(d_src, d_mats, (Ho, Wo), align_corners=False, channels_last=True),
                        ops.warp_im2col(d_src, d_mats, (Ho, Wo), stride=2)[0],
                        ops.warp_im2col(d_src, d_mats, (Ho, Wo), stride=1)[0])
        finally:
            ops._WARP_TMA = old
    assert np.abs(res[True][0].cpu().numpy() - ref).max() <= ATOL
    if kind in ("ring", "shrink", "zoom"):
        assert (ref != 0).any()
    for a, b, what in zip(res[True], res[False], ("nchw", "nhwc", "im2col s2", "im2col s1")):
        assert a.shape == b.shape and torch.equal(a, b), what
    # Inf / NaN in parts of the source that no valid tap reads must not leak (zero weights are never multiplied in)
    src2 = src.copy()
    src2[:, :, 0, 0] = np.inf
    o_tma = ops.warp_perspective(dev(src2, cuda), d_mats, (Ho, Wo), align_corners=False)
    old = ops._WARP_TMA
    try:
        ops._WARP_TMA = False
        o_old = ops.warp_perspective(dev(src2, cuda), d_mats, (Ho, Wo), align_corners=False)
    finally:
        ops._WARP_TMA = old
    assert torch.equal(torch.isfinite(o_tma), torch.isfinite(o_old))
    fin = torch.isfinite(o_old)
    assert torch.equal(o_tma[fin], o_old[fin])
