"""GPU: the Linear-layer GEMM path (mvd_linear_f32 -> the toolkit's cuBLASLt 12.9). The point to pin is ACCURACY: the
BF16x9 tensor-core emulation must be as accurate as native fp32, measured against an fp64 product."""
import pytest
import torch

from mvdetr_b200 import ops

pytestmark = pytest.mark.gpu


def _problem(rows, K, N, seed, cuda):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(rows, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    return x, w, b


@pytest.mark.parametrize("rows,K,N", [(4096, 128, 128), (4099, 128, 448), (2048, 128, 224), (3000, 128, 512),
                                      (3000, 512, 128), (1000, 1152, 128), (77, 896, 128)])
def test_linear_modes_are_fp32_accurate(cuda, rows, K, N):
    if not ops.linear_available():
        pytest.skip("cuBLASLt >= 12.9 not loadable on this box")
    x, w, b = _problem(rows, K, N, rows + N, cuda)
    exact = x.double() @ w.double().t() + b.double()
    err_torch = (torch.addmm(b, x, w.t()).double() - exact).abs().max().item()
    for mode in ("bf16x9", "fp32", "torch"):
        got = ops.linear(x, w, b, mode=mode)
        err = (got.double() - exact).abs().max().item()
        # emulation may not be worse than a small multiple of what the native fp32 GEMM does on the same data
        assert err <= max(4 * err_torch, 2e-6 * exact.abs().max().item()), (mode, err, err_torch)
    relu = ops.linear(x, w, b, relu=True, mode="bf16x9")
    assert (relu.double() - exact.clamp_min(0)).abs().max().item() <= max(4 * err_torch, 2e-6 * exact.abs().max().item())
    nobias = ops.linear(x, w, None, mode="bf16x9")
    assert (nobias.double() - (exact - b.double())).abs().max().item() <= max(4 * err_torch, 2e-6 * exact.abs().max().item())


def test_linear_torch_mode_needs_no_library(cuda):
    x, w, b = _problem(513, 128, 130, 5, cuda)  # N % 4 != 0 -> plain torch epilogue
    got = ops.linear(x, w, b, relu=True, mode="torch")
    assert torch.allclose(got, torch.relu(torch.addmm(b, x, w.t())), atol=1e-6)


@pytest.mark.parametrize("rows,K,N", [(75600, 128, 128), (4099, 128, 448), (2048, 128, 224), (3000, 128, 512),
                                      (3000, 512, 128), (1000, 1152, 128), (77, 896, 128), (130, 36, 20), (1, 4, 4),
                                      (129, 2304, 256), (20000, 128, 128), (641, 40, 132)])
@pytest.mark.parametrize("mode", ["bf16x3ss", "bf16x3ts", "f16x2", "tf32x3"])
def test_own_tensor_core_kernels_are_fp32_accurate(cuda, rows, K, N, mode):
    """Our tcgen05 kernels (operands split into bf16 / tf32 terms, accumulators in tensor memory) against an fp64 product:
    bf16x3 (six products) must be as accurate as the native fp32 GEMM, the 3xTF32 variant within a small multiple of it;
    bias / ReLU epilogue, row and column tails, K not a multiple of the 32-wide chunk, many tiles per CTA."""
    if mode != "tf32x3" and K % 8 != 0:
        pytest.skip("bf16x3 needs a 16-byte row pitch of the bf16 terms (K % 8 == 0); ops.linear then uses tf32x3")
    x, w, b = _problem(rows, K, N, rows + N + 1, cuda)
    exact = x.double() @ w.double().t() + b.double()
    err_torch = (torch.addmm(b, x, w.t()).double() - exact).abs().max().item()
    # bf16x3 (leading product and small terms in separate tensor-memory accumulators): at the native fp32 GEMM's level.
    # 3xTF32 (one accumulator, 3 roundings of the accumulator per k-step): a few 1e-6 relative, growing with K --
    # inside the 1e-4 bar; ops.linear only uses it when K % 8 != 0.
    scale = exact.abs().max().item()
    # f16x2 (two fp16 terms, three products): fp16 products carry 22 significant bits, of which the tensor core's adder drops
    # the lowest when it aligns them to the accumulator: measured 1.0-1.4x the native fp32 GEMM's error (r02n).
    tol = {"tf32x3": max(8 * err_torch, 1.5e-5 * scale), "f16x2": max(4 * err_torch, 4e-6 * scale)}.get(
        mode, max(2.5 * err_torch, 2e-6 * scale))
    got = ops.linear(x, w, b, mode=mode)
    assert (got.double() - exact).abs().max().item() <= tol
    if mode == "bf16x3ts":  # x terms in tensor memory: same products, same order, same accumulators as the shared-memory kernel
        assert torch.equal(got, ops.linear(x, w, b, mode="bf16x3ss"))
    relu = ops.linear(x, w, b, relu=True, mode=mode)
    assert (relu.double() - exact.clamp_min(0)).abs().max().item() <= tol
    out = torch.full((rows, N), float("nan"), device=cuda)
    nobias = ops.linear(x, w, None, mode=mode, out=out)
    assert nobias.data_ptr() == out.data_ptr()
    assert (nobias.double() - (exact - b.double())).abs().max().item() <= tol
    # weight updated in place: the cached hi/lo split must follow
    w.mul_(2.0)
    again = ops.linear(x, w, None, mode=mode)
    assert (again.double() - 2 * (exact - b.double())).abs().max().item() <= 2 * tol


def test_split_cache_is_keyed_on_the_tensor_object(cuda):
    """Regression (r02d): a cache keyed on data_ptr served the split of a DEAD weight to a new weight that the caching
    allocator placed at the same address."""
    g = torch.Generator().manual_seed(0)
    x = torch.randn(256, 64, generator=g).to(cuda)
    for i in range(4):
        w = torch.randn(64, 64, generator=g).to(cuda)   # same shape: the allocator reuses the block just freed
        for mode in ("bf16x3", "f16x2", "tf32x3"):
            got = ops.linear(x, w, None, mode=mode)
            assert (got.double() - x.double() @ w.double().t()).abs().max().item() <= 1e-4, (i, mode)
        del w


def test_f16x2_range_contract(cuda):
    """The default two-term fp16 kernel covers operands below 65504 and small ones down to fp16's subnormals with an
    absolute error far below the 1e-4 bar; beyond the range the result is non-finite (visible), and the three-term bf16
    kernel (MVDETR_B200_GEMM=bf16x3) takes the full fp32 range."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(300, 128, generator=g).to(cuda)
    w = (torch.randn(64, 128, generator=g) / 128 ** 0.5).to(cuda)
    exact = x.double() @ w.double().t()
    for scale in (1e-6, 1e-3, 1.0, 1e3):   # 1e3 * |x| <= ~5e3 < 65504
        got = ops.linear(x * scale, w, None, mode="f16x2")
        err = (got.double() - exact * scale).abs().max().item()
        assert err <= max(4e-6 * scale * exact.abs().max().item(), 1e-9), (scale, err)
    big = x.clone()
    big[5, 7] = 1e6
    assert not torch.isfinite(ops.linear(big, w, None, mode="f16x2")[5]).all()      # out of fp16 range: loud
    ok = ops.linear(big, w, None, mode="bf16x3")
    assert torch.isfinite(ok).all()
    assert (ok.double() - big.double() @ w.double().t()).abs().max().item() <= 1e-6 * 1e6


@pytest.mark.parametrize("NB,Hi,Wi,C,N,stride", [
    (7, 120, 360, 128, 128, 2),    # Wildtrack downsample conv: tile 2 x 60 output pixels
    (1, 120, 360, 128, 128, 1),    # Wildtrack upsample conv: tile 1 x 120
    (6, 160, 250, 128, 128, 2),    # MultiviewX: output width 125 = one tile row
    (2, 37, 64, 32, 20, 1),        # partial tiles in y (TY = 2 over 37 rows), narrow N, one channel chunk
    (3, 33, 48, 64, 132, 2),       # odd input height, two column tiles
])
def test_implicit_conv3x3_is_the_im2col_gemm_bit_for_bit(cuda, NB, Hi, Wi, C, N, stride):
    """The implicit-GEMM convolution (taps fetched by TMA from the channels-last image, zero padding = TMA out-of-bounds
    fill) against F.conv2d in fp64, and bit for bit against the same kernel fed the materialised (ky, kx, c) im2col matrix."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(NB * 1000 + Hi)
    x = torch.randn(NB, C, Hi, Wi, generator=g).to(cuda)
    w = (torch.randn(N, C, 3, 3, generator=g) / (9 * C) ** 0.5).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    x_cl = x.permute(0, 2, 3, 1).contiguous()
    w2d = w.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous()
    for mode in ("f16x2", "bf16x3"):
        old = ops._GEMM_MODE
        ops._GEMM_MODE = mode
        try:
            got = ops.conv3x3_nhwc(x_cl, w2d, b, stride=stride, relu=True)
        finally:
            ops._GEMM_MODE = old
        assert got is not None
        Ho, Wo = (Hi - 1) // stride + 1, (Wi - 1) // stride + 1
        exact = F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=1).clamp_min(0)
        exact = exact.permute(0, 2, 3, 1).reshape(NB * Ho * Wo, N)
        ref32 = F.conv2d(x, w, b, stride=stride, padding=1).clamp_min(0).permute(0, 2, 3, 1).reshape(NB * Ho * Wo, N)
        err, err32 = (got.double() - exact).abs().max().item(), (ref32.double() - exact).abs().max().item()
        assert err <= max(4 * err32, 4e-6 * exact.abs().max().item()), (mode, err, err32)
        # materialised im2col matrix, columns (ky, kx, c): unfold gives (c, ky, kx) -> reorder
        cols = F.unfold(x, 3, padding=1, stride=stride).view(NB, C, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(NB * Ho * Wo, 9 * C)
        via_matrix = ops.linear(cols.contiguous(), w2d, b, relu=True, mode=mode)
        assert torch.equal(got, via_matrix), mode
    # second output of the same epilogue: out + addend (the first encoder layer's query, src + pos)
    add = torch.randn(got.shape, generator=g).to(cuda)
    o1, o2 = ops.conv3x3_nhwc(x_cl, w2d, b, stride=stride, relu=True, add=add)
    assert torch.equal(o1, ops.conv3x3_nhwc(x_cl, w2d, b, stride=stride, relu=True)) and torch.equal(o2, o1 + add)


def test_implicit_conv3x3_reports_unsupported_widths(cuda):
    x_cl = torch.randn(1, 8, 131, 32, device=cuda)   # output width 131 is prime and > 128: no tile width
    w2d = torch.randn(16, 9 * 32, device=cuda)
    assert ops.conv3x3_nhwc(x_cl, w2d, None, stride=1) is None


def test_upsample_nhwc_matches_aten(cuda):
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 128, 60, 180, generator=g).to(cuda)
    x_cl = x.permute(0, 2, 3, 1).contiguous()
    ref = F.interpolate(x, size=(120, 360), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    got = ops.upsample_nhwc(x_cl, (120, 360))
    assert (got - ref).abs().max().item() <= 1e-6
    band = torch.zeros_like(got)
    ops.upsample_nhwc(x_cl, (120, 360), rows=(17, 40), out=band)
    assert torch.equal(band[:, 17:57], got[:, 17:57]) and band[:, :17].abs().max().item() == 0 and band[:, 57:].abs().max().item() == 0
    # same arithmetic as the im2col variant's centre tap
    A = ops.upsample_im2col(x_cl[:1].contiguous(), (120, 360))
    assert torch.equal(A.view(120, 360, 9, 128)[:, :, 4], got[0])
