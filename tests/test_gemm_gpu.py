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
