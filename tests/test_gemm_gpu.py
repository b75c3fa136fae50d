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
