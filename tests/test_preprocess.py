"""Input-side transform (ToTensor -> Normalize -> Resize, frameDataset.py:66-67).
CPU: the numpy restatement against the golden made by the torchvision Compose the reference uses.
GPU: the CUDA kernel (through the C ABI) against the same golden and, at the real size (1080x1920 -> 720x1280), against
torch's own antialiased interpolation on the host."""
import numpy as np
import pytest
import torch

from oracle import preprocess_ref as pr

CASES = ("down_1p5", "down_odd", "down_3x", "up", "same")
TOL = 5e-6  # fp32 rounding of the normalised values (|x| <= 2.7, ulp 2.4e-7) through two weighted sums


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_torchvision_golden(golden, name):
    g = golden("preprocess.npz")
    want = g[f"{name}.out"]
    assert np.abs(pr.resize_normalize(g[f"{name}.img"], want.shape[1:], antialias=True) - want).max() <= TOL
    assert np.abs(pr.resize_normalize(g[f"{name}.img"], want.shape[1:], antialias=False) - g[f"{name}.out_noaa"]).max() <= TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_kernel_matches_torchvision_golden(golden, cuda, name):
    from mvdetr_b200 import preprocess
    g = golden("preprocess.npz")
    img = torch.from_numpy(g[f"{name}.img"])[None].to(cuda)
    want = g[f"{name}.out"]
    got = preprocess.resize_normalize(img, want.shape[1:])
    assert np.abs(got[0].cpu().numpy() - want).max() <= TOL
    plain = preprocess.resize_normalize(img, want.shape[1:], antialias=False)
    assert np.abs(plain[0].cpu().numpy() - g[f"{name}.out_noaa"]).max() <= TOL


@pytest.mark.gpu
def test_kernel_full_size_views(cuda):
    """7 views of 1080x1920 -> 720x1280 (Wildtrack): view 0 and 6 against torch's antialiased bilinear on the host."""
    from mvdetr_b200 import preprocess
    g = torch.Generator().manual_seed(0)
    imgs = torch.randint(0, 256, (7, 1080, 1920, 3), generator=g, dtype=torch.uint8)
    size = preprocess.network_input_size((1080, 1920), 12)
    assert size == [720, 1280]
    got = preprocess.resize_normalize(imgs.to(cuda), size)
    assert got.shape == (7, 3, 720, 1280)
    mean = torch.tensor(preprocess.IMAGENET_MEAN).view(3, 1, 1)
    std = torch.tensor(preprocess.IMAGENET_STD).view(3, 1, 1)
    for v in (0, 6):
        x = (imgs[v].permute(2, 0, 1).float().div(255) - mean) / std
        want = torch.nn.functional.interpolate(x[None], size=size, mode="bilinear", align_corners=False, antialias=True)[0]
        assert (got[v].cpu() - want).abs().max().item() <= TOL


@pytest.mark.gpu
def test_argument_errors(cuda):
    from mvdetr_b200 import preprocess
    with pytest.raises(RuntimeError, match="CUDA"):
        preprocess.resize_normalize(torch.zeros(1, 4, 4, 3, dtype=torch.uint8), (2, 2))
    with pytest.raises(ValueError):
        preprocess.resize_normalize(torch.zeros(1, 3, 4, 4, dtype=torch.uint8, device=cuda), (2, 2))
    with pytest.raises(RuntimeError, match="not supported"):
        preprocess.resize_normalize(torch.zeros(1, 64, 64, 3, dtype=torch.uint8, device=cuda), (4, 4))  # 16x down, aa
