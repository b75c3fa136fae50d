"""CPU: argument validation of the host-side mirror matches the reference's error behaviour (no GPU needed)."""
import pytest
import torch

import mvdetr_b200
from mvdetr_b200 import ops


def _args(B=1, S=4, M=1, D=4, L=1, Lq=1, P=1, dtype=torch.float32):
    return [torch.zeros(B, S, M, D, dtype=dtype), torch.tensor([[2, 2]] * L), torch.zeros(L, dtype=torch.long),
            torch.zeros(B, Lq, M, L, P, 2, dtype=dtype), torch.zeros(B, Lq, M, L, P, dtype=dtype), 64]


def test_cpu_tensors_raise_like_the_reference():
    # ref: multiview_detector/models/ops/src/ms_deform_attn.h:38,60
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        ops.ms_deform_attn_forward(*_args())
    a = _args()
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        ops.ms_deform_attn_backward(*a[:5], torch.zeros(1, 1, 4), 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        ops.MSDeformAttnFunction.apply(*_args())


def test_shim_modules_expose_the_reference_names():
    import importlib
    import sys
    shims = mvdetr_b200.install_shims()
    assert sys.path[0] == shims
    msda = importlib.import_module("MultiScaleDeformableAttention")
    assert callable(msda.ms_deform_attn_forward) and callable(msda.ms_deform_attn_backward)
    kornia = importlib.import_module("kornia")
    assert kornia.warp_perspective is ops.warp_perspective


def test_warp_argument_errors():
    src, M = torch.zeros(2, 3, 4, 5), torch.eye(3).repeat(2, 1, 1)
    with pytest.raises(NotImplementedError):
        ops.warp_perspective(src, M, (4, 5), mode="nearest", align_corners=False)
    with pytest.raises(NotImplementedError):
        ops.warp_perspective(src, M, (4, 5))  # align_corners=None is kornia's legacy default (True)
    with pytest.raises(ValueError):
        ops.warp_perspective(src[0], M, (4, 5), align_corners=False)
    with pytest.raises(ValueError):
        ops.warp_perspective(src, M[:1], (4, 5), align_corners=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.warp_perspective(src, M, (4, 5), align_corners=False)
