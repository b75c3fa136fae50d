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


def test_mirror_loads_reference_state_dict_and_projection_chain(golden):
    """CPU: parameter names of the mirror are the reference's (strict load of a reference state_dict), buffers stay out
    of the state_dict, and the projection / reference-map helpers reproduce the reference's numbers."""
    import numpy as np
    from mvdetr_b200 import projection, synthetic
    from mvdetr_b200.world_feat import DeformTransWorldFeat
    g = golden("world_feat_mini.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g if k.startswith("sd.")}
    N, C = g["imgs_feat"].shape[:2]
    model = DeformTransWorldFeat(N, list(g["Rworld"]), C, hidden_dim=C, nhead=int(g["nhead"]), dim_feedforward=64,
                                 reference_points=torch.from_numpy(g["ref_points"]))
    model.load_state_dict(sd, strict=True)
    assert set(model.state_dict()) == set(sd)
    ds = synthetic.mini_scene()
    assert np.array_equal(projection.world_grid_projection_mats(ds).numpy(), g["proj_mats64"])
    ref = projection.create_reference_map(ds, 4).repeat([N, 1, 1, 1])
    assert torch.equal(ref, torch.from_numpy(g["ref_points"]))
    assert model.encoder.ref_table.shape[0] * N == ref.shape[0]
