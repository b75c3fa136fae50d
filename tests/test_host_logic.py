"""CPU: argument validation of the host-side mirror matches the reference's error behaviour (no GPU needed)."""
import os
import sys

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
    assert shims in sys.path[:2]
    msda = importlib.import_module("MultiScaleDeformableAttention")
    assert callable(msda.ms_deform_attn_forward) and callable(msda.ms_deform_attn_backward)
    kornia = importlib.import_module("kornia")  # no real kornia in this image: the stub package
    assert kornia.warp_perspective is ops.warp_perspective and kornia.__mvdetr_b200_stub__


def test_a_real_kornia_is_wrapped_not_shadowed(tmp_path):
    """with a real kornia installed, install_shims() must leave it importable and only route the hot-path
    call (fp32 CUDA, bilinear / zeros / align_corners=False) to our kernel; e.g. the reference's dataset calls
    kornia.warp_perspective(masks_cpu, M, size, 'nearest', align_corners=False) (frameDataset.py:80)."""
    import os
    import subprocess
    import sys
    pkg = tmp_path / "kornia"
    pkg.mkdir()
    (pkg / "__init__.py").write_text(
        "calls = []\n"
        "def warp_perspective(src, M, dsize, mode='bilinear', padding_mode='zeros', align_corners=None):\n"
        "    calls.append((mode, align_corners))\n"
        "    return 'real-kornia'\n"
        "def other():\n    return 'untouched'\n")
    code = (
        "import sys, torch\n"
        f"sys.path.insert(0, {str(tmp_path)!r})\n"
        "import mvdetr_b200\n"
        "mvdetr_b200.install_shims()\n"
        "import kornia\n"
        "assert not hasattr(kornia, '__mvdetr_b200_stub__') and kornia.other() == 'untouched'\n"
        "src, M = torch.zeros(1, 1, 4, 4), torch.eye(3)[None]\n"
        "assert kornia.warp_perspective(src, M, (4, 4), 'nearest', align_corners=False) == 'real-kornia'\n"
        "assert kornia.warp_perspective(src, M, (4, 4), align_corners=False) == 'real-kornia'  # CPU tensor\n"
        "assert kornia.calls == [('nearest', False), ('bilinear', False)]\n"
        "assert kornia.warp_perspective.__mvdetr_b200_wrapped__\n"
        "mvdetr_b200.install_shims()  # idempotent\n"
        "assert kornia.warp_perspective.__wrapped__.__module__ == 'kornia'\n"
        "print('OK')\n")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=repo)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


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


def test_conv_as_gemm_weight_layout_matches_the_im2col_definition():
    """CPU: DeformTransWorldFeat.gemm_weights() ((ky, kx, c_in) column order) applied to the im2col matrix defined in
    include/mvdetr_b200.h (built here with F.unfold) equals the module's own convolutions, merge conv included."""
    import torch
    import torch.nn.functional as F
    from mvdetr_b200.world_feat import DeformTransWorldFeat

    torch.manual_seed(0)
    N, C, Hg, Wg = 3, 8, 10, 14
    wf = DeformTransWorldFeat(N, [Hg, Wg], C, hidden_dim=C, nhead=2, dim_feedforward=16, n_points=4, stride=2).eval()
    Wd, Wm, Wu = wf.gemm_weights()

    def im2col(x, stride):
        BN, Cc = x.shape[:2]
        cols = F.unfold(x, 3, padding=1, stride=stride)
        return cols.view(BN, Cc, 9, -1).permute(0, 3, 2, 1).reshape(-1, 9 * Cc)

    x = torch.randn(N, C, Hg, Wg)
    want = wf.downsample(x)                                            # [N, C, Hd, Wd]
    Hd, Wd_ = want.shape[-2:]
    got = torch.relu(im2col(x, 2) @ Wd.t() + wf.downsample[0].bias)    # [N*Hd*Wd, C] token-major
    assert torch.allclose(got.view(N, Hd, Wd_, C).permute(0, 3, 1, 2), want, atol=1e-5)
    # merge 1x1 conv over (view, channel): cell-major rows [cells, N*C]
    mem = torch.randn(1, N * Hd * Wd_, C)
    want_m = wf.merge_linear(mem.view(1, N, Hd, Wd_, C).permute(0, 1, 4, 2, 3).reshape(1, N * C, Hd, Wd_))
    mem_cm = mem.view(N, Hd * Wd_, C).permute(1, 0, 2).reshape(Hd * Wd_, N * C)
    got_m = torch.relu(mem_cm @ Wm.t() + wf.merge_linear[0].bias)
    assert torch.allclose(got_m.view(1, Hd, Wd_, C).permute(0, 3, 1, 2), want_m, atol=1e-5)
    # upsample + 3x3 conv
    want_u = wf.upsample(want_m)
    up = F.interpolate(want_m, size=(Hg, Wg), mode="bilinear", align_corners=False)
    got_u = torch.relu(im2col(up, 1) @ Wu.t() + wf.upsample[1].bias)
    assert torch.allclose(got_u.view(1, Hg, Wg, C).permute(0, 3, 1, 2), want_u, atol=1e-5)
    # cache invalidation when a weight changes in place
    with torch.no_grad():
        wf.downsample[0].weight.mul_(2.0)
    assert torch.allclose(wf.gemm_weights()[0], 2.0 * Wd)


@pytest.mark.skipif(not os.path.isdir("/root/reference/multiview_detector"), reason="needs the reference checkout")
def test_mirror_matches_the_reference_classes_and_wraps_them():
    """Where the reference tree is importable (the build container): our mirror of the callers builds the same position
    table, the same initial parameters from the same RNG stream and the same state_dict keys as the reference classes,
    and from_reference() shares an existing reference module's parameters instead of copying them."""
    import types
    for name in ("MultiScaleDeformableAttention", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    import numpy as np
    import multiview_detector.models.trans_world_feat as rwf
    from mvdetr_b200 import world_feat as wf
    for hw, feats in (((60, 180), 64), ((7, 5), 8), ((1, 1), 4)):
        assert torch.equal(rwf.create_pos_embedding(np.array(hw), feats), wf.create_pos_embedding(np.array(hw), feats))
    ref_pts = torch.rand(3 * 6 * 10, 3, 4, 2)
    kw = dict(hidden_dim=32, nhead=8, dim_feedforward=64, n_points=4, stride=2, reference_points=ref_pts)
    torch.manual_seed(7)
    theirs = rwf.DeformTransWorldFeat(3, [12, 20], 32, **kw)
    torch.manual_seed(7)
    ours = wf.DeformTransWorldFeat(3, [12, 20], 32, **kw)
    sd_a, sd_b = theirs.state_dict(), ours.state_dict()
    assert sorted(sd_a) == sorted(sd_b) and all(torch.equal(sd_a[k], sd_b[k]) for k in sd_a)
    wrapped = wf.from_reference(theirs)
    shared = dict(wrapped.named_parameters())
    assert all(shared[n] is p for n, p in theirs.named_parameters())
