"""Generates the golden vectors under tests/golden/ by importing the REFERENCE's own Python code.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
The GPU box has no /root/reference, so the vectors are committed; nothing in tests/ reads the reference at run time.

What the reference contributes (hou-yz/MVDeTr @ 66cae15):
  * ms_deform_attn_core_pytorch            multiview_detector/models/ops/functions/ms_deform_attn_func.py:41-61
    -- run on CPU in fp64/fp32 on the shapes/seed/distributions of the reference's own op test
       (multiview_detector/models/ops/test.py:21-36), on BASELINE config 1, and on MVDeTr-shaped mini problems;
       its autograd gradients pin the backward.
  * MSDeformAttn (module)                  multiview_detector/models/ops/modules/ms_deform_attn.py:79-117
  * DeformTransWorldFeat + create_reference_map + get_worldcoord_from_imgcoord_mat
                                            multiview_detector/models/trans_world_feat.py:70-119, mvdetr.py:33-71,
                                            utils/projection.py:27-43
    with MSDeformAttnFunction rebound to the pure-PyTorch core (the CUDA extension cannot run here).
The warp has NO reference golden (kornia absent, parity unpinned); warp_*.npz are produced by the kornia restatement
in oracle/torch_port.py and pin only that our C oracle / CUDA kernel agree with that restatement.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)

# --- stubs so the reference modules import without its compiled extension / plotting deps ------------------
sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))
_mpl = types.ModuleType("matplotlib")
_plt = types.ModuleType("matplotlib.pyplot")
_mpl.pyplot = _plt
sys.modules.setdefault("matplotlib", _mpl)
sys.modules.setdefault("matplotlib.pyplot", _plt)
sys.path.insert(0, REF)

from multiview_detector.models.ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch  # noqa: E402
import multiview_detector.models.ops.modules.ms_deform_attn as ref_mod  # noqa: E402
from multiview_detector.models.trans_world_feat import DeformTransWorldFeat  # noqa: E402
from multiview_detector.utils.projection import get_worldcoord_from_imgcoord_mat  # noqa: E402


class _CoreAsFunction:
    """Stands in for MSDeformAttnFunction (ms_deform_attn_func.py:21) with the pure-PyTorch core."""

    @staticmethod
    def apply(value, shapes, start, loc, attn, im2col_step):
        return ms_deform_attn_core_pytorch(value, shapes, loc, attn)


ref_mod.MSDeformAttnFunction = _CoreAsFunction


def start_index(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def core_with_grads(value, shapes, loc, attn, grad_out):
    v, l, a = (t.clone().requires_grad_(True) for t in (value, loc, attn))
    out = ms_deform_attn_core_pytorch(v, shapes, l, a)
    out.backward(grad_out)
    return out.detach(), v.grad, l.grad, a.grad


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def gen_opstest():
    """Exact RNG stream of ops/test.py: seed 3, then the double check, then the float check."""
    N, M, D = 1, 2, 2
    Lq, L, P = 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    torch.manual_seed(3)
    out = {}
    for tag in ("f64", "f32"):
        value = torch.rand(N, S, M, D) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2)
        attn = torch.rand(N, Lq, M, L, P) + 1e-5
        attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
        if tag == "f64":
            value, loc, attn = value.double(), loc.double(), attn.double()
        o = ms_deform_attn_core_pytorch(value, shapes, loc, attn)
        out.update({f"value_{tag}": value, f"loc_{tag}": loc, f"attn_{tag}": attn, f"out_{tag}": o})
    # gradients for the D list of test.py:85-86 are checked by gradcheck on the GPU; here: autograd goldens at D=4
    value = (torch.rand(N, S, M, 4) * 0.01).double()
    loc = torch.rand(N, Lq, M, L, P, 2).double()
    attn = (torch.rand(N, Lq, M, L, P) + 1e-5).double()
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    go = torch.rand(N, Lq, M * 4).double()
    o, gv, gl, ga = core_with_grads(value, shapes, loc, attn, go)
    out.update(dict(g_value=value, g_loc=loc, g_attn=attn, g_grad_out=go, g_out=o, g_grad_value=gv, g_grad_loc=gl,
                    g_grad_attn=ga))
    save("msda_opstest.npz", shapes=shapes, start=start_index(shapes), **out)


def gen_config1():
    """BASELINE.json configs[0]: single query, N=2 'views' (batch), L=1, K=4, C=64, with out-of-range locations."""
    torch.manual_seed(3)
    B, M, D, Lq, L, P = 2, 1, 64, 1, 1, 4
    shapes = torch.as_tensor([(8, 8)], dtype=torch.long)
    S = 64
    value = (torch.rand(B, S, M, D) * 0.01).double()
    loc = (torch.rand(B, Lq, M, L, P, 2) * 1.4 - 0.2).double()
    attn = (torch.rand(B, Lq, M, L, P) + 1e-5).double()
    attn /= attn.sum(-1, keepdim=True).sum(-2, keepdim=True)
    go = torch.rand(B, Lq, M * D).double()
    o, gv, gl, ga = core_with_grads(value, shapes, loc, attn, go)
    o32 = ms_deform_attn_core_pytorch(value.float(), shapes, loc.float(), attn.float())
    save("msda_config1.npz", shapes=shapes, start=start_index(shapes), value=value, loc=loc, attn=attn, grad_out=go,
         out=o, out_f32=o32, grad_value=gv, grad_loc=gl, grad_attn=ga)


def gen_mvdetr_mini():
    """MVDeTr-shaped mini problem: L=3 views of a 6x10 grid, queries = all view pixels (Lq = S), M=2, D=16, P=4,
    locations = identity reference + offsets of a few pixels, some leaving the map; plus ragged levels, B=2."""
    torch.manual_seed(11)
    L, H, W, M, D, P = 3, 6, 10, 2, 16, 4
    shapes = torch.as_tensor([(H, W)] * L, dtype=torch.long)
    S = L * H * W
    Lq = S
    ys, xs = torch.meshgrid(torch.linspace(0.5, H - 0.5, H), torch.linspace(0.5, W - 0.5, W), indexing="ij")
    ref = torch.stack((xs / W, ys / H), -1).reshape(1, H * W, 1, 1, 1, 2).repeat(1, L, 1, 1, 1, 1)
    off = torch.randn(1, Lq, M, L, P, 2) * 2.5 / torch.tensor([W, H])
    value = torch.randn(1, S, M, D).double()
    loc = (ref + off).double()
    attn = torch.softmax(torch.randn(1, Lq, M, L * P), -1).view(1, Lq, M, L, P).double()
    go = torch.randn(1, Lq, M * D).double()
    o, gv, gl, ga = core_with_grads(value, shapes, loc, attn, go)
    save("msda_mvdetr_mini.npz", shapes=shapes, start=start_index(shapes), value=value, loc=loc, attn=attn,
         grad_out=go, out=o, grad_value=gv, grad_loc=gl, grad_attn=ga)

    # ragged levels, batch 2, Lq != S, odd head dim
    shapes = torch.as_tensor([(7, 5), (4, 9), (1, 3), (2, 2)], dtype=torch.long)
    S = int(shapes.prod(1).sum())
    B, Lq, M, D, L, P = 2, 13, 3, 5, 4, 3
    value = torch.randn(B, S, M, D).double()
    loc = (torch.rand(B, Lq, M, L, P, 2) * 1.3 - 0.15).double()
    attn = torch.softmax(torch.randn(B, Lq, M, L * P), -1).view(B, Lq, M, L, P).double()
    go = torch.randn(B, Lq, M * D).double()
    o, gv, gl, ga = core_with_grads(value, shapes, loc, attn, go)
    save("msda_ragged.npz", shapes=shapes, start=start_index(shapes), value=value, loc=loc, attn=attn, grad_out=go,
         out=o, grad_value=gv, grad_loc=gl, grad_attn=ga)


from mvdetr_b200.synthetic import mini_scene as mini_dataset  # noqa: E402  (scene generator shared with the tests)


def gen_world_feat_mini():
    """Reference DeformTransWorldFeat + MVDeTr's projection chain on a mini 3-camera scene (weights saved)."""
    from oracle import torch_port as tp
    ds = mini_dataset()
    N, C = ds.num_cam, 32
    # --- projection chain exactly as MVDeTr.__init__ / forward compute it (mvdetr.py:82-95,155-161)
    world_zoom = np.diag([ds.world_reduce, ds.world_reduce, 1])
    Rworldgrid_from_worldcoord = np.linalg.inv(ds.base.worldcoord_from_worldgrid_mat @ world_zoom @
                                               ds.base.world_indexing_from_xy_mat)
    w_from_i = [get_worldcoord_from_imgcoord_mat(ds.base.intrinsic_matrices[c], ds.base.extrinsic_matrices[c], 0)
                for c in range(N)]
    proj_mats64 = torch.stack([torch.from_numpy(Rworldgrid_from_worldcoord @ w_from_i[c]) for c in range(N)])
    torch.manual_seed(5)
    M_aug = torch.eye(3).view(1, 1, 3, 3).repeat(1, N, 1, 1)
    M_aug[0, 1] = torch.tensor([[0.9, 0.0, 12.0], [0.0, 0.9, -7.0], [0.0, 0.0, 1.0]])  # a scale+shift augmentation
    inv_aff = torch.inverse(M_aug.view(N, 3, 3))
    img_from_Rimg = inv_aff @ torch.from_numpy(np.diag([ds.img_reduce, ds.img_reduce, 1])).view(1, 3, 3).repeat(
        N, 1, 1).float()
    proj_mats = proj_mats64.view(N, 3, 3).float() @ img_from_Rimg

    # --- reference create_reference_map, imported lazily because mvdetr.py imports kornia at module top
    kornia_stub = types.ModuleType("kornia")
    kornia_stub.warp_perspective = lambda src, M, dsize, align_corners=False: tp.warp_perspective(src, M, dsize)
    sys.modules.setdefault("kornia", kornia_stub)
    from multiview_detector.models.mvdetr import create_reference_map
    ref_points = create_reference_map(ds, 4).repeat([N, 1, 1, 1])

    model = DeformTransWorldFeat(N, ds.Rworld_shape, C, hidden_dim=C, nhead=4, dim_feedforward=64, n_points=4,
                                 stride=2, reference_points=ref_points).eval()
    with torch.no_grad():
        for layer in model.encoder.layers:  # make offsets / weights query dependent (default init has zero weight)
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.05)
            layer.self_attn.attention_weights.weight.normal_(0, 0.2)
        imgs_feat = torch.randn(N, C, *ds.Rimg_shape)
        world_in = tp.warp_perspective(imgs_feat, proj_mats, tuple(ds.Rworld_shape))  # kornia restatement (unpinned)
        out = model(world_in.view(1, N, C, *ds.Rworld_shape))
    sd = {k: v for k, v in model.state_dict().items()}
    save("world_feat_mini.npz", proj_mats=proj_mats, proj_mats64=proj_mats64, ref_points=ref_points,
         imgs_feat=imgs_feat, world_in=world_in, out=out, Rworld=np.array(ds.Rworld_shape),
         Rimg=np.array(ds.Rimg_shape), nhead=4, **{"sd." + k: v for k, v in sd.items()})

    # --- one MSDeformAttn module call in isolation (pins the fused loc/softmax entry point)
    attn_mod = model.encoder.layers[0].self_attn
    Hd, Wd = ds.Rworld_shape[0] // 2, ds.Rworld_shape[1] // 2
    Lq = N * Hd * Wd
    torch.manual_seed(6)
    query, src = torch.randn(1, Lq, C), torch.randn(1, Lq, C)
    shapes = torch.as_tensor([[Hd, Wd]] * N, dtype=torch.long)
    with torch.no_grad():
        o = attn_mod(query, ref_points.unsqueeze(0), src, shapes, start_index(shapes))
        value = attn_mod.value_proj(src).view(1, Lq, 4, C // 4)
        offsets = attn_mod.sampling_offsets(query).view(1, Lq, 4, N, 4, 2)
        logits = attn_mod.attention_weights(query).view(1, Lq, 4, N * 4)
        attn = torch.softmax(logits, -1).view(1, Lq, 4, N, 4)
        norm = torch.stack([shapes[..., 1], shapes[..., 0]], -1)
        loc = ref_points.unsqueeze(0)[:, :, None] + offsets / norm[None, None, None, :, None, :]
        core = ms_deform_attn_core_pytorch(value, shapes, loc, attn)
    save("msda_module_mini.npz", query=query, src=src, shapes=shapes, start=start_index(shapes), value=value,
         offsets=offsets, logits=logits, attn=attn, loc=loc, ref_table=ref_points[:Hd * Wd], core_out=core,
         module_out=o)


def gen_warp():
    """Kornia-restatement outputs (NOT reference outputs; parity unpinned) for C-oracle / CUDA agreement tests."""
    from oracle import torch_port as tp
    torch.manual_seed(7)
    src = torch.randn(3, 5, 18, 32)
    mats = torch.tensor([
        [[1.9, 0.1, 3.0], [0.05, 1.2, -2.0], [1e-3, 2e-3, 1.0]],
        [[0.8, -0.3, 10.0], [0.2, 1.1, 4.0], [-2e-3, 1e-3, 1.0]],
        [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]],
    ])
    dsize = (24, 40)
    out = tp.warp_perspective(src, mats, dsize)
    s = src.clone().requires_grad_(True)
    g = torch.randn_like(out)
    tp.warp_perspective(s, mats, dsize).backward(g)
    save("warp_small.npz", src=src, mats=mats, dsize=np.array(dsize), out=out, grad_out=g, grad_src=s.grad)


if __name__ == "__main__":
    torch.set_num_threads(4)
    gen_opstest()
    gen_config1()
    gen_mvdetr_mini()
    gen_world_feat_mini()
    gen_warp()
