"""torch-CPU restatement of the reference's own CPU-capable path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This is what `bench.py --impl reference` and the `cpu_baseline` leg time (kind "port": /root/reference does not
exist on the GPU box, and the reference's C++ CPU entry points are stubs that throw:
multiview_detector/models/ops/src/cpu/ms_deform_attn_cpu.cpp:26,39).

  msda_core            follows ms_deform_attn_core_pytorch   multiview_detector/models/ops/functions/ms_deform_attn_func.py:41-61
  warp_perspective     follows kornia.warp_perspective as called at multiview_detector/models/mvdetr.py:194-195
                       (third-party, restated from its published algorithm; PARITY UNPINNED, see oracle/warp_ref.c)
  msda_module_forward  follows MSDeformAttn.forward          multiview_detector/models/ops/modules/ms_deform_attn.py:79-117
  encoder_layer_forward / world_feat_forward
                       follow DeformableTransformerEncoderLayer.forward (deformable_transformer.py:75-85) and
                       DeformTransWorldFeat.forward (trans_world_feat.py:87-110), functional over a state_dict
                       with the reference's parameter names.
"""
import math

import torch
import torch.nn.functional as F


def msda_core(value, shapes, loc, attn):
    """value [B,S,M,D]; shapes: list of (H,W); loc [B,Lq,M,L,P,2] in [0,1]; attn [B,Lq,M,L,P] -> [B,Lq,M*D]."""
    B, S, M, D = value.shape
    Lq, P = loc.shape[1], loc.shape[4]
    out = value.new_zeros(B * M, D, Lq)
    begin = 0
    for lvl, (H, W) in enumerate(shapes):
        H, W = int(H), int(W)
        plane = value[:, begin:begin + H * W].permute(0, 2, 3, 1).reshape(B * M, D, H, W)
        begin += H * W
        grid = (loc[:, :, :, lvl] * 2 - 1).permute(0, 2, 1, 3, 4).reshape(B * M, Lq, P, 2)
        taps = F.grid_sample(plane, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
        w = attn[:, :, :, lvl].permute(0, 2, 1, 3).reshape(B * M, 1, Lq, P)
        out += (taps * w).sum(-1)
    return out.view(B, M * D, Lq).transpose(1, 2).contiguous()


def _pixel_to_normalized(h, w, dtype):
    eps = 1e-14
    wd = eps if w == 1 else w - 1.0
    hd = eps if h == 1 else h - 1.0
    return torch.tensor([[2.0 / wd, 0.0, -1.0], [0.0, 2.0 / hd, -1.0], [0.0, 0.0, 1.0]], dtype=dtype)


def normalized_inverse_homography(mat, src_hw, dst_hw):
    """kornia normalize_homography + inverse, in mat's dtype (fp32 in the reference)."""
    n_src = _pixel_to_normalized(src_hw[0], src_hw[1], mat.dtype).to(mat.device)
    n_dst = _pixel_to_normalized(dst_hw[0], dst_hw[1], mat.dtype).to(mat.device)
    dst_from_src = n_dst @ (mat @ torch.inverse(n_src))
    return torch.inverse(dst_from_src)


def warp_grid(mat, src_hw, dst_hw):
    T = normalized_inverse_homography(mat, src_hw, dst_hw)
    Ho, Wo = dst_hw
    xs = torch.linspace(-1, 1, Wo, dtype=mat.dtype, device=mat.device)
    ys = torch.linspace(-1, 1, Ho, dtype=mat.dtype, device=mat.device)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    pts = torch.stack([gx, gy, torch.ones_like(gx)], -1).reshape(1, Ho * Wo, 3)
    hom = torch.bmm(pts.expand(mat.shape[0], -1, -1), T.transpose(1, 2))
    z = hom[..., 2:3]
    eps = 1e-8
    scale = torch.where(z.abs() > eps, 1.0 / (z + eps), torch.ones_like(z))
    return (hom[..., :2] * scale).view(mat.shape[0], Ho, Wo, 2)


def warp_perspective(src, mat, dsize):
    """src [BN,C,Hi,Wi], mat [BN,3,3] src-pixel -> dst-pixel, dsize (Ho,Wo); bilinear, zeros, align_corners=False."""
    grid = warp_grid(mat.to(src.dtype), tuple(src.shape[-2:]), tuple(dsize))
    return F.grid_sample(src, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


def msda_module_forward(sd, prefix, query, ref_points, src, shapes, n_heads, n_points, msda_fn=None,
                        shapes_dev=None):
    """MSDeformAttn.forward with per-(level, point) reference points [B,Lq,L,P,2] (MVDeTr's modification,
    ms_deform_attn.py:104-107). sd: state_dict, prefix e.g. 'encoder.layers.0.self_attn.'.
    msda_fn(value, shapes_dev, start_dev, loc, attn): replaces msda_core (bench.py plugs the reference's own CUDA op
    in here); shapes_dev = ([L,2] int64 spatial_shapes, [L] level_start_index, sync) on the device as the reference's
    caller builds them (trans_world_feat.py:95-96); sync=True reproduces its device->host assert
    (ms_deform_attn.py:94)."""
    B, Lq, C = query.shape
    L = len(shapes)
    if shapes_dev is not None and shapes_dev[2]:
        assert (shapes_dev[0][:, 0] * shapes_dev[0][:, 1]).sum() == src.shape[1]
    value = F.linear(src, sd[prefix + "value_proj.weight"], sd[prefix + "value_proj.bias"])
    value = value.view(B, src.shape[1], n_heads, C // n_heads)
    off = F.linear(query, sd[prefix + "sampling_offsets.weight"], sd[prefix + "sampling_offsets.bias"])
    off = off.view(B, Lq, n_heads, L, n_points, 2)
    aw = F.linear(query, sd[prefix + "attention_weights.weight"], sd[prefix + "attention_weights.bias"])
    aw = F.softmax(aw.view(B, Lq, n_heads, L * n_points), -1).view(B, Lq, n_heads, L, n_points)
    norm = torch.tensor([[w, h] for h, w in shapes], dtype=query.dtype, device=query.device)
    loc = ref_points[:, :, None] + off / norm[None, None, None, :, None, :]
    if msda_fn is not None:
        out = msda_fn(value, shapes_dev[0], shapes_dev[1], loc, aw)
    else:
        out = msda_core(value, shapes, loc, aw)
    return F.linear(out, sd[prefix + "output_proj.weight"], sd[prefix + "output_proj.bias"])


def encoder_layer_forward(sd, prefix, src, pos, ref_points, shapes, n_heads, n_points, msda_fn=None, shapes_dev=None):
    """Eval-mode (dropout = identity) DeformableTransformerEncoderLayer.forward."""
    C = src.shape[-1]
    a = msda_module_forward(sd, prefix + "self_attn.", src + pos, ref_points, src, shapes, n_heads, n_points,
                            msda_fn, shapes_dev)
    src = F.layer_norm(src + a, (C,), sd[prefix + "norm1.weight"], sd[prefix + "norm1.bias"])
    f = F.linear(F.relu(F.linear(src, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"])),
                 sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"])
    return F.layer_norm(src + f, (C,), sd[prefix + "norm2.weight"], sd[prefix + "norm2.bias"])


def sine_pos_embedding(hw, num_pos_feats, temperature=10000.0):
    """create_pos_embedding (trans_world_feat.py:15-37), normalize=True, scale=2*pi -> [1, 2*num_pos_feats, H, W]."""
    H, W = int(hw[0]), int(hw[1])
    ones = torch.ones(1, H, W)
    y = ones.cumsum(1, dtype=torch.float32)
    x = ones.cumsum(2, dtype=torch.float32)
    y = y / (y[:, -1:, :] + 1e-6) * (2 * math.pi)
    x = x / (x[:, :, -1:] + 1e-6) * (2 * math.pi)
    i = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / num_pos_feats)
    px, py = x[..., None] / dim_t, y[..., None] / dim_t
    px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), dim=4).flatten(3)
    py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), dim=4).flatten(3)
    return torch.cat((py, px), dim=3).permute(0, 3, 1, 2)


def world_feat_forward(sd, x, ref_points, n_heads=8, n_points=4, n_layers=3, stride=2, pos_embedding=None,
                       msda_fn=None, sync_assert=False):
    """DeformTransWorldFeat.forward (eval mode) over state_dict `sd` (reference key names, no prefix).
    x [1,N,C,H,W]; ref_points [N*Hd*Wd, N, P, 2].
    As in the reference, `pos_embedding` [1,C,Hd,Wd] (built once; trans_world_feat.py:78) and `ref_points` are moved
    to x.device on EVERY call (trans_world_feat.py:93, deformable_transformer.py:48): hand in host tensors to time
    the path as shipped, device tensors to time its kernels only. sync_assert reproduces ms_deform_attn.py:94."""
    B, N, C, H, W = x.shape
    y = F.relu(F.conv2d(x.view(B * N, C, H, W), sd["downsample.0.weight"], sd["downsample.0.bias"], stride=stride,
                        padding=1))
    Cd, Hd, Wd = y.shape[1:]
    src = y.view(B, N, Cd, Hd, Wd).permute(0, 1, 3, 4, 2).reshape(B, N * Hd * Wd, Cd)
    if pos_embedding is None:
        pos_embedding = sine_pos_embedding((Hd, Wd), Cd // 2)
    pos = pos_embedding.to(x.device).flatten(2).transpose(1, 2).unsqueeze(1)  # [1,1,HW,C]
    pos = (pos + sd["lvl_embedding"].view(B, N, 1, Cd)).reshape(B, N * Hd * Wd, Cd)
    shapes = [(Hd, Wd)] * N
    shapes_dev = None
    if msda_fn is not None or sync_assert:
        sh = torch.as_tensor(shapes, dtype=torch.long, device=x.device)
        shapes_dev = (sh, torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1])), sync_assert)
    ref = ref_points.unsqueeze(0).repeat([B, 1, 1, 1, 1]).to(x.device)
    out = src
    for i in range(n_layers):
        out = encoder_layer_forward(sd, f"encoder.layers.{i}.", out, pos, ref, shapes, n_heads, n_points, msda_fn,
                                    shapes_dev)
    mem = out.view(B, N, Hd, Wd, Cd).permute(0, 1, 4, 2, 3).reshape(B, N * Cd, Hd, Wd)
    merged = F.relu(F.conv2d(mem, sd["merge_linear.0.weight"], sd["merge_linear.0.bias"]))
    up = F.interpolate(merged, size=(H, W), mode="bilinear", align_corners=False)
    return F.relu(F.conv2d(up, sd["upsample.1.weight"], sd["upsample.1.bias"], padding=1))
