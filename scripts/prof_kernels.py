"""Launches each hot-path kernel once on Wildtrack-shaped tensors produced by the model's own first layer
(same inputs as bench.py's kernel_breakdown). Meant to run under `ncu --set full -k regex:'msda_|warp_'`;
prints nothing that is a benchmark number."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import bench  # noqa: E402
from mvdetr_b200 import ops  # noqa: E402


def main():
    device = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    wl = bench.WORKLOADS["wildtrack"]
    ds, fusion = bench.build_fusion(device, wl)
    wf = fusion.world_feat
    N, (Hg, Wg) = ds.num_cam, ds.Rworld_shape
    Hd, Wd = Hg // 2, Wg // 2
    Lq = S = N * Hd * Wd
    H, P, C = wl["heads"], wl["points"], wl["hidden"]
    D = C // H
    g = torch.Generator(device="cpu").manual_seed(1)
    feat = torch.randn(N, C, *ds.Rimg_shape, generator=g).to(device)
    proj = fusion.projection(torch.eye(3).view(1, 1, 3, 3).repeat(1, N, 1, 1)).to(device)
    with torch.no_grad():
        world_cl = ops.warp_perspective(feat, proj, (Hg, Wg), align_corners=False, channels_last=True)  # nhwc kernel
        world = ops.warp_perspective(feat, proj, (Hg, Wg), align_corners=False)  # nchw kernel
        x = wf.downsample(world)
        src = x.view(1, N, C, Hd, Wd).permute(0, 1, 3, 4, 2).reshape(1, S, C)
        pos = (wf.pos_embedding.flatten(2).transpose(1, 2).unsqueeze(1) + wf.lvl_embedding.view(1, N, 1, C)
               ).view(1, S, C)
        am = wf.encoder.layers[0].self_attn
        value = am.value_proj(src).view(1, S, H, D).contiguous()
        offsets = am.sampling_offsets(src + pos).view(1, Lq, H, N, P, 2).contiguous()
        logits = am.attention_weights(src + pos).view(1, Lq, H, N * P).contiguous()
        geo = wf._level_geometry(N, Hd, Wd, device)
        table = wf.encoder.ref_table
        out, attn, loc = ops.msda_fused_forward(value, geo.shapes, geo.start, offsets, logits, table, want_aux=True)
        ops.msda_fused_forward(value, geo.shapes, geo.start, offsets, logits, table)
        ops.msda_fused_forward(value, geo.shapes, geo.start, offsets, logits, table, grid_hw=(Hd, Wd), ref_table_lm=wf.encoder.ref_table_lm)
        # the frame's own launch: bias-free GEMM outputs + in-kernel biases
        q2 = (src + pos).view(Lq, C)
        raw_off = ops.linear(q2, am.sampling_offsets.weight).view(1, Lq, H, N, P, 2)
        raw_log = ops.linear(q2, am.attention_weights.weight).view(1, Lq, H, N * P)
        ops.msda_fused_forward(value, geo.shapes, geo.start, raw_off, raw_log, table, grid_hw=(Hd, Wd),
                               ref_table_lm=wf.encoder.ref_table_lm, off_bias=am.sampling_offsets.bias,
                               logit_bias=am.attention_weights.bias)
        A, _ = ops.warp_im2col(feat, proj, (Hg, Wg), stride=2)                      # transpose + warp_im2col_kernel
        ops.upsample_im2col(torch.randn(1, Hd, Wd, C, device=device), (Hg, Wg))     # upsample_im2col_kernel
        ops.add_layer_norm(src.view(S, C).contiguous(), torch.randn(S, C, device=device), wf.encoder.layers[0].norm1.weight,
                           wf.encoder.layers[0].norm1.bias, 1e-5, res_bias=am.output_proj.bias)
        ops.ms_deform_attn_forward(value, geo.shapes, geo.start, loc, attn, 64)
        ops.ms_deform_attn_backward(value, geo.shapes, geo.start, loc, attn, torch.randn_like(out), 64)
        gw = torch.randn_like(world)
        torch.cuda.synchronize()
    src_g = feat.clone().requires_grad_(True)
    ops.warp_perspective(src_g, proj, (Hg, Wg), align_corners=False).backward(gw)
    torch.cuda.synchronize()
    print("prof_kernels done", float(out.abs().mean()), float(world_cl.abs().mean()))


if __name__ == "__main__":
    main()
