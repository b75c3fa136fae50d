"""One launch of the fused view-grid MSDA kernel at the 4K stress shape (BASELINE configs[3]: 8 views, 120x360 tokens per
view, 8 heads x 32 channels, 8 points), default-initialised offsets -- for an ncu capture."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import bench  # noqa: E402
from mvdetr_b200 import ops  # noqa: E402


def main():
    device = torch.device("cuda:0")
    wl = bench.WORKLOADS["stress4k"]
    ds, fusion = bench.build_fusion(device, wl)
    wf = fusion.world_feat
    N, (Hg, Wg) = ds.num_cam, ds.Rworld_shape
    Hd, Wd = Hg // 2, Wg // 2
    Lq = S = N * Hd * Wd
    H, P, C = wl["heads"], wl["points"], wl["hidden"]
    D = C // H
    g = torch.Generator(device="cpu").manual_seed(1)
    with torch.no_grad():
        src = torch.randn(1, S, C, generator=g).to(device)
        am = wf.encoder.layers[0].self_attn
        value = ops.linear(src.view(S, C), am.value_proj.weight, am.value_proj.bias).view(1, S, H, D)
        raw_off, raw_log = am.offsets_and_logits(src.view(Lq, C), 1, Lq)
        geo = wf._level_geometry(N, Hd, Wd, device)
        for _ in range(2):
            out = ops.msda_fused_forward(value, geo.shapes, geo.start, raw_off, raw_log, wf.encoder.ref_table,
                                         grid_hw=(Hd, Wd), ref_table_lm=wf.encoder.ref_table_lm,
                                         off_bias=am.sampling_offsets.bias, logit_bias=am.attention_weights.bias)
        torch.cuda.synchronize()
    print("done", float(out.abs().mean()))


if __name__ == "__main__":
    main()
