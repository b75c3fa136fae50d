"""Test-side tool (uses the oracle, so it lives under tests/). Error budget of the full-size frame: ours (GPU, fp32-level GEMMs) vs the CPU oracle in fp32 and in fp64, and the fp32
oracle vs the fp64 oracle (how much of the 1e-4 parity bar the fp32 REFERENCE arithmetic itself consumes)."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mvdetr_b200 import ops, synthetic  # noqa: E402
from mvdetr_b200.fusion import FrameRunner, MultiviewFusion  # noqa: E402
from oracle import cpu_oracle as co  # noqa: E402
from oracle import torch_port as tp  # noqa: E402


def main():
    cuda = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    hidden, heads, points = 128, 8, 4
    for scene in ("wildtrack_like", "multiviewx_like"):
        torch.manual_seed(1234)
        ds = getattr(synthetic, scene)(seed=0)
        fusion = MultiviewFusion(ds, base_dim=hidden, hidden_dim=hidden, nhead=heads, n_points=points)
        with torch.no_grad():
            for layer in fusion.world_feat.encoder.layers:
                layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
                layer.self_attn.attention_weights.weight.normal_(0, 0.05)
        fusion = fusion.to(cuda).eval()
        N, (Hg, Wg) = ds.num_cam, ds.Rworld_shape
        g = torch.Generator().manual_seed(3)
        feat = torch.randn(N, hidden, *ds.Rimg_shape, generator=g)
        proj = fusion.projection(torch.eye(3).view(1, 1, 3, 3).repeat(1, N, 1, 1))
        runner = FrameRunner(fusion, tuple(feat.shape), cuda, use_graph=True, depth=2)
        runner.load(feat.to(cuda), proj.to(cuda), slot=1)
        out = runner.step(1)
        torch.cuda.synchronize()
        out = out.cpu()
        world = torch.from_numpy(co.warp_forward(feat.numpy(), proj.numpy(), (Hg, Wg)))
        sd = {k: v.detach().cpu() for k, v in fusion.world_feat.state_dict().items()}
        rp = fusion.world_feat.encoder.reference_points.cpu()
        with torch.no_grad():
            want32 = tp.world_feat_forward(sd, world.view(1, N, hidden, Hg, Wg), rp, n_heads=heads, n_points=points)
            sd64 = {k: v.double() for k, v in sd.items()}
            want64 = tp.world_feat_forward(sd64, world.double().view(1, N, hidden, Hg, Wg), rp.double(), n_heads=heads,
                                           n_points=points)
        print(json.dumps({"scene": scene, "gemm": ops._GEMM_MODE, "max_abs_out": float(want64.abs().max()),
                          "ours_vs_oracle_fp32": float((out - want32).abs().max()),
                          "ours_vs_oracle_fp64": float((out.double() - want64).abs().max()),
                          "oracle_fp32_vs_fp64": float((want32.double() - want64).abs().max())}), flush=True)


if __name__ == "__main__":
    main()
