"""BASELINE configs[3] and [4] on one B200: deformable-attention stress shapes and the forward/backward throughput sweep.
    python scripts/sweep.py [--out gpurun_out/sweep.jsonl] [--quick]
Each line: shape, our kernel's device time (CUDA events, median of `iters`, 512 MB L2 flush before every launch),
algorithmic GB/s (SURVEY 8d byte counts) and fraction of the measured HBM peak, the reference's own CUDA op on the same
tensors when oracle/_ref was built, and max |ours - reference| (parity at sizes the CPU oracle cannot reach)."""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import bench  # noqa: E402
from mvdetr_b200 import ops  # noqa: E402
from tests.gpu_util import ref_cuda_ext, start_index  # noqa: E402


def problem(shapes_hw, Lq, M, D, P, dev, seed=0, viewgrid=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    shapes = torch.as_tensor(shapes_hw, dtype=torch.long)
    S, L = int(shapes.prod(1).sum()), len(shapes_hw)
    value = torch.randn(1, S, M, D, device=dev)
    if viewgrid:  # identity reference grid + a few pixels of head/point dependent offset (MVDeTr-like locality)
        H, W = shapes_hw[0]
        R = Lq // (H * W)
        ys, xs = torch.meshgrid(torch.linspace(0.5, H - 0.5, H, device=dev), torch.linspace(0.5, W - 0.5, W, device=dev),
                                indexing="ij")
        ref = torch.stack((xs / W, ys / H), -1).reshape(1, H * W, 1, 1, 1, 2).repeat(1, R, 1, 1, 1, 1)
        loc = ref + torch.randn(1, Lq, M, L, P, 2, device=dev) * 2.5 / torch.tensor([W, H], device=dev)
    else:
        loc = torch.rand(1, Lq, M, L, P, 2, device=dev)
    attn = torch.softmax(torch.randn(1, Lq, M, L * P, device=dev), -1).view(1, Lq, M, L, P)
    go = torch.randn(1, Lq, M * D, device=dev)
    return value, shapes.to(dev), start_index(shapes).to(dev), loc.contiguous(), attn.contiguous(), go


def run_case(name, shapes_hw, Lq, M, D, P, dev, peak, iters, flush, viewgrid=False, do_bwd=True):
    value, shapes, start, loc, attn, go = problem(shapes_hw, Lq, M, D, P, dev, viewgrid=viewgrid)
    S, L = value.shape[1], len(shapes_hw)
    fb = bench.msda_algorithmic_bytes(1, S, M, D, L, Lq, P)
    bb = 4 * (3 * S * M * D + Lq * M * D + 2 * 3 * Lq * M * L * P)
    rec = {"case": name, "S": S, "Lq": Lq, "M": M, "D": D, "L": L, "P": P, "C": M * D, "layout": "viewgrid" if viewgrid else "generic"}
    t, tmin = bench.time_kernel_events(lambda: ops.ms_deform_attn_forward(value, shapes, start, loc, attn, 64), iters, flush)
    rec["fwd"] = {"us": t, "us_min": tmin, "bytes": fb, "GBps": fb / t / 1e3, "frac": fb / t / 1e3 / peak}
    if do_bwd:
        t, tmin = bench.time_kernel_events(lambda: ops.ms_deform_attn_backward(value, shapes, start, loc, attn, go, 64),
                                           max(3, iters // 2), flush)
        rec["bwd"] = {"us": t, "us_min": tmin, "bytes": bb, "GBps": bb / t / 1e3, "frac": bb / t / 1e3 / peak}
    ext = ref_cuda_ext()
    if ext is not None:
        t, _ = bench.time_kernel_events(lambda: ext.ms_deform_attn_forward(value, shapes, start, loc, attn, 64),
                                        max(3, iters // 2), flush)
        rec["ref_cuda_fwd_us"] = t
        ours = ops.ms_deform_attn_forward(value, shapes, start, loc, attn, 64)
        theirs = ext.ms_deform_attn_forward(value, shapes, start, loc, attn, 64)
        rec["fwd_max_abs_diff_vs_ref_cuda"] = (ours - theirs).abs().max().item()
        if do_bwd:
            t, _ = bench.time_kernel_events(lambda: ext.ms_deform_attn_backward(value, shapes, start, loc, attn, go, 64),
                                            max(3, iters // 2), flush)
            rec["ref_cuda_bwd_us"] = t
            mine = ops.ms_deform_attn_backward(value, shapes, start, loc, attn, go, 64)
            ref = ext.ms_deform_attn_backward(value, shapes, start, loc, attn, go, 64)
            rec["bwd_max_rel_diff_vs_ref_cuda"] = max(
                ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item() for a, b in zip(mine, ref))
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "sweep.jsonl"))
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peak, _ = bench.measured_peak_hbm()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    iters = 5 if args.quick else 12
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    recs = []

    def emit(rec):
        recs.append(rec)
        print(json.dumps(rec), flush=True)

    # ---- configs[4]: sweep N_query 1k..256k x C 64..512 (single-view-ish pyramid: L=4 square levels, P=4, M=8) ----
    for n in ([1 << 10, 1 << 14, 1 << 18] if args.quick else [1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18]):
        side = int(round((n * 64 / 85) ** 0.5))  # levels side, side/2, side/4, side/8: S ~ n
        shapes_hw = [(max(1, side >> k), max(1, side >> k)) for k in range(4)]
        S = sum(h * w for h, w in shapes_hw)
        for C in (64, 128, 256, 512):
            emit(run_case(f"sweep_n{n}_c{C}", shapes_hw, S, 8, C // 8, 4, dev, peak, iters, flush))
    # ---- configs[3]: 8-view 4K stress, C=256 (M=8, D=32), K=8 ----
    emit(run_case("stress_L4_120x360", [(120, 360)] * 4, 4 * 120 * 360, 8, 32, 8, dev, peak, iters, flush, viewgrid=True))
    emit(run_case("stress_L8_views_120x360", [(120, 360)] * 8, 8 * 120 * 360, 8, 32, 8, dev, peak, max(3, iters // 2), flush,
                  viewgrid=True))
    # its warp: [8,256,180,320] -> [8,256,240,720]
    feat = torch.randn(8, 256, 180, 320, device=dev)
    sx, sy = 720 / 320, 240 / 180
    mats = torch.tensor([[sx, 0.05, 3.0], [0.02, sy, -2.0], [1e-5, 2e-5, 1.0]], device=dev).repeat(8, 1, 1)
    wb = bench.warp_algorithmic_bytes(8, 256, 180, 320, 240, 720)
    for nm, src, cl in (("warp_4k_nchw_to_nhwc", feat, True), ("warp_4k_nchw_to_nchw", feat, False),
                        ("warp_4k_cl_to_nhwc", feat.contiguous(memory_format=torch.channels_last), True)):
        t, tmin = bench.time_kernel_events(lambda: ops.warp_perspective(src, mats, (240, 720), align_corners=False,
                                                                        channels_last=cl), iters, flush)
        emit({"case": nm, "us": t, "us_min": tmin, "bytes": wb, "GBps": wb / t / 1e3, "frac": wb / t / 1e3 / peak})
    # ---- MultiviewX layer shape (configs[2] per-GPU kernel): 6 views of 80x125 ----
    emit(run_case("multiviewx_layer", [(80, 125)] * 6, 6 * 80 * 125, 8, 16, 4, dev, peak, iters, flush, viewgrid=True))
    with open(args.out, "w") as f:
        for r in recs:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
