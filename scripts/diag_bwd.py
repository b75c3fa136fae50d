"""Diagnostic (GPU): who is right when our backward and the reference's CUDA backward disagree at large sizes?
Both fp32 results are compared, tensor by tensor, with an fp64 run of the same problem (our fp64 scalar kernel AND the
reference's fp64 kernel, which the reference's own gradcheck contract trusts)."""
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from mvdetr_b200 import ops  # noqa: E402
from scripts.sweep import problem  # noqa: E402
from tests.gpu_util import ref_cuda_ext  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def main():
    dev = torch.device("cuda:0")
    ext = ref_cuda_ext()
    cases = {"n65536_c512": ([(222, 222), (111, 111), (55, 55), (27, 27)], None, 8, 64, 4, False),
             "n65536_c256": ([(222, 222), (111, 111), (55, 55), (27, 27)], None, 8, 32, 4, False),
             "stress_L4": ([(120, 360)] * 4, 4 * 120 * 360, 8, 32, 8, True)}
    for name, (shapes_hw, Lq, M, D, P, vg) in cases.items():
        S = sum(h * w for h, w in shapes_hw)
        value, shapes, start, loc, attn, go = problem(shapes_hw, Lq or S, M, D, P, dev, viewgrid=vg)
        ours = ops.ms_deform_attn_backward(value, shapes, start, loc, attn, go, 64)
        ours2 = ops.ms_deform_attn_backward(value, shapes, start, loc, attn, go, 64)
        v64, l64, a64, g64 = value.double(), loc.double(), attn.double(), go.double()
        ours64 = ops.ms_deform_attn_backward(v64, shapes, start, l64, a64, g64, 64)
        rec = {"case": name}
        names = ("grad_value", "grad_loc", "grad_attn")
        rec["ours_vs_ours64"] = {n: rel(a, b) for n, a, b in zip(names, ours, ours64)}
        rec["ours_run_to_run"] = {n: rel(a, b) for n, a, b in zip(names, ours, ours2)}
        if ext is not None:
            ref = ext.ms_deform_attn_backward(value, shapes, start, loc, attn, go, 64)
            ref_again = ext.ms_deform_attn_backward(value, shapes, start, loc, attn, go, 64)
            ref64 = ext.ms_deform_attn_backward(v64, shapes, start, l64, a64, g64, 64)
            rec["ref_vs_ref64"] = {n: rel(a, b) for n, a, b in zip(names, ref, ref64)}
            rec["ref_run_to_run"] = {n: rel(a, b) for n, a, b in zip(names, ref, ref_again)}
            rec["ours64_vs_ref64"] = {n: rel(a, b) for n, a, b in zip(names, ours64, ref64)}
            rec["ours_vs_ref"] = {n: rel(a, b) for n, a, b in zip(names, ours, ref)}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
