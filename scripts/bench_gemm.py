"""Per-shape timing of the frame's Linear / im2col-conv GEMMs: our tcgen05 3xTF32 kernel vs cuBLASLt BF16x9 (incl. its
operand-scan kernels, which CUDA events around the call see) vs torch.mm, with max error against fp64.
   python scripts/bench_gemm.py > gpurun_out/gemm.jsonl"""
import json
import os
import statistics
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from mvdetr_b200 import ops  # noqa: E402

SHAPES = [("conv_down", 75600, 1152, 128, True), ("value/out_proj", 75600, 128, 128, False),
          ("offsets", 75600, 128, 448, False), ("logits", 75600, 128, 224, False), ("linear1", 75600, 128, 512, True),
          ("linear2", 75600, 512, 128, False), ("merge", 10800, 896, 128, True), ("conv_up", 43200, 1152, 128, True),
          ("per_rank_8gpu", 10800, 128, 128, False)]


def timed(fn, flush, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts)


def main():
    modes = tuple(os.environ.get("BENCH_GEMM_MODES", "f16x2,bf16x3ts,bf16x3ss,tf32x3,bf16x9,torch").split(","))
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator().manual_seed(0)
    for name, rows, K, N, relu in SHAPES:
        x = torch.randn(rows, K, generator=g).to(dev)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
        b = torch.randn(N, generator=g).to(dev)
        exact = x[:4096].double() @ w.double().t() + b.double()
        if relu:
            exact = exact.clamp_min(0)
        rec = {"rings": os.environ.get("MVD_GEMM_RINGS", "default"), "name": name, "rows": rows, "K": K, "N": N, "relu": relu,
               "hbm_floor_us": 4 * (rows * K + rows * N + N * K) / 6549.8e3}
        for mode in modes:
            try:
                out = ops.linear(x, w, b, relu=relu, mode=mode)
                rec[mode + "_err"] = (out[:4096].double() - exact).abs().max().item()
                rec[mode + "_us"] = timed(lambda: ops.linear(x, w, b, relu=relu, mode=mode), flush)
            except Exception as e:
                rec[mode + "_error"] = repr(e)[:200]
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
