"""Times the encoder layer's six Linear GEMMs at Wildtrack size in the three modes of ops.linear (GPU dev helper).
Prints one JSON line per (shape, mode): device time (CUDA events, median of 20, L2 flushed), TFLOP/s, max abs error
against an fp64 product on a 4096-row sample."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mvdetr_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
print(json.dumps({"cublasLt_version": ops.linear_available()}))
rows = 75600
for name, K, N, relu, bias in (("value_proj", 128, 128, False, True), ("sampling_offsets", 128, 448, False, False),
                               ("attention_weights", 128, 224, False, False), ("output_proj", 128, 128, False, False),
                               ("linear1", 128, 512, True, True), ("linear2", 512, 128, False, False),
                               ("downsample_as_gemm", 1152, 128, True, True), ("merge_as_gemm_10800rows", 896, 128, True, True)):
    r = 10800 if "merge" in name else rows
    x = torch.randn(r, K, device=dev)
    w = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.randn(N, device=dev) if bias else None
    exact = x[:4096].double() @ w.double().t() + (b.double() if bias else 0)
    if relu:
        exact = exact.clamp_min(0)
    for mode in ("bf16x9", "fp32", "torch"):
        t, tmin = bench.time_kernel_events(lambda: ops.linear(x, w, b, relu=relu, mode=mode), 20, flush)
        err = (ops.linear(x, w, b, relu=relu, mode=mode)[:4096].double() - exact).abs().max().item()
        print(json.dumps({"gemm": name, "rows": r, "K": K, "N": N, "mode": mode, "us": t, "us_min": tmin,
                          "TFLOPs": 2.0 * r * K * N / t / 1e6, "max_abs_err_vs_fp64": err}), flush=True)
