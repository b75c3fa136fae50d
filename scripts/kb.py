"""Dev helper (GPU): per-kernel device times of the hot-path kernels on the bench workload (bench.kernel_breakdown)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
ds, fusion = bench.build_fusion(dev)
kb = bench.kernel_breakdown(fusion, ds, dev, iters=int(sys.argv[1]) if len(sys.argv) > 1 else 20)
for k, v in kb.items():
    print(k, json.dumps(v))
