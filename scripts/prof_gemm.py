"""One launch of the default GEMM on the frame's big-K shapes (for an ncu capture: which warp role waits on which barrier)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvdetr_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for rows, K, N, relu in [(75600, 1152, 128, True), (75600, 512, 128, False), (75600, 128, 128, False), (75600, 128, 512, True)]:
    x = torch.randn(rows, K, generator=g).to(dev)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    for _ in range(2):
        ops.linear(x, w, b, relu=relu)
    torch.cuda.synchronize()
print("done")
