"""First-run check of the implicit-GEMM 3x3 convolution against F.conv2d and the im2col route."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mvdetr_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
ok = True
for NB, Hi, Wi, C, N, stride in [(1, 8, 32, 32, 16, 1), (2, 37, 64, 32, 20, 1), (3, 33, 48, 64, 132, 2), (7, 120, 360, 128, 128, 2),
                                 (1, 120, 360, 128, 128, 1)]:
    x = torch.randn(NB, C, Hi, Wi, generator=g).to(dev)
    w = (torch.randn(N, C, 3, 3, generator=g) / (9 * C) ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    x_cl = x.permute(0, 2, 3, 1).contiguous()
    w2d = w.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous()
    got = ops.conv3x3_nhwc(x_cl, w2d, b, stride=stride, relu=False)
    torch.cuda.synchronize()
    Ho, Wo = (Hi - 1) // stride + 1, (Wi - 1) // stride + 1
    exact = F.conv2d(x.double(), w.double(), b.double(), stride=stride, padding=1).permute(0, 2, 3, 1).reshape(NB * Ho * Wo, N)
    err = (got.double() - exact).abs().max().item()
    print(NB, Hi, Wi, C, N, stride, "err", err, flush=True)
    ok = ok and err < 1e-4
raise SystemExit(0 if ok else 1)
