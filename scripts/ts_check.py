"""First-run check of the TS-form GEMM (x terms in tensor memory) against fp64 and against the shared-memory kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvdetr_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
ok = True
for rows, K, N in [(300, 128, 448), (129, 64, 128), (1000, 512, 128), (257, 288, 256), (5, 8, 4), (75600, 128, 512)]:
    x = torch.randn(rows, K, generator=g).to(dev); w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev); b = torch.randn(N, generator=g).to(dev)
    a = ops.linear(x, w, b, mode="bf16x3ts"); s = ops.linear(x, w, b, mode="bf16x3ss")
    torch.cuda.synchronize()
    e = (a.double() - (x.double() @ w.double().t() + b.double())).abs().max().item()
    h = ops.linear(x, w, b, mode="f16x2")
    eh = (h.double() - (x.double() @ w.double().t() + b.double())).abs().max().item()
    et = (torch.addmm(b, x, w.t()).double() - (x.double() @ w.double().t() + b.double())).abs().max().item()
    print(rows, K, N, "f16x2 err", eh, "torch fp32 err", et, flush=True)
    ok = ok and eh < 1e-4
    print(rows, K, N, "err", e, "equal_to_ss", torch.equal(a, s), "max diff to ss", (a - s).abs().max().item(), flush=True)
    ok = ok and e < 1e-4
raise SystemExit(0 if ok else 1)
