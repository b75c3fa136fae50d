"""Dev experiment (GPU): where the dense glue of the fusion stage spends its time and which torch-level settings
help. Not part of the product or the tests; prints a table to stdout."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mvdetr_b200 import ops

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(name, fn, iters=10):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print(f"{name:70s} avg {sum(ts)/len(ts):9.1f} us   min {min(ts):9.1f} us", flush=True)


R = 75600
x128 = torch.randn(R, 128, device=dev)
x512 = torch.randn(R, 512, device=dev)
with torch.no_grad():
    for nout in (128, 224, 448, 512):
        lin = torch.nn.Linear(128, nout).to(dev)
        timeit(f"F.linear 128->{nout}", lambda: lin(x128))
        wt = lin.weight.t().contiguous()
        timeit(f"addmm (W^T contiguous) 128->{nout}", lambda: torch.addmm(lin.bias, x128, wt))
        timeit(f"mm no bias 128->{nout}", lambda: torch.mm(x128, wt))
        timeit(f"_addmm_activation relu 128->{nout}", lambda: torch._addmm_activation(lin.bias, x128, lin.weight.t()))
    lin = torch.nn.Linear(512, 128).to(dev)
    timeit("F.linear 512->128", lambda: lin(x512))
    for lib in ("cublas", "cublaslt"):
        try:
            torch.backends.cuda.preferred_blas_library(lib)
            lin2 = torch.nn.Linear(128, 448).to(dev)
            timeit(f"[{lib}] F.linear 128->448", lambda: lin2(x128))
        except Exception as e:
            print(lib, "failed", e)
    torch.backends.cuda.preferred_blas_library("default")
    # concatenated projections: offsets(448)+logits(224) in one GEMM
    lin3 = torch.nn.Linear(128, 672).to(dev)
    timeit("F.linear 128->672 (offsets+logits fused)", lambda: lin3(x128))

    ln = torch.nn.LayerNorm(128).to(dev)
    y = torch.randn_like(x128)
    timeit("torch add + LayerNorm", lambda: ln(x128 + y))
    timeit("ops.add_layer_norm", lambda: ops.add_layer_norm(x128, y, ln.weight, ln.bias, ln.eps))
    print("LN max diff", (ln(x128 + y) - ops.add_layer_norm(x128, y, ln.weight, ln.bias, ln.eps)).abs().max().item())

    # convs
    xw = torch.randn(7, 128, 120, 360, device=dev)
    xw_cl = xw.contiguous(memory_format=torch.channels_last)
    conv = torch.nn.Conv2d(128, 128, 3, 2, 1).to(dev)
    for bench in (False, True):
        torch.backends.cudnn.benchmark = bench
        timeit(f"downsample conv NCHW  benchmark={bench}", lambda: conv(xw))
        timeit(f"downsample conv NHWC  benchmark={bench}", lambda: conv(xw_cl))
    xu = torch.randn(1, 128, 120, 360, device=dev)
    conv2 = torch.nn.Conv2d(128, 128, 3, 1, 1).to(dev)
    for bench in (False, True):
        torch.backends.cudnn.benchmark = bench
        timeit(f"final 3x3 conv NCHW benchmark={bench}", lambda: conv2(xu))
        timeit(f"final 3x3 conv NHWC benchmark={bench}", lambda: conv2(xu.contiguous(memory_format=torch.channels_last)))
    xm = torch.randn(1, 896, 60, 180, device=dev)
    conv3 = torch.nn.Conv2d(896, 128, 1).to(dev)
    timeit("merge 1x1 conv NCHW", lambda: conv3(xm))
    mem = torch.randn(7, 10800, 128, device=dev)
    w3 = conv3.weight.view(128, 7, 128)
    def merge_mm():
        out = torch.addmm(conv3.bias, mem[0], w3[:, 0].t())
        for n in range(1, 7):
            out.addmm_(mem[n], w3[:, n].t())
        return out
    timeit("merge as 7 accumulated GEMMs on [N,HW,C] tokens (no permute)", merge_mm)
    timeit("merge as one GEMM after permute-copy", lambda: torch.addmm(conv3.bias, mem.permute(1, 0, 2).reshape(10800, 896), conv3.weight.view(128, 896).t()))
    up = torch.nn.Upsample((120, 360), mode="bilinear", align_corners=False)
    xs = torch.randn(1, 128, 60, 180, device=dev)
    timeit("upsample bilinear", lambda: up(xs))
    # TF32 for context only
    torch.backends.cuda.matmul.allow_tf32 = True
    lin = torch.nn.Linear(128, 448).to(dev)
    timeit("[tf32 context only] F.linear 128->448", lambda: lin(x128))
