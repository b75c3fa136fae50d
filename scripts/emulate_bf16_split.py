"""CPU study for the round-2 GEMM kernel: how many of the 9 partial products of a 3-way bf16 split are needed for
fp32-level accuracy? (bf16 x bf16 products are exact in fp32, so a torch fp32 matmul of the split terms emulates a
tensor-core MMA with fp32 accumulation.)   python scripts/emulate_bf16_split.py
Result on the encoder's GEMM shapes: 6 products (i + j <= 2) match all 9 to the last digit and beat the native fp32
GEMM; 3 products are ~25x worse."""
import torch

torch.manual_seed(0)


def split3(x):
    h = x.to(torch.bfloat16)
    r = x - h.float()
    m = r.to(torch.bfloat16)
    return h.float(), m.float(), (r - m.float()).to(torch.bfloat16).float()


for M, K, N in ((2048, 128, 448), (2048, 128, 512), (2048, 512, 128), (1024, 1152, 128)):
    x, w = torch.randn(M, K), torch.randn(N, K) / K ** 0.5
    exact = x.double() @ w.double().t()
    xs, ws = split3(x), split3(w)
    line = [f"M{M} K{K} N{N}: native fp32 {((x @ w.t()).double() - exact).abs().max().item():.2e}"]
    for name, keep in (("9", lambda i, j: True), ("6", lambda i, j: i + j <= 2), ("3", lambda i, j: i + j <= 1)):
        acc = torch.zeros(M, N)
        for i, j in sorted(((i, j) for i in range(3) for j in range(3) if keep(i, j)), key=lambda t: -(t[0] + t[1])):
            acc += xs[i] @ ws[j].t()
        line.append(f"{name} products {(acc.double() - exact).abs().max().item():.2e}")
    print(" | ".join(line))
