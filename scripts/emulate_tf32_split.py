"""numpy emulation of the 3xTF32 split used by csrc/gemm_tf32.cu, against fp64 -- how many partial products are needed?
a = hi + lo, hi = RN_tf32(a) (ties away, as cvt.rna.tf32.f32), lo = a - hi (exact); the tensor core TRUNCATES each
operand to tf32 (reads the top 19 bits) and accumulates exact products in fp32.
   python scripts/emulate_tf32_split.py"""
import numpy as np


def rna_tf32(x):
    b = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    b = (b + 0x1000) & 0xFFFFE000          # add half an ulp of the 13 dropped bits, truncate (ties away from zero)
    return b.astype(np.uint32).view(np.float32)


def trunc_tf32(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def split(x):
    hi = rna_tf32(x)
    lo = (x.astype(np.float32) - hi).astype(np.float32)
    return hi, trunc_tf32(lo)


def main():
    rng = np.random.RandomState(0)
    for rows, K, N in ((512, 128, 128), (256, 512, 128), (128, 1152, 128)):
        x = rng.randn(rows, K).astype(np.float32)
        w = (rng.randn(N, K) / np.sqrt(K)).astype(np.float32)
        ref = x.astype(np.float64) @ w.astype(np.float64).T
        xh, xl = split(x)
        wh, wl = split(w)
        f = lambda a, b: a.astype(np.float64) @ b.astype(np.float64).T   # products exact, fp32 accumulation ~ 1e-7 extra
        p3 = f(xh, wh) + f(xh, wl) + f(xl, wh)
        p4 = p3 + f(xl, wl)
        p1 = f(trunc_tf32(x), trunc_tf32(w))
        nat = (x @ w.T).astype(np.float64)
        scale = np.abs(ref).max()
        print(f"rows={rows} K={K} N={N}: max|err|/max|ref|  tf32x1 {np.abs(p1 - ref).max() / scale:.2e}  "
              f"tf32x3 {np.abs(p3 - ref).max() / scale:.2e}  tf32x4 {np.abs(p4 - ref).max() / scale:.2e}  "
              f"native fp32 (numpy) {np.abs(nat - ref).max() / scale:.2e}")


if __name__ == "__main__":
    main()
