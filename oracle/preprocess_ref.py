"""numpy restatement of the reference's per-view image transform -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

  resize_normalize   follows T.Compose([T.ToTensor(), T.Normalize(mean, std), T.Resize(size)])
                     ref: multiview_detector/datasets/frameDataset.py:66-67
T.Resize on a tensor is F.interpolate(mode='bilinear', align_corners=False, antialias=True) in the torchvision of this
image (0.26; antialias defaults to True since 0.17) -- ATen _upsample_bilinear2d_aa, restated below from its published
algorithm (aten/src/ATen/native/cpu/UpSampleKernel.cpp: _compute_indices_min_size_weights_aa, triangle filter).
Pinned by tests/golden/preprocess.npz (outputs of that torchvision Compose, tests/golden/make_golden_preprocess.py).
"""
import numpy as np

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def _aa_weights(out_size, in_size):
    f = np.float32
    scale = f(in_size) / f(out_size)
    support = scale if scale >= 1 else f(1)
    invscale = f(1) / scale if scale >= 1 else f(1)
    taps = []
    for i in range(out_size):
        center = scale * (f(i) + f(0.5))
        lo = max(int(center - support + f(0.5)), 0)
        n = min(int(center + support + f(0.5)), in_size) - lo
        x = np.abs((np.arange(n, dtype=np.float32) + f(lo) - center + f(0.5)) * invscale)
        w = np.where(x < 1, f(1) - x, f(0)).astype(np.float32)
        tot = w.sum(dtype=np.float32)
        taps.append((lo, (w / tot).astype(np.float32) if tot != 0 else w))
    return taps


def resize_normalize(img_u8, size, mean=MEAN, std=STD, antialias=True):
    """img_u8 [H,W,3] uint8 -> [3,Ho,Wo] float32."""
    x = img_u8.astype(np.float32) / np.float32(255)
    x = ((x - np.asarray(mean, dtype=np.float32)) / np.asarray(std, dtype=np.float32)).astype(np.float32)
    x = x.transpose(2, 0, 1)
    Hi, Wi = x.shape[1:]
    Ho, Wo = size
    if antialias:
        tx, ty = _aa_weights(Wo, Wi), _aa_weights(Ho, Hi)
        tmp = np.stack([(x[:, :, lo:lo + len(w)] * w).sum(-1, dtype=np.float32) for lo, w in tx], -1)      # [3,Hi,Wo]
        return np.stack([(tmp[:, lo:lo + len(w), :] * w[None, :, None]).sum(1, dtype=np.float32) for lo, w in ty], 1)
    f = np.float32
    sy, sx = f(Hi) / f(Ho), f(Wi) / f(Wo)
    fy = np.maximum(sy * (np.arange(Ho, dtype=np.float32) + f(0.5)) - f(0.5), 0).astype(np.float32)
    fx = np.maximum(sx * (np.arange(Wo, dtype=np.float32) + f(0.5)) - f(0.5), 0).astype(np.float32)
    y0 = np.minimum(fy.astype(np.int64), Hi - 1)
    x0 = np.minimum(fx.astype(np.int64), Wi - 1)
    y1, x1 = np.minimum(y0 + 1, Hi - 1), np.minimum(x0 + 1, Wi - 1)
    ly, lx = (fy - y0).astype(np.float32)[None, :, None], (fx - x0).astype(np.float32)[None, None, :]
    hy, hx = f(1) - ly, f(1) - lx
    g = lambda yy, xx: x[:, yy][:, :, xx]
    return (hy * (hx * g(y0, x0) + lx * g(y0, x1)) + ly * (hx * g(y1, x0) + lx * g(y1, x1))).astype(np.float32)
