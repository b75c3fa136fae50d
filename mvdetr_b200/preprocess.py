"""Input side of the path: decoded camera frames -> network input on the GPU.

Host-side mirror of the reference dataset's per-view transform
    ref: multiview_detector/datasets/frameDataset.py:66-67
         T.Compose([T.ToTensor(), T.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225)),
                    T.Resize((H * 8 // img_reduce, W * 8 // img_reduce))])
over the C ABI (mvd_resize_normalize_u8): the reference runs it per view on the CPU in its data-loader workers; here the
uint8 frames are uploaded as they were decoded (3 bytes per pixel instead of 12) and one kernel produces the
[N, 3, Ho, Wo] fp32 tensor MVDeTr.forward consumes (mvdetr.py:151-153).
"""
import ctypes

import torch

from . import _C
from .ops import _on_device, _stream

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def resize_normalize(imgs_u8, size, mean=IMAGENET_MEAN, std=IMAGENET_STD, antialias=True, out=None):
    """imgs_u8 [N, H, W, 3] uint8 CUDA (HWC, RGB as PIL decodes) -> [N, 3, size[0], size[1]] fp32.
    antialias=True is what T.Resize does on tensors in torchvision >= 0.17 (the reference pins no version);
    antialias=False reproduces older torchvision (plain bilinear)."""
    if imgs_u8.dim() != 4 or imgs_u8.shape[-1] != 3:
        raise ValueError(f"imgs_u8 must be [N,H,W,3], got {tuple(imgs_u8.shape)}")
    if not (imgs_u8.is_cuda and imgs_u8.is_contiguous() and imgs_u8.dtype == torch.uint8):
        raise RuntimeError("resize_normalize: contiguous uint8 CUDA tensor required (no CPU fallback)")
    N, Hi, Wi, _ = imgs_u8.shape
    Ho, Wo = int(size[0]), int(size[1])
    if out is None:
        out = torch.empty((N, 3, Ho, Wo), dtype=torch.float32, device=imgs_u8.device)
    elif not (out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (N, 3, Ho, Wo)):
        raise RuntimeError("resize_normalize: out must be a contiguous fp32 CUDA tensor [N,3,Ho,Wo]")
    m = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s = (ctypes.c_float * 3)(*[float(v) for v in std])
    with _on_device(imgs_u8):
        rc = _C.lib.mvd_resize_normalize_u8(imgs_u8.data_ptr(), N, Hi, Wi, Ho, Wo, ctypes.cast(m, ctypes.c_void_p),
                                            ctypes.cast(s, ctypes.c_void_p), 1 if antialias else 0, out.data_ptr(),
                                            _stream(imgs_u8))
    _C.check(rc, "mvd_resize_normalize_u8")
    return out


def network_input_size(img_shape, img_reduce):
    """(H * 8 // img_reduce, W * 8 // img_reduce): the size the reference resizes every view to (frameDataset.py:67)."""
    return [int(img_shape[0]) * 8 // int(img_reduce), int(img_shape[1]) * 8 // int(img_reduce)]
