"""mvdetr_b200 -- B200 (sm_100a) implementation of MVDeTr's per-frame multiview fusion hot path.

Only what the path needs: csrc/ (CUDA kernels + C ABI), the ctypes binding, and the host-side mirror of the
reference's operator interface (ops.py), its encoder callers (world_feat.py), the view-sharded multi-GPU path
(sharded.py) and the drop-in shim modules (shims/). Importing the package requires the built library; there is no
CPU fallback.
"""
import os

from . import _C  # noqa: F401  (raises ImportError with build instructions if the .so is missing)
from .ops import (MSDeformAttnFunction, add_layer_norm, ms_deform_attn_backward, ms_deform_attn_forward, msda_fused_forward,
                  msda_viewgrid_forward, warp_perspective)

__all__ = ["MSDeformAttnFunction", "ms_deform_attn_forward", "ms_deform_attn_backward", "msda_fused_forward",
           "msda_viewgrid_forward", "warp_perspective", "add_layer_norm", "install_shims"]

__version__ = "0.1.0"


def _real_kornia_available(stub_dir):
    import importlib.machinery
    import sys
    paths = [p for p in sys.path if os.path.abspath(p or ".") != stub_dir]
    return importlib.machinery.PathFinder.find_spec("kornia", paths) is not None


def _hot_path_warp_call(src, mode, padding_mode, align_corners):
    import torch
    return (torch.is_tensor(src) and src.is_cuda and src.dtype == torch.float32 and src.dim() == 4 and
            mode == "bilinear" and padding_mode == "zeros" and align_corners is False)


def install_shims(kornia="auto"):
    """Makes the UNMODIFIED reference import our kernels.

    * `import MultiScaleDeformableAttention as MSDA` (ms_deform_attn_func.py:18): mvdetr_b200/shims goes first on
      sys.path (that name has no other provider than the reference's own build).
    * `import kornia` (mvdetr.py:7): a real kornia is never shadowed. kornia="auto": when a real kornia is importable,
      its `warp_perspective` is wrapped -- fp32 CUDA calls with mode='bilinear', padding_mode='zeros',
      align_corners=False (the hot-path call, mvdetr.py:194-195) run on our kernel, every other call (CPU masks in
      'nearest' mode at frameDataset.py:80, the visualisation scripts) goes to kornia itself; only when kornia is NOT
      installed the minimal stub package (shims/_kornia_stub) is added. kornia=False leaves kornia alone entirely;
      kornia="stub" forces the stub (tests)."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    shims = os.path.join(here, "shims")
    stub = os.path.join(shims, "_kornia_stub")
    if shims not in sys.path:
        sys.path.insert(0, shims)
    if kornia is False or kornia is None:
        return shims
    if kornia == "stub" or not _real_kornia_available(stub):
        if stub not in sys.path:
            sys.path.insert(0, stub)
        return shims
    import kornia as real
    if not getattr(real.warp_perspective, "__mvdetr_b200_wrapped__", False):
        original = real.warp_perspective

        def warp_perspective_dispatch(src, M, dsize, mode="bilinear", padding_mode="zeros", align_corners=None, *a, **k):
            if not a and not k and _hot_path_warp_call(src, mode, padding_mode, align_corners):
                return warp_perspective(src, M, dsize, mode=mode, padding_mode=padding_mode, align_corners=False)
            return original(src, M, dsize, mode, padding_mode, align_corners, *a, **k)

        warp_perspective_dispatch.__mvdetr_b200_wrapped__ = True
        warp_perspective_dispatch.__wrapped__ = original
        real.warp_perspective = warp_perspective_dispatch
    return shims
