"""mvdetr_b200 -- B200 (sm_100a) implementation of MVDeTr's per-frame multiview fusion hot path.

Only what the path needs: csrc/ (CUDA kernels + C ABI), the ctypes binding, and the host-side mirror of the
reference's operator interface (ops.py), its encoder callers (world_feat.py), the view-sharded multi-GPU path
(sharded.py) and the drop-in shim modules (shims/). Importing the package requires the built library; there is no
CPU fallback.
"""
from . import _C  # noqa: F401  (raises ImportError with build instructions if the .so is missing)
from .ops import (MSDeformAttnFunction, add_layer_norm, ms_deform_attn_backward, ms_deform_attn_forward, msda_fused_forward,
                  msda_viewgrid_forward, warp_perspective)

__all__ = ["MSDeformAttnFunction", "ms_deform_attn_forward", "ms_deform_attn_backward", "msda_fused_forward",
           "msda_viewgrid_forward", "warp_perspective", "add_layer_norm", "install_shims"]

__version__ = "0.1.0"


def install_shims():
    """Puts mvdetr_b200/shims first on sys.path so the UNMODIFIED reference imports our code:
    `import MultiScaleDeformableAttention as MSDA` (ms_deform_attn_func.py:18) and `import kornia` (mvdetr.py:7)."""
    import os
    import sys
    shims = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
    if shims not in sys.path:
        sys.path.insert(0, shims)
    return shims
