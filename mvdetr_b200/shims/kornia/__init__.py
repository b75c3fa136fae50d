"""Minimal `kornia` stand-in exposing the one function MVDeTr's hot path calls
(ref: multiview_detector/models/mvdetr.py:7,194-195): warp_perspective, backed by the sm_100a warp kernel.
Only for environments where the real kornia is absent; it implements bilinear / zeros / align_corners=False."""
from mvdetr_b200.ops import warp_perspective  # noqa: F401
