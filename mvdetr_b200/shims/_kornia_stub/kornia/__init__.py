"""Minimal `kornia` stand-in for environments where the real kornia is NOT installed (mvdetr_b200.install_shims() puts
this directory on sys.path only then; a real kornia is never shadowed -- its warp_perspective is wrapped instead).

It exposes the one function MVDeTr's hot path calls (ref: multiview_detector/models/mvdetr.py:7,194-195):
warp_perspective for fp32 CUDA tensors, mode='bilinear', padding_mode='zeros', align_corners=False, backed by the
sm_100a warp kernels. The reference's other kornia calls (nearest-mode warps of CPU masks at dataset construction,
ref: multiview_detector/datasets/frameDataset.py:80, and the visualisation scripts) are outside the hot path and need
the real package: they raise NotImplementedError here, naming this fact."""
from mvdetr_b200.ops import warp_perspective  # noqa: F401

__mvdetr_b200_stub__ = True
