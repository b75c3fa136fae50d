"""Drop-in for the reference's compiled extension module `MultiScaleDeformableAttention`
(ref: multiview_detector/models/ops/src/vision.cpp:13-15; imported at
 multiview_detector/models/ops/functions/ms_deform_attn_func.py:18).

Put this directory on sys.path (mvdetr_b200.install_shims()) and the unmodified reference model runs on the
sm_100a kernels."""
from mvdetr_b200.ops import ms_deform_attn_backward, ms_deform_attn_forward  # noqa: F401
