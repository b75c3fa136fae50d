"""Output side of the path: ground-plane heatmap -> detections on the GPU.

Host-side mirror of the test loop's post-processing
    ref: multiview_detector/trainer.py:121-136   (mvdet_decode -> threshold -> nms -> res rows)
    ref: multiview_detector/utils/decode.py:80-93 (mvdet_decode), multiview_detector/utils/nms.py:7-44 (nms)
over the C ABI (mvd_decode_candidates_f32, mvd_distance_nms_f32). The reference copies the heatmap and offset maps to
the host every frame and suppresses in a Python while-loop; here both steps are CUDA kernels and only the kept
detections are copied back.

  decode_candidates(heatmap, offset, reduce, cls_thres, indexing)   -> Candidates (device tensors, row-major order after nms)
  distance_nms(cands, dist_thres=20, top_k=None)                    -> (keep [B,cap] int32, keep_count [B] int32)
  detect(heatmap, offset, reduce, cls_thres, indexing, ...)         -> per batch element [count, 2] positions (host),
                                                                       the rows the reference writes to its result file
"""
import math

import torch

from . import _C
from .ops import _on_device, _stream


class Candidates:
    """Thresholded cells of a batch of heatmaps. After distance_nms() `cell`, `pos`, `score` hold, per batch element,
    the first count[b] candidates in row-major cell order: exactly the reference's `positions[b, ids]`, `scores[b, ids, 0]`."""

    def __init__(self, B, cap, device):
        self.B, self.cap = B, cap
        self.count = torch.empty(B, dtype=torch.int32, device=device)
        self.cell = torch.empty((B, cap), dtype=torch.int32, device=device)
        self.pos = torch.empty((B, cap, 2), dtype=torch.float32, device=device)
        self.score = torch.empty((B, cap), dtype=torch.float32, device=device)
        self.ordered = False


def decode_candidates(heatmap, offset=None, reduce=4, cls_thres=0.6, indexing="ij", cap=None):
    """heatmap [B,1,H,W] LOGITS (the sigmoid of trainer.py:121 happens in the kernel), offset [B,2,H,W] or None.
    indexing 'xy' keeps (x, y) positions, anything else swaps to (row, col) as trainer.py:127-130 does."""
    if heatmap.dim() != 4 or heatmap.shape[1] != 1:
        raise ValueError(f"heatmap must be [B,1,H,W], got {tuple(heatmap.shape)}")
    B, _, H, W = heatmap.shape
    if not heatmap.is_cuda:
        raise RuntimeError("mvdetr_b200.detect: heatmap must be a CUDA tensor (no CPU fallback)")
    tensors = [("heatmap", heatmap)] + ([("offset", offset)] if offset is not None else [])
    for name, t in tensors:
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError(f"decode_candidates: {name} must be a contiguous fp32 CUDA tensor")
    if offset is not None and tuple(offset.shape) != (B, 2, H, W):
        raise ValueError(f"offset must be [B,2,H,W], got {tuple(offset.shape)}")
    cap = H * W if cap is None else int(cap)
    c = Candidates(B, cap, heatmap.device)
    with _on_device(heatmap):
        rc = _C.lib.mvd_decode_candidates_f32(heatmap.data_ptr(), offset.data_ptr() if offset is not None else None,
                                              B, H, W, float(reduce), float(cls_thres), 0 if indexing == "xy" else 1,
                                              cap, c.count.data_ptr(), c.cell.data_ptr(), c.pos.data_ptr(),
                                              c.score.data_ptr(), _stream(heatmap))
    _C.check(rc, "mvd_decode_candidates_f32")
    return c


def distance_nms(cands, dist_thres=20.0, top_k=None):
    """Greedy distance NMS over decode_candidates' output. Reorders `cands` into row-major cell order (in new tensors)
    and returns (keep [B,cap] int32 candidate numbers in the order kept, keep_count [B] int32)."""
    B, cap, dev = cands.B, cands.cap, cands.count.device
    k = 0 if top_k is None or (isinstance(top_k, float) and math.isinf(top_k)) else int(top_k)
    ws_bytes = int(_C.lib.mvd_distance_nms_workspace_bytes(B, cap))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    o_cell, o_pos, o_score = torch.empty_like(cands.cell), torch.empty_like(cands.pos), torch.empty_like(cands.score)
    keep = torch.empty((B, cap), dtype=torch.int32, device=dev)
    keep_count = torch.empty(B, dtype=torch.int32, device=dev)
    with _on_device(cands.count):
        rc = _C.lib.mvd_distance_nms_f32(cands.count.data_ptr(), cands.cell.data_ptr(), cands.pos.data_ptr(),
                                         cands.score.data_ptr(), B, cap, float(dist_thres), k, ws.data_ptr(), ws_bytes,
                                         o_cell.data_ptr(), o_pos.data_ptr(), o_score.data_ptr(), keep.data_ptr(),
                                         keep_count.data_ptr(), _stream(cands.count))
    _C.check(rc, "mvd_distance_nms_f32")
    cands.cell, cands.pos, cands.score, cands.ordered = o_cell, o_pos, o_score, True
    return keep, keep_count


def detect(heatmap, offset=None, reduce=4, cls_thres=0.6, indexing="ij", dist_thres=20.0, top_k=None,
           return_scores=False):
    """The reference's per-frame post-processing (trainer.py:121-136) on the GPU. Returns one [count, 2] fp32 host
    tensor of kept positions per batch element (with return_scores: tuples (positions, scores))."""
    c = decode_candidates(heatmap, offset, reduce, cls_thres, indexing)
    keep, keep_count = distance_nms(c, dist_thres, top_k)
    counts = keep_count.cpu().tolist()            # the one synchronising read: B ints
    out = []
    for b, n in enumerate(counts):
        idx = keep[b, :n].long()
        pos = c.pos[b].index_select(0, idx).cpu()
        out.append((pos, c.score[b].index_select(0, idx).cpu()) if return_scores else pos)
    return out
