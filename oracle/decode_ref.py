"""numpy restatement of the reference's detection post-processing -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

  decode_threshold   follows mvdet_decode           multiview_detector/utils/decode.py:80-93
                     + the per-frame threshold       multiview_detector/trainer.py:121-133
  distance_nms       follows nms                     multiview_detector/utils/nms.py:7-44
Pinned by tests/golden/decode_nms.npz (outputs of the reference's own functions, tests/golden/make_golden_decode.py).
"""
import numpy as np


def sigmoid32(x):
    x = np.asarray(x, dtype=np.float32)
    return (np.float32(1) / (np.float32(1) + np.exp(-x, dtype=np.float32))).astype(np.float32)


def decode_threshold(heat, offset, reduce, cls_thres, indexing):
    """heat [H,W] logits, offset [2,H,W] or None -> (pos [n,2], score [n], cell [n]) in row-major cell order."""
    H, W = heat.shape
    score = sigmoid32(heat).reshape(-1)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    if offset is not None:
        x = (xs + offset[0]).astype(np.float32)
        y = (ys + offset[1]).astype(np.float32)
    else:
        x, y = xs + np.float32(0.5), ys + np.float32(0.5)
    x, y = (x * np.float32(reduce)).astype(np.float32), (y * np.float32(reduce)).astype(np.float32)
    xy = np.stack([x.reshape(-1), y.reshape(-1)], 1)
    pos = xy if indexing == "xy" else xy[:, ::-1]
    ids = score > np.float32(cls_thres)
    return np.ascontiguousarray(pos[ids]), score[ids], np.nonzero(ids)[0]


def distance_nms(pos, score, dist_thres=20.0, top_k=0):
    """-> kept candidate numbers in the order kept. Visits by descending score; among equal scores the larger
    candidate number first (a stable ascending sort read from the back, nms.py:24-32)."""
    n = len(score)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    order = np.argsort(score, kind="stable")
    if top_k and top_k < n:
        order = order[-top_k:]
    order = list(order)
    keep = []
    while order:
        i = order.pop()
        keep.append(i)
        if not order:
            break
        rest = np.asarray(order)
        d = pos[i].astype(np.float32) - pos[rest].astype(np.float32)
        dist = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32))
        order = list(rest[dist > np.float32(dist_thres)])
    return np.asarray(keep, dtype=np.int64)
