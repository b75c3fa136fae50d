"""Golden vectors for the decode + NMS output stage, produced by the REFERENCE's own functions
(multiview_detector/utils/decode.py:80-93 mvdet_decode, multiview_detector/utils/nms.py:7-44 nms, combined exactly as
multiview_detector/trainer.py:121-136 does) on the deterministic inputs of tests/decode_cases.py.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_decode.py
Also pins the reference's evaluator KAT: its bundled demo result file (evaluation/test-demo.txt) is turned into
heatmaps frame by frame, pushed through the reference decode + NMS, and the detections that come out are scored by
the reference's pyeval against evaluation/gt-demo.txt -> MODA 88.4454 (SURVEY 8f-3 / BASELINE.md 5)."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
sys.modules.setdefault("MultiScaleDeformableAttention", types.ModuleType("MultiScaleDeformableAttention"))

from multiview_detector.utils.decode import mvdet_decode  # noqa: E402
from multiview_detector.utils.nms import nms  # noqa: E402
from tests import decode_cases as dc  # noqa: E402


def reference_postprocess(heat, off, reduce, cls_thres, indexing, dist_thres, top_k):
    """trainer.py:121-136 for B = 1, on the host like the reference."""
    heat_t = torch.from_numpy(heat)
    off_t = torch.from_numpy(off) if off is not None else None
    xys = mvdet_decode(torch.sigmoid(heat_t), off_t, reduce=reduce)
    grid_xy, scores = xys[:, :, :2], xys[:, :, 2:3]
    positions = grid_xy if indexing == "xy" else grid_xy[:, :, [1, 0]]
    ids = scores[0].squeeze(-1) > cls_thres
    pos, s = positions[0, ids], scores[0, ids, 0]
    keep, count = nms(pos, s, dist_thres, np.inf if top_k == 0 else top_k)
    return pos.numpy(), s.numpy(), keep[:count].numpy().astype(np.int64)


def main():
    out = {}
    for name in dc.CASES:
        heat, off, kw = dc.case_inputs(name)
        pos, s, keep = reference_postprocess(heat, off, **kw)
        out[f"{name}.pos"], out[f"{name}.score"], out[f"{name}.keep"] = pos, s, keep
        print(f"{name}: {len(s)} candidates, {len(keep)} kept")
    # ---- evaluator KAT on the reference's bundled demo files ----
    ev = os.path.join(REF, "multiview_detector", "evaluation")
    demo = np.loadtxt(os.path.join(ev, "test-demo.txt")).astype(np.int64)
    frames = np.unique(demo[:, 0])
    res_rows = []
    for k, f in enumerate(frames):
        rows = demo[demo[:, 0] == f][:, 1:]
        heat, off = dc.demo_frame_maps(rows, seed=int(f))
        pos, s, keep = reference_postprocess(heat, off, 4, 0.6, "ij", 20.0, 0)
        kept = pos[keep]
        assert sorted(map(tuple, kept.astype(np.int64).tolist())) == sorted(map(tuple, rows.tolist())), f
        res_rows.append(np.concatenate([np.full((len(kept), 1), f, dtype=np.float32), kept], axis=1))
    res = np.concatenate(res_rows)
    res_path = "/tmp/mvd_demo_res.txt"
    np.savetxt(res_path, res, "%d")
    from multiview_detector.evaluation.pyeval.evaluateDetection import evaluateDetection_py
    recall, precision, moda, modp = evaluateDetection_py(res_path, os.path.join(ev, "gt-demo.txt"), "Wildtrack")
    print(f"demo KAT through decode+nms: MODA {moda:.4f} MODP {modp:.4f} precision {precision:.4f} recall {recall:.4f}")
    out["demo.rows"] = demo
    out["demo.res"] = res
    out["demo.moda_modp_prec_recall"] = np.array([moda, modp, precision, recall], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "decode_nms.npz"), **out)


if __name__ == "__main__":
    main()
