"""CPU: the numpy restatement of the reference's decode + NMS (oracle/decode_ref.py) against the golden vectors made
by the reference's own functions (tests/golden/decode_nms.npz), and the evaluator KAT recorded with them."""
import numpy as np
import pytest

from oracle import decode_ref
from tests import decode_cases as dc


@pytest.mark.parametrize("name", sorted(dc.CASES))
def test_oracle_matches_reference_golden(golden, name):
    g = golden("decode_nms.npz")
    heat, off, kw = dc.case_inputs(name)
    pos, score, _ = decode_ref.decode_threshold(heat[0, 0], None if off is None else off[0], kw["reduce"],
                                                kw["cls_thres"], kw["indexing"])
    assert pos.shape == g[f"{name}.pos"].shape
    assert np.array_equal(pos, g[f"{name}.pos"])                       # positions: exactly rounded fp32 arithmetic
    assert np.abs(score - g[f"{name}.score"]).max(initial=0) <= 2e-7   # sigmoid: libm vs ATen's vectorised exp
    keep = decode_ref.distance_nms(g[f"{name}.pos"], g[f"{name}.score"], kw["dist_thres"], kw["top_k"])
    assert np.array_equal(keep, g[f"{name}.keep"])


def test_demo_file_roundtrip_and_evaluator_kat(golden):
    """The reference's bundled demo detections (evaluation/test-demo.txt), rebuilt as heatmaps and pushed through
    decode + NMS, come back unchanged; the reference's own evaluator scored exactly that list at MODA 88.4454
     when the golden was made."""
    g = golden("decode_nms.npz")
    assert abs(g["demo.moda_modp_prec_recall"][0] - 88.4454) < 5e-4
    demo = g["demo.rows"]
    out = []
    for f in np.unique(demo[:, 0]):
        rows = demo[demo[:, 0] == f][:, 1:]
        heat, off = dc.demo_frame_maps(rows, seed=int(f))
        pos, score, _ = decode_ref.decode_threshold(heat[0, 0], off[0], 4, 0.6, "ij")
        keep = decode_ref.distance_nms(pos, score, 20.0, 0)
        out.append(np.concatenate([np.full((len(keep), 1), f, dtype=np.float32), pos[keep]], axis=1))
    assert np.array_equal(np.concatenate(out), g["demo.res"])
