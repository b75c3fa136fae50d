"""GPU: decode + distance NMS kernels (through the C ABI, mvdetr_b200.detect) against the outputs of the REFERENCE's own
mvdet_decode / nms (tests/golden/decode_nms.npz) -- candidate order, kept indices and positions bit-exact, scores to
2e-7 -- and against the numpy oracle on further seeds."""
import numpy as np
import pytest
import torch

from mvdetr_b200 import detect
from oracle import decode_ref
from tests import decode_cases as dc
from tests.gpu_util import dev

pytestmark = pytest.mark.gpu


def _run(heat, off, kw, device):
    c = detect.decode_candidates(dev(heat, device), None if off is None else dev(off, device), kw["reduce"],
                                 kw["cls_thres"], kw["indexing"])
    keep, keep_count = detect.distance_nms(c, kw["dist_thres"], kw["top_k"])
    n, k = int(c.count[0]), int(keep_count[0])
    return (c.pos[0, :n].cpu().numpy(), c.score[0, :n].cpu().numpy(), c.cell[0, :n].cpu().numpy(),
            keep[0, :k].cpu().numpy().astype(np.int64))


@pytest.mark.parametrize("name", sorted(dc.CASES))
def test_matches_reference_golden(golden, cuda, name):
    g = golden("decode_nms.npz")
    heat, off, kw = dc.case_inputs(name)
    pos, score, cell, keep = _run(heat, off, kw, cuda)
    assert pos.shape == g[f"{name}.pos"].shape, "different candidate set"
    assert np.array_equal(pos, g[f"{name}.pos"])
    assert np.all(np.diff(cell) > 0)                                    # row-major order, as the boolean mask gives
    assert np.abs(score - g[f"{name}.score"]).max(initial=0) <= 2e-7
    assert np.array_equal(keep, g[f"{name}.keep"])


def test_demo_file_roundtrip(golden, cuda):
    """All 40 frames of the reference's bundled demo result file through detect(): the same rows the reference's
    pipeline produced (and its evaluator scored at MODA 88.4454)."""
    g = golden("decode_nms.npz")
    demo = g["demo.rows"]
    frames = np.unique(demo[:, 0])
    heats, offs = zip(*(dc.demo_frame_maps(demo[demo[:, 0] == f][:, 1:], seed=int(f)) for f in frames))
    heat, off = np.concatenate(heats), np.concatenate(offs)            # one batched call: B = 40
    res = detect.detect(dev(heat, cuda), dev(off, cuda), reduce=4, cls_thres=0.6, indexing="ij", dist_thres=20.0)
    rows = np.concatenate([np.concatenate([np.full((len(p), 1), f, dtype=np.float32), p.numpy()], axis=1)
                           for f, p in zip(frames, res)])
    assert np.array_equal(rows, g["demo.res"])


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_random_maps_vs_oracle(cuda, seed):
    rng = np.random.RandomState(seed)
    B, H, W = 3, 37, 53
    heat = rng.uniform(-4, 3, size=(B, 1, H, W)).astype(np.float32)
    off = rng.uniform(-0.5, 1.5, size=(B, 2, H, W)).astype(np.float32)
    c = detect.decode_candidates(dev(heat, cuda), dev(off, cuda), 4, 0.55, "xy")
    keep, kc = detect.distance_nms(c, 9.0, None)
    for b in range(B):
        pos, score, cell = decode_ref.decode_threshold(heat[b, 0], off[b], 4, 0.55, "xy")
        n = int(c.count[b])
        got_cell = c.cell[b, :n].cpu().numpy()
        common = np.intersect1d(got_cell, cell)
        assert len(common) >= max(len(cell), n) - 2          # a score within 1 ulp of the threshold may differ
        if n == len(cell) and np.array_equal(got_cell, cell):
            assert np.array_equal(c.pos[b, :n].cpu().numpy(), pos)
            want = decode_ref.distance_nms(pos, c.score[b, :n].cpu().numpy(), 9.0, 0)
            assert np.array_equal(keep[b, :int(kc[b])].cpu().numpy(), want)


def test_argument_errors(cuda):
    with pytest.raises(RuntimeError, match="CUDA"):
        detect.decode_candidates(torch.zeros(1, 1, 4, 4))
    with pytest.raises(ValueError):
        detect.decode_candidates(torch.zeros(1, 2, 4, 4, device=cuda))
