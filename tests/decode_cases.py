"""Deterministic heatmap / offset inputs for the decode + NMS tests, shared by the golden generator
(tests/golden/make_golden_decode.py, which runs the REFERENCE's mvdet_decode / nms on them) and the tests (which
rebuild the same inputs instead of storing megabytes of maps). Only exactly rounded fp32 operations are used
(add, mul, min/max), so every machine builds bit-identical arrays."""
import numpy as np


def blob_maps(H, W, peaks, seed, bg=-6.0, noise=0.5, with_offset=True):
    """Logit map [1,1,H,W]: paraboloid blobs (peak height ph, width ~3 cells) on a noisy background; offsets [1,2,H,W]
    uniform in [0,1). peaks: list of (y, x, height)."""
    rng = np.random.RandomState(seed)
    heat = (bg + noise * rng.uniform(-1, 1, size=(H, W))).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    for (py, px, ph) in peaks:
        d2 = (ys - np.float32(py)) * (ys - np.float32(py)) + (xs - np.float32(px)) * (xs - np.float32(px))
        # per-cell jitter keeps the scores distinct: the reference's sort is not stable, so the visiting order of EQUAL
        # scores is unspecified there (ours: larger candidate number first)
        blob = np.float32(ph) - np.float32(0.9) * d2 + np.float32(0.02) * rng.uniform(-1, 1, size=(H, W)).astype(np.float32)
        heat = np.maximum(heat, blob.astype(np.float32))
    off = rng.uniform(0, 1, size=(1, 2, H, W)).astype(np.float32) if with_offset else None
    return heat.reshape(1, 1, H, W), off


def random_peaks(H, W, n, seed, lo=0.5, hi=6.0):
    rng = np.random.RandomState(seed + 1000)
    return [(int(rng.randint(0, H)), int(rng.randint(0, W)), float(np.float32(rng.uniform(lo, hi)))) for _ in range(n)]


# name -> (H, W, number of peaks, seed, reduce, cls_thres, indexing, dist_thres, top_k (0 = all), with_offset)
CASES = {
    "wildtrack": (120, 360, 60, 1, 4, 0.4, "ij", 20.0, 0, True),
    "wildtrack_thres06": (120, 360, 60, 2, 4, 0.6, "ij", 20.0, 0, True),
    "multiviewx": (160, 250, 80, 3, 4, 0.4, "xy", 20.0, 0, True),
    "no_offset_topk": (40, 50, 30, 4, 4, 0.3, "xy", 12.0, 25, False),
    "crowded": (128, 128, 1500, 5, 4, 0.2, "ij", 20.0, 0, True),    # > 4096 candidates: global-memory sort path
    "empty": (32, 48, 0, 6, 4, 0.6, "ij", 20.0, 0, True),
    "single": (8, 8, 1, 7, 2, 0.5, "xy", 20.0, 0, True),
}


def case_inputs(name):
    H, W, n, seed, reduce, thres, indexing, dist, top_k, with_off = CASES[name]
    heat, off = blob_maps(H, W, random_peaks(H, W, n, seed), seed, with_offset=with_off)
    return heat, off, dict(reduce=reduce, cls_thres=thres, indexing=indexing, dist_thres=dist, top_k=top_k)


def demo_frame_maps(rows, H=120, W=360, reduce=4, seed=0):
    """Heatmap for one frame of the reference's bundled demo result file (evaluation/test-demo.txt): every detection
    (row, col) in grid units becomes a peak at cell (row/reduce, col/reduce) with zero offset (so it decodes to exactly
    that position), surrounded by weaker decoy cells that NMS must suppress, on a sub-threshold background."""
    rng = np.random.RandomState(seed)
    heat = (-7.0 + rng.uniform(-1, 1, size=(H, W))).astype(np.float32)
    off = rng.uniform(0, 1, size=(1, 2, H, W)).astype(np.float32)
    order = rng.permutation(len(rows))
    for rank, i in enumerate(order):
        r, c = int(rows[i][0]) // reduce, int(rows[i][1]) // reduce
        peak = np.float32(2.0 + 0.05 * rank)
        for dy in range(-2, 3):
            for dx in range(-2, 3):
                y, x = r + dy, c + dx
                if 0 <= y < H and 0 <= x < W and (dy or dx):
                    heat[y, x] = max(heat[y, x], np.float32(peak - 1.5 - 0.1 * (abs(dy) + abs(dx))))
    for rank, i in enumerate(order):  # peaks last: a neighbour's decoy never overwrites a detection
        r, c = int(rows[i][0]) // reduce, int(rows[i][1]) // reduce
        heat[r, c] = np.float32(2.0 + 0.05 * rank)
        off[0, :, r, c] = 0.0
    return heat.reshape(1, 1, H, W), off
