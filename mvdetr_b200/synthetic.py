"""Synthetic calibration + inputs with the shapes of the reference's datasets (there is no network and no dataset
on the boxes). The objects returned expose exactly the attributes the reference model reads from its
`frameDataset` (ref: multiview_detector/models/mvdetr.py:34,46-56,78-95):
    Rimg_shape, Rworld_shape, img_reduce, world_reduce, num_cam,
    base.{intrinsic_matrices, extrinsic_matrices, worldcoord_from_worldgrid_mat, world_indexing_from_xy_mat,
          worldcoord_unit, indexing}

  wildtrack_like()   7 cameras, 1080p, 480x1440 grid of 2.5 cm cells, ij indexing   ref: datasets/Wildtrack.py:21-32
  multiviewx_like()  6 cameras, 1080p, 640x1000 grid of 2.5 cm cells (metres), xy   ref: datasets/MultiviewX.py:21-32
  stress4k_like()    8 cameras, 2160x3840, 960x2880 grid of 2.5 cm cells (BASELINE configs[3])
  mini_scene()       the small 3-camera scene behind tests/golden/world_feat_mini.npz
"""
import types

import numpy as np


def _look_at(eye, target):
    fwd = (target - eye) / np.linalg.norm(target - eye)
    right = np.cross(fwd, [0.0, 0.0, 1.0])
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd])
    return np.concatenate([R, (-R @ eye)[:, None]], axis=1)


def ring_scene(num_cam, Rworld_shape, Rimg_shape, world_reduce, img_reduce, worldcoord_from_worldgrid_mat, indexing,
               worldcoord_unit, focal_px, height_range, radius_scale=0.9, seed=0):
    """Pinhole cameras on a ring around the ground rectangle covered by the grid, each looking at its centre
    (+- jitter). Units follow `worldcoord_from_worldgrid_mat` (cm for Wildtrack, m for MultiviewX)."""
    rng = np.random.RandomState(seed)
    base = types.SimpleNamespace()
    base.worldcoord_unit = worldcoord_unit
    base.indexing = indexing
    base.world_indexing_from_xy_mat = (np.array([[0, 1, 0], [1, 0, 0], [0, 0, 1]], dtype=float) if indexing == "ij"
                                       else np.eye(3))
    base.worldcoord_from_worldgrid_mat = np.asarray(worldcoord_from_worldgrid_mat, dtype=float)
    nrow, ncol = Rworld_shape[0] * world_reduce, Rworld_shape[1] * world_reduce
    corners_grid = np.array([[0, 0, 1], [nrow, ncol, 1]], dtype=float).T if indexing == "ij" else \
        np.array([[0, 0, 1], [ncol, nrow, 1]], dtype=float).T
    corners = base.worldcoord_from_worldgrid_mat @ corners_grid
    lo, hi = corners[:2].min(1), corners[:2].max(1)
    centre, extent = (lo + hi) / 2, (hi - lo)
    H_img, W_img = Rimg_shape[0] * img_reduce, Rimg_shape[1] * img_reduce
    Ks, Rts = [], []
    for cam in range(num_cam):
        ang = 2 * np.pi * cam / num_cam + rng.uniform(-0.2, 0.2)
        radius = radius_scale * extent.max()
        eye = np.array([centre[0] + radius * np.cos(ang), centre[1] + radius * np.sin(ang),
                        rng.uniform(*height_range)])
        jitter = extent.max() * 0.02
        target = np.array([centre[0] + rng.uniform(-jitter, jitter), centre[1] + rng.uniform(-jitter, jitter), 0.0])
        Ks.append(np.array([[focal_px, 0, W_img / 2], [0, focal_px, H_img / 2], [0, 0, 1.0]]))
        Rts.append(_look_at(eye, target))
    base.intrinsic_matrices, base.extrinsic_matrices = Ks, Rts
    ds = types.SimpleNamespace()
    ds.base, ds.num_cam = base, num_cam
    ds.Rworld_shape, ds.Rimg_shape = list(Rworld_shape), list(Rimg_shape)
    ds.world_reduce, ds.img_reduce = world_reduce, img_reduce
    return ds


def wildtrack_like(seed=0, world_reduce=4, img_reduce=12):
    """7 x 1080p views, Rimg 90x160, Rworld 120x360 (main.py:179-181 defaults), grid origin (-300,-900) cm."""
    grid = [[0, 2.5, -300], [2.5, 0, -900], [0, 0, 1]]
    return ring_scene(7, (480 // world_reduce, 1440 // world_reduce), (1080 // img_reduce, 1920 // img_reduce),
                      world_reduce, img_reduce, grid, "ij", 0.01, focal_px=1750.0, height_range=(200.0, 400.0),
                      radius_scale=0.75, seed=seed)


def multiviewx_like(seed=0, world_reduce=4, img_reduce=12):
    """6 x 1080p views, Rworld 160x250, 16 m x 25 m plane in metres, xy indexing."""
    grid = [[0.025, 0, 0], [0, 0.025, 0], [0, 0, 1]]
    return ring_scene(6, (640 // world_reduce, 1000 // world_reduce), (1080 // img_reduce, 1920 // img_reduce),
                      world_reduce, img_reduce, grid, "xy", 1.0, focal_px=1750.0, height_range=(2.0, 4.0),
                      radius_scale=0.75, seed=seed)


def stress4k_like(seed=0, world_reduce=4, img_reduce=12):
    """8 x 4K views, Rimg 180x320, Rworld 240x720: the "8-view 4K, C=256, K=8" stress shape (the model on top uses
    hidden 256, 8 heads of D=32, 8 points)."""
    grid = [[0, 2.5, -600], [2.5, 0, -1800], [0, 0, 1]]
    return ring_scene(8, (960 // world_reduce, 2880 // world_reduce), (2160 // img_reduce, 3840 // img_reduce),
                      world_reduce, img_reduce, grid, "ij", 0.01, focal_px=3500.0, height_range=(300.0, 600.0),
                      radius_scale=0.75, seed=seed)


def mini_scene(num_cam=3, Rworld=(24, 40), Rimg=(18, 32), world_reduce=4, img_reduce=12, seed=0):
    """Small scene used for the committed goldens (tests/golden/make_golden.py)."""
    nrow, ncol = Rworld[0] * world_reduce, Rworld[1] * world_reduce
    cell = 2.5
    grid = [[0, cell, -ncol * cell / 2], [cell, 0, -nrow * cell / 2], [0, 0, 1]]
    return ring_scene(num_cam, Rworld, Rimg, world_reduce, img_reduce, grid, "ij", 0.01,
                      focal_px=0.9 * Rimg[1] * img_reduce, height_range=(250.0, 400.0), radius_scale=0.9, seed=seed)
