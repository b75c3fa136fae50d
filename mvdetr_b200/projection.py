"""Camera geometry of the hot path (host side, fp64 numpy, once per model / once per frame for the 3x3 chain).

  imgcoord_from_worldcoord_mat / worldcoord_from_imgcoord_mat   ref: multiview_detector/utils/projection.py:27-43
  project_points                                                 ref: multiview_detector/utils/projection.py:4-14
  world_grid_projection_mats   the `self.proj_mats` chain        ref: multiview_detector/models/mvdetr.py:82-95
  frame_projection_mats        the per-frame chain               ref: multiview_detector/models/mvdetr.py:155-161
  create_reference_map         per-(camera, point) reference table  ref: multiview_detector/models/mvdetr.py:33-71

`dataset` is any object exposing what the reference reads: Rworld_shape, world_reduce, img_reduce, num_cam and
base.{intrinsic_matrices, extrinsic_matrices, worldcoord_from_worldgrid_mat, world_indexing_from_xy_mat,
worldcoord_unit}.
"""
import numpy as np
import torch


def imgcoord_from_worldcoord_mat(intrinsic, extrinsic, z=0.0):
    """3x3 homography: ground-plane point at height z (world units) -> image pixel. K @ [R|t] @ lift(z)."""
    lift = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, z], [0, 0, 1.0]])
    return np.asarray(intrinsic, dtype=np.float64) @ np.asarray(extrinsic, dtype=np.float64) @ lift


def worldcoord_from_imgcoord_mat(intrinsic, extrinsic, z=0.0):
    return np.linalg.inv(imgcoord_from_worldcoord_mat(intrinsic, extrinsic, z))


def project_points(mat, pts):
    """Apply a 3x3 homography to [n,2] points (rows) with the homogeneous divide."""
    pts = np.asarray(pts, dtype=np.float64)
    hom = np.concatenate([pts, np.ones((pts.shape[0], 1))], axis=1) @ np.asarray(mat, dtype=np.float64).T
    return hom[:, :2] / hom[:, 2:3]


def _reduced_worldgrid_from_worldcoord(dataset, reduce):
    zoom = np.diag([reduce, reduce, 1.0])
    return np.linalg.inv(dataset.base.worldcoord_from_worldgrid_mat @ zoom @ dataset.base.world_indexing_from_xy_mat)


def world_grid_projection_mats(dataset, z=0.0):
    """[num_cam,3,3] fp64 tensor: image pixel (xy) -> reduced world-grid cell (xy), per camera."""
    to_grid = _reduced_worldgrid_from_worldcoord(dataset, dataset.world_reduce)
    mats = [to_grid @ worldcoord_from_imgcoord_mat(dataset.base.intrinsic_matrices[c],
                                                   dataset.base.extrinsic_matrices[c],
                                                   z / dataset.base.worldcoord_unit) for c in range(dataset.num_cam)]
    return torch.from_numpy(np.stack(mats))


def frame_projection_mats(proj_mats64, M, img_reduce):
    """Per-frame chain: feature-map pixel -> reduced world grid, given the augmentation matrices M [B,N,3,3]
    (image -> augmented image). fp32 result [B*N,3,3] on M's device, same operation order as the reference
    (torch.inverse in fp32, then the fp64 table cast to float)."""
    B, N = M.shape[:2]
    inv_aug = torch.inverse(M.reshape(B * N, 3, 3).float())
    scale = torch.diag(torch.tensor([img_reduce, img_reduce, 1.0], dtype=torch.float32, device=M.device))
    img_from_feat = inv_aug @ scale
    return proj_mats64.to(M.device).repeat(B, 1, 1).float() @ img_from_feat


def create_reference_map(dataset, n_points=4, downsample=2):
    """[Hd*Wd, num_cam, n_points, 2] normalised (x/Wd, y/Hd) reference points: for camera c and height z_i, where the
    ground cell's column of height z_i lands when it is re-projected through camera c onto the z=0 plane
    ("shadow" positions). n_points=4 uses z=0 only, i.e. the pixel-centre grid itself."""
    H, W = dataset.Rworld_shape
    H, W = H // downsample, W // downsample
    ys, xs = np.meshgrid(np.linspace(0.5, H - 0.5, H, dtype=np.float32), np.linspace(0.5, W - 0.5, W, dtype=np.float32),
                         indexing="ij")
    cells = np.stack([xs, ys], -1).reshape(-1, 2)
    if n_points == 4:
        heights = [0, 0, 0, 0]
    elif n_points == 8:
        heights = [-0.4, -0.2, 0, 0, 0.2, 0.4, 1, 1.8]
    else:
        raise ValueError("n_points must be 4 or 8")
    to_grid = _reduced_worldgrid_from_worldcoord(dataset, dataset.world_reduce * downsample)
    table = torch.zeros([H * W, dataset.num_cam, n_points, 2])
    for cam in range(dataset.num_cam):
        K, Rt = dataset.base.intrinsic_matrices[cam], dataset.base.extrinsic_matrices[cam]
        ground = to_grid @ worldcoord_from_imgcoord_mat(K, Rt)
        for i, z in enumerate(heights):
            at_z = to_grid @ worldcoord_from_imgcoord_mat(K, Rt, z / dataset.base.worldcoord_unit)
            in_image = project_points(np.linalg.inv(at_z), cells)
            table[:, cam, i, :] = torch.from_numpy(project_points(ground, in_image))
    table[..., 0] /= W
    table[..., 1] /= H
    return table
