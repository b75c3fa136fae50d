"""Self-contained builder of the benchmark workloads for the CPU (reference) arm -- TEST INFRASTRUCTURE, NOT PRODUCT.

`bench.py --impl reference` must not import the product package (its process would map libmvdetr_b200.so), so the
synthetic calibration, the projection chain, the reference-point table and a random-init state_dict with the
reference's parameter names are restated here with numpy / torch only:

  scene()                  ring of pinhole cameras around the ground grid (same recipe and seed as
                           mvdetr_b200/synthetic.py; tests/test_bench_cpu.py checks both give identical matrices)
  world_grid_proj_mats()   ref: multiview_detector/models/mvdetr.py:82-95   (+ utils/projection.py:27-43)
  frame_proj_mats()        ref: multiview_detector/models/mvdetr.py:155-161
  reference_map()          ref: multiview_detector/models/mvdetr.py:33-71
  state_dict()             parameter names/shapes/initialisation of DeformTransWorldFeat
                           ref: multiview_detector/models/trans_world_feat.py:70-119,
                                multiview_detector/models/ops/modules/ms_deform_attn.py:62-77
  problem(workload)        everything one CPU step needs (oracle.torch_port.world_feat_forward arguments)
"""
import math

import numpy as np
import torch

# name -> (num_cam, image HxW, img_reduce, grid rows x cols (full res), world_reduce, grid->world matrix, indexing,
#          unit, camera height range, base_dim = hidden, heads, points)
WORKLOADS = {
    "wildtrack": dict(num_cam=7, img=(1080, 1920), img_reduce=12, grid=(480, 1440), world_reduce=4,
                      grid_mat=[[0, 2.5, -300], [2.5, 0, -900], [0, 0, 1]], indexing="ij", unit=0.01,
                      heights=(200.0, 400.0), hidden=128, heads=8, points=4),
    "multiviewx": dict(num_cam=6, img=(1080, 1920), img_reduce=12, grid=(640, 1000), world_reduce=4,
                       grid_mat=[[0.025, 0, 0], [0, 0.025, 0], [0, 0, 1]], indexing="xy", unit=1.0,
                       heights=(2.0, 4.0), hidden=128, heads=8, points=4),
    "stress4k": dict(num_cam=8, img=(2160, 3840), img_reduce=12, grid=(960, 2880), world_reduce=4,
                     grid_mat=[[0, 2.5, -600], [2.5, 0, -1800], [0, 0, 1]], indexing="ij", unit=0.01,
                     heights=(300.0, 600.0), hidden=256, heads=8, points=8),
}
FOCAL_PER_1080 = 1750.0


def _look_at(eye, target):
    fwd = (target - eye) / np.linalg.norm(target - eye)
    right = np.cross(fwd, [0.0, 0.0, 1.0])
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd])
    return np.concatenate([R, (-R @ eye)[:, None]], axis=1)


def scene(workload, seed=0):
    w = WORKLOADS[workload]
    rng = np.random.RandomState(seed)
    grid_mat = np.asarray(w["grid_mat"], dtype=float)
    nrow, ncol = w["grid"]
    corners_grid = (np.array([[0, 0, 1], [nrow, ncol, 1]], dtype=float).T if w["indexing"] == "ij" else
                    np.array([[0, 0, 1], [ncol, nrow, 1]], dtype=float).T)
    corners = grid_mat @ corners_grid
    lo, hi = corners[:2].min(1), corners[:2].max(1)
    centre, extent = (lo + hi) / 2, hi - lo
    H_img, W_img = w["img"]
    focal = FOCAL_PER_1080 * H_img / 1080.0
    Ks, Rts = [], []
    for cam in range(w["num_cam"]):
        ang = 2 * np.pi * cam / w["num_cam"] + rng.uniform(-0.2, 0.2)
        radius = 0.75 * extent.max()
        eye = np.array([centre[0] + radius * np.cos(ang), centre[1] + radius * np.sin(ang),
                        rng.uniform(*w["heights"])])
        jitter = extent.max() * 0.02
        target = np.array([centre[0] + rng.uniform(-jitter, jitter), centre[1] + rng.uniform(-jitter, jitter), 0.0])
        Ks.append(np.array([[focal, 0, W_img / 2], [0, focal, H_img / 2], [0, 0, 1.0]]))
        Rts.append(_look_at(eye, target))
    perm = np.array([[0, 1, 0], [1, 0, 0], [0, 0, 1]], dtype=float) if w["indexing"] == "ij" else np.eye(3)
    return dict(w, K=Ks, Rt=Rts, grid_mat=grid_mat, perm=perm,
                Rworld=(nrow // w["world_reduce"], ncol // w["world_reduce"]),
                Rimg=(H_img // w["img_reduce"], W_img // w["img_reduce"]))


def _world_from_img(K, Rt, z=0.0):
    lift = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, z], [0, 0, 1.0]])
    return np.linalg.inv(np.asarray(K, dtype=np.float64) @ np.asarray(Rt, dtype=np.float64) @ lift)


def _grid_from_world(sc, reduce):
    return np.linalg.inv(sc["grid_mat"] @ np.diag([reduce, reduce, 1.0]) @ sc["perm"])


def world_grid_proj_mats(sc, z=0.0):
    to_grid = _grid_from_world(sc, sc["world_reduce"])
    return torch.from_numpy(np.stack([to_grid @ _world_from_img(sc["K"][c], sc["Rt"][c], z / sc["unit"])
                                      for c in range(sc["num_cam"])]))


def frame_proj_mats(proj64, M, img_reduce):
    B, N = M.shape[:2]
    inv_aug = torch.inverse(M.reshape(B * N, 3, 3).float())
    scale = torch.diag(torch.tensor([img_reduce, img_reduce, 1.0]))
    return proj64.repeat(B, 1, 1).float() @ (inv_aug @ scale)


def _project(mat, pts):
    hom = np.concatenate([pts, np.ones((pts.shape[0], 1))], axis=1) @ np.asarray(mat, dtype=np.float64).T
    return hom[:, :2] / hom[:, 2:3]


def reference_map(sc, downsample=2):
    H, W = sc["Rworld"][0] // downsample, sc["Rworld"][1] // downsample
    ys, xs = np.meshgrid(np.linspace(0.5, H - 0.5, H, dtype=np.float32),
                         np.linspace(0.5, W - 0.5, W, dtype=np.float32), indexing="ij")
    cells = np.stack([xs, ys], -1).reshape(-1, 2).astype(np.float64)
    heights = [0, 0, 0, 0] if sc["points"] == 4 else [-0.4, -0.2, 0, 0, 0.2, 0.4, 1, 1.8]
    to_grid = _grid_from_world(sc, sc["world_reduce"] * downsample)
    table = torch.zeros([H * W, sc["num_cam"], len(heights), 2])
    for cam in range(sc["num_cam"]):
        ground = to_grid @ _world_from_img(sc["K"][cam], sc["Rt"][cam])
        for i, z in enumerate(heights):
            at_z = to_grid @ _world_from_img(sc["K"][cam], sc["Rt"][cam], z / sc["unit"])
            table[:, cam, i, :] = torch.from_numpy(_project(ground, _project(np.linalg.inv(at_z), cells)))
    table[..., 0] /= W
    table[..., 1] /= H
    return table


def _xavier(g, *shape):
    fan_out = shape[0] * int(np.prod(shape[2:])) if len(shape) > 2 else shape[0]
    fan_in = shape[1] * int(np.prod(shape[2:])) if len(shape) > 2 else shape[1]
    a = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(*shape, generator=g) * 2 - 1) * a


def state_dict(sc, seed=0, ffn=512, layers=3, query_dependent=True):
    """Random-init weights with the reference's key names. query_dependent: sampling_offsets.weight ~ N(0, 0.01) and
    attention_weights.weight ~ N(0, 0.05) as the GPU arm sets them."""
    g = torch.Generator().manual_seed(seed)
    C, N, M, P = sc["hidden"], sc["num_cam"], sc["heads"], sc["points"]
    sd = {"downsample.0.weight": _xavier(g, C, C, 3, 3), "downsample.0.bias": torch.zeros(C),
          "lvl_embedding": torch.randn(N, C, generator=g),
          "merge_linear.0.weight": _xavier(g, C, C * N, 1, 1), "merge_linear.0.bias": torch.zeros(C),
          "upsample.1.weight": _xavier(g, C, C, 3, 3), "upsample.1.bias": torch.zeros(C)}
    thetas = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
    grid = torch.stack([thetas.cos(), thetas.sin()], -1)
    grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(M, 1, 1, 2).repeat(1, N, P, 1)
    for i in range(P):
        grid[:, :, i, :] *= i + 1
    for i in range(layers):
        p = f"encoder.layers.{i}."
        sd[p + "self_attn.sampling_offsets.weight"] = (torch.randn(M * N * P * 2, C, generator=g) * 0.01
                                                      if query_dependent else torch.zeros(M * N * P * 2, C))
        sd[p + "self_attn.sampling_offsets.bias"] = grid.reshape(-1).clone()
        sd[p + "self_attn.attention_weights.weight"] = (torch.randn(M * N * P, C, generator=g) * 0.05
                                                       if query_dependent else torch.zeros(M * N * P, C))
        sd[p + "self_attn.attention_weights.bias"] = torch.zeros(M * N * P)
        for name in ("value_proj", "output_proj"):
            sd[p + f"self_attn.{name}.weight"] = _xavier(g, C, C)
            sd[p + f"self_attn.{name}.bias"] = torch.zeros(C)
        sd[p + "norm1.weight"], sd[p + "norm1.bias"] = torch.ones(C), torch.zeros(C)
        sd[p + "norm2.weight"], sd[p + "norm2.bias"] = torch.ones(C), torch.zeros(C)
        sd[p + "linear1.weight"], sd[p + "linear1.bias"] = _xavier(g, ffn, C), torch.zeros(ffn)
        sd[p + "linear2.weight"], sd[p + "linear2.bias"] = _xavier(g, C, ffn), torch.zeros(C)
    return sd


def problem(workload="wildtrack", seed=0, strip=1):
    """-> dict(sc, sd, ref [N*Hd*Wd, N, P, 2], proj [N,3,3] fp32, feat [N, C, Hf, Wf]).
    strip > 1: bounded sample -- all views, but only the first 1/strip of the ground grid's columns (every stage of the
    path is linear in the number of ground cells)."""
    sc = scene(workload, seed)
    N = sc["num_cam"]
    if strip > 1:
        sc["Rworld"] = (sc["Rworld"][0], sc["Rworld"][1] // strip)
    ref = reference_map(sc).repeat([N, 1, 1, 1])
    proj = frame_proj_mats(world_grid_proj_mats(sc), torch.eye(3).view(1, 1, 3, 3).repeat(1, N, 1, 1), sc["img_reduce"])
    g = torch.Generator().manual_seed(seed)
    feat = torch.randn(N, sc["hidden"], *sc["Rimg"], generator=g)
    return dict(sc=sc, sd=state_dict(sc, seed), ref=ref, proj=proj, feat=feat)
