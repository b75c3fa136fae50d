"""CPU: the oracle (plain-C restatement + torch port) against the golden vectors produced by the REFERENCE's own
Python code (tests/golden/make_golden.py). This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import cpu_oracle as co
from oracle import torch_port as tp

CASES = ["msda_config1.npz", "msda_mvdetr_mini.npz", "msda_ragged.npz"]


def test_opstest_golden_fp64_and_fp32(golden):
    g = golden("msda_opstest.npz")
    o64 = co.msda_forward(g["value_f64"], g["shapes"], g["start"], g["loc_f64"], g["attn_f64"])
    assert np.allclose(o64, g["out_f64"], rtol=1e-5, atol=1e-8)  # reference tolerance, ops/test.py:40
    o32 = co.msda_forward(g["value_f32"], g["shapes"], g["start"], g["loc_f32"], g["attn_f32"])
    assert o32.dtype == np.float32
    assert np.allclose(o32, g["out_f32"], rtol=1e-2, atol=1e-3)  # ops/test.py:56
    assert np.abs(o32 - g["out_f32"]).max() <= 1e-4  # north-star tolerance
    gv, gl, ga = co.msda_backward(g["g_grad_out"], g["g_value"], g["shapes"], g["start"], g["g_loc"], g["g_attn"])
    assert np.allclose(gv, g["g_grad_value"], rtol=1e-9, atol=1e-12)
    assert np.allclose(gl, g["g_grad_loc"], rtol=1e-9, atol=1e-12)
    assert np.allclose(ga, g["g_grad_attn"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_reference_core_fp64(golden, name):
    g = golden(name)
    out = co.msda_forward(g["value"], g["shapes"], g["start"], g["loc"], g["attn"])
    assert np.allclose(out, g["out"], rtol=1e-9, atol=1e-12)
    gv, gl, ga = co.msda_backward(g["grad_out"], g["value"], g["shapes"], g["start"], g["loc"], g["attn"])
    assert np.allclose(gv, g["grad_value"], rtol=1e-9, atol=1e-11)
    assert np.allclose(gl, g["grad_loc"], rtol=1e-9, atol=1e-10)
    assert np.allclose(ga, g["grad_attn"], rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_fp32_within_1e4(golden, name):
    g = golden(name)
    f = np.float32
    out = co.msda_forward(g["value"].astype(f), g["shapes"], g["start"], g["loc"].astype(f), g["attn"].astype(f))
    assert np.abs(out - g["out"]).max() <= 1e-4


@pytest.mark.parametrize("name", CASES)
def test_torch_port_matches_reference_core(golden, name):
    g = golden(name)
    out = tp.msda_core(torch.from_numpy(g["value"]), g["shapes"].tolist(), torch.from_numpy(g["loc"]),
                       torch.from_numpy(g["attn"]))
    assert np.allclose(out.numpy(), g["out"], rtol=1e-10, atol=1e-12)


def test_empty_contribution_edge_cases():
    """All samples outside the map -> exact zeros; a sample exactly on the last pixel centre keeps one tap."""
    shapes = np.array([[3, 4]], dtype=np.int64)
    start = np.zeros(1, dtype=np.int64)
    value = np.arange(12 * 2, dtype=np.float64).reshape(1, 12, 1, 2) + 1
    loc = np.full((1, 1, 1, 1, 2, 2), 5.0)
    attn = np.full((1, 1, 1, 1, 2), 0.5)
    assert np.all(co.msda_forward(value, shapes, start, loc, attn) == 0)
    loc[..., 0] = (3 + 0.5) / 4  # x: pixel centre of the last column
    loc[..., 1] = (2 + 0.5) / 3  # y: pixel centre of the last row
    out = co.msda_forward(value, shapes, start, loc, attn)
    assert np.allclose(out[0, 0], value[0, 11, 0])
    loc[..., 0] = -0.5 / 4 + 1e-9  # w_im just above -1: only the right column taps survive, weight ~0
    out = co.msda_forward(value, shapes, start, loc, attn)
    assert np.all(np.abs(out) < 1e-6)


def test_prep_matches_reference_module_arithmetic(golden):
    m = golden("msda_module_mini.npz")
    loc, attn = co.msda_prep(m["offsets"], m["logits"], m["ref_table"], m["shapes"])
    assert np.abs(loc - m["loc"]).max() <= 1e-6
    assert np.abs(attn - m["attn"]).max() <= 1e-6
    out = co.msda_forward(m["value"], m["shapes"], m["start"], loc, attn)
    assert np.abs(out - m["core_out"]).max() <= 1e-4


def test_world_feat_port_matches_reference_model(golden):
    g = golden("world_feat_mini.npz")
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g if k.startswith("sd.")}
    x = torch.from_numpy(g["world_in"])
    N, C = x.shape[:2]
    out = tp.world_feat_forward(sd, x.view(1, N, C, *x.shape[2:]), torch.from_numpy(g["ref_points"]),
                                n_heads=int(g["nhead"]))
    assert (out - torch.from_numpy(g["out"])).abs().max().item() <= 1e-4


def test_warp_c_oracle_vs_kornia_restatement(golden):
    """Parity UNPINNED against the reference (kornia absent): this only ties the two restatements together."""
    g = golden("warp_small.npz")
    dsize = tuple(int(v) for v in g["dsize"])
    out = co.warp_forward(g["src"], g["mats"], dsize)
    assert np.abs(out - g["out"]).max() <= 1e-4
    T = tp.normalized_inverse_homography(torch.from_numpy(g["mats"]), g["src"].shape[-2:], dsize).numpy()
    out_t = co.warp_forward(g["src"], g["mats"], dsize, T=T)
    assert np.abs(out_t - g["out"]).max() <= 2e-5
    gs = co.warp_backward(g["grad_out"], g["mats"], g["src"].shape[-2:])
    assert np.abs(gs - g["grad_src"]).max() <= 2e-4
    # closed form of the kornia quirk: with M = I the sampled source coordinate of dst pixel u is
    # ix = u*Wi/(Wi-1) - 0.5 (an align_corners=True normalisation fed to an align_corners=False sampler).
    src = g["src"][2:3, :1]
    Hi, Wi = src.shape[-2:]
    same = co.warp_forward(src, np.eye(3, dtype=np.float32)[None], (Hi, Wi))
    ref = np.zeros_like(same)
    for v in range(Hi):
        for u in range(Wi):
            ix, iy = u * Wi / (Wi - 1) - 0.5, v * Hi / (Hi - 1) - 0.5
            x0, y0 = int(np.floor(ix)), int(np.floor(iy))
            for yy, wy in ((y0, y0 + 1 - iy), (y0 + 1, iy - y0)):
                for xx, wx in ((x0, x0 + 1 - ix), (x0 + 1, ix - x0)):
                    if 0 <= yy < Hi and 0 <= xx < Wi:
                        ref[0, 0, v, u] += src[0, 0, yy, xx] * wy * wx
    assert np.abs(same - ref).max() <= 1e-5
