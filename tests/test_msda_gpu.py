"""GPU parity tests for multi-scale deformable attention: CUDA (through the C ABI) vs the CPU oracle, the committed
golden vectors produced by the reference's Python core, the reference's own CUDA op (when oracle/_ref was built),
and size-independent properties at BASELINE.json's full Wildtrack size."""
import os

import numpy as np
import pytest
import torch

from mvdetr_b200 import ops
from oracle import cpu_oracle as co
from tests.gpu_util import dev, make_problem, ref_cuda_ext, viewgrid_problem

pytestmark = pytest.mark.gpu

FP32_ATOL = 1e-4  # BASELINE.json north_star: "within 1e-4 fp32"


def run_fwd(value, shapes, start, loc, attn, step=64):
    return ops.MSDeformAttnFunction.apply(value, shapes, start, loc, attn, step)


def test_reference_opstest_forward_double_and_float(golden, cuda):
    """ops/test.py:31-60 with the reference's shapes, seed and tolerances, against the reference core's outputs."""
    g = golden("msda_opstest.npz")
    shapes, start = dev(g["shapes"], cuda), dev(g["start"], cuda)
    o64 = run_fwd(dev(g["value_f64"], cuda), shapes, start, dev(g["loc_f64"], cuda), dev(g["attn_f64"], cuda), 2)
    assert torch.allclose(o64.cpu(), torch.from_numpy(g["out_f64"]))  # rtol 1e-5, atol 1e-8
    o32 = run_fwd(dev(g["value_f32"], cuda), shapes, start, dev(g["loc_f32"], cuda), dev(g["attn_f32"], cuda), 2)
    assert torch.allclose(o32.cpu(), torch.from_numpy(g["out_f32"]), rtol=1e-2, atol=1e-3)
    assert (o32.cpu() - torch.from_numpy(g["out_f32"])).abs().max() <= FP32_ATOL


@pytest.mark.parametrize("name", ["msda_config1.npz", "msda_mvdetr_mini.npz", "msda_ragged.npz"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_golden_forward_backward(golden, cuda, name, dtype):
    g = golden(name)
    value, loc, attn, go = (dev(g[k], cuda, dtype) for k in ("value", "loc", "attn", "grad_out"))
    shapes, start = dev(g["shapes"], cuda), dev(g["start"], cuda)
    value.requires_grad_(True), loc.requires_grad_(True), attn.requires_grad_(True)
    out = run_fwd(value, shapes, start, loc, attn)
    out.backward(go)
    if dtype == torch.float64:
        tol = dict(rtol=1e-9, atol=1e-10)
    else:
        tol = dict(rtol=0, atol=FP32_ATOL)
    assert np.allclose(out.detach().cpu().numpy(), g["out"], **tol)
    # grad_loc is piecewise constant in the sampling position: a sample within fp32 rounding of a cell boundary
    # (msda_mvdetr_mini has one at w_im = 3 - 6e-8) lands in the neighbouring cell in fp32. Mask those samples when
    # comparing fp32 results with the fp64 golden; the same-precision C oracle is compared unmasked.
    shp = g["shapes"].astype(np.float64)
    px = g["loc"] * shp[None, None, None, :, None, ::-1] - 0.5
    near = (np.abs(px - np.round(px)) < 1e-4).any(-1)
    n32 = [g[k].astype(np.float32) for k in ("grad_out", "value", "loc", "attn")]
    same_prec = co.msda_backward(n32[0], n32[1], g["shapes"], g["start"], n32[2], n32[3])
    for got, key, sp in zip((value.grad, loc.grad, attn.grad), ("grad_value", "grad_loc", "grad_attn"), same_prec):
        ref = g[key]
        scale = max(1.0, float(np.abs(ref).max()))
        got = got.cpu().numpy()
        if dtype == torch.float64:
            assert np.abs(got - ref).max() <= 1e-9 * scale, key
            continue
        assert np.abs(got - sp).max() <= FP32_ATOL * scale, (key, "vs fp32 C oracle")
        err = np.abs(got - ref)
        if key == "grad_loc":
            err = err[~near]
        if key != "grad_value" or not near.any():  # grad_value scatter targets also move with the cell
            assert err.max() <= FP32_ATOL * scale, (key, err.max())


@pytest.mark.parametrize("D", [16, 30, 32, 64, 71, 1025, 2048, 3096])
def test_reference_gradcheck_contract(cuda, D):
    """ops/test.py:63-86: fp64 gradcheck on the reference's D list (every reference backward variant) plus D=16."""
    N, M, Lq, L, P = 1, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long, device=cuda)
    start = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = 30
    torch.manual_seed(3)
    value = (torch.rand(N, S, M, D) * 0.01).double().to(cuda).requires_grad_(True)
    loc = torch.rand(N, Lq, M, L, P, 2).double().to(cuda).requires_grad_(True)
    attn = torch.rand(N, Lq, M, L, P) + 1e-5
    attn = (attn / attn.sum(-1, keepdim=True).sum(-2, keepdim=True)).double().to(cuda).requires_grad_(True)
    assert torch.autograd.gradcheck(ops.MSDeformAttnFunction.apply, (value, shapes, start, loc, attn, 2),
                                    nondet_tol=1e-9)


@pytest.mark.parametrize("D,M,P", [(4, 3, 2), (8, 8, 4), (16, 8, 4), (32, 8, 8), (64, 2, 4), (128, 1, 3), (12, 2, 4),
                                   (7, 3, 1)])
def test_seeded_random_vs_c_oracle_fp32(cuda, D, M, P):
    """Every vec4 instantiation and the scalar fallback, ragged levels, B=2 (B > im2col_step exercised below)."""
    shapes_list = [(13, 17), (7, 9), (4, 5), (1, 1)]
    prob = make_problem(2, shapes_list, M, D, 37, P, seed=100 + D)
    value, shapes, start, loc, attn, go = (t.to(cuda) for t in prob)
    value.requires_grad_(True), loc.requires_grad_(True), attn.requires_grad_(True)
    out = run_fwd(value, shapes, start, loc, attn)
    out.backward(go)
    n = [t.detach().cpu().numpy() for t in prob]
    ref = co.msda_forward(n[0], n[1], n[2], n[3], n[4])
    assert np.abs(out.detach().cpu().numpy() - ref).max() <= FP32_ATOL
    gv, gl, ga = co.msda_backward(n[5], n[0], n[1], n[2], n[3], n[4])
    for got, want in ((value.grad, gv), (loc.grad, gl), (attn.grad, ga)):
        scale = max(1.0, float(np.abs(want).max()))
        assert np.abs(got.cpu().numpy() - want).max() <= FP32_ATOL * scale


def test_im2col_step_semantics(cuda):
    """batch % min(batch, im2col_step) must be 0 (ms_deform_attn_cuda.cu:50-52); chunking never changes results."""
    prob = [t.to(cuda) for t in make_problem(4, [(5, 6)], 2, 16, 9, 4, seed=5)]
    a = run_fwd(*prob[:5], 64)
    b = run_fwd(*prob[:5], 2)
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError, match="must divide"):
        run_fwd(*prob[:5], 3)


def test_error_behaviour_on_gpu(cuda):
    prob = [t.to(cuda) for t in make_problem(1, [(5, 6)], 2, 16, 9, 4, seed=5)]
    with pytest.raises(RuntimeError, match="contiguous"):
        run_fwd(prob[0].transpose(1, 2).contiguous().transpose(1, 2), *prob[1:5])
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        run_fwd(prob[0], prob[1].cpu(), *prob[2:5])
    with pytest.raises(RuntimeError, match="not implemented for"):
        run_fwd(prob[0].half(), prob[1], prob[2], prob[3].half(), prob[4].half())


def test_all_samples_outside_give_exact_zero(cuda):
    value, shapes, start, loc, attn, _ = (t.to(cuda) for t in make_problem(1, [(6, 7), (3, 3)], 4, 16, 11, 4, seed=9))
    out = run_fwd(value, shapes, start, loc + 3.0, attn)
    assert torch.count_nonzero(out) == 0


def test_matches_reference_cuda_extension(cuda):
    """On-GPU parity with the reference's own CUDA op (built from its sources for sm_100a into oracle/_ref)."""
    ext = ref_cuda_ext()
    if ext is None:
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    value, shapes, start, loc, attn, go = (t.to(cuda) for t in viewgrid_problem(7, 30, 45, 8, 16, 4, seed=1))
    ours = ops.ms_deform_attn_forward(value, shapes, start, loc, attn, 64)
    theirs = ext.ms_deform_attn_forward(value, shapes, start, loc, attn, 64)
    assert (ours - theirs).abs().max().item() <= FP32_ATOL
    g_ours = ops.ms_deform_attn_backward(value, shapes, start, loc, attn, go, 64)
    g_theirs = ext.ms_deform_attn_backward(value, shapes, start, loc, attn, go, 64)
    for a, b in zip(g_ours, g_theirs):
        assert (a - b).abs().max().item() <= FP32_ATOL * max(1.0, b.abs().max().item())
    # fp64 too
    o64 = ops.ms_deform_attn_forward(value.double(), shapes, start, loc.double(), attn.double(), 64)
    t64 = ext.ms_deform_attn_forward(value.double(), shapes, start, loc.double(), attn.double(), 64)
    assert torch.allclose(o64, t64, rtol=1e-10, atol=1e-12)


def test_full_wildtrack_size_properties(cuda):
    """BASELINE config 2 size (Lq = S = 75600, M=8, D=16, L=7, P=4): properties that need no CPU reference.
    (a) constant value field and in-range samples => out = const * sum(attn) = const;
    (b) linearity in value and in attn;
    (c) a strided sample of queries agrees with the C oracle."""
    L, H, W, M, D, P = 7, 60, 180, 8, 16, 4
    value, shapes, start, loc, attn, _ = (t.to(cuda) for t in viewgrid_problem(L, H, W, M, D, P, seed=2))
    out = run_fwd(value, shapes, start, loc, attn)
    assert out.shape == (1, L * H * W, M * D) and torch.isfinite(out).all()
    # (a) interior queries only (their +-4 px samples stay inside the map)
    const = torch.full_like(value, 0.75)
    oc = run_fwd(const, shapes, start, loc.clamp(0.1, 0.9), attn)
    assert (oc - 0.75).abs().max().item() <= 1e-5
    # (b)
    v2 = torch.randn_like(value)
    lhs = run_fwd(value * 0.5 + v2 * 2.0, shapes, start, loc, attn)
    rhs = out * 0.5 + run_fwd(v2, shapes, start, loc, attn) * 2.0
    assert (lhs - rhs).abs().max().item() <= 1e-4
    assert (run_fwd(value, shapes, start, loc, attn * 3.0) - out * 3.0).abs().max().item() <= 1e-4
    # (c) every 997th query through the C oracle
    idx = torch.arange(0, L * H * W, 997, device=cuda)
    ref = co.msda_forward(value.cpu().numpy(), shapes.cpu().numpy(), start.cpu().numpy(),
                          loc[:, idx].cpu().numpy(), attn[:, idx].cpu().numpy())
    assert np.abs(out[:, idx].cpu().numpy() - ref).max() <= FP32_ATOL


def test_full_wildtrack_size_backward_properties(cuda):
    """Backward at full size: sum of grad_value equals sum over samples of attn*grad_out*(in-range bilinear weights),
    checked through <grad_value, 1> = d/dt out(value + t*1) . grad_out, plus an oracle check on a query sample."""
    L, H, W, M, D, P = 7, 60, 180, 8, 16, 4
    value, shapes, start, loc, attn, go = (t.to(cuda) for t in viewgrid_problem(L, H, W, M, D, P, seed=3))
    gv, gl, ga = ops.ms_deform_attn_backward(value, shapes, start, loc, attn, go, 64)
    ones = torch.ones_like(value)
    directional = (run_fwd(ones, shapes, start, loc, attn) * go).sum().item()  # out is linear in value
    assert abs(gv.sum().item() - directional) <= 1e-3 * max(1.0, abs(directional))
    # grad_attn[q,m,l,p] = <grad_out[q,m,:], bilinear(value)>  =>  sum_lp attn*grad_attn = <grad_out, out>
    out = run_fwd(value, shapes, start, loc, attn)
    lhs = (attn * ga).sum().item()
    rhs = (out * go).sum().item()
    assert abs(lhs - rhs) <= 1e-3 * max(1.0, abs(rhs))
    idx = torch.arange(0, L * H * W, 1999, device=cuda)
    n = lambda t: t.cpu().numpy()  # noqa: E731
    _, gl_ref, ga_ref = co.msda_backward(n(go[:, idx]), n(value), n(shapes), n(start), n(loc[:, idx]), n(attn[:, idx]))
    assert np.abs(n(gl[:, idx]) - gl_ref).max() <= FP32_ATOL * max(1.0, np.abs(gl_ref).max())
    assert np.abs(n(ga[:, idx]) - ga_ref).max() <= FP32_ATOL * max(1.0, np.abs(ga_ref).max())


def test_fused_entry_point_matches_module_arithmetic(golden, cuda):
    """mvd_msda_fused_fwd_f32 vs the reference module's own loc/softmax arithmetic + core (golden from the reference)."""
    m = golden("msda_module_mini.npz")
    value, offsets, logits, ref = (dev(m[k], cuda) for k in ("value", "offsets", "logits", "ref_table"))
    shapes, start = dev(m["shapes"], cuda), dev(m["start"], cuda)
    out, attn, loc = ops.msda_fused_forward(value, shapes, start, offsets, logits, ref, want_aux=True)
    assert np.abs(out.cpu().numpy() - m["core_out"]).max() <= FP32_ATOL
    assert np.abs(attn.cpu().numpy() - m["attn"]).max() <= 1e-6
    assert np.abs(loc.cpu().numpy() - m["loc"]).max() <= 1e-6
    out2 = ops.msda_fused_forward(value, shapes, start, offsets, logits, ref)
    assert torch.equal(out, out2)


def test_host_buffer_entry_point(cuda):
    """The *_host C entry point (used for e2e timing) gives the same bytes as the device entry point."""
    from mvdetr_b200 import _C
    prob = make_problem(1, [(9, 11), (5, 6)], 8, 16, 50, 4, seed=21)
    value, shapes, start, loc, attn, _ = prob
    out_host = torch.empty(1, 50, 128)
    rc = _C.lib.mvd_msda_fwd_f32_host(value.data_ptr(), shapes.data_ptr(), start.data_ptr(), loc.data_ptr(),
                                      attn.data_ptr(), 1, value.shape[1], 8, 16, 2, 50, 4, out_host.data_ptr(), None)
    assert rc == 0, _C.error_string(rc)
    out_dev = run_fwd(*(t.to(cuda) for t in prob[:5]))
    assert torch.equal(out_host, out_dev.cpu())


@pytest.mark.parametrize("L,H,W,M,D,P,R,B,offset_px", [
    (7, 30, 45, 8, 16, 4, 7, 1, 3.0),     # MVDeTr layout, offsets inside the staged window
    (7, 30, 45, 8, 16, 4, 7, 1, 40.0),    # most samples leave the window -> masked global path, same results
    (3, 13, 21, 2, 16, 4, 5, 2, 8.0),     # partial tiles, R != L, batch 2, mixed window/global
    (6, 9, 50, 4, 32, 8, 6, 1, 5.0),      # D=32, P=8 instantiation
    (2, 5, 7, 3, 8, 4, 1, 1, 2.0),        # D=8, a single replica of a grid smaller than one tile
    (4, 17, 33, 2, 16, 8, 20, 1, 6.0),    # many replicas -> smaller tile
    (7, 30, 45, 8, 16, 4, 1, 1, 4.0),     # one replica per rank of the view-sharded path -> taller (8 x 16) tile
    (7, 30, 45, 8, 16, 4, 2, 1, 4.0),     # two replicas (4 GPUs)
    (8, 20, 36, 8, 32, 8, 8, 1, 8.0),     # BASELINE configs[3] layout: D=32 (two lanes per pair), P=8 -> halo 9
    (4, 20, 36, 4, 32, 4, 1, 1, 4.0),     # D=32, P=4, one replica
])
def test_viewgrid_kernel_vs_c_oracle(cuda, L, H, W, M, D, P, R, B, offset_px):
    """The TMA-staged view-grid kernel (host-int geometry) against the C oracle and bit-level against nothing less:
    window path, out-of-window fallback, zero-filled borders, partial tiles."""
    probs = [viewgrid_problem(L, H, W, M, D, P, seed=50 + b, R=R, offset_px=offset_px) for b in range(B)]
    value, loc, attn = (torch.cat([p[i] for p in probs]).to(cuda) for i in (0, 3, 4))
    shapes, start = probs[0][1], probs[0][2]
    out = ops.msda_viewgrid_forward(value, loc, attn, H, W)
    ref = co.msda_forward(value.cpu().numpy(), shapes.numpy(), start.numpy(), loc.cpu().numpy(), attn.cpu().numpy())
    assert np.abs(out.cpu().numpy() - ref).max() <= FP32_ATOL
    # the drop-in op picks the same kernel from the device-side shapes and must agree exactly
    via_op = ops.ms_deform_attn_forward(value, shapes.to(cuda), start.to(cuda), loc, attn, 64)
    assert torch.equal(via_op, out)


def test_viewgrid_fused_matches_unfused_and_generic(cuda):
    """Fused (offsets/logits/ref-table) view-grid kernel == generic fused kernel's aux outputs fed to the plain op."""
    L, H, W, M, D, P = 7, 20, 36, 8, 16, 4
    g = torch.Generator().manual_seed(77)
    S = Lq = L * H * W
    value = torch.randn(1, S, M, D, generator=g).to(cuda)
    offsets = (torch.randn(1, Lq, M, L, P, 2, generator=g) * 3).to(cuda)
    logits = torch.randn(1, Lq, M, L * P, generator=g).to(cuda)
    ys, xs = torch.meshgrid(torch.linspace(0.5, H - 0.5, H), torch.linspace(0.5, W - 0.5, W), indexing="ij")
    table = torch.stack((xs / W, ys / H), -1).reshape(H * W, 1, 1, 2).repeat(1, L, P, 1).contiguous().to(cuda)
    shapes = torch.as_tensor([[H, W]] * L, dtype=torch.long, device=cuda)
    start = torch.arange(L, device=cuda) * (H * W)
    o_gen, attn, loc = ops.msda_fused_forward(value, shapes, start, offsets, logits, table, want_aux=True)
    o_vg = ops.msda_fused_forward(value, shapes, start, offsets, logits, table, grid_hw=(H, W))
    # aux outputs are only produced by the generic kernel (the view-grid kernel defers the softmax normalisation)
    _, attn2, loc2 = ops.msda_fused_forward(value, shapes, start, offsets, logits, table, want_aux=True, grid_hw=(H, W))
    assert torch.equal(attn, attn2) and torch.equal(loc, loc2)
    assert (o_vg - o_gen).abs().max().item() <= 1e-5
    ref = co.msda_forward(value.cpu().numpy(), shapes.cpu().numpy(), start.cpu().numpy(), loc.cpu().numpy(),
                          attn.cpu().numpy())
    assert np.abs(o_vg.cpu().numpy() - ref).max() <= FP32_ATOL
    # query subset (view-sharded ranks): rows of views 2..4 only
    hw = H * W
    sub = ops.msda_fused_forward(value, shapes, start, offsets[:, 2 * hw:5 * hw].contiguous(),
                                 logits[:, 2 * hw:5 * hw].contiguous(), table, grid_hw=(H, W))
    assert torch.equal(sub, o_vg[:, 2 * hw:5 * hw])
    # offsets / logits as column ranges of one [Lq, M*L*P*3] buffer (one GEMM for both Linear layers): read in place
    no, nl = M * L * P * 2, M * L * P
    both = torch.cat((offsets.view(Lq, no), logits.view(Lq, nl)), 1).contiguous()
    packed = ops.msda_fused_forward(value, shapes, start, both[:, :no].view(1, Lq, M, L, P, 2),
                                    both[:, no:].view(1, Lq, M, L * P), table, grid_hw=(H, W))
    assert torch.equal(packed, o_vg)
    # ... and through the generic kernel (no grid): densified copies, same results as the dense call
    packed_gen = ops.msda_fused_forward(value, shapes, start, both[:, :no].view(1, Lq, M, L, P, 2),
                                        both[:, no:].view(1, Lq, M, L * P), table)
    assert torch.equal(packed_gen, o_gen)


@pytest.mark.parametrize("grid", [True, False])
def test_fused_bias_folding_is_bit_identical(cuda, grid):
    """off_bias / logit_bias added inside the kernel == the same biases added by the Linear layers beforehand
    (fl(raw + bias) either way), for the view-grid and the generic fused kernel."""
    L, H, W, M, D, P = 5, 12, 20, 4, 16, 4
    g = torch.Generator().manual_seed(123)
    S = Lq = L * H * W
    value = torch.randn(1, S, M, D, generator=g).to(cuda)
    raw_off = (torch.randn(1, Lq, M, L, P, 2, generator=g) * 2).to(cuda)
    raw_log = torch.randn(1, Lq, M, L * P, generator=g).to(cuda)
    off_bias = (torch.randn(M * L * P * 2, generator=g) * 3).to(cuda)
    logit_bias = torch.randn(M * L * P, generator=g).to(cuda)
    ys, xs = torch.meshgrid(torch.linspace(0.5, H - 0.5, H), torch.linspace(0.5, W - 0.5, W), indexing="ij")
    table = torch.stack((xs / W, ys / H), -1).reshape(H * W, 1, 1, 2).repeat(1, L, P, 1).contiguous().to(cuda)
    shapes = torch.as_tensor([[H, W]] * L, dtype=torch.long, device=cuda)
    start = torch.arange(L, device=cuda) * (H * W)
    kw = dict(grid_hw=(H, W)) if grid else {}
    pre = ops.msda_fused_forward(value, shapes, start, raw_off + off_bias.view(1, 1, M, L, P, 2),
                                 raw_log + logit_bias.view(1, 1, M, L * P), table, **kw)
    folded = ops.msda_fused_forward(value, shapes, start, raw_off, raw_log, table, off_bias=off_bias,
                                    logit_bias=logit_bias, **kw)
    assert torch.equal(pre, folded)
    _, attn, loc = ops.msda_fused_forward(value, shapes, start, raw_off, raw_log, table, want_aux=True,
                                          off_bias=off_bias, logit_bias=logit_bias)
    ref = co.msda_forward(value.cpu().numpy(), shapes.cpu().numpy(), start.cpu().numpy(), loc.cpu().numpy(),
                          attn.cpu().numpy())
    assert np.abs(folded.cpu().numpy() - ref).max() <= FP32_ATOL


@pytest.mark.parametrize("C", [128, 256, 64])
def test_add_layer_norm_with_deferred_bias(cuda, C):
    g = torch.Generator().manual_seed(C)
    x, res = torch.randn(1000, C, generator=g).to(cuda), torch.randn(1000, C, generator=g).to(cuda)
    w, b, rb = (torch.randn(C, generator=g).to(cuda) for _ in range(3))
    want = torch.nn.functional.layer_norm(x + (res + rb), (C,), w, b, 1e-5)
    got = ops.add_layer_norm(x, res, w, b, 1e-5, res_bias=rb)
    assert (got - want).abs().max().item() <= 2e-5
    assert torch.equal(ops.add_layer_norm(x, res + rb, w, b, 1e-5), got)


@pytest.mark.parametrize("relu", [False, True])
def test_bias_act_inplace(cuda, relu):
    g = torch.Generator().manual_seed(7)
    x, b = torch.randn(777, 512, generator=g).to(cuda), torch.randn(512, generator=g).to(cuda)
    want = torch.relu(x + b) if relu else x + b
    got = ops.bias_act_(x.clone(), b, relu=relu)
    assert torch.equal(got, want)


def test_add_layer_norm_cell_major_output(cuda):
    g = torch.Generator().manual_seed(11)
    N, cells, C = 7, 60, 128
    x, res = torch.randn(N * cells, C, generator=g).to(cuda), torch.randn(N * cells, C, generator=g).to(cuda)
    w, b = torch.randn(C, generator=g).to(cuda), torch.randn(C, generator=g).to(cuda)
    plain = ops.add_layer_norm(x, res, w, b, 1e-5)
    perm = ops.add_layer_norm(x, res, w, b, 1e-5, perm_inner=cells)
    assert torch.equal(perm.view(cells, N, C), plain.view(N, cells, C).permute(1, 0, 2))


def test_empty_query_set_behaves_like_the_reference(cuda):
    """Lq = 0: the reference returns its zero-initialised (empty) output and zero gradients
    (ms_deform_attn_cuda.cu:54,121-123); no kernel is launched here."""
    value, shapes, start, loc, attn, go = make_problem(1, [(4, 5), (2, 3)], 2, 8, 3, 2, seed=1, device=cuda)
    l0, a0, g0 = loc[:, :0].contiguous(), attn[:, :0].contiguous(), go[:, :0].contiguous()
    out = ops.ms_deform_attn_forward(value, shapes, start, l0, a0, 64)
    assert tuple(out.shape) == (1, 0, 16)
    gv, gl, ga = ops.ms_deform_attn_backward(value, shapes, start, l0, a0, g0, 64)
    assert gv.shape == value.shape and float(gv.abs().sum()) == 0.0 and gl.numel() == 0 and ga.numel() == 0


@pytest.mark.skipif(os.environ.get("MVDETR_B200_BWD_VIEWGRID", "0") != "1",
                    reason="experimental view-grid backward: opt-in (MVDETR_B200_BWD_VIEWGRID=1), not yet validated")
@pytest.mark.parametrize("L,H,W,M,D,P,R,offset_px", [(7, 30, 45, 8, 16, 4, 7, 3.0), (7, 30, 45, 8, 16, 4, 7, 40.0),
                                                     (3, 13, 21, 2, 16, 4, 5, 8.0), (6, 9, 50, 4, 32, 8, 6, 5.0)])
def test_experimental_viewgrid_backward_matches_generic(cuda, L, H, W, M, D, P, R, offset_px):
    value, shapes, start, loc, attn, go = (t.to(cuda) for t in viewgrid_problem(L, H, W, M, D, P, seed=9, R=R,
                                                                               offset_px=offset_px))
    got = ops.ms_deform_attn_backward(value, shapes, start, loc, attn, go, 64)      # opt-in kernel (env set)
    old = ops._BWD_VIEWGRID
    try:
        ops._BWD_VIEWGRID = False
        want = ops.ms_deform_attn_backward(value, shapes, start, loc, attn, go, 64)  # generic kernel
    finally:
        ops._BWD_VIEWGRID = old
    for a, b in zip(got, want):
        assert (a - b).abs().max().item() <= 1e-4 * max(1.0, b.abs().max().item())
