#!/usr/bin/env python
"""bench.py -- multiview fusion frames/s on synthetic input (BASELINE.json configs[1] by default).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload wildtrack|multiviewx|stress4k]
        our arm (CUDA kernels behind the C ABI)
  python bench.py --impl reference [...]
        the reference's CPU-capable path (oracle port, no product import, all host threads)
  torchrun --nproc-per-node N ... bench.py --gpus N ...
        N>1: camera views sharded across ranks

One step = one frame through the fusion stage of MVDeTr.forward: perspective warp of the per-view feature maps onto
the ground grid followed by DeformTransWorldFeat (3 deformable-attention encoder layers), i.e.
ref multiview_detector/models/mvdetr.py:194-202. The backbone is out of scope; features are synthetic.
Workloads (BASELINE.json configs[1..3]):
  wildtrack   7 views, features [7,128,90,160] -> ground grid 120x360, C=128, 8 heads x D=16, 4 points   (headline)
  multiviewx  6 views, features [6,128,90,160] -> ground grid 160x250, same model; on 8 GPUs ranks 6-7 hold no view
  stress4k    8 views of 4K, features [8,256,180,320] -> ground grid 240x720, C=256, 8 heads x D=32, 8 points
Prints ONE JSON line (see the keys below); everything else goes to stderr.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

LAYERS = 3
WORKLOADS = {
    "wildtrack": dict(label="wildtrack_7view_1080p_resnet18feat_deform_trans", scene="wildtrack_like", hidden=128,
                      heads=8, points=4, cpu_strip=1),
    "multiviewx": dict(label="multiviewx_6view_1080p_resnet18feat_deform_trans", scene="multiviewx_like", hidden=128,
                       heads=8, points=4, cpu_strip=1),
    "stress4k": dict(label="stress_8view_4k_c256_d32_p8_deform_trans", scene="stress4k_like", hidden=256, heads=8,
                     points=8, cpu_strip=8),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak_hbm():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy kernel)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def msda_algorithmic_bytes(B, S, M, D, L, Lq, P, fused_ref_rows=0):
    """value + loc/offsets + attn/logits + out, fp32 (+ the compact reference table when fused)."""
    n = B * S * M * D + B * Lq * M * L * P * 2 + B * Lq * M * L * P + B * Lq * M * D + fused_ref_rows * L * P * 2
    return 4 * n


def warp_algorithmic_bytes(BN, C, Hi, Wi, Ho, Wo):
    """source read once + warped grid written once (what the reference's op boundary moves)."""
    return 4 * BN * C * (Hi * Wi + Ho * Wo) + 36 * BN


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:  # region shorter than one sampling period: one synchronous query right after it
            try:
                q = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                f = [x.strip() for x in q.strip().splitlines()[0].split(",")]
                return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]),
                        "reasons": [nm for nm, val in zip(names, f[3:7]) if val.lower().startswith("active")],
                        "samples": 0, "note": "timed region shorter than the sampling period; sampled right after it"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_config(wl, sc_num_cam, feat_shape, world_grid):
    return {"workload": wl["label"], "views": sc_num_cam, "feat": list(feat_shape), "world_grid": list(world_grid),
            "layers": LAYERS, "heads": wl["heads"], "points": wl["points"], "hidden": wl["hidden"]}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU-capable path restated in oracle/ (kind "port"). Imports NOTHING from the product
# package: the problem (calibration, projection chain, reference table, weights) comes from oracle/ref_problem.py.
# ------------------------------------------------------------------------------------------------------------
def cpu_threads():
    """All host cores, explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently turn the CPU
    arm into a single-thread run."""
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def cpu_problem(wl_name, seed=0):
    from oracle import ref_problem
    from oracle import torch_port as tp
    wl = WORKLOADS[wl_name]
    p = ref_problem.problem(wl_name, seed=seed, strip=wl["cpu_strip"])
    Hg, Wg = p["sc"]["Rworld"]
    p["pos"] = tp.sine_pos_embedding((Hg // 2, Wg // 2), wl["hidden"] // 2)  # built once, as the reference's ctor does
    p["wl"] = wl
    return p


def cpu_step(p):
    from oracle import torch_port as tp
    sc, wl = p["sc"], p["wl"]
    with torch.no_grad():
        world = tp.warp_perspective(p["feat"], p["proj"], tuple(sc["Rworld"]))
        return tp.world_feat_forward(p["sd"], world.view(1, sc["num_cam"], wl["hidden"], *sc["Rworld"]), p["ref"],
                                     n_heads=wl["heads"], n_points=wl["points"], n_layers=LAYERS,
                                     pos_embedding=p["pos"])


def cpu_sample_text(p):
    sc, wl = p["sc"], p["wl"]
    full = "the FULL frame" if wl["cpu_strip"] == 1 else f"a 1/{wl['cpu_strip']} column strip of the ground grid"
    return (f"{full}: all {sc['num_cam']} views [{sc['num_cam']},{wl['hidden']},{sc['Rimg'][0]},{sc['Rimg'][1]}] -> "
            f"{sc['Rworld'][0]}x{sc['Rworld'][1]} ground cells, 3 encoder layers, merge + upsample; fp32 torch CPU")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_threads()
    p = cpu_problem(args.workload)
    strip = p["wl"]["cpu_strip"]
    for _ in range(max(1, args.warmup)):
        cpu_step(p)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(p)
    dt = (time.perf_counter() - t0) / args.steps
    fps = (1.0 / strip) / dt
    sc = p["sc"]
    full = ref_full_dims(args.workload)
    line = {"impl": "reference", "metric": "multiview_frames_per_sec", "value": fps, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(p["wl"], sc["num_cam"], (sc["num_cam"], p["wl"]["hidden"], *sc["Rimg"]), full),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": cpu_sample_text(p), "host_cpus": os.cpu_count(),
                             "frames_per_step": 1.0 / strip},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def ref_full_dims(wl_name):
    from oracle import ref_problem
    w = ref_problem.WORKLOADS[wl_name]
    return [w["grid"][0] // w["world_reduce"], w["grid"][1] // w["world_reduce"]]


def cpu_baseline_leg(wl_name, reps=3):
    cores = cpu_threads()
    p = cpu_problem(wl_name)
    strip = p["wl"]["cpu_strip"]
    cpu_step(p)
    t0 = time.perf_counter()
    for _ in range(reps):
        cpu_step(p)
    dt = (time.perf_counter() - t0) / reps
    return {"value": (1.0 / strip) / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{reps} reps of {cpu_sample_text(p)}", "host_cpus": os.cpu_count(),
            "seconds_per_frame_equiv": dt * strip}


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def build_fusion(device, wl, seed=0):
    """Random-init fusion stage with the workload's geometry; offsets/attention made query dependent (SURVEY 8d config
    2: default init has zero weight => every query samples the same fixed ring, an unrealistically cache-friendly
    pattern)."""
    from mvdetr_b200 import synthetic
    from mvdetr_b200.fusion import MultiviewFusion
    torch.manual_seed(seed)
    ds = getattr(synthetic, wl["scene"])(seed=seed)
    fusion = MultiviewFusion(ds, base_dim=wl["hidden"], hidden_dim=wl["hidden"], nhead=wl["heads"],
                             n_points=wl["points"])
    with torch.no_grad():
        for layer in fusion.world_feat.encoder.layers:
            layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
            layer.self_attn.attention_weights.weight.normal_(0, 0.05)
    return ds, fusion.to(device).eval()


def synthetic_frames(ds, hidden, n, seed=0, pin=False):
    g = torch.Generator().manual_seed(seed)
    feats, Ms = [], []
    for _ in range(n):
        f = torch.randn(ds.num_cam, hidden, *ds.Rimg_shape, generator=g)
        feats.append(f.pin_memory() if pin else f)
        Ms.append(torch.eye(3).view(1, 1, 3, 3).repeat(1, ds.num_cam, 1, 1))
    return feats, Ms


def time_kernel_events(fn, iters, flush=None):
    """(median, min) device time of fn() in microseconds: CUDA events on the current stream, 3 untimed warm-ups,
    optional L2 flush (a write larger than L2) before every timed launch."""
    for _ in range(3):
        fn()
    evs = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) * 1e3 for a, b in evs]
    return statistics.median(ts), min(ts)


def load_ref_ext():
    ref_so = os.path.join(REPO, "oracle", "_ref", "MultiScaleDeformableAttention.so")
    if not os.path.exists(ref_so):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("MultiScaleDeformableAttention", ref_so)
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    return ext


def kernel_breakdown(fusion, ds, wl, device, iters=20):
    """Per-kernel device times for the hot-path kernels on this workload's real tensors (L2 flushed between
    launches with a 512 MB memset), plus the reference's own CUDA op on the same inputs when oracle/_ref exists."""
    from mvdetr_b200 import ops
    wf = fusion.world_feat
    HIDDEN, HEADS, POINTS = wl["hidden"], wl["heads"], wl["points"]
    N, (Hg, Wg) = ds.num_cam, ds.Rworld_shape
    Hd, Wd = Hg // 2, Wg // 2
    Lq = S = N * Hd * Wd
    D = HIDDEN // HEADS
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=device)
    g = torch.Generator(device="cpu").manual_seed(1)
    feat = torch.randn(N, HIDDEN, *ds.Rimg_shape, generator=g).to(device)
    proj = fusion.projection(torch.eye(3).view(1, 1, 3, 3).repeat(1, N, 1, 1)).to(device)
    res = {}
    with torch.no_grad():
        wb = warp_algorithmic_bytes(N, HIDDEN, *ds.Rimg_shape, Hg, Wg)

        def warp_entry(fn, **extra):
            t, tmin = time_kernel_events(fn, iters, flush)
            return dict({"us": t, "us_min": tmin, "bytes": wb, "GBps": wb / t / 1e3}, **extra)

        # NCHW features in (the backbone's layout), channels-last world grid out
        res["warp"] = warp_entry(lambda: ops.warp_perspective(feat, proj, (Hg, Wg), align_corners=False,
                                                              channels_last=True),
                                 launches=ops.warp_launch_names(feat, channels_last=True))
        # what the frame runner launches: the same warp, written straight into the downsample conv's im2col matrix.
        # `bytes` / GBps use SURVEY 8(d)'s ALGORITHMIC warp bytes (source + warped grid once); `written_bytes` is what
        # the kernel really stores ([tokens, 9C] matrix = 2.25 copies of the grid)
        ib = 4 * N * HIDDEN * ds.Rimg_shape[0] * ds.Rimg_shape[1] + 4 * N * Hd * Wd * 9 * HIDDEN + 36 * N
        t, tmin = time_kernel_events(lambda: ops.warp_im2col(feat, proj, (Hg, Wg), stride=2), iters, flush)
        res["warp_im2col"] = {"us": t, "us_min": tmin, "bytes": wb, "GBps": wb / t / 1e3, "moved_bytes": ib,
                              "moved_GBps": ib / t / 1e3, "launches": ops.warp_launch_names(feat, im2col=True)}
        feat_cl = feat.contiguous(memory_format=torch.channels_last)
        res["warp_cl_src"] = warp_entry(lambda: ops.warp_perspective(feat_cl, proj, (Hg, Wg), align_corners=False,
                                                                     channels_last=True),
                                        launches=ops.warp_launch_names(feat_cl, channels_last=True))
        res["warp_nchw_contract"] = warp_entry(lambda: ops.warp_perspective(feat, proj, (Hg, Wg),
                                                                            align_corners=False),
                                               launches=ops.warp_launch_names(feat, channels_last=False))
        # realistic MSDA inputs: run the model's own first layer projections on a real frame
        world = ops.warp_perspective(feat, proj, (Hg, Wg), align_corners=False).view(1, N, HIDDEN, Hg, Wg)
        x = wf.downsample(world.view(N, HIDDEN, Hg, Wg))
        del world
        src = x.view(1, N, HIDDEN, Hd, Wd).permute(0, 1, 3, 4, 2).reshape(1, S, HIDDEN)
        pos = (wf.pos_embedding.flatten(2).transpose(1, 2).unsqueeze(1) + wf.lvl_embedding.view(1, N, 1, HIDDEN)
               ).view(1, S, HIDDEN)
        attn_mod = wf.encoder.layers[0].self_attn
        value = attn_mod.value_proj(src).view(1, S, HEADS, D).contiguous()
        offsets = attn_mod.sampling_offsets(src + pos).view(1, Lq, HEADS, N, POINTS, 2).contiguous()
        logits = attn_mod.attention_weights(src + pos).view(1, Lq, HEADS, N * POINTS).contiguous()
        geo = wf._level_geometry(N, Hd, Wd, torch.device(device))
        table = wf.encoder.ref_table
        out, attn, loc = ops.msda_fused_forward(value, geo.shapes, geo.start, offsets, logits, table, want_aux=True)
        t, tmin = time_kernel_events(lambda: ops.msda_fused_forward(value, geo.shapes, geo.start, offsets, logits,
                                                                    table), iters, flush)
        fb = msda_algorithmic_bytes(1, S, HEADS, D, N, Lq, POINTS, fused_ref_rows=table.shape[0])
        res["msda_fused_fwd_generic"] = {"us": t, "us_min": tmin, "bytes": fb, "GBps": fb / t / 1e3}
        # exactly the frame's launch: raw (bias-free) GEMM outputs + the two Linear biases added in the kernel
        q2 = (src + pos).view(Lq, HIDDEN)
        raw_off = ops.linear(q2, attn_mod.sampling_offsets.weight).view(1, Lq, HEADS, N, POINTS, 2)
        raw_log = ops.linear(q2, attn_mod.attention_weights.weight).view(1, Lq, HEADS, N * POINTS)
        t, tmin = time_kernel_events(lambda: ops.msda_fused_forward(value, geo.shapes, geo.start, raw_off, raw_log,
                                                                    table, grid_hw=(Hd, Wd),
                                                                    ref_table_lm=wf.encoder.ref_table_lm,
                                                                    off_bias=attn_mod.sampling_offsets.bias,
                                                                    logit_bias=attn_mod.attention_weights.bias),
                                     iters, flush)
        res["msda_fused_fwd"] = {"us": t, "us_min": tmin, "bytes": fb + 4 * 3 * HEADS * N * POINTS, "GBps": fb / t / 1e3}
        vg = ops.msda_fused_forward(value, geo.shapes, geo.start, offsets, logits, table, grid_hw=(Hd, Wd),
                                    ref_table_lm=wf.encoder.ref_table_lm)
        res["viewgrid_vs_generic_max_abs_diff"] = (vg - out).abs().max().item()
        t, tmin = time_kernel_events(lambda: ops.ms_deform_attn_forward(value, geo.shapes, geo.start, loc, attn, 64),
                                     iters, flush)
        ub = msda_algorithmic_bytes(1, S, HEADS, D, N, Lq, POINTS)
        res["msda_fwd"] = {"us": t, "us_min": tmin, "bytes": ub, "GBps": ub / t / 1e3}
        go = torch.randn_like(out)
        t, tmin = time_kernel_events(lambda: ops.ms_deform_attn_backward(value, geo.shapes, geo.start, loc, attn, go,
                                                                         64), max(5, iters // 2), flush)
        bb = 4 * (3 * S * HEADS * D + Lq * HEADS * D + 2 * 3 * Lq * HEADS * N * POINTS)
        res["msda_bwd"] = {"us": t, "us_min": tmin, "bytes": bb, "GBps": bb / t / 1e3,
                           "kernel": ops.msda_bwd_kernel_name(value, geo.hw, Lq)}
        if ops._BWD_BANDED:  # the same kernel walking the pairs in query order (round 1), for comparison
            ops._BWD_BANDED = False
            try:
                t, tmin = time_kernel_events(lambda: ops.ms_deform_attn_backward(value, geo.shapes, geo.start, loc, attn,
                                                                                 go, 64), max(5, iters // 2), flush)
            finally:
                ops._BWD_BANDED = True
            res["msda_bwd_query_order"] = {"us": t, "us_min": tmin, "GBps": bb / t / 1e3}
        try:  # comparator only; never fatal
            ext = load_ref_ext()
            if ext is not None:
                t, tmin = time_kernel_events(lambda: ext.ms_deform_attn_forward(value, geo.shapes, geo.start, loc,
                                                                                attn, 64), iters, flush)
                res["ref_cuda_msda_fwd"] = {"us": t, "us_min": tmin, "GBps": ub / t / 1e3}
                t, tmin = time_kernel_events(lambda: ext.ms_deform_attn_backward(value, geo.shapes, geo.start, loc,
                                                                                 attn, go, 64), max(5, iters // 2),
                                             flush)
                res["ref_cuda_msda_bwd"] = {"us": t, "us_min": tmin, "GBps": bb / t / 1e3}
                diff = (ext.ms_deform_attn_forward(value, geo.shapes, geo.start, loc, attn, 64) - out).abs().max()
                res["ref_cuda_max_abs_diff"] = diff.item()
        except Exception as e:
            res["ref_cuda_error"] = repr(e)[:200]
    del flush
    return res


def ref_cuda_frame(fusion, ds, wl, device, ours_out, feat, proj, iters=10):
    """SURVEY 8(d) "Reference comparators (i)": the reference's frame on the SAME B200 -- its DeformTransWorldFeat
    arithmetic (trans_world_feat.py:87-110, restated functionally in oracle/torch_port.py over the same weights),
    kornia-style warp through F.grid_sample, cuDNN convolutions / cuBLAS Linear layers through torch, and the
    reference's OWN CUDA op (oracle/_ref, its sources compiled for sm_100a) for the deformable attention.
      as_shipped   with the per-frame host->device uploads of the position embedding (trans_world_feat.py:93) and the
                   repeated reference-point table (deformable_transformer.py:48) and the device->host assert
                   (ms_deform_attn.py:94); wall clock around synchronised frames
      kernels_only those tensors already resident, no assert; CUDA events
    COMPARATOR ONLY: nothing here is on the product path."""
    ext = load_ref_ext()
    if ext is None:
        return {"unavailable": "oracle/_ref/MultiScaleDeformableAttention.so not built"}
    from oracle import torch_port as tp
    wf = fusion.world_feat
    N, (Hg, Wg) = ds.num_cam, ds.Rworld_shape
    sd = {k: v.detach() for k, v in wf.state_dict().items()}
    ref_host = wf.encoder.reference_points.detach().cpu()        # [N*Hd*Wd, N, P, 2], as mvdetr.py:129-130 builds it
    pos_host = wf.pos_embedding.detach().cpu()
    ref_dev, pos_dev = ref_host.to(device), pos_host.to(device)

    def msda_fn(value, shapes, start, loc, attn):
        return ext.ms_deform_attn_forward(value.contiguous(), shapes, start, loc.contiguous(), attn.contiguous(), 64)

    def frame(ref, pos, sync_assert):
        world = tp.warp_perspective(feat, proj, (Hg, Wg))
        return tp.world_feat_forward(sd, world.view(1, N, wl["hidden"], Hg, Wg), ref, n_heads=wl["heads"],
                                     n_points=wl["points"], n_layers=LAYERS, pos_embedding=pos, msda_fn=msda_fn,
                                     sync_assert=sync_assert)

    res = {}
    with torch.no_grad():
        for _ in range(3):
            out = frame(ref_dev, pos_dev, False)
        res["max_abs_diff_vs_ours"] = (out - ours_out).abs().max().item()
        t, tmin = time_kernel_events(lambda: frame(ref_dev, pos_dev, False), iters)
        res["kernels_only"] = {"ms": t / 1e3, "ms_min": tmin / 1e3, "frames_per_sec": 1e6 / t}
        for _ in range(2):
            frame(ref_host, pos_host, True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            frame(ref_host, pos_host, True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / iters
        res["as_shipped"] = {"ms": dt * 1e3, "frames_per_sec": 1.0 / dt,
                             "h2d_bytes_per_frame": ref_host.numel() * 4 + pos_host.numel() * 4}
    res["what"] = ("reference DeformTransWorldFeat arithmetic + F.grid_sample warp + the reference's own CUDA op "
                   "(oracle/_ref) + cuDNN/cuBLAS fp32 through torch, same weights and inputs, same GPU; features "
                   "already on the device in both variants")
    return res


def gemm_accuracy_check(device, hidden):
    """The dense layers run on split operands (fp16 / bf16 terms on the tensor cores): measure, live, the largest error of
    ops.linear against an fp64 product on a layer-shaped problem, next to the same figure for torch's native fp32 GEMM
    (TF32 off). dtype "f32" in the line means fp32-LEVEL results, and this is the evidence."""
    from mvdetr_b200 import ops
    g = torch.Generator().manual_seed(7)
    out = {}
    for name, K, N in (("linear_128", hidden, 4 * hidden), ("conv_9x128", 9 * hidden, hidden)):
        x = torch.randn(8192, K, generator=g).to(device)
        w = (torch.randn(N, K, generator=g) / K ** 0.5).to(device)
        b = torch.randn(N, generator=g).to(device)
        exact = x.double() @ w.double().t() + b.double()
        out[name] = {"max_err_ours": (ops.linear(x, w, b).double() - exact).abs().max().item(),
                     "max_err_torch_fp32": (torch.addmm(b, x, w.t()).double() - exact).abs().max().item(),
                     "max_abs_value": exact.abs().max().item()}
    return out


def our_launches_per_frame(fusion, lt_gemm):
    """OUR kernels per frame: the warp (1 launch with the TMA kernel on the NCHW source, else relayout + gather); per
    encoder layer 1 fused MSDA + 2 add_layernorm (the 2nd also emits the next layer's query) + the 5 Linear GEMMs (offsets
    and logits share one) when
    they run on our tcgen05 kernel (cuBLASLt launches are NOT counted; torch.mm mode adds 2 bias_act); on the GEMM conv
    path the downsample / merge / upsample-conv GEMMs (ours), upsample_im2col and the final NHWC->NCHW transpose."""
    from mvdetr_b200 import ops
    own = ops._GEMM_MODE in ("bf16x3", "f16x2", "tf32x3")
    n = ops.warp_launch_count(im2col=fusion.gemm_path)
    n += LAYERS * (3 + (5 if own else 0) + (0 if (own or lt_gemm) else 2))  # offsets + logits: one GEMM
    if fusion.gemm_path:
        n += 2 + (3 if own else 0)
    return n


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference "
                         "for the CPU arm)")
    wl = WORKLOADS[args.workload]
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    torch.backends.cuda.matmul.allow_tf32 = False  # dense glue stays full fp32, like the reference
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True  # as the reference sets it (main.py:48); tuned before graph capture
    if world > 1:
        import datetime
        import threading
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=180))
        # hard watchdog: a wedged collective must end the process (non-zero), never hang the box
        wd = threading.Timer(600.0, lambda: (log("bench.py: watchdog fired, aborting"), os._exit(3)))
        wd.daemon = True
        wd.start()
        if os.environ.get("MVD_BENCH_TRACE"):  # debugging aid: dump every thread's Python stack if we stall
            import faulthandler
            faulthandler.dump_traceback_later(float(os.environ["MVD_BENCH_TRACE"]), exit=True, file=sys.stderr)

    from mvdetr_b200.fusion import FrameRunner
    HIDDEN = wl["hidden"]
    ds, fusion = build_fusion(device, wl)
    BN = ds.num_cam
    feat_shape = (BN, HIDDEN, *ds.Rimg_shape)
    if world > 1:
        from mvdetr_b200.sharded import ShardedFrameRunner
        runner = ShardedFrameRunner(fusion, feat_shape, device, rank, world)
        mode = runner.mode
    else:
        runner = FrameRunner(fusion, feat_shape, device, use_graph=True, depth=2)
        mode = "single"

    n_frames = 4 if args.workload != "stress4k" else 2  # distinct synthetic frames cycled through (> L2 in total)
    feats_pinned, Ms = synthetic_frames(ds, HIDDEN, n_frames, seed=0, pin=True)
    feats_dev = [f.to(device) for f in feats_pinned]
    projs_dev = [fusion.projection(M).to(device) for M in Ms]
    out_shape = (1, HIDDEN, *ds.Rworld_shape)

    def barrier():
        # drain this rank's streams BEFORE the collective: the graph replays on runner.compute contain all-gathers on
        # the same communicator, and a barrier kernel racing them on another stream can order differently per rank
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    log(f"[rank {rank}] runner ready: mode={mode}")
    # ---- device-resident throughput (`value`): inputs already in HBM, alternate the two slots ----
    for s in range(runner.depth):
        runner.load(feats_dev[s % n_frames], projs_dev[s % n_frames], slot=s)
    for i in range(args.warmup):
        runner.step(i % runner.depth)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()  # `ncu --profile-from-start off` sees exactly the timed steps (no warm-up / autotune)
    with torch.cuda.stream(runner.compute):
        e0.record()
    for i in range(args.steps):
        runner.step(i % runner.depth)
    with torch.cuda.stream(runner.compute):
        e1.record()
    barrier()
    torch.cuda.profiler.stop()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms / args.steps
    log(f"[rank {rank}] timed region done: {ms_per_step:.3f} ms/step")
    frames_per_step = runner.frames_per_step if world > 1 else 1
    value = frames_per_step * 1e3 / ms_per_step

    # ---- end to end: pinned host features in, fused world feature out, every step ----
    outs_pinned = [torch.empty(out_shape).pin_memory() for _ in range(n_frames)]
    k = args.steps
    fl = [feats_pinned[i % n_frames] for i in range(k)]
    ml = [Ms[i % n_frames] for i in range(k)]
    ol = [outs_pinned[i % n_frames] for i in range(k)]
    runner.run_host_frames(fl[:max(2, args.warmup)], ml[:max(2, args.warmup)], ol[:max(2, args.warmup)])
    barrier()
    t0 = time.perf_counter()
    runner.run_host_frames(fl, ml, ol)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = t.item()
    e2e_fps = frames_per_step * k / e2e_s
    log(f"[rank {rank}] e2e done: {e2e_fps:.1f} frames/s")
    h2d = feats_pinned[0].numel() * 4 + BN * 36
    d2h = outs_pinned[0].numel() * 4

    line = None
    from mvdetr_b200 import ops as _ops
    lt = _ops.linear_available()
    lt_gemm = bool(lt) and _ops._GEMM_MODE != "torch"
    gemm_mode = _ops.gemm_mode_text()
    if rank == 0:
        kb = kernel_breakdown(fusion, ds, wl, device, iters=20 if args.workload != "stress4k" else 6)
        peak, peak_src = measured_peak_hbm()
        dom = kb["msda_fused_fwd"]
        D = HIDDEN // wl["heads"]
        roofline = {"kernel": f"msda_vg_kernel<{D},{wl['points']},FUSED> (mvd_msda_fused_fwd_viewgrid_f32), "
                              f"{LAYERS} launches/step",
                    "bound": "hbm", "achieved": dom["GBps"], "peak": peak, "unit": "GB/s",
                    "frac": dom["GBps"] / peak, "traffic": None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": dom["bytes"], "us_per_launch": dom["us"],
                    "timing": "CUDA events on the launch stream, median of 20, 512 MB L2 flush before every launch"}
        implicit = fusion.gemm_path and _ops.conv3x3_implicit_ok(HIDDEN, HIDDEN)  # feature channels = hidden (synthetic.py)
        wk = kb["warp_im2col"] if (fusion.gemm_path and not implicit) else kb["warp"]  # the launch the frame issues
        hot_us = wk["us"] + LAYERS * dom["us"]
        prof = {}
        if args.workload == "wildtrack":
            try:  # DRAM traffic / pipe utilisation of the dominant kernel from the committed ncu --set full capture
                with open(os.path.join(REPO, "profiles", "ncu_dominant_kernel.json")) as f:
                    prof = json.load(f)
            except Exception:
                pass
        roofline["traffic"] = prof.get("dram_bytes_per_launch")
        roofline["traffic_source"] = prof.get("source")
        roofline["onchip_frac"] = prof.get("l1tex_data_pipe_frac")
        roofline["onchip_note"] = ("the kernel's binding resource is the SM L1/shared-memory data pipe (4 corners x D x 4 "
                                   "B per sample through a 128 B/clk/SM pipe; DESIGN.md 4.1); onchip_frac = l1tex "
                                   "data-pipe utilisation from the ncu capture")
        cpu = cpu_baseline_leg(args.workload) if world == 1 else None
        rcf = None
        if world == 1:
            try:
                with torch.no_grad():
                    ours_out = fusion.fuse(feats_dev[0], projs_dev[0]).clone()
                rcf = ref_cuda_frame(fusion, ds, wl, device, ours_out, feats_dev[0], projs_dev[0],
                                     iters=10 if args.workload != "stress4k" else 3)
                if "kernels_only" in rcf:
                    rcf["ours_over_ref_kernels_only"] = value / rcf["kernels_only"]["frames_per_sec"]
                    rcf["ours_over_ref_as_shipped"] = value / rcf["as_shipped"]["frames_per_sec"]
            except Exception as e:  # comparator only; never fatal
                rcf = {"error": repr(e)[:300]}
        cfg = workload_config(wl, BN, feat_shape, ds.Rworld_shape)
        try:
            cfg["gemm_error_vs_fp64"] = gemm_accuracy_check(device, HIDDEN)
        except Exception as e:  # evidence only; never fatal
            cfg["gemm_error_vs_fp64"] = repr(e)[:200]
        cfg.update({"mode": mode, "cuda_graph": True, "tf32": False, "gemm": gemm_mode,
                    "convs": (("3x3 convs as implicit GEMMs on our tcgen05 kernel (taps fetched by TMA, no im2col matrix)"
                               if world == 1 else "downsample conv as implicit GEMM on our tcgen05 kernel; row-band tail "
                               "through the im2col GEMM (ours)") if implicit else "3x3 convs as im2col GEMMs (ours)")
                    if fusion.gemm_path else "cuDNN fp32",
                    "l2": "2 alternating frame slots; per-step working set ~1.5 GB >> 126 MB L2"})
        line = {"metric": "multiview_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                # one frame's work is split across the ranks (views) => total work fixed as N grows, at every N
                "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": cfg,
                "roofline": roofline,
                "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "pipeline": "pinned host -> H2D stream / graph replay / D2H stream, 2 slots"},
                # rank 0's own kernels; the view-sharded path adds the multicast copy of the final tokens (the symmetric-
                # memory barrier kernels and NCCL's are not ours and not counted)
                "gpu_launches": (our_launches_per_frame(fusion, lt_gemm) + (1 if world > 1 and "multicast" in mode else 0))
                                * args.steps,
                "clocks": clocks,
                "hot_path": {"warp_us": wk["us"], "warp_kernel": wk.get("launches"), "msda_fused_fwd_us": dom["us"],
                             "frames_per_sec_kernels_only": 1e6 / hot_us,
                             "warp_algorithmic_bytes": wk["bytes"], "warp_GBps": wk["GBps"],
                             "warp_frac": wk["GBps"] / peak,
                             "warp_frac_note": "SURVEY 8(d) algorithmic bytes (source + warped grid once) / time of the "
                                               "launch(es) the frame issues for the stage"},
                "kernels": kb}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if rcf is not None:
            line["ref_cuda_frame"] = rcf
    if line is not None:
        emit(line)
    if world > 1:
        # Tear down in the order NCCL needs: the captured graphs hold the communicator's kernels, and
        # destroy_process_group() with live graphs never returns (seen on 2xB200: both ranks stuck in it).
        dist.barrier()
        runner.graphs = [None] * runner.depth
        del runner
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout; everything else (ours, NCCL's version banner, library
    chatter) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)   # keep the real stdout for the result line ...
    os.dup2(2, 1)            # ... and send every other write to fd 1 (C libraries included) to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 200 (ours), 20 (reference: ~1 s per CPU frame)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("MVD_BENCH_WORKLOAD", "wildtrack"),
                    choices=sorted(WORKLOADS))
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 200 if args.impl == "ours" else 20
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
