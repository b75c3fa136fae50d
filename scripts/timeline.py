"""Per-kernel device timeline of the frame (1 GPU) or of the view-sharded frame (torchrun, N ranks), through
torch.profiler (CUPTI activity records; nsys is not in the image). Rank 0 prints, per kernel name, launches and total /
mean device time over the profiled steps, the span of one step and the idle time inside it, and writes a Chrome trace.

  python scripts/timeline.py [--steps 6] [--out gpurun_out/timeline]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      scripts/timeline.py --out gpurun_out/timeline_8gpu
"""
import argparse
import collections
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402  (build_fusion / synthetic_frames / WORKLOADS)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--workload", default="wildtrack")
    ap.add_argument("--out", default="gpurun_out/timeline")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    wl = bench.WORKLOADS[args.workload]
    ds, fusion = bench.build_fusion(dev, wl)
    feat_shape = (ds.num_cam, wl["hidden"], *ds.Rimg_shape)
    if world > 1:
        from mvdetr_b200.sharded import ShardedFrameRunner
        runner = ShardedFrameRunner(fusion, feat_shape, dev, rank, world)
    else:
        from mvdetr_b200.fusion import FrameRunner
        runner = FrameRunner(fusion, feat_shape, dev, use_graph=True, depth=2)
    feats, Ms = bench.synthetic_frames(ds, wl["hidden"], 2, seed=0)
    for s in range(2):
        runner.load(feats[s].to(dev), fusion.projection(Ms[s]).to(dev), slot=s)
    for i in range(4):
        runner.step(i % 2)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(args.steps):
            runner.step(i % 2)
        torch.cuda.synchronize()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        prof.export_chrome_trace(args.out + ".trace.json")
        ev = [e for e in json.load(open(args.out + ".trace.json"))["traceEvents"]
              if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
        ev.sort(key=lambda e: e["ts"])
        per = collections.OrderedDict()
        for e in ev:
            d = per.setdefault(e["name"][:110], [0, 0.0])
            d[0] += 1
            d[1] += e["dur"]
        span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
        # busy time = union of kernel intervals (streams overlap)
        busy, end = 0.0, ev[0]["ts"]
        for e in ev:
            s0, s1 = max(e["ts"], end), e["ts"] + e["dur"]
            if s1 > s0:
                busy += s1 - s0
                end = s1
        lines = [f"# rank 0 of {world}, workload {args.workload}, {args.steps} steps: span {span / args.steps:.1f} us/step, "
                 f"device busy {busy / args.steps:.1f} us/step, idle {(span - busy) / args.steps:.1f} us/step",
                 f"{'kernel':112s} {'n/step':>7s} {'us/step':>9s} {'us each':>9s}"]
        for name, (n, tot) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            lines.append(f"{name:112s} {n / args.steps:7.1f} {tot / args.steps:9.1f} {tot / n:9.1f}")
        text = "\n".join(lines)
        open(args.out + ".txt", "w").write(text + "\n")
        print(text)
        if os.path.getsize(args.out + ".trace.json") > 20 << 20:
            os.remove(args.out + ".trace.json")
    if world > 1:
        dist.barrier()
        runner.graphs = [None] * runner.depth
        del runner
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
