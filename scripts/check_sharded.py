"""GPU check of the view-sharded path (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/check_sharded.py
Every rank computes the whole frame on its own GPU (single-GPU path) and its share of the view-sharded path, for the
Wildtrack-shaped (7 views) and the MultiviewX-shaped (6 views, odd grid width) scenes, and compares. Results equal up
to the reduction order inside cuBLAS (row partitions of the same GEMMs); tolerance 1e-4 (north star)."""
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from mvdetr_b200 import synthetic  # noqa: E402
from mvdetr_b200.fusion import MultiviewFusion  # noqa: E402
from mvdetr_b200.sharded import ShardedFrameRunner  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ok = True
    # the one-view scene leaves every rank but the first WITHOUT views (the 8-GPU run of a 7-view scene has one such rank)
    for name, make in (("wildtrack", synthetic.wildtrack_like), ("multiviewx", synthetic.multiviewx_like),
                       ("one_view", lambda seed: synthetic.mini_scene(num_cam=1, seed=seed))):
        torch.manual_seed(0)
        ds = make(seed=0)
        fusion = MultiviewFusion(ds, base_dim=128, hidden_dim=128, nhead=8, n_points=4)
        with torch.no_grad():
            for layer in fusion.world_feat.encoder.layers:
                layer.self_attn.sampling_offsets.weight.normal_(0, 0.01)
                layer.self_attn.attention_weights.weight.normal_(0, 0.05)
        fusion = fusion.to(dev).eval()
        g = torch.Generator().manual_seed(1)
        feat = torch.randn(ds.num_cam, 128, *ds.Rimg_shape, generator=g).to(dev)
        M = torch.eye(3).view(1, 1, 3, 3).repeat(1, ds.num_cam, 1, 1)
        proj = fusion.projection(M).to(dev)
        with torch.no_grad():
            single = fusion.fuse(feat, proj)
        for use_graph in (False, True):
            runner = ShardedFrameRunner(fusion, tuple(feat.shape), dev, rank, world, use_graph=use_graph)
            runner.load(feat, proj, slot=0)
            out = runner.step(0)
            torch.cuda.synchronize()
            err = (out - single).abs().max().item()
            scale = single.abs().max().item()
            errs = [None] * world
            dist.all_gather_object(errs, err)
            if rank == 0:
                print(f"{name}: views={ds.num_cam} world={world} mode={runner.mode} graph={use_graph} "
                      f"max|sharded-single| per rank={['%.2e' % e for e in errs]} (|out| max {scale:.3f})", flush=True)
            ok = ok and max(errs) <= 1e-4 * max(1.0, scale)
            del runner
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDED CHECK", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
