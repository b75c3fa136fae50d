"""The per-frame multiview fusion path as one module: projection chain -> perspective warp -> DeformTransWorldFeat.

This is the slice of MVDeTr.forward between the backbone and the world heads
  ref: multiview_detector/models/mvdetr.py:155-161 (per-frame projection matrices), :192-202 (warp + world_feat)
with the model set-up of mvdetr.py:82-95,129-132 (projection table, reference map).

  MultiviewFusion   nn.Module; forward(imgs_feat [B*N,C,Hf,Wf], M [B,N,3,3]) -> world_feat [B,hidden,Hg,Wg]
  FrameRunner       steady-state executor for inference: static device buffers, the whole (warp -> world_feat)
                    step captured in a CUDA graph, and a double-buffered host pipeline (H2D / compute / D2H on three
                    streams) for end-to-end frames from pinned host memory.
"""
import os

import torch
from torch import nn

from . import ops
from .projection import create_reference_map, frame_projection_mats, world_grid_projection_mats
from .world_feat import DeformTransWorldFeat

_FUSED_QUERY = os.environ.get("MVDETR_B200_FUSED_QUERY", "0") == "1"


class MultiviewFusion(nn.Module):
    def __init__(self, dataset, base_dim=128, z=0, hidden_dim=128, nhead=8, dim_feedforward=512, n_points=4,
                 channels_last_warp=True):
        super().__init__()
        self.num_cam = dataset.num_cam
        self.Rworld_shape = [int(v) for v in dataset.Rworld_shape]
        self.img_reduce = dataset.img_reduce
        self.channels_last_warp = channels_last_warp
        self.gemm_path = os.environ.get("MVDETR_B200_CONV", "gemm") != "cudnn"  # A/B switch: cuDNN convolutions
        # fp64 table kept on the host like the reference's plain attribute (mvdetr.py:93-95)
        self.proj_mats = world_grid_projection_mats(dataset, z)
        reference_points = create_reference_map(dataset, n_points).repeat([dataset.num_cam, 1, 1, 1])
        self.world_feat = DeformTransWorldFeat(dataset.num_cam, dataset.Rworld_shape, base_dim, hidden_dim=hidden_dim,
                                               nhead=nhead, dim_feedforward=dim_feedforward, n_points=n_points,
                                               stride=2, reference_points=reference_points)

    def projection(self, M):
        """M [B,N,3,3] (host or device) -> [B*N,3,3] fp32 feature-pixel -> world-grid homographies."""
        return frame_projection_mats(self.proj_mats, M, self.img_reduce)

    def fuse(self, imgs_feat, proj_mats):
        """imgs_feat [B*N,C,Hf,Wf] and device-resident proj_mats [B*N,3,3] -> [B,hidden,Hg,Wg]. No host work:
        this is the part FrameRunner captures in a CUDA graph."""
        BN, C = imgs_feat.shape[:2]
        B = BN // self.num_cam
        Hg, Wg = self.Rworld_shape
        if B == 1 and self.gemm_path and self.world_feat.fast_path_ok(imgs_feat):
            # inference: convs as tensor-core GEMMs. Default: warp to a channels-last grid (one TMA-staged launch on the
            # NCHW source), downsample conv as an IMPLICIT GEMM fetching its taps by TMA -- no im2col matrix in memory
            if ops.conv3x3_implicit_ok(C, self.world_feat.hidden_dim):
                g_cl = ops.warp_perspective(imgs_feat, proj_mats, (Hg, Wg), align_corners=False, channels_last=True)
                # MVDETR_B200_FUSED_QUERY=1: the conv epilogue also writes tokens + pos (the first layer's query). Measured
                # r02u: +42 us on the conv against 15 us for the separate element-wise add (whose operands sit in L2), so
                # it is off by default
                fq = _FUSED_QUERY
                res = self.world_feat.tokens_from_warped(g_cl, with_query=fq)
                if res is not None:
                    tokens, q0 = res if fq else (res, None)
                    return self.world_feat.forward_from_tokens(tokens, self.num_cam, (Hg - 1) // 2 + 1,
                                                               (Wg - 1) // 2 + 1, query0=q0)
            # im2col route: warp straight into the downsample conv's im2col matrix
            A, (Hd, Wd) = ops.warp_im2col(imgs_feat, proj_mats, (Hg, Wg), stride=2)
            return self.world_feat.forward_from_im2col(A, self.num_cam, Hd, Wd)
        cl = self.channels_last_warp and not torch.is_grad_enabled()
        world = ops.warp_perspective(imgs_feat, proj_mats, (Hg, Wg), align_corners=False, channels_last=cl)
        if cl:  # [BN,Hg,Wg,C] storage viewed as NCHW: the stride-2 conv then runs channels-last, no permute-copy
            world = world.permute(0, 3, 1, 2)
        return self.world_feat(world.view(B, self.num_cam, C, Hg, Wg) if not cl else
                               world.unflatten(0, (B, self.num_cam)))

    def forward(self, imgs_feat, M):
        proj = self.projection(M).to(imgs_feat.device, non_blocking=True)
        return self.fuse(imgs_feat, proj)


class FrameRunner:
    """Inference executor over a MultiviewFusion: `step()` replays one frame on device-resident inputs;
    `run_host_frames()` streams frames from pinned host memory through a 2-deep H2D/compute/D2H pipeline."""

    def __init__(self, fusion, feat_shape, device, use_graph=True, depth=2):
        self.fusion = fusion.to(device).eval()
        self.device = torch.device(device)
        self.depth = depth
        BN = feat_shape[0]
        self.feat = [torch.zeros(feat_shape, device=device) for _ in range(depth)]
        self.proj = [torch.eye(3, device=device).repeat(BN, 1, 1) for _ in range(depth)]
        self.out = [None] * depth
        self.graphs = [None] * depth
        self.compute = torch.cuda.Stream(device=device)
        self.s_in = torch.cuda.Stream(device=device)
        self.s_out = torch.cuda.Stream(device=device)
        self.use_graph = use_graph
        self.recapture()

    def _param_versions(self):
        return tuple((p.data_ptr(), p._version) for p in self.fusion.parameters())

    def recapture(self):
        """(Re)runs the warm-up and captures the CUDA graphs. The graphs bake in derived copies of some parameters (conv
        weights in GEMM layout, position + level embedding); after load_state_dict() or any in-place weight update call
        this again -- step() refuses to replay a graph captured from other parameter versions."""
        depth, device = self.depth, self.device
        self.graphs = [None] * depth
        with torch.no_grad():
            with torch.cuda.stream(self.compute):
                for i in range(depth):
                    for _ in range(2):  # warm-up: cuDNN/cuBLAS heuristics, lazy buffers
                        self.out[i] = self.fusion.fuse(self.feat[i], self.proj[i])
                self.compute.synchronize()
                if self.use_graph:
                    for i in range(depth):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=self.compute):
                            self.out[i] = self.fusion.fuse(self.feat[i], self.proj[i])
                        self.graphs[i] = g
        self._captured_versions = self._param_versions()
        torch.cuda.synchronize(device)

    def load(self, imgs_feat, proj_mats, slot=0):
        self.feat[slot].copy_(imgs_feat)
        self.proj[slot].copy_(proj_mats)

    def step(self, slot=0):
        """One frame on the inputs already in slot `slot`; asynchronous on self.compute; returns the output buffer."""
        with torch.cuda.stream(self.compute):
            if self.graphs[slot] is not None:
                if self._param_versions() != self._captured_versions:
                    raise RuntimeError("FrameRunner: model parameters changed since the CUDA graphs were captured "
                                       "(load_state_dict / in-place update); call recapture() first")
                self.graphs[slot].replay()
            else:
                with torch.no_grad():
                    self.out[slot] = self.fusion.fuse(self.feat[slot], self.proj[slot])
        return self.out[slot]

    def run_host_frames(self, feats_pinned, Ms, outs_pinned):
        """feats_pinned[i] [BN,C,Hf,Wf] pinned, Ms[i] [B,N,3,3] host, outs_pinned[i] pinned [B,hidden,Hg,Wg].
        Frame i uses slot i % depth. Returns after the last device->host copy has completed."""
        n = len(feats_pinned)
        ev_in = [torch.cuda.Event() for _ in range(self.depth)]
        ev_done = [torch.cuda.Event() for _ in range(self.depth)]
        ev_out = [torch.cuda.Event() for _ in range(self.depth)]
        for i in range(n):
            s = i % self.depth
            proj = self.fusion.projection(Ms[i])  # host 3x3 chain (fp32), as the reference does per frame
            with torch.cuda.stream(self.s_in):
                if i >= self.depth:
                    self.s_in.wait_event(ev_done[s])  # slot's previous frame no longer reads its input
                self.feat[s].copy_(feats_pinned[i], non_blocking=True)
                self.proj[s].copy_(proj, non_blocking=True)
                ev_in[s].record(self.s_in)
            self.compute.wait_event(ev_in[s])
            if i >= self.depth:
                self.compute.wait_event(ev_out[s])  # slot's previous result has been copied out
            self.step(s)
            ev_done[s].record(self.compute)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_done[s])
                outs_pinned[i].copy_(self.out[s], non_blocking=True)
                ev_out[s].record(self.s_out)
        self.s_out.synchronize()
