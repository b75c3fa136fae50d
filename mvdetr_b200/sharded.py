"""View-sharded multi-GPU execution of the fusion stage: one process per GPU, camera views split across ranks.

What shards: backbone features, the perspective warp and the stride-2 downsample conv are independent per
view (the reference folds views into the batch dim: multiview_detector/models/mvdetr.py:153,177-178,194;
trans_world_feat.py:89), and so are the encoder's per-token Linear / LayerNorm / FFN. Views couple only through the
deformable attention's `value` (every query samples all L = num_cam levels). Hence, per frame, on rank r owning the
contiguous view range [v0, v1):

    warp(views v0..v1) -> downsample conv -> tokens src_r [(v1-v0)*Hd*Wd, C]
    for each of the 3 encoder layers:
        value_r = value_proj(src_r)                       (local rows only)
        value   = ALL-GATHER(value_r)   <-- the one exchange step per layer (5.5 MB / view at Wildtrack size)
        src_r   = layer(src_r, queries of the local views sampling the full `value`)
    memory = ALL-GATHER(src_r)                            (for the merge conv over all views)
    tail, sharded by ROWS of the ground grid: every rank runs the (cheap) 1x1 merge over all cells, then upsample +
    3x3 conv for its own band of output rows only (the band's halo is computed, not exchanged), and one ALL-GATHER of the
    [rows, C] bands puts the fused feature on every rank (rank 0's copy is the one the caller reads)

Inside each layer the exchange of the value rows runs on a side stream while the compute stream runs the
sampling-offset / attention-logit GEMMs of the local queries (fork/join with events; both streams are captured into the
same CUDA graph).

The first all-gather is the north star's "all-gather of warped world-grid features before the transformer" (the
warped features, after the per-view conv and the per-token value projection, both of which commute with the gather).
Results equal the single-GPU path up to the reduction order inside cuBLAS (row partitions of the same GEMMs).

Views do not have to divide the world size: ranks take ceil(N/world) views each, trailing ranks may own fewer or
none; gather buffers are padded to the maximum and the valid rows form a contiguous prefix (no compaction copy).

`gather_rows` runs on any torch.distributed backend (NCCL on the GPUs; gloo in the CPU tests of the host logic).
"""
import os

import torch
import torch.distributed as dist

from . import ops


class ViewPartition:
    """Contiguous split of `num_views` over `world` ranks: rank r owns views [lo(r), hi(r))."""

    def __init__(self, num_views, world):
        if num_views <= 0 or world <= 0:
            raise ValueError("num_views and world must be positive")
        self.num_views, self.world = int(num_views), int(world)
        self.per_rank = -(-self.num_views // self.world)  # ceil

    def lo(self, rank):
        return min(self.num_views, rank * self.per_rank)

    def hi(self, rank):
        return min(self.num_views, (rank + 1) * self.per_rank)

    def count(self, rank):
        return self.hi(rank) - self.lo(rank)

    def active_ranks(self):
        return [r for r in range(self.world) if self.count(r) > 0]


def gather_rows(local_rows, partition, rows_per_view, rank, out=None, group=None):
    """All-gather of per-view row blocks. local_rows [count(rank)*rows_per_view, C] -> [num_views*rows_per_view, C]
    (a view of the padded gather buffer `out` [world, per_rank*rows_per_view, C], allocated when None)."""
    C = local_rows.shape[-1]
    pad_rows = partition.per_rank * rows_per_view
    if out is None:
        out = local_rows.new_empty((partition.world, pad_rows, C))
    mine = out[rank]
    n = local_rows.shape[0]
    if mine.data_ptr() != local_rows.data_ptr():
        mine[:n].copy_(local_rows)
    if partition.world > 1:
        dist.all_gather_into_tensor(out.view(-1), mine.reshape(-1), group=group)
    return out.view(partition.world * pad_rows, C)[:partition.num_views * rows_per_view]


class SymmetricBuffers:
    """Gather buffers in NVLink symmetric memory (torch.distributed._symmetric_memory): every rank allocates the same
    layout, the rendezvous maps the peers' copies and -- on NVSwitch systems -- ONE multicast address per buffer whose
    stores the switch replicates into every GPU's copy. With them the all-gather is not a collective call any more:
    the producing kernel's epilogue stores straight to the multicast address (ops.linear_multicast / multicast_copy)
    and a barrier kernel (signal pads in the same symmetric memory) separates producers from consumers.
    `available` is False when symmetric memory or multicast cannot be set up; callers then use NCCL."""

    def __init__(self, device, group=None):
        self.device, self.group = device, group
        self.bufs, self.available, self.why = {}, False, ""
        if os.environ.get("MVDETR_B200_FUSED_GATHER", "1") == "0":
            self.why = "disabled by MVDETR_B200_FUSED_GATHER=0"
            return
        try:
            import torch.distributed._symmetric_memory as symm
            self.symm = symm
            g = group if group is not None else dist.group.WORLD
            self.group_name = g.group_name
            probe = symm.empty(1024, dtype=torch.float32, device=device)
            hdl = symm.rendezvous(probe, self.group_name)
            if not hdl.multicast_ptr:
                self.why = "no NVLink multicast support on this system"
                return
            self.available = True
        except Exception as e:  # symmetric memory unavailable in this build / on this fabric
            self.why = f"{type(e).__name__}: {e}"[:200]

    def get(self, key, shape):
        """-> (local tensor [shape], multicast base address, handle); allocated and rendezvoused once per key."""
        ent = self.bufs.get(key)
        if ent is None:
            t = self.symm.empty(tuple(shape), dtype=torch.float32, device=self.device)
            t.zero_()
            hdl = self.symm.rendezvous(t, self.group_name)
            ent = self.bufs[key] = (t, int(hdl.multicast_ptr), hdl)
        return ent


class ShardedFusion:
    """Inference-only view-sharded forward over a MultiviewFusion's parameters (every rank holds the full, identical
    parameter set; only activations are sharded)."""

    def __init__(self, fusion, rank, world, group=None):
        self.fusion = fusion
        self.wf = fusion.world_feat
        self.rank, self.world, self.group = rank, world, group
        self.part = ViewPartition(fusion.num_cam, world)
        self.v0, self.v1 = self.part.lo(rank), self.part.hi(rank)
        self._bufs = {}
        self._side = None       # side stream for the GEMMs that overlap the value all-gather
        self.shard_tail = True  # False: every rank computes the whole tail (round-1 behaviour, A/B switch)
        self.symm = None        # SymmetricBuffers once the first CUDA frame runs (fused multicast all-gather)

    def fused_gather(self, device):
        """True when the all-gathers run as multicast stores from our kernels' epilogues instead of ncclAllGather."""
        if self.symm is None:
            self.symm = SymmetricBuffers(device, self.group) if (self.world > 1 and device.type == "cuda" and
                                                                 ops._GEMM_MODE in ("bf16x3", "f16x2")) else False
        return bool(self.symm) and self.symm.available

    def _buf(self, key, shape, like):
        b = self._bufs.get(key)
        if b is None or tuple(b.shape) != tuple(shape) or b.device != like.device:
            b = self._bufs[key] = torch.zeros(shape, dtype=like.dtype, device=like.device)
        return b

    def tokens(self, feat_local, proj_local):
        """warp + downsample of the local views -> [nv*Hd*Wd, C] token rows (and Hd, Wd)."""
        wf = self.wf
        Hg, Wg = self.fusion.Rworld_shape
        nv = feat_local.shape[0]
        C = wf.hidden_dim
        if nv == 0:
            Hd, Wd = wf.pos_embedding.shape[-2:]
            return feat_local.new_zeros((0, C)), Hd, Wd
        if self.fusion.gemm_path and wf.fast_path_ok(feat_local):
            if ops.conv3x3_implicit_ok(feat_local.shape[1], C):  # same route (and bits) as the single-GPU frame
                g_cl = ops.warp_perspective(feat_local, proj_local, (Hg, Wg), align_corners=False, channels_last=True)
                t = wf.tokens_from_warped(g_cl)
                if t is not None:
                    return t, (Hg - 1) // 2 + 1, (Wg - 1) // 2 + 1
            A, (Hd, Wd) = ops.warp_im2col(feat_local, proj_local, (Hg, Wg), stride=2)
            return wf.tokens_from_im2col(A), Hd, Wd
        world = ops.warp_perspective(feat_local, proj_local, (Hg, Wg), align_corners=False, channels_last=True)
        x = wf.downsample(world.permute(0, 3, 1, 2))
        Hd, Wd = x.shape[-2:]
        return x.permute(0, 2, 3, 1).reshape(nv * Hd * Wd, C), Hd, Wd

    def encode(self, src_local, Hd, Wd, slot=0):
        """Query-sharded encoder + replicated merge/upsample. src_local [nv*Hd*Wd, C] -> [1, hidden, Hg, Wg]."""
        wf, part, rank = self.wf, self.part, self.rank
        N, C, hw = self.fusion.num_cam, wf.hidden_dim, Hd * Wd
        S = N * hw
        dev = src_local.device
        geo = wf._level_geometry(N, Hd, Wd, dev)
        table = wf.encoder.ref_table
        if table is None:
            raise RuntimeError("ShardedFusion needs the compact reference table (MVDeTr's create_reference_map layout)")
        nq = src_local.shape[0]
        q0 = self.v0 * hw
        # position + level embedding of the local queries: static at inference, cached per parameter version
        pkey = (id(wf.lvl_embedding), wf.lvl_embedding._version, q0, nq, dev)
        if getattr(self, "_pos_key", None) != pkey:
            self._pos = (wf.pos_embedding.flatten(2).transpose(1, 2).unsqueeze(1) +
                         wf.lvl_embedding.detach().view(1, N, 1, C)).view(S, C)[q0:q0 + nq].contiguous()
            self._pos_key = pkey
        pos = self._pos
        gbuf = self._buf(("gather", slot), (part.world, part.per_rank * hw, C), src_local)
        src = src_local
        cur = torch.cuda.current_stream(dev) if src_local.is_cuda else None
        if cur is not None and self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        fused = self.fused_gather(dev)
        next_query = None
        for li, layer in enumerate(wf.encoder.layers):
            attn = layer.self_attn
            M, L, P = attn.n_heads, attn.n_levels, attn.n_points
            mine = gbuf[rank][:nq]
            # value projection first (main stream); then the exchange of the value rows runs on the side stream while
            # the main stream computes the two query GEMMs, which only need local rows. (The GEMMs are persistent
            # kernels that fill every SM: overlapping two of THEM gains nothing -- r02h -- but the all-gather / barrier
            # kernels need almost no SM resources.)
            if fused:
                # GEMM whose epilogue is the all-gather: value rows leave through the multicast address of this rank's
                # slot and land in every GPU's buffer; the barrier kernel makes them visible to all consumers.
                # Two buffers alternate by layer, so a fast rank's stores for layer l+1 never hit a buffer a slow rank
                # still reads for layer l (it cannot be more than one barrier ahead).
                sbuf, mc, hdl = self.symm.get(("gather", slot, li % 2), (part.world, part.per_rank * hw, C))
                if nq:
                    ops.linear_multicast(src, attn.value_proj.weight, attn.value_proj.bias,
                                         mc + rank * part.per_rank * hw * C * 4)
            elif nq:
                ops.linear(src, attn.value_proj.weight, attn.value_proj.bias, out=mine)  # local rows -> gather buffer

            def exchange():
                if fused:
                    hdl.barrier(channel=0)
                    return sbuf.view(part.world * part.per_rank * hw, C)[:S]
                return gather_rows(mine, part, hw, rank, out=gbuf, group=self.group)

            if cur is not None:
                fork, join = torch.cuda.Event(), torch.cuda.Event()
                fork.record(cur)
                with torch.cuda.stream(self._side):
                    self._side.wait_event(fork)
                    value = exchange()
                    join.record(self._side)
            else:
                value = exchange()
            if nq:
                query = next_query if next_query is not None else src + pos
                # bias-free GEMMs; the biases are applied inside our kernels (world_feat.MSDeformAttn.forward)
                offsets, logits = attn.offsets_and_logits(query, 1, nq)   # one GEMM, column ranges of its output
            if cur is not None:
                cur.wait_event(join)
            if nq:
                out = ops.msda_fused_forward(value.view(1, S, M, C // M), geo.shapes, geo.start, offsets, logits,
                                             table, grid_hw=(Hd, Wd), ref_table_lm=wf.encoder.ref_table_lm,
                                             off_bias=attn.sampling_offsets.bias, logit_bias=attn.attention_weights.bias)
                src2 = ops.linear(out.view(nq, C), attn.output_proj.weight)
                src = ops.add_layer_norm(src, src2, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps,
                                         res_bias=attn.output_proj.bias)
                hidden = ops.linear(src, layer.linear1.weight, layer.linear1.bias, relu=True)
                src2 = ops.linear(hidden, layer.linear2.weight)
                if li + 1 < len(wf.encoder.layers):  # this LayerNorm also emits the next layer's query, src + pos
                    src, next_query = ops.add_layer_norm(src, src2, layer.norm2.weight, layer.norm2.bias,
                                                         layer.norm2.eps, res_bias=layer.linear2.bias, pos=pos)
                else:
                    src = ops.add_layer_norm(src, src2, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps,
                                             res_bias=layer.linear2.bias)
        if fused and self.fusion.gemm_path and wf.fast_path_ok(src_local):
            # the final tokens go out CELL-major (row = ground cell, columns = (view, channel)): every rank receives the
            # merge conv's GEMM operand directly, no permute-copy (trans_world_feat.py:107-108)
            cbuf, mc, hdl = self.symm.get(("cells", slot), (hw, N * C))
            if nq:
                ops.multicast_copy(src.contiguous(), mc, inner=hw, outer_total=N, outer0=self.v0)
            hdl.barrier(channel=0)
            if self.shard_tail:
                return self.sharded_tail(cbuf, Hd, Wd, slot)
            return wf.tail_from_cell_major(cbuf, Hd, Wd)
        if fused:
            n_layers = len(wf.encoder.layers)
            sbuf, mc, hdl = self.symm.get(("gather", slot, n_layers % 2), (part.world, part.per_rank * hw, C))
            if nq:
                ops.multicast_copy(src.contiguous(), mc + rank * part.per_rank * hw * C * 4)
            hdl.barrier(channel=0)
            memory = sbuf.view(part.world * part.per_rank * hw, C)[:S]
        else:
            memory = gather_rows(src, part, hw, rank, out=gbuf, group=self.group)
        if self.fusion.gemm_path and wf.fast_path_ok(memory):
            mem_cm = memory.view(N, hw, C).permute(1, 0, 2).reshape(hw, N * C)  # view-major -> cell-major rows
            if self.shard_tail and part.world > 1:
                return self.sharded_tail(mem_cm, Hd, Wd, slot)
            return wf.tail_from_cell_major(mem_cm, Hd, Wd)
        merged = wf.merge_linear(memory.view(1, N, Hd, Wd, C).permute(0, 1, 4, 2, 3).reshape(1, N * C, Hd, Wd))
        return wf.upsample(merged)

    def sharded_tail(self, mem_cm, Hd, Wd, slot):
        """merge 1x1 conv over all cells on every rank (one small GEMM), then upsample + 3x3 conv for this rank's band
        of ground-plane rows, all-gather of the bands, NHWC -> NCHW. Same arithmetic per output element as
        DeformTransWorldFeat.tail_from_cell_major (row partitions of the same GEMM)."""
        wf, world, rank = self.wf, self.world, self.rank
        _, Wm, Wu = wf.gemm_weights()
        C = wf.hidden_dim
        Hg, Wg = wf.Rworld_shape
        band = -(-Hg // world)  # ceil: rows per rank; trailing ranks may own fewer or none
        r0, r1 = min(Hg, rank * band), min(Hg, (rank + 1) * band)
        merged = ops.linear(mem_cm.contiguous(), Wm, wf.merge_linear[0].bias, relu=True)   # [cells, C] = NHWC map
        if self.fused_gather(merged.device):
            obuf, mc, hdl = self.symm.get(("tail", slot), (world, band * Wg, C))
            if r1 > r0:  # the band's conv GEMM stores its rows into every GPU's copy of the output
                A = ops.upsample_im2col(merged.view(1, Hd, Wd, C), (Hg, Wg), rows=(r0, r1 - r0))
                ops.linear_multicast(A, Wu, wf.upsample[1].bias, mc + rank * band * Wg * C * 4, relu=True)
            hdl.barrier(channel=0)
        else:
            obuf = self._buf(("tail", slot), (world, band * Wg, C), merged)
            if r1 > r0:
                A = ops.upsample_im2col(merged.view(1, Hd, Wd, C), (Hg, Wg), rows=(r0, r1 - r0))
                ops.linear(A, Wu, wf.upsample[1].bias, relu=True, out=obuf[rank][:(r1 - r0) * Wg])
            dist.all_gather_into_tensor(obuf.view(-1), obuf[rank].reshape(-1), group=self.group)
        out_cl = obuf.view(world * band * Wg, C)[:Hg * Wg]   # bands are contiguous row ranges: valid rows form a prefix
        return ops.transpose_last2(out_cl.view(1, Hg * Wg, C)).view(1, C, Hg, Wg)

    def fuse(self, feat_local, proj_local, slot=0):
        src, Hd, Wd = self.tokens(feat_local, proj_local)
        return self.encode(src, Hd, Wd, slot)


class ShardedFrameRunner:
    """FrameRunner's interface (load / step / run_host_frames) for the view-sharded path. `load` and
    `run_host_frames` take FULL frames and keep this rank's views only; the fused world feature is valid on every rank
    (rank 0's copy is the one the bench reads back)."""

    frames_per_step = 1

    def __init__(self, fusion, feat_shape, device, rank, world, use_graph=True, depth=2, group=None):
        self.device = torch.device(device)
        self.fusion = fusion.to(device).eval()
        self.sf = ShardedFusion(self.fusion, rank, world, group)
        self.rank, self.world, self.depth = rank, world, depth
        self.mode = f"views{fusion.num_cam}_over_{world}ranks_allgather"
        v0, v1 = self.sf.v0, self.sf.v1
        self.feat = [torch.zeros((v1 - v0, *feat_shape[1:]), device=device) for _ in range(depth)]
        self.proj = [torch.eye(3, device=device).repeat(v1 - v0, 1, 1) for _ in range(depth)]
        self.out = [None] * depth
        self.graphs = [None] * depth
        self.compute = torch.cuda.Stream(device=device)
        self.s_in = torch.cuda.Stream(device=device)
        self.s_out = torch.cuda.Stream(device=device)
        with torch.no_grad(), torch.cuda.stream(self.compute):
            for i in range(depth):
                for _ in range(2):
                    self.out[i] = self.sf.fuse(self.feat[i], self.proj[i], slot=i)
            self.compute.synchronize()
            if use_graph:
                try:
                    for i in range(depth):
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g, stream=self.compute):
                            self.out[i] = self.sf.fuse(self.feat[i], self.proj[i], slot=i)
                        self.graphs[i] = g
                except Exception as e:  # NCCL capture unsupported on this stack: stay eager, say so
                    self.graphs = [None] * depth
                    self.mode += f"_eager({type(e).__name__})"
                    torch.cuda.synchronize(device)
        torch.cuda.synchronize(device)
        symm = self.sf.symm
        self.mode += ("_multicast_epilogue" if symm and symm.available else
                      "_nccl" + (f"({symm.why})" if symm else ""))

    def load(self, imgs_feat, proj_mats, slot=0):
        v0, v1 = self.sf.v0, self.sf.v1
        self.feat[slot].copy_(imgs_feat[v0:v1])
        self.proj[slot].copy_(proj_mats[v0:v1])

    def step(self, slot=0):
        with torch.cuda.stream(self.compute):
            if self.graphs[slot] is not None:
                self.graphs[slot].replay()
            else:
                with torch.no_grad():
                    self.out[slot] = self.sf.fuse(self.feat[slot], self.proj[slot], slot=slot)
        return self.out[slot]

    def run_host_frames(self, feats_pinned, Ms, outs_pinned):
        """Same pipeline as FrameRunner.run_host_frames; each rank uploads only its own views (in deployment every
        GPU's host process receives its own cameras) and rank 0 reads the fused result back."""
        v0, v1 = self.sf.v0, self.sf.v1
        n = len(feats_pinned)
        ev_in = [torch.cuda.Event() for _ in range(self.depth)]
        ev_done = [torch.cuda.Event() for _ in range(self.depth)]
        ev_out = [torch.cuda.Event() for _ in range(self.depth)]
        for i in range(n):
            s = i % self.depth
            proj = self.fusion.projection(Ms[i])[v0:v1]
            with torch.cuda.stream(self.s_in):
                if i >= self.depth:
                    self.s_in.wait_event(ev_done[s])
                if v1 > v0:
                    self.feat[s].copy_(feats_pinned[i][v0:v1], non_blocking=True)
                    self.proj[s].copy_(proj, non_blocking=True)
                ev_in[s].record(self.s_in)
            self.compute.wait_event(ev_in[s])
            if i >= self.depth:
                self.compute.wait_event(ev_out[s])
            self.step(s)
            ev_done[s].record(self.compute)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_done[s])
                if self.rank == 0:
                    outs_pinned[i].copy_(self.out[s], non_blocking=True)
                ev_out[s].record(self.s_out)
        self.s_out.synchronize()
        self.compute.synchronize()
