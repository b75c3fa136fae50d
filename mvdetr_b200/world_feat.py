"""Host-side mirror of the callers of the hot path: MSDeformAttn -> DeformableTransformerEncoder(Layer) ->
DeformTransWorldFeat, with the reference's constructor signatures, forward signatures and state_dict keys, so a
reference checkpoint loads unchanged (SURVEY 5 "checkpoint").

  MSDeformAttn                       ref: multiview_detector/models/ops/modules/ms_deform_attn.py:31-117
  DeformableTransformerEncoderLayer  ref: multiview_detector/models/deformable_transformer.py:55-85
  DeformableTransformerEncoder       ref: multiview_detector/models/deformable_transformer.py:22-52
  DeformTransWorldFeat               ref: multiview_detector/models/trans_world_feat.py:70-119
  create_pos_embedding               ref: multiview_detector/models/trans_world_feat.py:15-37

What differs from the reference is only where time went for no reason (SURVEY 3.1 "hot spots (host)"):
  * pos_embedding, reference_points, spatial_shapes, level_start_index live on the device as non-persistent
    buffers (the reference re-uploads 16.9 MB + 5.5 MB per frame: deformable_transformer.py:48,
    trans_world_feat.py:93-95) and the device->host assert of ms_deform_attn.py:94 is done on host ints;
  * without autograd, the elementwise tail of MSDeformAttn.forward (softmax over L*P, loc = ref + off/(W,H),
    ms_deform_attn.py:100-107) runs inside the CUDA kernel (ops.msda_fused_forward) reading the compact
    [Hd*Wd, L, P, 2] reference table instead of its num_cam-fold repeat (mvdetr.py:130);
  * the whole forward is CUDA-graph capturable (no host syncs, no per-call host->device copies).
Dense layers (Linear / LayerNorm / Conv2d) stay cuBLAS / cuDNN through torch -- out of the hot-path scope.
"""
import copy
import math
import weakref

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import constant_, normal_, xavier_uniform_

from . import ops


def create_pos_embedding(img_size, num_pos_feats=64, temperature=10000, normalize=True, scale=None):
    """Sine position embedding [1, 2*num_pos_feats, H, W], bit-identical to the table the reference builds once in its
    constructor (ref: multiview_detector/models/trans_world_feat.py:15-37): channel c < num_pos_feats encodes the ROW,
    the others the COLUMN; 1-based coordinates are scaled to (0, scale], divided by temperature^(2*floor(i/2)/F) and
    passed through sin (even i) / cos (odd i)."""
    if scale is not None and not normalize:
        raise ValueError("normalize should be True if scale is passed")
    scale = 2 * math.pi if scale is None else scale
    H, W = int(img_size[0]), int(img_size[1])
    F_ = int(num_pos_feats)
    i = torch.arange(F_, dtype=torch.float32)
    period = temperature ** (2 * torch.div(i, 2, rounding_mode="floor") / F_)

    def axis_code(n):
        coord = torch.arange(1, n + 1, dtype=torch.float32)        # what a cumulative sum of ones gives
        if normalize:
            coord = coord / (coord[-1] + 1e-6) * scale
        phase = coord[:, None] / period                             # [n, F]
        code = torch.empty_like(phase)
        code[:, 0::2] = phase[:, 0::2].sin()
        code[:, 1::2] = phase[:, 1::2].cos()
        return code                                                 # [n, F]

    rows = axis_code(H)[:, None, :].expand(H, W, F_)
    cols = axis_code(W)[None, :, :].expand(H, W, F_)
    # stored channels-last and viewed as [1, 2F, H, W], like the reference's table: flatten(2).transpose(1, 2) of it
    # is then a contiguous [1, H*W, 2F] (DeformTransWorldFeat.forward views it)
    return torch.cat((rows, cols), -1).unsqueeze(0).permute(0, 3, 1, 2)


class LevelGeometry:
    """Host + device copies of (spatial_shapes, level_start_index) made once, so no per-frame upload or readback."""

    def __init__(self, shapes_hw, device):
        self.hw = [(int(h), int(w)) for h, w in shapes_hw]
        self.len_in = sum(h * w for h, w in self.hw)
        self.shapes = torch.as_tensor(self.hw, dtype=torch.long, device=device)
        starts = np.concatenate([[0], np.cumsum([h * w for h, w in self.hw])[:-1]])
        self.start = torch.as_tensor(starts, dtype=torch.long, device=device)
        self.uniform = all(s == self.hw[0] for s in self.hw)


class MSDeformAttn(nn.Module):
    """Parameter names, shapes and initialisation of ref ops/modules/ms_deform_attn.py:31-77 (checkpoints load as is)."""

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads:
            raise ValueError(f"d_model must be divisible by n_heads, but got {d_model} and {n_heads}")
        self.im2col_step = 64
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        samples = n_heads * n_levels * n_points
        for name, width in (("sampling_offsets", 2 * samples), ("attention_weights", samples),
                            ("value_proj", d_model), ("output_proj", d_model)):
            setattr(self, name, nn.Linear(d_model, width))
        self._reset_parameters()

    def _reset_parameters(self):
        """Offsets start as a ring: head m looks along angle 2*pi*m/M (scaled so the larger component is 1), point k
        sits k + 1 pixels out, the same for every level; attention logits start at zero; projections Xavier."""
        M, L, P = self.n_heads, self.n_levels, self.n_points
        angle = torch.arange(M, dtype=torch.float32) * (2.0 * math.pi / M)
        ray = torch.stack([angle.cos(), angle.sin()], -1)
        ray = ray / ray.abs().max(-1, keepdim=True)[0]
        steps = torch.arange(1, P + 1, dtype=torch.float32)
        ring = (ray[:, None, None, :] * steps[None, None, :, None]).expand(M, L, P, 2)
        with torch.no_grad():
            self.sampling_offsets.weight.zero_()
            self.sampling_offsets.bias = nn.Parameter(ring.reshape(-1).clone())
            self.attention_weights.weight.zero_()
            self.attention_weights.bias.zero_()
            for proj in (self.value_proj, self.output_proj):
                xavier_uniform_(proj.weight)
                proj.bias.zero_()

    def offsets_and_logits(self, q2, N, Len_q, packed=True):
        """Raw (bias-free) outputs of sampling_offsets and attention_weights for the query rows q2 [N*Len_q, C]:
        -> ([N,Len_q,M,L,P,2], [N,Len_q,M,L*P]). packed: ONE GEMM over the concatenated weights (the query rows are read
        and split once, one launch less); the two results are column ranges of its output, which the view-grid MSDA
        kernel reads in place. ref: ms_deform_attn.py:100-101."""
        M, L, P = self.n_heads, self.n_levels, self.n_points
        no, nl = M * L * P * 2, M * L * P
        if not packed:
            return (ops.linear(q2, self.sampling_offsets.weight).view(N, Len_q, M, L, P, 2),
                    ops.linear(q2, self.attention_weights.weight).view(N, Len_q, M, L * P))
        ws = (self.sampling_offsets.weight, self.attention_weights.weight)
        key = getattr(self, "_ol_key", None)
        if not (key is not None and all(r() is w and v == w._version for (r, v), w in zip(key, ws))):
            self._ol_w = torch.cat([w.detach() for w in ws], 0).contiguous()   # [M*L*P*3, C]
            self._ol_key = [(weakref.ref(w), w._version) for w in ws]
        both = ops.linear(q2, self._ol_w)                                       # [N*Len_q, M*L*P*3]
        return both[:, :no].view(N, Len_q, M, L, P, 2), both[:, no:].view(N, Len_q, M, L * P)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None, geometry=None, ref_table=None, ref_table_lm=None, defer_output_bias=False):
        """Reference signature (first six arguments). Extensions used by our encoder: `geometry` (LevelGeometry,
        avoids the device->host check), `ref_table` ([Lr,L,P,2] compact reference points: enables the fused
        kernel when autograd is off) and `defer_output_bias` (fused path only: returns output_proj WITHOUT its bias,
        which the caller adds inside the residual+LayerNorm kernel)."""
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        if geometry is not None:
            assert geometry.len_in == Len_in
        else:
            assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in
        M, L, P = self.n_heads, self.n_levels, self.n_points

        # the fused kernels are instantiated for power-of-two head dims; any other D (e.g. hidden 192 / 8 heads = 24)
        # takes the reference's own arithmetic below through the 6-argument op, which accepts every D
        fused = (ref_table is not None and not torch.is_grad_enabled() and input_flatten.dtype == torch.float32 and
                 input_flatten.is_cuda and (self.d_model // self.n_heads) in (4, 8, 16, 32, 64, 128))
        if fused:
            value = ops.linear(input_flatten.reshape(N * Len_in, -1), self.value_proj.weight, self.value_proj.bias)
        else:
            value = self.value_proj(input_flatten)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(N, Len_in, M, self.d_model // M)
        assert fused or not defer_output_bias, "defer_output_bias is only valid on the fused inference path"
        if fused:
            # both Linear layers as bias-free GEMMs (ops.linear: tensor-core fp32 emulation when the toolkit's cuBLASLt
            # is available); the biases are added inside the MSDA kernel before the same arithmetic as the reference
            q2 = query.reshape(N * Len_q, -1)
            grid_hw = geometry.hw[0] if geometry is not None and geometry.uniform else None
            sampling_offsets, attention_weights = self.offsets_and_logits(q2, N, Len_q, packed=grid_hw is not None)
            output = ops.msda_fused_forward(value.contiguous(), input_spatial_shapes, input_level_start_index,
                                            sampling_offsets, attention_weights, ref_table,
                                            grid_hw=grid_hw, ref_table_lm=ref_table_lm,
                                            off_bias=self.sampling_offsets.bias, logit_bias=self.attention_weights.bias)
            if defer_output_bias:
                return ops.linear(output.view(N * Len_q, -1), self.output_proj.weight).view(N, Len_q, -1)
            return self.output_proj(output)

        # training / autograd path: the reference module's arithmetic (ms_deform_attn.py:100-117) through the 6-argument op
        if reference_points.shape[-1] != 2:
            raise ValueError(f"Last dim of reference_points must be 2, but get {reference_points.shape[-1]} instead.")
        if reference_points.dim() == 4:  # upstream Deformable-DETR layout [N, Lq, L, 2]: one point per level
            reference_points = reference_points.unsqueeze(3)
        wh = input_spatial_shapes.flip(-1)                                              # (W_l, H_l) per level
        offsets = self.sampling_offsets(query).view(N, Len_q, M, L, P, 2)
        locations = reference_points.unsqueeze(2) + offsets / wh[None, None, None, :, None, :]
        weights = F.softmax(self.attention_weights(query).view(N, Len_q, M, L * P), -1).view(N, Len_q, M, L, P)
        output = ops.MSDeformAttnFunction.apply(value.contiguous(), input_spatial_shapes, input_level_start_index,
                                                locations.contiguous(), weights.contiguous(), self.im2col_step)
        return self.output_proj(output)


class DeformableTransformerEncoderLayer(nn.Module):
    """Sub-module names of ref deformable_transformer.py:55-73: self_attn, norm1, linear1, linear2, norm2 (+ dropouts)."""

    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        self.self_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.linear1, self.linear2 = nn.Linear(d_model, d_ffn), nn.Linear(d_ffn, d_model)
        self.norm1, self.norm2 = nn.LayerNorm(d_model), nn.LayerNorm(d_model)
        self.dropout1, self.dropout2, self.dropout3 = (nn.Dropout(dropout) for _ in range(3))

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward(self, src, pos, reference_points, spatial_shapes, level_start_index, padding_mask=None,
                geometry=None, ref_table=None, ref_table_lm=None, perm_inner=0, query=None, emit_query=False):
        """perm_inner > 0 (inference fast path only): the layer's output rows leave cell-major, see add_layer_norm.
        query: src + pos already formed by the previous layer's LayerNorm kernel; emit_query (inference fast path):
        returns (out, out + pos), the next layer's query, from this layer's last LayerNorm kernel."""
        fast = (not torch.is_grad_enabled() and not self.training and src.is_cuda and src.dtype == torch.float32
                and src.shape[-1] % 4 == 0 and src.shape[-1] <= 1024)
        defer = fast and ref_table is not None and (src.shape[-1] // self.self_attn.n_heads) in (4, 8, 16, 32, 64, 128)
        src2 = self.self_attn(query if query is not None else self.with_pos_embed(src, pos), reference_points, src,
                              spatial_shapes,
                              level_start_index, padding_mask, geometry=geometry, ref_table=ref_table,
                              ref_table_lm=ref_table_lm, defer_output_bias=defer)
        if fast:
            # inference: residual (+ deferred Linear bias) + LayerNorm in one kernel, bias+ReLU in the GEMM epilogue,
            # linear2 bias-free with its bias folded into the second LayerNorm kernel; same arithmetic
            src = ops.add_layer_norm(src.contiguous(), src2.contiguous(), self.norm1.weight, self.norm1.bias,
                                     self.norm1.eps, res_bias=self.self_attn.output_proj.bias if defer else None)
            hidden = ops.linear(src.view(-1, src.shape[-1]), self.linear1.weight, self.linear1.bias, relu=True)
            src2 = ops.linear(hidden, self.linear2.weight).view(src.shape)
            if emit_query and perm_inner == 0 and pos is not None:
                return ops.add_layer_norm(src, src2, self.norm2.weight, self.norm2.bias, self.norm2.eps,
                                          res_bias=self.linear2.bias, pos=pos.expand_as(src).contiguous())
            out = ops.add_layer_norm(src, src2, self.norm2.weight, self.norm2.bias, self.norm2.eps,
                                     res_bias=self.linear2.bias, perm_inner=perm_inner)
            return (out, None) if emit_query else out
        assert perm_inner == 0, "perm_inner is only valid on the inference fast path"
        src = self.norm1(src + self.dropout1(src2))
        src2 = self.linear2(self.dropout2(F.relu(self.linear1(src))))
        out = self.norm2(src + self.dropout3(src2))
        return (out, None) if emit_query else out


class DeformableTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers, reference_points=None):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        # reference keeps this as a plain CPU attribute and uploads it every call (deformable_transformer.py:27,48)
        self.register_buffer("reference_points", reference_points, persistent=False)
        self.register_buffer("ref_table", None, persistent=False)
        self.register_buffer("ref_table_lm", None, persistent=False)  # level-major copy [L, rows, P, 2]

    def set_compact_table(self, rows):
        """If reference_points is `k` identical copies of its first `rows` rows (MVDeTr: mvdetr.py:130), keep the
        compact [rows, L, P, 2] table for the fused kernel."""
        rp = self.reference_points
        if rp is None or rp.dim() != 4 or rp.shape[0] % rows != 0:
            return False
        if not torch.equal(rp.view(-1, rows, *rp.shape[1:]), rp[:rows].unsqueeze(0).expand(rp.shape[0] // rows, -1, -1,
                                                                                           -1, -1)):
            return False
        self.ref_table = rp[:rows].contiguous().float()
        self.ref_table_lm = self.ref_table.permute(1, 0, 2, 3).contiguous()
        return True

    @staticmethod
    def get_reference_points(spatial_shapes_hw, valid_ratios, device):
        """Upstream Deformable-DETR fallback (no table given): pixel centres of every level, normalised by the valid
        part of the level, then expressed in every level's valid ratio -> [B, sum(H*W), L, 2]
        (ref: deformable_transformer.py:29-41)."""
        per_level = []
        for lvl, (H_, W_) in enumerate(spatial_shapes_hw):
            cy = torch.linspace(0.5, H_ - 0.5, H_, dtype=torch.float32, device=device)
            cx = torch.linspace(0.5, W_ - 0.5, W_, dtype=torch.float32, device=device)
            gy, gx = torch.meshgrid(cy, cx, indexing="ij")
            x = gx.reshape(1, -1) / (valid_ratios[:, None, lvl, 0] * W_)
            y = gy.reshape(1, -1) / (valid_ratios[:, None, lvl, 1] * H_)
            per_level.append(torch.stack((x, y), -1))
        return torch.cat(per_level, 1)[:, :, None] * valid_ratios[:, None]

    def forward(self, src, spatial_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None,
                geometry=None, perm_inner_last=0, query0=None):
        output = src
        if self.reference_points is None:
            hw = geometry.hw if geometry is not None else spatial_shapes.tolist()
            reference_points = self.get_reference_points(hw, valid_ratios, device=src.device)
        else:
            reference_points = self.reference_points.unsqueeze(0).expand(src.shape[0], -1, -1, -1, -1)
        query = query0  # src + pos when the producer of src formed it (fast path), else the layer adds it
        for i, layer in enumerate(self.layers):
            last = i == self.num_layers - 1
            res = layer(output, pos, reference_points, spatial_shapes, level_start_index, padding_mask,
                        geometry=geometry, ref_table=self.ref_table, ref_table_lm=self.ref_table_lm,
                        perm_inner=perm_inner_last if last else 0, query=query, emit_query=not last)
            output, query = res if not last else (res, None)
        return output


class DeformTransWorldFeat(nn.Module):
    def __init__(self, num_cam, Rworld_shape, base_dim, hidden_dim=128, dropout=0.1, nhead=8, dim_feedforward=512,
                 n_points=4, stride=2, reference_points=None):
        super().__init__()
        self.num_cam, self.hidden_dim, self.stride = num_cam, hidden_dim, stride
        self.Rworld_shape = [int(v) for v in Rworld_shape]
        self.downsample = nn.Sequential(nn.Conv2d(base_dim, hidden_dim, 3, stride, 1), nn.ReLU())
        encoder_layer = DeformableTransformerEncoderLayer(hidden_dim, dim_feedforward, dropout, n_levels=num_cam,
                                                          n_heads=nhead, n_points=n_points)
        self.encoder = DeformableTransformerEncoder(encoder_layer, 3, reference_points)
        self.register_buffer("pos_embedding",
                             create_pos_embedding(np.array(self.Rworld_shape) // stride, hidden_dim // 2),
                             persistent=False)
        self.lvl_embedding = nn.Parameter(torch.Tensor(num_cam, hidden_dim))
        self.merge_linear = nn.Sequential(nn.Conv2d(hidden_dim * num_cam, hidden_dim, 1), nn.ReLU())
        self.upsample = nn.Sequential(nn.Upsample(self.Rworld_shape, mode="bilinear", align_corners=False),
                                      nn.Conv2d(hidden_dim, hidden_dim, 3, 1, 1), nn.ReLU())
        self._reset_parameters()
        self._geometry = None
        Hd, Wd = self.pos_embedding.shape[-2:]
        self.encoder.set_compact_table(Hd * Wd)

    def _reset_parameters(self):
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)
        for m in self.modules():
            if isinstance(m, MSDeformAttn):
                m._reset_parameters()
        normal_(self.lvl_embedding)

    def _level_geometry(self, N, H, W, device):
        g = self._geometry
        if g is None or g.shapes.device != device or g.hw != [(H, W)] * N:
            g = self._geometry = LevelGeometry([(H, W)] * N, device)
        return g

    # ------------------------------------------------------------------------------------------------------------
    # Inference fast path: both 3x3 convolutions and the 1x1 merge as Linear GEMMs over im2col / cell-major rows
    # (csrc/im2col.cu). Same arithmetic as forward() up to fp32 summation order; used by MultiviewFusion.fuse and
    # ShardedFusion when autograd is off.
    # ------------------------------------------------------------------------------------------------------------
    def gemm_weights(self):
        """Conv weights reshaped for the GEMM path ((ky, kx, c_in) column order), cached until a weight changes."""
        ws = [c.weight for c in (self.downsample[0], self.merge_linear[0], self.upsample[1])]
        key = getattr(self, "_gw_key", None)  # (weak refs to the source Parameters, their versions): object identity,
        fresh = key is not None and all(r() is w and v == w._version for (r, v), w in zip(key, ws))  # never addresses
        if not fresh:
            self._gw = [w.detach().permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous() for w in ws]
            self._gw_key = [(weakref.ref(w), w._version) for w in ws]
        return self._gw

    def fast_path_ok(self, x):
        return (not torch.is_grad_enabled() and not self.training and x.is_cuda and x.dtype == torch.float32 and
                self.hidden_dim % 4 == 0 and x.shape[1] % 4 == 0 and self.stride == 2 and
                self.encoder.ref_table is not None)

    def tokens_from_warped(self, g_cl, with_query=False):
        """g_cl [N, Hg, Wg, C_in] channels-last warped grid -> downsample conv (3x3, stride 2) + ReLU as an implicit GEMM
        (no im2col matrix) -> [tokens, hidden], or None when the library cannot take the shape. with_query: returns
        (tokens, tokens + pos): the first encoder layer's query leaves the same epilogue."""
        Wd, _, _ = self.gemm_weights()
        if not with_query:
            return ops.conv3x3_nhwc(g_cl, Wd, self.downsample[0].bias, stride=2, relu=True)
        N, Hg, Wg = g_cl.shape[0], g_cl.shape[1], g_cl.shape[2]
        pos = self._pos_rows(1, N, (Hg - 1) // 2 + 1, (Wg - 1) // 2 + 1)
        return ops.conv3x3_nhwc(g_cl, Wd, self.downsample[0].bias, stride=2, relu=True, add=pos)

    def _pos_rows(self, B, N, Hd, Wd):
        """Position + level embedding rows [B, N*Hd*Wd, C] (static at inference; cached per embedding version)."""
        C = self.hidden_dim
        key = (id(self.lvl_embedding), self.lvl_embedding._version, id(self.pos_embedding), self.pos_embedding.device)
        if getattr(self, "_pos_key", None) != key:
            self._pos = (self.pos_embedding.flatten(2).transpose(1, 2).unsqueeze(1) +
                         self.lvl_embedding.detach().view([B, N, 1, C])).view([B, N * Hd * Wd, C]).contiguous()
            self._pos_key = key
        return self._pos

    def tokens_from_im2col(self, A):
        """A [tokens, 9*C_in] (ops.warp_im2col) -> downsample conv + ReLU as one GEMM -> [tokens, hidden]."""
        Wd, _, _ = self.gemm_weights()
        return ops.linear(A, Wd, self.downsample[0].bias, relu=True)

    def encode_tokens(self, src, N, Hd, Wd, perm_inner_last=0, query0=None):
        """src [1, N*Hd*Wd, hidden] view-major tokens -> encoder output (cell-major rows when perm_inner_last). query0:
        src + pos when the producer of src already formed it."""
        B, _, C = src.shape
        pos = self._pos_rows(B, N, Hd, Wd)
        geo = self._level_geometry(N, Hd, Wd, src.device)
        return self.encoder(src, geo.shapes, geo.start, None, pos, geometry=geo, perm_inner_last=perm_inner_last,
                            query0=query0)

    def tail_from_cell_major(self, mem_cm, Hd, Wd):
        """mem_cm [Hd*Wd, N*hidden] (row = ground cell, columns = (view, channel)) -> merge 1x1 conv + ReLU, bilinear
        upsample, 3x3 conv + ReLU -> [1, hidden, Hg, Wg] (contiguous NCHW like the reference)."""
        _, Wm, Wu = self.gemm_weights()
        C = self.hidden_dim
        Hg, Wg = self.Rworld_shape
        merged = ops.linear(mem_cm, Wm, self.merge_linear[0].bias, relu=True)            # [cells, C] = NHWC map
        out_cl = None
        if ops.conv3x3_implicit_ok(C, Wu.shape[0]):  # upsample -> implicit-GEMM conv: no [Hg*Wg, 9C] matrix
            up = ops.upsample_nhwc(merged.view(1, Hd, Wd, C), (Hg, Wg))
            out_cl = ops.conv3x3_nhwc(up, Wu, self.upsample[1].bias, stride=1, relu=True)
        if out_cl is None:
            A = ops.upsample_im2col(merged.view(1, Hd, Wd, C), (Hg, Wg))                  # [Hg*Wg, 9C]
            out_cl = ops.linear(A, Wu, self.upsample[1].bias, relu=True)                 # [Hg*Wg, C]
        return ops.transpose_last2(out_cl.view(1, Hg * Wg, C)).view(1, C, Hg, Wg)

    def forward_from_tokens(self, tokens, N, Hd, Wd, query0=None):
        """Whole stage from the downsample conv's output tokens [N*Hd*Wd, hidden] (B = 1): -> [1, hidden, Hg, Wg]."""
        src = tokens.view(1, N * Hd * Wd, self.hidden_dim)
        if query0 is not None:
            query0 = query0.view_as(src)
        mem_cm = self.encode_tokens(src, N, Hd, Wd, perm_inner_last=Hd * Wd, query0=query0)
        return self.tail_from_cell_major(mem_cm.view(Hd * Wd, N * self.hidden_dim), Hd, Wd)

    def forward_from_im2col(self, A, N, Hd, Wd):
        """Whole stage from the warp's im2col matrix (B = 1, as the reference): -> [1, hidden, Hg, Wg]."""
        src = self.tokens_from_im2col(A).view(1, N * Hd * Wd, self.hidden_dim)
        mem_cm = self.encode_tokens(src, N, Hd, Wd, perm_inner_last=Hd * Wd)
        return self.tail_from_cell_major(mem_cm.view(Hd * Wd, N * self.hidden_dim), Hd, Wd)

    def forward(self, x, visualize=False):
        """x [B, N, C, H, W] (any strides; channels-last per view avoids the reference's permute-copy at
        trans_world_feat.py:92) -> [B, hidden, H, W]. As in the reference only B=1 is valid (its
        lvl_embedding.view([B, N, 1, C]) at :94)."""
        B, N, C, H, W = x.shape
        x = self.downsample(x.reshape(B * N, C, H, W))
        _, _, H, W = x.shape
        src_flatten = x.view(B, N, C, H, W).permute(0, 1, 3, 4, 2).reshape(B, N * H * W, C)
        lvl_pos_embed_flatten = (self.pos_embedding.flatten(2).transpose(1, 2).unsqueeze(1) +
                                 self.lvl_embedding.view([B, N, 1, C])).view([B, N * H * W, C])
        geo = self._level_geometry(N, H, W, x.device)
        valid_ratios = None if self.encoder.reference_points is not None else torch.ones([B, N, 2], device=x.device)
        memory = self.encoder(src_flatten, geo.shapes, geo.start, valid_ratios, lvl_pos_embed_flatten, geometry=geo)
        merged_feat = self.merge_linear(memory.view(B, N, H, W, C).permute(0, 1, 4, 2, 3).reshape(B, N * C, H, W))
        return self.upsample(merged_feat)


def from_reference(ref_world_feat):
    """Wraps an instance of the UNMODIFIED reference `DeformTransWorldFeat` (ref: trans_world_feat.py:70-119): returns our
    mirror built from the reference module's own hyper-parameters and SHARING its Parameter objects (no copy), so
    `model.world_feat = from_reference(model.world_feat)` accelerates an existing reference model in place -- training
    keeps updating the same tensors, checkpoints keep their keys."""
    conv = ref_world_feat.downsample[0]
    layer = ref_world_feat.encoder.layers[0]
    attn = layer.self_attn
    num_cam, hidden = ref_world_feat.lvl_embedding.shape
    size = ref_world_feat.upsample[0].size
    rp = getattr(ref_world_feat.encoder, "reference_points", None)
    ours = DeformTransWorldFeat(num_cam, [int(v) for v in size], conv.in_channels, hidden_dim=hidden,
                                dropout=layer.dropout1.p, nhead=attn.n_heads, dim_feedforward=layer.linear1.out_features,
                                n_points=attn.n_points, stride=conv.stride[0], reference_points=rp)
    for name, param in ref_world_feat.named_parameters():
        owner = ours
        *path, leaf = name.split(".")
        for part in path:
            owner = getattr(owner, part)
        owner._parameters[leaf] = param
    return ours.to(ref_world_feat.lvl_embedding.device).train(ref_world_feat.training)
