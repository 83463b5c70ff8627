"""Point Transformer V3 scene backbone (the `ptv3_pretraining` config) on B200.

Module mirror of /root/reference/pointcept/models/point_transformer_v3/point_transformer_v3m1_base.py (RPE 32-51,
SerializedAttention 54-225, MLP 228-251, Block 254-341, SerializedPooling 344-447, SerializedUnpooling 450-485, Embedding
488-519, PointTransformerV3 522-779) and of `Point` (pointcept/models/utils/structure.py:13-140), with the same constructor
arguments, sub-module names and parameter layouts, so the reference's checkpoints load.  What runs underneath:

  * `Point.serialization`: z / z-trans codes from `up3d_zorder_keys` (csrc/serialize.cu; hilbert raises);
  * `Point.sparsify` / the xCPE and stem convolutions: `unipre3d_b200.sparse.SubMConv3d` (csrc/sparse_conv.cu) on an
    order-preserving sparse tensor (PTv3 keeps `point.feat` and `sparse_conv_feat.features` row-aligned);
  * `SerializedAttention`: padded patches of K = 48 points; exact-softmax attention (the reference's non-flash branch) in
    the activation dtype, or `enable_flash=True` -> torch's fused scaled-dot-product kernels on the (patches, heads, K, d)
    batch in bf16 (the reference calls flash_attn's varlen kernel in fp16 there);
  * `SerializedPooling`: `torch_scatter.segment_csr` restated with index_reduce / index_add on the sorted clusters.

Quirks kept: after the embedding the reference rebuilds the Point with `offset = [N_total]` (one segment) while `batch`
keeps the scene indices (699-717), so the patch padding treats all scenes as one sequence; pooling shuffles the order list
with `torch.randperm` on the CPU generator.  `PointFusion` (use_fusion) is not built -> raises.
"""
from __future__ import annotations

import math
from functools import partial

import torch
import torch.nn as nn

from .sparse import SparseConvTensor, SubMConv3d


def offset2bincount(offset):
    return torch.diff(offset, prepend=offset.new_zeros(1))


def offset2batch(offset):
    bc = offset2bincount(offset)
    return torch.arange(len(bc), device=offset.device, dtype=torch.long).repeat_interleave(bc)


def batch2offset(batch):
    return torch.cumsum(batch.bincount(), dim=0).long()


def segment_csr(src, indptr, reduce):
    """torch_scatter.segment_csr for sorted segments: out[s] = reduce(src[indptr[s]:indptr[s+1]])."""
    n_seg = indptr.numel() - 1
    counts = torch.diff(indptr)
    seg = torch.repeat_interleave(torch.arange(n_seg, device=src.device), counts)
    if reduce in ("sum", "mean"):
        out = torch.zeros((n_seg,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add_(0, seg, src)
        if reduce == "mean":
            out = out / counts.clamp(min=1).view(-1, *([1] * (src.dim() - 1))).to(src.dtype)
        return out
    init = float("-inf") if reduce == "max" else float("inf")
    out = torch.full((n_seg,) + tuple(src.shape[1:]), init, dtype=src.dtype, device=src.device)
    return out.index_reduce_(0, seg, src, "amax" if reduce == "max" else "amin", include_self=True)


class Point(dict):
    """addict.Dict-style point-cloud record (structure.py:13-140): attribute access on a dict."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if "batch" not in self and "offset" in self:
            self["batch"] = offset2batch(self["offset"])
        elif "offset" not in self and "batch" in self:
            self["offset"] = batch2offset(self["batch"])

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def serialization(self, order=("z",), depth=None, shuffle_orders=False):
        from . import serialization as ser
        assert "batch" in self
        if "grid_coord" not in self:
            assert {"grid_size", "coord"}.issubset(self.keys())
            self["grid_coord"] = torch.div(self.coord - self.coord.min(0)[0], self.grid_size, rounding_mode="trunc").int()
        if depth is None:
            depth = int(self.grid_coord.max()).bit_length()
        self["serialized_depth"] = depth
        assert depth * 3 + len(self.offset).bit_length() <= 63 and depth <= 16
        code = torch.stack([ser.encode(self.grid_coord, self.batch, depth, order=o) for o in order])
        order_ = torch.argsort(code)
        inverse = torch.zeros_like(order_).scatter_(
            dim=1, index=order_, src=torch.arange(0, code.shape[1], device=order_.device).repeat(code.shape[0], 1))
        if shuffle_orders:
            perm = torch.randperm(code.shape[0])
            code, order_, inverse = code[perm], order_[perm], inverse[perm]
        self["serialized_code"], self["serialized_order"], self["serialized_inverse"] = code, order_, inverse

    def sparsify(self, pad=96):
        assert {"feat", "batch"}.issubset(self.keys())
        if "grid_coord" not in self:
            assert {"grid_size", "coord"}.issubset(self.keys())
            self["grid_coord"] = torch.div(self.coord - self.coord.min(0)[0], self.grid_size, rounding_mode="trunc").int()
        indices = torch.cat([self.batch.unsqueeze(-1).int(), self.grid_coord.int()], dim=1).contiguous()
        self["sparse_conv_feat"] = SparseConvTensor(self.feat, indices, None, None, keep_order=True)


class PointModule(nn.Module):
    pass


class PointSequential(PointModule):
    """pointcept/models/modules.py:18-90."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for idx, module in enumerate(args):
            self.add_module(str(idx), module)
        for name, module in kwargs.items():
            self.add_module(name, module)

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        self.add_module(str(len(self._modules)) if name is None else name, module)

    def forward(self, input):
        for module in self._modules.values():
            if isinstance(module, PointModule):
                input = module(input)
            elif isinstance(module, SubMConv3d):
                if isinstance(input, Point):
                    input.sparse_conv_feat = module(input.sparse_conv_feat)
                    input.feat = input.sparse_conv_feat.features
                else:
                    input = module(input)
            else:
                if isinstance(input, Point):
                    input.feat = module(input.feat)
                    if "sparse_conv_feat" in input:
                        input.sparse_conv_feat = input.sparse_conv_feat.replace_feature(input.feat)
                elif isinstance(input, SparseConvTensor):
                    input = input.replace_feature(module(input.features))
                else:
                    input = module(input)
        return input


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask.div_(keep)


class RPE(nn.Module):
    def __init__(self, patch_size, num_heads):
        super().__init__()
        self.patch_size, self.num_heads = patch_size, num_heads
        self.pos_bnd = int((4 * patch_size) ** (1 / 3) * 2)
        self.rpe_num = 2 * self.pos_bnd + 1
        self.rpe_table = nn.Parameter(torch.zeros(3 * self.rpe_num, num_heads))
        nn.init.trunc_normal_(self.rpe_table, std=0.02)

    def forward(self, coord):
        idx = coord.clamp(-self.pos_bnd, self.pos_bnd) + self.pos_bnd + torch.arange(3, device=coord.device) * self.rpe_num
        out = self.rpe_table.index_select(0, idx.reshape(-1))
        return out.view(idx.shape + (-1,)).sum(3).permute(0, 3, 1, 2)


class SerializedAttention(PointModule):
    def __init__(self, channels, num_heads, patch_size, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0,
                 order_index=0, enable_rpe=False, enable_flash=True, upcast_attention=True, upcast_softmax=True):
        super().__init__()
        assert channels % num_heads == 0
        self.channels, self.num_heads = channels, num_heads
        self.scale = qk_scale or (channels // num_heads) ** -0.5
        self.order_index = order_index
        self.upcast_attention, self.upcast_softmax = upcast_attention, upcast_softmax
        self.enable_rpe, self.enable_flash = enable_rpe, enable_flash
        if enable_flash:
            assert enable_rpe is False and upcast_attention is False and upcast_softmax is False
            self.patch_size = patch_size
            self.attn_drop = attn_drop
        else:
            self.patch_size_max = patch_size
            self.patch_size = 0
            self.attn_drop = nn.Dropout(attn_drop)
        self.qkv = nn.Linear(channels, channels * 3, bias=qkv_bias)
        self.proj = nn.Linear(channels, channels)
        self.proj_drop = nn.Dropout(proj_drop)
        self.softmax = nn.Softmax(dim=-1)
        self.rpe = RPE(patch_size, num_heads) if enable_rpe else None

    @torch.no_grad()
    def get_rel_pos(self, point, order):
        K = self.patch_size
        key = f"rel_pos_{self.order_index}"
        if key not in point:
            gc = point.grid_coord[order].reshape(-1, K, 3)
            point[key] = gc.unsqueeze(2) - gc.unsqueeze(1)
        return point[key]

    @torch.no_grad()
    def get_padding_and_inverse(self, point):
        """point_transformer_v3m1_base.py:118-173: pad every segment to a multiple of the patch size by repeating the
        tail of its second-to-last patch.  (Index arithmetic on the host, as the reference: a handful of segments.)"""
        if not {"pad", "unpad", "cu_seqlens_key"}.issubset(point.keys()):
            offset = point.offset
            K = self.patch_size
            bincount = offset2bincount(offset)
            bincount_pad = torch.div(bincount + K - 1, K, rounding_mode="trunc") * K
            mask_pad = bincount > K
            bincount_pad = ~mask_pad * bincount + mask_pad * bincount_pad
            _offset = nn.functional.pad(offset, (1, 0)).tolist()
            _offset_pad = nn.functional.pad(torch.cumsum(bincount_pad, dim=0), (1, 0)).tolist()
            bc, bcp = bincount.tolist(), bincount_pad.tolist()
            pad = torch.arange(_offset_pad[-1], device=offset.device)
            unpad = torch.arange(_offset[-1], device=offset.device)
            cu = []
            for i in range(len(bc)):
                unpad[_offset[i]:_offset[i + 1]] += _offset_pad[i] - _offset[i]
                if bc[i] != bcp[i]:
                    pad[_offset_pad[i + 1] - K + (bc[i] % K):_offset_pad[i + 1]] = \
                        pad[_offset_pad[i + 1] - 2 * K + (bc[i] % K):_offset_pad[i + 1] - K]
                pad[_offset_pad[i]:_offset_pad[i + 1]] -= _offset_pad[i] - _offset[i]
                cu.append(torch.arange(_offset_pad[i], _offset_pad[i + 1], step=K, dtype=torch.int32, device=offset.device))
            point["pad"], point["unpad"] = pad, unpad
            point["cu_seqlens_key"] = nn.functional.pad(torch.concat(cu), (0, 1), value=_offset_pad[-1])
        return point["pad"], point["unpad"], point["cu_seqlens_key"]

    def forward(self, point):
        if not self.enable_flash:
            self.patch_size = min(int(offset2bincount(point.offset).min()), self.patch_size_max)
        H, K, C = self.num_heads, self.patch_size, self.channels
        pad, unpad, _ = self.get_padding_and_inverse(point)
        order = point.serialized_order[self.order_index][pad]
        inverse = unpad[point.serialized_inverse[self.order_index]]
        qkv = self.qkv(point.feat)[order]
        q, k, v = qkv.reshape(-1, K, 3, H, C // H).permute(2, 0, 3, 1, 4).unbind(dim=0)        # (N', H, K, C')
        if not self.enable_flash:
            if self.upcast_attention:
                q, k = q.float(), k.float()
            attn = (q * self.scale) @ k.transpose(-2, -1)
            if self.enable_rpe:
                attn = attn + self.rpe(self.get_rel_pos(point, order))
            if self.upcast_softmax:
                attn = attn.float()
            attn = self.attn_drop(self.softmax(attn)).to(qkv.dtype)
            feat = (attn @ v).transpose(1, 2).reshape(-1, C)
        else:
            # every patch is a full K-point sequence after padding, so the varlen call of the reference is a plain batched
            # attention over (patches, heads, K, C'): fused SDPA in bf16 (the reference: flash_attn in fp16)
            lp = torch.bfloat16 if qkv.is_cuda else qkv.dtype
            o = nn.functional.scaled_dot_product_attention(q.to(lp), k.to(lp), v.to(lp), scale=self.scale,
                                                           dropout_p=self.attn_drop if self.training else 0.0)
            feat = o.transpose(1, 2).reshape(-1, C).to(qkv.dtype)
        feat = feat[inverse]
        point.feat = self.proj_drop(self.proj(feat))
        return point


class MLP(nn.Module):
    def __init__(self, in_channels, hidden_channels=None, out_channels=None, act_layer=nn.GELU, drop=0.0):
        super().__init__()
        out_channels = out_channels or in_channels
        hidden_channels = hidden_channels or in_channels
        self.fc1 = nn.Linear(in_channels, hidden_channels)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_channels, out_channels)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class Block(PointModule):
    def __init__(self, channels, num_heads, patch_size=48, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, attn_drop=0.0,
                 proj_drop=0.0, drop_path=0.0, norm_layer=nn.LayerNorm, act_layer=nn.GELU, pre_norm=True, order_index=0,
                 cpe_indice_key=None, enable_rpe=False, enable_flash=True, upcast_attention=True, upcast_softmax=True):
        super().__init__()
        self.channels, self.pre_norm = channels, pre_norm
        self.cpe = PointSequential(SubMConv3d(channels, channels, kernel_size=3, bias=True, indice_key=cpe_indice_key),
                                   nn.Linear(channels, channels), norm_layer(channels))
        self.norm1 = PointSequential(norm_layer(channels))
        self.attn = SerializedAttention(channels=channels, patch_size=patch_size, num_heads=num_heads, qkv_bias=qkv_bias,
                                        qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=proj_drop, order_index=order_index,
                                        enable_rpe=enable_rpe, enable_flash=enable_flash, upcast_attention=upcast_attention,
                                        upcast_softmax=upcast_softmax)
        self.norm2 = PointSequential(norm_layer(channels))
        self.mlp = PointSequential(MLP(in_channels=channels, hidden_channels=int(channels * mlp_ratio),
                                       out_channels=channels, act_layer=act_layer, drop=proj_drop))
        self.drop_path = PointSequential(DropPath(drop_path) if drop_path > 0.0 else nn.Identity())

    def forward(self, point: Point):
        shortcut = point.feat
        point = self.cpe(point)
        point.feat = shortcut + point.feat
        shortcut = point.feat
        if self.pre_norm:
            point = self.norm1(point)
        point = self.drop_path(self.attn(point))
        point.feat = shortcut + point.feat
        if not self.pre_norm:
            point = self.norm1(point)
        shortcut = point.feat
        if self.pre_norm:
            point = self.norm2(point)
        point = self.drop_path(self.mlp(point))
        point.feat = shortcut + point.feat
        if not self.pre_norm:
            point = self.norm2(point)
        point.sparse_conv_feat = point.sparse_conv_feat.replace_feature(point.feat)
        return point


class SerializedPooling(PointModule):
    def __init__(self, in_channels, out_channels, stride=2, norm_layer=None, act_layer=None, reduce="max",
                 shuffle_orders=True, traceable=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        assert stride == 2 ** (math.ceil(stride) - 1).bit_length()
        self.stride, self.reduce, self.shuffle_orders, self.traceable = stride, reduce, shuffle_orders, traceable
        assert reduce in ["sum", "mean", "min", "max"]
        self.proj = nn.Linear(in_channels, out_channels)
        self.norm = PointSequential(norm_layer(out_channels)) if norm_layer is not None else None
        self.act = PointSequential(act_layer()) if act_layer is not None else None

    def forward(self, point: Point):
        pooling_depth = (math.ceil(self.stride) - 1).bit_length()
        if pooling_depth > point.serialized_depth:
            pooling_depth = 0
        code = point.serialized_code >> pooling_depth * 3
        _, cluster, counts = torch.unique(code[0], sorted=True, return_inverse=True, return_counts=True)
        _, indices = torch.sort(cluster)
        idx_ptr = torch.cat([counts.new_zeros(1), torch.cumsum(counts, dim=0)])
        head_indices = indices[idx_ptr[:-1]]
        code = code[:, head_indices]
        order = torch.argsort(code)
        inverse = torch.zeros_like(order).scatter_(
            dim=1, index=order, src=torch.arange(0, code.shape[1], device=order.device).repeat(code.shape[0], 1))
        if self.shuffle_orders:
            perm = torch.randperm(code.shape[0])
            code, order, inverse = code[perm], order[perm], inverse[perm]
        d = dict(feat=segment_csr(self.proj(point.feat)[indices], idx_ptr, reduce=self.reduce),
                 coord=segment_csr(point.coord[indices], idx_ptr, reduce="mean"),
                 grid_coord=point.grid_coord[head_indices] >> pooling_depth,
                 serialized_code=code, serialized_order=order, serialized_inverse=inverse,
                 serialized_depth=point.serialized_depth - pooling_depth, batch=point.batch[head_indices])
        if self.traceable:
            d["pooling_inverse"], d["pooling_parent"] = cluster, point
        point = Point(d)
        if self.norm is not None:
            point = self.norm(point)
        if self.act is not None:
            point = self.act(point)
        point.sparsify()
        return point


class SerializedUnpooling(PointModule):
    def __init__(self, in_channels, skip_channels, out_channels, norm_layer=None, act_layer=None, traceable=False):
        super().__init__()
        self.proj = PointSequential(nn.Linear(in_channels, out_channels))
        self.proj_skip = PointSequential(nn.Linear(skip_channels, out_channels))
        if norm_layer is not None:
            self.proj.add(norm_layer(out_channels))
            self.proj_skip.add(norm_layer(out_channels))
        if act_layer is not None:
            self.proj.add(act_layer())
            self.proj_skip.add(act_layer())
        self.traceable = traceable

    def forward(self, point):
        parent, inverse = point.pop("pooling_parent"), point.pop("pooling_inverse")
        point = self.proj(point)
        parent = self.proj_skip(parent)
        parent.feat = parent.feat + point.feat[inverse]
        if self.traceable:
            parent["unpooling_parent"] = point
        return parent


class Embedding(PointModule):
    def __init__(self, in_channels, embed_channels, norm_layer=None, act_layer=None):
        super().__init__()
        self.in_channels, self.embed_channels = in_channels, embed_channels
        self.stem = PointSequential(conv=SubMConv3d(in_channels, embed_channels, kernel_size=5, padding=1, bias=False,
                                                    indice_key="stem"))
        if norm_layer is not None:
            self.stem.add(norm_layer(embed_channels), name="norm")
        if act_layer is not None:
            self.stem.add(act_layer(), name="act")

    def forward(self, point: Point):
        return self.stem(point)


class PointTransformerV3(PointModule):
    def __init__(self, in_channels=6, order=("z", "z-trans"), stride=(2, 2, 2, 2), enc_depths=(2, 2, 2, 6, 2),
                 enc_channels=(32, 64, 128, 256, 512), enc_num_head=(2, 4, 8, 16, 32), enc_patch_size=(48, 48, 48, 48, 48),
                 dec_depths=(2, 2, 2, 2), dec_channels=(64, 64, 128, 256), dec_num_head=(4, 4, 8, 16),
                 dec_patch_size=(48, 48, 48, 48), mlp_ratio=4, qkv_bias=True, qk_scale=None, attn_drop=0.0, proj_drop=0.0,
                 drop_path=0.3, pre_norm=True, shuffle_orders=True, enable_rpe=False, enable_flash=True,
                 upcast_attention=False, upcast_softmax=False, cls_mode=False, pdnorm_bn=False, pdnorm_ln=False, cfg=None,
                 **unused_pdnorm_kwargs):
        super().__init__()
        if pdnorm_bn or pdnorm_ln:
            raise NotImplementedError("PDNorm (point prompt training) is not on the pre-training path")
        self.cfg = cfg
        self.num_stages = len(enc_depths)
        self.order = [order] if isinstance(order, str) else list(order)
        self.cls_mode, self.shuffle_orders, self.channels = cls_mode, shuffle_orders, enc_channels
        assert self.num_stages == len(stride) + 1 == len(enc_channels) == len(enc_num_head) == len(enc_patch_size)
        assert cls_mode or self.num_stages == len(dec_depths) + 1 == len(dec_channels) + 1 == len(dec_num_head) + 1
        self.use_fusion = bool(cfg is not None and hasattr(cfg.opt, "use_fusion") and cfg.opt.use_fusion)
        bn_layer = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        ln_layer, act_layer = nn.LayerNorm, nn.GELU
        self.embedding = Embedding(in_channels=in_channels, embed_channels=enc_channels[0], norm_layer=bn_layer,
                                   act_layer=act_layer)
        blk = dict(mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=proj_drop,
                   norm_layer=ln_layer, act_layer=act_layer, pre_norm=pre_norm, enable_rpe=enable_rpe,
                   enable_flash=enable_flash, upcast_attention=upcast_attention, upcast_softmax=upcast_softmax)
        enc_dp = [x.item() for x in torch.linspace(0, drop_path, sum(enc_depths))]
        self.enc = PointSequential()
        for s in range(self.num_stages):
            dp = enc_dp[sum(enc_depths[:s]):sum(enc_depths[:s + 1])]
            enc = PointSequential()
            if s > 0:
                enc.add(SerializedPooling(in_channels=enc_channels[s - 1], out_channels=enc_channels[s], stride=stride[s - 1],
                                          norm_layer=bn_layer, act_layer=act_layer), name="down")
            for i in range(enc_depths[s]):
                enc.add(Block(channels=enc_channels[s], num_heads=enc_num_head[s], patch_size=enc_patch_size[s],
                              drop_path=dp[i], order_index=i % len(self.order), cpe_indice_key=f"stage{s}", **blk),
                        name=f"block{i}")
            if len(enc) != 0:
                self.enc.add(module=enc, name=f"enc{s}")
        if not cls_mode:
            dec_dp = [x.item() for x in torch.linspace(0, drop_path, sum(dec_depths))]
            self.dec = PointSequential()
            dec_channels = list(dec_channels) + [enc_channels[-1]]
            for s in reversed(range(self.num_stages - 1)):
                dp = dec_dp[sum(dec_depths[:s]):sum(dec_depths[:s + 1])]
                dp.reverse()
                dec = PointSequential()
                dec.add(SerializedUnpooling(in_channels=dec_channels[s + 1], skip_channels=enc_channels[s],
                                            out_channels=dec_channels[s], norm_layer=bn_layer, act_layer=act_layer), name="up")
                for i in range(dec_depths[s]):
                    dec.add(Block(channels=dec_channels[s], num_heads=dec_num_head[s], patch_size=dec_patch_size[s],
                                  drop_path=dp[i], order_index=i % len(self.order), cpe_indice_key=f"stage{s}", **blk),
                            name=f"block{i}")
                self.dec.add(module=dec, name=f"dec{s}")

    def forward(self, data_dict, img_features=None, unprojected_coords=None, fusion_mlps=None):
        if self.use_fusion:
            raise NotImplementedError("PointTransformerV3: PointFusion (fusion/point_fusion.py) is not built; "
                                      "set opt.use_fusion=false")
        original_point = Point(data_dict)
        original_point.serialization(order=self.order, shuffle_orders=self.shuffle_orders)
        original_point.sparsify()
        original_point = self.embedding(original_point)
        now = {}
        data_dict["coord"] = original_point.coord
        original_point.feat = original_point.sparse_conv_feat.features
        # the reference rebuilds the Point as ONE segment (699-717); scene indices stay in `batch`
        original_point.offset = torch.tensor([len(original_point.batch)], device=original_point.batch.device)
        for key in data_dict.keys():
            now[key] = original_point[key]
        now["batch"] = original_point.batch
        point = Point(now)
        point.serialization(order=self.order, shuffle_orders=self.shuffle_orders)
        point.sparsify()
        point = self.enc(point)
        if not self.cls_mode:
            point = self.dec(point)
        return point
