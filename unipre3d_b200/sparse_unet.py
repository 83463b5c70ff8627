"""SpUNet scene backbone (the `sparseunet_pretraining` config) on the B200 sparse-convolution engine.

Module mirror of /root/reference/pointcept/models/sparse_unet/spconv_unet_v1m1_base.py: `BasicBlock` (25-105) and
`SpUNetBase` (107-363) with the same constructor arguments, sub-module names (conv_input, down, up, enc, dec, final) and
parameter layouts, so the reference's checkpoints load; the spconv layers are `unipre3d_b200.sparse` (csrc/sparse_conv.cu).
`forward(input_dict, img_features, unprojected_coords, fusion_mlps)` keeps the reference signature and returns a
`SparseConvTensor`; rows are in sorted-voxel order (`.features_in_input_order()` restores the caller's order).

Not built: `PointFusion` (fusion/point_fusion.py:36-131 -- its voxeliser is a CPU/numpy GridSample with train-mode random
sampling, pointcept/datasets/transform_with_extrinsic.py:1179-1327); with `cfg.opt.use_fusion` the forward raises.
"""
from __future__ import annotations

from collections import OrderedDict
from functools import partial

import torch
import torch.nn as nn

from .sparse import SparseConv3d, SparseConvTensor, SparseInverseConv3d, SparseSequential, SubMConv3d


def offset2batch(offset: torch.Tensor) -> torch.Tensor:
    """pointcept.models.utils.offset2batch: cumulative point counts -> per-point batch index."""
    counts = torch.diff(offset, prepend=offset.new_zeros(1))
    return torch.repeat_interleave(torch.arange(len(offset), device=offset.device), counts)


class BasicBlock(nn.Module):
    is_sparse_module = True
    expansion = 1

    def __init__(self, in_channels, embed_channels, stride=1, norm_fn=None, indice_key=None, bias=False):
        super().__init__()
        assert norm_fn is not None
        if in_channels == embed_channels:
            self.proj = SparseSequential(nn.Identity())
        else:
            self.proj = SparseSequential(SubMConv3d(in_channels, embed_channels, kernel_size=1, bias=False),
                                         norm_fn(embed_channels))
        self.conv1 = SubMConv3d(in_channels, embed_channels, kernel_size=3, stride=stride, padding=1, bias=bias,
                                indice_key=indice_key)
        self.bn1 = norm_fn(embed_channels)
        self.relu = nn.ReLU()
        self.conv2 = SubMConv3d(embed_channels, embed_channels, kernel_size=3, stride=stride, padding=1, bias=bias,
                                indice_key=indice_key)
        self.bn2 = norm_fn(embed_channels)
        self.stride = stride

    def forward(self, x):
        residual = x
        out = self.conv1(x)
        out = out.replace_feature(self.relu(self.bn1(out.features)))
        out = self.conv2(out)
        out = out.replace_feature(self.bn2(out.features))
        out = out.replace_feature(self.relu(out.features + self.proj(residual).features))
        return out


class SpUNetBase(nn.Module):
    def __init__(self, in_channels, num_classes, cfg=None, base_channels=32, channels=(32, 64, 128, 256, 256, 128, 96, 96),
                 layers=(2, 3, 4, 6, 2, 2, 2, 2), cls_mode=False):
        super().__init__()
        assert len(layers) % 2 == 0 and len(layers) == len(channels)
        self.cfg = cfg
        self.use_fusion = bool(cfg is not None and hasattr(cfg.opt, "use_fusion") and cfg.opt.use_fusion)
        self.in_channels, self.num_classes, self.base_channels = in_channels, num_classes, base_channels
        self.channels, self.layers, self.num_stages, self.cls_mode = channels, layers, len(layers) // 2, cls_mode
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        self.conv_input = SparseSequential(
            SubMConv3d(in_channels, base_channels, kernel_size=5, padding=1, bias=False, indice_key="stem"),
            norm_fn(base_channels), nn.ReLU())
        enc_channels, dec_channels = base_channels, channels[-1]
        self.down, self.up, self.enc = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.dec = nn.ModuleList() if not cls_mode else None
        for s in range(self.num_stages):
            self.down.append(SparseSequential(
                SparseConv3d(enc_channels, channels[s], kernel_size=2, stride=2, bias=False, indice_key=f"spconv{s + 1}"),
                norm_fn(channels[s]), nn.ReLU()))
            self.enc.append(SparseSequential(OrderedDict(
                [(f"block{i}", BasicBlock(channels[s], channels[s], norm_fn=norm_fn, indice_key=f"subm{s + 1}"))
                 for i in range(layers[s])])))
            if not cls_mode:
                self.up.append(SparseSequential(
                    SparseInverseConv3d(channels[len(channels) - s - 2], dec_channels, kernel_size=2, bias=False,
                                        indice_key=f"spconv{s + 1}"),
                    norm_fn(dec_channels), nn.ReLU()))
                self.dec.append(SparseSequential(OrderedDict(
                    [(f"block{i}", BasicBlock(dec_channels + enc_channels if i == 0 else dec_channels, dec_channels,
                                              norm_fn=norm_fn, indice_key=f"subm{s}"))
                     for i in range(layers[len(channels) - s - 1])])))
            enc_channels = channels[s]
            dec_channels = channels[len(channels) - s - 2]
        final_in = channels[-1] if not cls_mode else channels[self.num_stages - 1]
        self.final = SubMConv3d(final_in, num_classes, kernel_size=1, padding=1, bias=True) if num_classes > 0 else nn.Identity()
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m):
        if isinstance(m, (nn.Linear, SubMConv3d)):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.BatchNorm1d):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward(self, input_dict, img_features=None, unprojected_coords=None, fusion_mlps=None) -> SparseConvTensor:
        if self.use_fusion:
            raise NotImplementedError("SpUNetBase: PointFusion (fusion/point_fusion.py) is not built; set opt.use_fusion=false")
        grid_coord, feat, offset = input_dict["grid_coord"], input_dict["feat"], input_dict["offset"]
        batch = offset2batch(offset)
        x = SparseConvTensor(feat, torch.cat([batch.unsqueeze(-1).int(), grid_coord.int()], dim=1).contiguous(),
                             spatial_shape=None, batch_size=int(len(offset)))
        x = self.conv_input(x)
        skips = [x]
        for s in range(self.num_stages):
            x = self.down[s](x)
            x = self.enc[s](x)
            skips.append(x)
        x = skips.pop(-1)
        if not self.cls_mode:
            for s in reversed(range(self.num_stages)):
                x = self.up[s](x)
                skip = skips.pop(-1)
                x = x.replace_feature(torch.cat((x.features, skip.features), dim=1))
                x = self.dec[s](x)
        x = self.final(x) if isinstance(self.final, SubMConv3d) else x
        if self.cls_mode:
            b = x.indices[:, 0].long()
            n = int(b.max()) + 1
            s = torch.zeros((n, x.features.shape[1]), device=x.features.device).index_add_(0, b, x.features)
            x = x.replace_feature(s / torch.bincount(b, minlength=n).clamp(min=1).unsqueeze(1))
        return x
