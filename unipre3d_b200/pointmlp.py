"""PointMLP backbone (encoder + feature-propagation decoder) of the render-loss path -- SURVEY.md §8a row B6,
BASELINE.json configs[2] (`pointmlp_pretraining`): one Gaussian per INPUT point (P = N = 8192).

Mirrors the operator surface and the state-dict layout of
/root/reference/openpoints/models/backbone/pointmlp.py:116-639 (LocalGrouper, ConvBNReLU1D, ConvBNReLURes1D,
PreExtraction, PosExtraction, PointNetFeaturePropagation, PointMLPEncoder, pointMLP factory), so
`point_network.encoder.*` checkpoint keys load unchanged.

B200 notes (same arithmetic, different dataflow)
  * every Conv1d(kernel 1) runs as a GEMM over channels-last rows (B*S*k, C); BatchNorm1d on (rows, C) sees the same
    per-channel statistics as on (B*S, C, k);
  * kNN / 3-NN: `up3d_knn` (warp per query, shared-memory tiles) instead of the reference's (S x N) distance matrix +
    `topk` / FULL `sort` (pointmlp.py:102-113, 397-403) -- 128 MB per object at the first stage are never written;
  * FPS: `up3d_fps` (register-resident cloud).

Quirk reproduced on purpose: with `in_channels = 4` the reference hands the (B,N,4) tensor to its FPS kernel, which
indexes it as (B,N,3) (`dataset[k*3+..]`, base `b*n*3`: sampling_gpu.cu:117-131) -- i.e. FPS runs on the first 3N floats
re-read as N xyz triples, offset by b*3N.  `_fps_like_reference` feeds our FPS exactly those floats so the sampled
indices match.  Distances for kNN / 3-NN are taken over all C channels, as `square_distance` does.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import pointops
from .fusion import FeatureFusion


def get_activation(activation):
    a = activation.lower()
    table = {"gelu": nn.GELU, "rrelu": lambda: nn.RReLU(inplace=True), "selu": lambda: nn.SELU(inplace=True),
             "silu": lambda: nn.SiLU(inplace=True), "hardswish": lambda: nn.Hardswish(inplace=True),
             "leakyrelu": lambda: nn.LeakyReLU(inplace=True)}
    return table[a]() if a in table else nn.ReLU(inplace=True)


PREFETCH_GEOMETRY = True      # tests flip it to compare against the in-line order of the reference


def _pw(conv: nn.Conv1d, rows: torch.Tensor) -> torch.Tensor:
    """Conv1d(kernel 1, groups 1) on channels-last rows."""
    return F.linear(rows, conv.weight.squeeze(-1), conv.bias)


def _fps_like_reference(xyz: torch.Tensor, npoint: int) -> torch.Tensor:
    B, N, C = xyz.shape
    if C == 3:
        return pointops.furthest_point_sample(xyz.contiguous(), npoint)
    flat = xyz.contiguous().reshape(-1)
    view = torch.as_strided(flat, (B, N, 3), (N * 3, 3, 1)).contiguous()
    return pointops.furthest_point_sample(view, npoint)


def index_points(points, idx):
    """points (B,N,C), idx (B,S[,k]) int64 -> (B,S[,k],C)"""
    B = points.shape[0]
    bidx = torch.arange(B, device=points.device).view(B, *([1] * (idx.dim() - 1)))
    return points[bidx, idx]


class LocalGrouper(nn.Module):
    def __init__(self, channel, sample_ratio, kneighbors, use_xyz=True, normalize="center", **kwargs):
        super().__init__()
        self.sample_ratio, self.kneighbors, self.use_xyz = sample_ratio, kneighbors, use_xyz
        self.normalize = normalize.lower() if normalize is not None else None
        if self.normalize not in ["center", "anchor"]:
            self.normalize = None
        if self.normalize is not None:
            add_channel = 3 if self.use_xyz else 0
            self.affine_alpha = nn.Parameter(torch.ones([1, 1, 1, channel + add_channel]))
            self.affine_beta = nn.Parameter(torch.zeros([1, 1, 1, channel + add_channel]))

    def geometry(self, xyz):
        """The part of forward() that depends on the coordinates only: (fps_idx, new_xyz, knn idx)."""
        B, N, C = xyz.shape
        S = N // self.sample_ratio
        xyz = xyz.contiguous()
        fps_idx = _fps_like_reference(xyz, S).long()
        new_xyz = index_points(xyz, fps_idx)
        return fps_idx, new_xyz, pointops.knn_point(self.kneighbors, xyz, new_xyz)

    def forward(self, xyz, points, geom=None):
        """xyz (B,N,C), points (B,N,d) -> new_xyz (B,S,C), new_points (B,S,k,2d[+..]); `geom` = a prefetched geometry()"""
        B, N, C = xyz.shape
        S = N // self.sample_ratio
        xyz = xyz.contiguous()
        fps_idx, new_xyz, idx = geom if geom is not None else self.geometry(xyz)
        new_points = index_points(points, fps_idx)
        grouped_points = index_points(points, idx)                                  # (B,S,k,d)
        if self.use_xyz:
            grouped_points = torch.cat([grouped_points, index_points(xyz, idx)], dim=-1)
        if self.normalize is not None:
            if self.normalize == "center":
                mean = torch.mean(grouped_points, dim=2, keepdim=True)
            else:
                mean = torch.cat([new_points, new_xyz], dim=-1) if self.use_xyz else new_points
                mean = mean.unsqueeze(dim=-2)
            centred = grouped_points - mean
            std = torch.std(centred.reshape(B, -1).float(), dim=-1, keepdim=True).unsqueeze(dim=-1).unsqueeze(dim=-1)
            grouped_points = centred / (std + 1e-5)
            grouped_points = self.affine_alpha * grouped_points + self.affine_beta
        new_points = torch.cat([grouped_points,
                                new_points.view(B, S, 1, -1).expand(-1, -1, self.kneighbors, -1)], dim=-1)
        return new_xyz, new_points


class ConvBNReLU1D(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, bias=True, activation="relu"):
        super().__init__()
        self.act = get_activation(activation)
        self.net = nn.Sequential(nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, bias=bias),
                                 nn.BatchNorm1d(out_channels), self.act)

    def forward(self, rows):
        """rows: (R, C_in) channels-last"""
        return self.net[2](self.net[1](_pw(self.net[0], rows)))


class ConvBNReLURes1D(nn.Module):
    def __init__(self, channel, kernel_size=1, groups=1, res_expansion=1.0, bias=True, activation="relu"):
        super().__init__()
        if groups != 1:
            raise NotImplementedError("grouped residual blocks are not used by the pointMLP factory (groups=1)")
        self.act = get_activation(activation)
        mid = int(channel * res_expansion)
        self.net1 = nn.Sequential(nn.Conv1d(channel, mid, kernel_size=kernel_size, groups=groups, bias=bias),
                                  nn.BatchNorm1d(mid), self.act)
        self.net2 = nn.Sequential(nn.Conv1d(mid, channel, kernel_size=kernel_size, bias=bias), nn.BatchNorm1d(channel))

    def forward(self, rows):
        y = self.net1[2](self.net1[1](_pw(self.net1[0], rows)))
        y = self.net2[1](_pw(self.net2[0], y))
        return self.act(y + rows)


class PreExtraction(nn.Module):
    def __init__(self, channels, out_channels, blocks=1, groups=1, res_expansion=1, bias=True, activation="relu",
                 use_xyz=True):
        super().__init__()
        in_channels = 3 + 2 * channels if use_xyz else 2 * channels
        self.transfer = ConvBNReLU1D(in_channels, out_channels, bias=bias, activation=activation)
        self.operation = nn.Sequential(*[ConvBNReLURes1D(out_channels, groups=groups, res_expansion=res_expansion,
                                                         bias=bias, activation=activation) for _ in range(blocks)])

    def forward(self, x):
        """x (B,S,k,d) -> (B,S,out) channels-last (the reference returns (B,out,S))"""
        b, n, s, d = x.size()
        rows = self.operation(self.transfer(x.reshape(b * n * s, d)))
        return rows.reshape(b * n, s, -1).max(dim=1)[0].reshape(b, n, -1)


class PosExtraction(nn.Module):
    def __init__(self, channels, blocks=1, groups=1, res_expansion=1, bias=True, activation="relu"):
        super().__init__()
        self.operation = nn.Sequential(*[ConvBNReLURes1D(channels, groups=groups, res_expansion=res_expansion,
                                                         bias=bias, activation=activation) for _ in range(blocks)])

    def forward(self, x):
        """x (B,S,C) channels-last"""
        b, n, c = x.shape
        return self.operation(x.reshape(b * n, c)).reshape(b, n, c)


class PointNetFeaturePropagation(nn.Module):
    def __init__(self, in_channel, out_channel, blocks=1, groups=1, res_expansion=1.0, bias=True, activation="relu",
                 has_MLP=True):
        super().__init__()
        if has_MLP:
            self.fuse = ConvBNReLU1D(in_channel, out_channel, 1, bias=bias)
            self.extraction = PosExtraction(out_channel, blocks, groups=groups, res_expansion=res_expansion, bias=bias,
                                            activation=activation)
        self.has_MLP = has_MLP

    def forward(self, xyz1, xyz2, points1, points2, nn3=None):
        """xyz1 (B,N,C) dense, xyz2 (B,S,C) sparse, points1 (B,N,D') or None, points2 (B,S,D'') -> (B,N,D''')
        (all channels-last).  3-NN inverse-distance interpolation, pointmlp.py:397-409."""
        B, N, _ = xyz1.shape
        S = xyz2.shape[1]
        if S == 1:
            interpolated = points2.expand(-1, N, -1)
        else:
            # nn3: prefetched knn_point(3, xyz2, xyz1, return_dist=True)
            idx, dists = nn3 if nn3 is not None else pointops.knn_point(3, xyz2, xyz1, return_dist=True)   # (B,N,3)
            dist_recip = 1.0 / (dists + 1e-8)
            weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
            interpolated = torch.sum(index_points(points2, idx) * weight.view(B, N, 3, 1).to(points2.dtype), dim=2)
        new_points = torch.cat([points1, interpolated], dim=-1) if points1 is not None else interpolated
        if self.has_MLP:
            b, n, c = new_points.shape
            rows = self.fuse(new_points.reshape(b * n, c))
            new_points = self.extraction.operation(rows).reshape(b, n, -1)
        return new_points


class PointMLPEncoder(nn.Module):
    def __init__(self, in_channels=4, embed_dim=64, groups=1, res_expansion=1.0, activation="relu", bias=False,
                 use_xyz=True, normalize="anchor", dim_expansion=(2, 2, 2, 2), pre_blocks=(2, 2, 2, 2),
                 pos_blocks=(2, 2, 2, 2), k_neighbors=(32, 32, 32, 32), reducers=(2, 2, 2, 2), de_blocks=(2, 2, 2, 2),
                 de_dims=(512, 256, 128, 128), **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.stages = len(pre_blocks)
        self.embedding = ConvBNReLU1D(in_channels, embed_dim, bias=bias, activation=activation)
        self.use_fusion = kwargs.get("use_fusion", True)
        assert len(pre_blocks) == len(k_neighbors) == len(reducers) == len(pos_blocks) == len(dim_expansion)
        self.local_grouper_list = nn.ModuleList()
        self.pre_blocks_list = nn.ModuleList()
        self.pos_blocks_list = nn.ModuleList()
        last_channel = embed_dim
        channels = [embed_dim]
        for i in range(len(pre_blocks)):
            out_channel = last_channel * dim_expansion[i]
            channels.append(out_channel)
            self.local_grouper_list.append(LocalGrouper(last_channel, reducers[i], k_neighbors[i], use_xyz, normalize))
            self.pre_blocks_list.append(PreExtraction(last_channel, out_channel, pre_blocks[i], groups=groups,
                                                      res_expansion=res_expansion, bias=bias, activation=activation,
                                                      use_xyz=use_xyz))
            self.pos_blocks_list.append(PosExtraction(out_channel, pos_blocks[i], groups=groups,
                                                      res_expansion=res_expansion, bias=bias, activation=activation))
            last_channel = out_channel
        self.out_channels = last_channel
        self.decode_list = nn.ModuleList()
        en_dims = list(reversed(channels))
        de_dims = [en_dims[0]] + list(de_dims)
        assert len(en_dims) == len(de_dims) == len(de_blocks) + 1
        for i in range(len(de_dims) - 1):
            self.decode_list.append(PointNetFeaturePropagation(de_dims[i] + en_dims[i + 1], de_dims[i + 1],
                                                               blocks=de_blocks[i], groups=groups,
                                                               res_expansion=res_expansion, bias=bias,
                                                               activation=activation))
        self.channels = channels
        self.act = get_activation(activation)

    def forward(self, p, image_features, c2w_projection_matrix, feature_mlps, intrinsic):
        """-> (features (B,N,C_out) after fusion [channels-last, as the reference returns it], p (B,N,C_in))"""
        x = None
        if isinstance(p, dict):
            p, x = p["pos"], p.get("x", None)
        b, n, c = p.shape
        rows = p.reshape(b * n, c) if x is None else x.transpose(1, 2).reshape(b * n, -1)
        # Everything that depends on the coordinates only -- the four sequential farthest-point samplings (4095 + 2047 +
        # 1023 + 511 dependent rounds: ~5 ms of a 37 ms step), the kNN tables and the decoder's 3-NN tables -- is issued
        # on a side stream now and joined stage by stage, so it overlaps the Conv-BN-ReLU blocks of the earlier stages.
        geoms, nn3s, ready = [None] * self.stages, [None] * len(self.decode_list), None
        if p.is_cuda and PREFETCH_GEOMETRY:
            from .fused_encoder import _SIDE_STREAMS
            main = torch.cuda.current_stream(p.device)
            key = ("pmlp", p.device.index if p.device.index is not None else torch.cuda.current_device())
            if key not in _SIDE_STREAMS:
                _SIDE_STREAMS[key] = torch.cuda.Stream(device=p.device)
            side = _SIDE_STREAMS[key]
            side.wait_stream(main)
            ready = []
            with torch.cuda.stream(side), torch.no_grad():
                q, pts = p, [p]
                for i in range(self.stages):
                    geoms[i] = self.local_grouper_list[i].geometry(q)
                    q = geoms[i][1]
                    pts.append(q)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    ready.append(ev)
                rev = list(reversed(pts))
                for i in range(len(self.decode_list)):
                    if rev[i].shape[1] > 1:
                        nn3s[i] = pointops.knn_point(3, rev[i], rev[i + 1], return_dist=True)
                ev = torch.cuda.Event()
                ev.record(side)
                ready.append(ev)
            p.record_stream(side)
        x = self.embedding(rows).reshape(b, n, -1)                                  # (B,N,D) channels-last
        p_list, x_list = [p], [x]
        for i in range(self.stages):
            if ready is not None:
                torch.cuda.current_stream(p.device).wait_event(ready[i])
            p, g = self.local_grouper_list[i](p, x, geoms[i])
            x = self.pos_blocks_list[i](self.pre_blocks_list[i](g))
            p_list.append(p)
            x_list.append(x)
        if ready is not None:
            torch.cuda.current_stream(p.device).wait_event(ready[-1])
        p_list.reverse()
        x_list.reverse()
        x = x_list[0]
        last = len(self.decode_list) - 1
        for i in range(len(self.decode_list)):
            x = self.decode_list[i](p_list[i + 1], p_list[i], x_list[i + 1], x, nn3s[i])
            if self.use_fusion and feature_mlps is not None and i == last:
                x = FeatureFusion(feature_mlps)(x, p_list[i + 1][..., :3], image_features, c2w_projection_matrix,
                                                intrinsic)
        return x, p_list[-1]


def pointMLP(num_classes=40, cfg=None, **kwargs) -> PointMLPEncoder:
    """pointmlp.py:621-639"""
    return PointMLPEncoder(in_channels=cfg.model.in_channels, num_classes=num_classes, embed_dim=64, groups=1,
                           res_expansion=1.0, activation="relu", bias=False, use_xyz=False, normalize="anchor",
                           dim_expansion=[2, 2, 2, 2], pre_blocks=[2, 2, 2, 2], pos_blocks=[2, 2, 2, 2],
                           k_neighbors=[24, 24, 24, 24], reducers=[2, 2, 2, 2], de_dims=[512, 256, 128, 128], **kwargs)
