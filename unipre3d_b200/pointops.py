"""Point ops with the reference's Python surface
(/root/reference/openpoints/models/layers/subsample.py:77-160, group.py:76-203): same names, argument order,
dtypes and autograd behaviour, backed by the sm_100a kernels through the C ABI.
"""
from __future__ import annotations

from typing import Tuple

import torch
from torch.autograd import Function

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        """xyz (B,N,3) contiguous f32, N > npoint -> (B,npoint) int32 indices (first index 0)."""
        require_cuda(xyz)
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        output = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
        temp = None
        if N > _lib.lib.up3d_fps_max_resident_points():
            temp = torch.empty((B, N), dtype=torch.float32, device=xyz.device)
        with torch.cuda.device(xyz.device):
            check(_lib.lib.up3d_fps(B, N, int(npoint), ptr(xyz), ptr(temp), ptr(output), stream_ptr()), launches=1)
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B,C,N), idx (B,npoint) int32 -> (B,C,npoint)."""
        require_cuda(features, idx)
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, npoint = idx.size()
        _, Cc, N = features.size()
        output = torch.empty((B, Cc, npoint), dtype=torch.float32, device=features.device)
        with torch.cuda.device(features.device):
            check(_lib.lib.up3d_gather_points(B, Cc, N, npoint, ptr(features), ptr(idx), ptr(output), stream_ptr()), 1)
        ctx.for_backwards = (idx, Cc, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, Cc, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_out = grad_out.contiguous().float()
        grad_features = torch.empty((B, Cc, N), dtype=torch.float32, device=grad_out.device)
        with torch.cuda.device(grad_out.device):
            check(_lib.lib.up3d_gather_points_grad(B, Cc, N, npoint, ptr(grad_out), ptr(idx), ptr(grad_features),
                                                   stream_ptr()), 1)
        return grad_features, None


gather_operation = GatherOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius: float, nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor) -> torch.Tensor:
        """xyz (B,N,3), new_xyz (B,npoint,3) -> idx (B,npoint,nsample) int32."""
        require_cuda(xyz, new_xyz)
        assert new_xyz.is_contiguous()
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        npoint = new_xyz.size(1)
        idx = torch.empty((B, npoint, nsample), dtype=torch.int32, device=xyz.device)
        with torch.cuda.device(xyz.device):
            check(_lib.lib.up3d_ball_query(B, N, npoint, float(radius), int(nsample), ptr(new_xyz), ptr(xyz), ptr(idx),
                                           stream_ptr()), 1)
        ctx.mark_non_differentiable(idx)
        return idx

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """features (B,C,N), idx (B,npoint,nsample) int32 -> (B,C,npoint,nsample)."""
        require_cuda(features, idx)
        features = features.float()
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, nfeatures, nsample = idx.size()
        _, Cc, N = features.size()
        output = torch.empty((B, Cc, nfeatures, nsample), dtype=torch.float32, device=features.device)
        with torch.cuda.device(features.device):
            check(_lib.lib.up3d_group_points(B, Cc, N, nfeatures, nsample, ptr(features), ptr(idx), ptr(output),
                                             stream_ptr()), 1)
        ctx.for_backwards = (idx, N)
        return output

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        idx, N = ctx.for_backwards
        B, Cc, npoint, nsample = grad_out.size()
        grad_out = grad_out.contiguous().float()
        grad_features = torch.empty((B, Cc, N), dtype=torch.float32, device=grad_out.device)
        with torch.cuda.device(grad_out.device):
            check(_lib.lib.up3d_group_points_grad(B, Cc, N, npoint, nsample, ptr(grad_out), ptr(idx),
                                                  ptr(grad_features), stream_ptr()), 2)
        return grad_features, None


grouping_operation = GroupingOperation.apply


def subsample_group(xyz: torch.Tensor, num_groups: int, group_size: int, radius: float, return_idx: bool = False,
                    out=None):
    """Fused SubsampleGroup.forward (group_embed.py:39-57 with the ball-query grouper): FPS -> centres ->
    ball query -> grouped, centred coordinates.  xyz (B,N,3) -> neighborhood (B,3,G,K), center (B,G,3).
    xyz carries no gradient on the reference's path (raw input coordinates), so neither do the outputs.
    `out` = (neighborhood, center): contiguous fp32 tensors to write into."""
    require_cuda(xyz)
    xyz = xyz.contiguous().float()
    B, N, _ = xyz.shape
    with torch.no_grad():
        fidx = furthest_point_sample(xyz, num_groups)
        if out is not None:
            neigh, center = out
            if not (tuple(neigh.shape) == (B, 3, num_groups, group_size) and tuple(center.shape) == (B, num_groups, 3)
                    and neigh.dtype == center.dtype == torch.float32 and neigh.is_contiguous() and center.is_contiguous()):
                raise ValueError("subsample_group: `out` must be contiguous fp32 (B,3,G,K) and (B,G,3) tensors")
        else:
            center = torch.empty((B, num_groups, 3), dtype=torch.float32, device=xyz.device)
            neigh = torch.empty((B, 3, num_groups, group_size), dtype=torch.float32, device=xyz.device)
        idx = torch.empty((B, num_groups, group_size), dtype=torch.int32, device=xyz.device) if return_idx else None
        with torch.cuda.device(xyz.device):
            check(_lib.lib.up3d_subsample_group(B, N, num_groups, group_size, float(radius), ptr(xyz), ptr(fidx),
                                                ptr(center), ptr(neigh), ptr(idx), stream_ptr()), 1)
    if return_idx:
        return neigh, center, fidx, idx
    return neigh, center


def knn_point(nsample: int, xyz: torch.Tensor, new_xyz: torch.Tensor, return_dist: bool = False):
    """pointMLP's knn_point (openpoints/models/backbone/pointmlp.py:102-113): indices of the `nsample` nearest points
    of `xyz` (B,N,C) for every query in `new_xyz` (B,S,C), C in {3,4}; int64 (B,S,nsample), ascending distance (the
    reference's order is unspecified: topk(sorted=False)).  No distance matrix is materialised."""
    require_cuda(xyz, new_xyz)
    xyz, new_xyz = xyz.contiguous().float(), new_xyz.contiguous().float()
    B, N, Cd = xyz.shape
    S = new_xyz.shape[1]
    with torch.no_grad():
        idx = torch.empty((B, S, nsample), dtype=torch.int32, device=xyz.device)
        dist = torch.empty((B, S, nsample), dtype=torch.float32, device=xyz.device) if return_dist else None
        with torch.cuda.device(xyz.device):
            check(_lib.lib.up3d_knn(B, N, S, int(nsample), int(Cd), ptr(xyz), ptr(new_xyz), ptr(idx), ptr(dist),
                                    stream_ptr()), 1)
    return (idx.long(), dist) if return_dist else idx.long()
