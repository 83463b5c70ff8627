"""The tokenizer's per-group mini-PointNet (backbone.Encoder == /root/reference/openpoints/models/backbone/
transformer.py:210-243) as ONE autograd node on B200, train mode.

    x (B,3,G,K) -> Conv1d(3,128)+BN+ReLU [fused, K=3 layer computed from the input] -> Conv1d(128,256) [GEMM] -> f
      fg = max_K f [group_max]                         (transformer.py:235)
      [fg || f] -> Conv1d(512,512): W3 [fg || f] = W3g fg + W3l f -- the concat (236) is never materialised: the global
      half is a (B*G)-row GEMM broadcast over the K rows inside the BatchNorm passes, the local half a K=256 GEMM
      -> BN + ReLU [gbn_stats / finalize / gbn_apply_relu] -> Conv1d(512,C) [GEMM] -> max_K [group_max] -> (B,G,C)

The four dense GEMMs (and their 2 backward GEMMs each) are library calls; the memory-bound passes are
csrc/pointnet.cu.  ~30 launches per step instead of ~80, no 33 MB concat, no separate ReLU / cast / bias-gradient
reductions.  Train-mode BatchNorm semantics are nn.BatchNorm1d's (batch statistics over all B*G*K rows, running
statistics updated with momentum and the unbiased variance; SyncBatchNorm: the partial sums are all-reduced).
Conv biases in front of a train-mode BatchNorm receive an exactly zero gradient (the reference computes rounding
noise there).  Eval mode and CPU tensors use the module path (backbone.Encoder.forward).
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr
from .fused_encoder import SideStream, _flag, _wgrad

_NUM_SMS = 148


def _world(bn) -> int:
    if isinstance(bn, torch.nn.SyncBatchNorm) and dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


def _reduce(partials, C):
    """per-CTA partial sums (n, rows, C) -> sums (rows, C), rows = 1 or 2, fp64 accumulation."""
    rows = partials.shape[1]
    sums = torch.empty((rows, C), dtype=torch.float32, device=partials.device)
    check(_lib.lib.up3d_bn_reduce_sums(partials.shape[0], rows, C, ptr(partials), ptr(sums), stream_ptr()), 1)
    return sums


def _bn_forward_stats(partials, C, rows_per_partial, rows, bn, world):
    """shifted partials (n,3,C) -> stats (4,C) + running-statistics update.  SyncBatchNorm: every rank reduces its rows to
    (mean, M2), the triples are all-gathered and merged again (the same pairwise fp64 merge)."""
    L = _lib.lib
    stats = torch.empty((4, C), dtype=torch.float32, device=partials.device)
    mom = 0.1 if bn.momentum is None else float(bn.momentum)
    n = partials.shape[0]
    if world > 1:
        local = torch.empty((3, C), dtype=torch.float32, device=partials.device)
        check(L.up3d_bn_reduce_finalize(n, C, ptr(partials), rows_per_partial, rows, None, None, 0.0, 0.0, None, None, None,
                                        ptr(local), None, stream_ptr()), 1)
        partials = torch.empty((world, 3, C), dtype=torch.float32, device=local.device)
        dist.all_gather_into_tensor(partials, local)
        n, rows_per_partial, rows = world, rows, rows * world
    check(L.up3d_bn_reduce_finalize(n, C, ptr(partials), rows_per_partial, rows, ptr(bn.weight), ptr(bn.bias), float(bn.eps),
                                    mom, ptr(bn.running_mean), ptr(bn.running_var), ptr(bn.num_batches_tracked), None,
                                    ptr(stats), stream_ptr()), 1)
    return stats


def _bn_backward_sums(partials, C, world):
    """-> (local sums (2,C): the dbeta / dgamma of this rank, sums used to normalise dz: all-reduced under SyncBatchNorm
    -- torch.nn.SyncBatchNorm likewise all-reduces sum_dy / sum_dy_xmu but keeps the weight gradients local)."""
    local = _reduce(partials, C)
    if world == 1:
        return local, local
    glob = local.clone()
    dist.all_reduce(glob)
    return local, glob


class _PNMeta:
    def __init__(self, B, G, K, act, bn1, bn2, weights):
        self.B, self.G, self.K, self.act, self.bn1, self.bn2 = B, G, K, act, bn1, bn2
        self.weights = weights          # (W2, b2, W3, W4, b4) in the GEMM operand dtype, 2-D


class MiniPointNetFn(torch.autograd.Function):
    """params: W1 (128,3,1), b1, g1, be1, W2 (256,128,1), b2, W3 (512,512,1), b3, g2, be2, W4 (C,512,1), b4."""

    @staticmethod
    def forward(ctx, nb, meta: _PNMeta, *params):
        W1, b1, g1, be1, W2, b2, W3, b3, g2, be2, W4, b4 = params
        B, G, K, act = meta.B, meta.G, meta.K, meta.act
        require_cuda(nb)
        nb = nb.contiguous().float()
        Gt, R, GK = B * G, B * G * K, G * K
        C1, C2, C3, C4 = W1.shape[0], W2.shape[0], W3.shape[0], W4.shape[0]
        if C1 != 128 or W3.shape[1] != 2 * C2:
            raise RuntimeError("fused mini-PointNet: expects the reference layout Conv1d(3,128) ... Conv1d(2*C2, C3)")
        W2c, b2c, W3c, W4c, b4c = meta.weights
        dev, L, fl = nb.device, _lib.lib, _flag(act)
        w1, wd1, wd2 = W1.detach().reshape(C1, 3).contiguous(), _world(meta.bn1), _world(meta.bn2)
        with torch.cuda.device(dev), torch.autocast("cuda", enabled=False):
            # ---- layer 1 (+BN+ReLU), computed from the 3-channel input
            tile = int(L.up3d_pn_stats_tile_rows())
            part1 = torch.empty(((R + tile - 1) // tile, 3, C1), dtype=torch.float32, device=dev)
            check(L.up3d_pn_conv1_stats(R, GK, ptr(nb), ptr(w1), ptr(b1), ptr(part1), stream_ptr()), 1)
            stats1 = _bn_forward_stats(part1, C1, tile, R, meta.bn1, wd1)
            y1 = torch.empty((R, C1), dtype=act, device=dev)
            check(L.up3d_pn_conv1_bn_relu(fl, R, GK, ptr(nb), ptr(w1), ptr(b1), ptr(stats1), ptr(y1), stream_ptr()), 1)
            # ---- layer 2 + group max
            f = F.linear(y1, W2c, b2c)                                              # (R, C2)
            fg = torch.empty((Gt, C2), dtype=act, device=dev)
            arg2 = torch.empty((Gt, C2), dtype=torch.int32, device=dev)
            check(L.up3d_group_max(fl, Gt, K, C2, ptr(f), ptr(fg), ptr(arg2), stream_ptr()), 1)
            # ---- layer 3 on [global || local] without the concat
            W3g, W3l = W3c[:, :C2], W3c[:, C2:]
            if act == torch.float32:
                gpart = fg @ W3g.t()
            else:
                gpart = torch.mm(fg, W3g.t(), out_dtype=torch.float32)              # (Gt, C3) fp32
            zl = f @ W3l.t()                                                        # (R, C3)
            # groups per CTA of the group-tile passes.  bf16: one wave of CTAs, each streaming its groups through a
            # shared-memory ring (csrc/pointnet.cu, STAGED); fp32: short CTAs with direct loads
            gpc = max(1, -(-Gt // _NUM_SMS)) if act == torch.bfloat16 else (2 if Gt >= 4 * _NUM_SMS else 1)
            n2 = (Gt + gpc - 1) // gpc
            part2 = torch.empty((n2, 3, C3), dtype=torch.float32, device=dev)
            check(L.up3d_gbn_stats(fl, Gt, K, C3, gpc, ptr(zl), ptr(gpart), ptr(b3), ptr(part2), stream_ptr()), 1)
            stats2 = _bn_forward_stats(part2, C3, gpc * K, R, meta.bn2, wd2)
            y3 = torch.empty((R, C3), dtype=act, device=dev)
            check(L.up3d_gbn_apply_relu(fl, Gt, K, C3, gpc, ptr(zl), ptr(gpart), ptr(b3), ptr(stats2), ptr(y3), stream_ptr()), 1)
            # ---- layer 4 + group max
            z4 = F.linear(y3, W4c, b4c)                                             # (R, C4)
            tok = torch.empty((Gt, C4), dtype=act, device=dev)
            arg4 = torch.empty((Gt, C4), dtype=torch.int32, device=dev)
            check(L.up3d_group_max(fl, Gt, K, C4, ptr(z4), ptr(tok), ptr(arg4), stream_ptr()), 1)
        ctx.meta, ctx.gpc, ctx.dims = meta, gpc, (Gt, R, GK, C1, C2, C3, C4)
        ctx.save_for_backward(nb, w1, b1, b3, stats1, y1, f, fg, arg2, gpart, zl, stats2, y3, arg4)
        return tok.view(B, G, C4)

    @staticmethod
    def backward(ctx, dtok):
        meta, gpc = ctx.meta, ctx.gpc
        Gt, R, GK, C1, C2, C3, C4 = ctx.dims
        K, act = meta.K, meta.act
        nb, w1, b1, b3, stats1, y1, f, fg, arg2, gpart, zl, stats2, y3, arg4 = ctx.saved_tensors
        W2c, b2c, W3c, W4c, b4c = meta.weights
        dev, L, fl = dtok.device, _lib.lib, _flag(act)
        wd1, wd2 = _world(meta.bn1), _world(meta.bn2)
        with torch.cuda.device(dev), torch.autocast("cuda", enabled=False):
            dt = dtok.reshape(Gt, C4)
            gb4 = dt.sum(0, dtype=torch.float32)
            dt = dt.contiguous().to(act)
            # ---- layer 4
            dz4 = torch.empty((R, C4), dtype=act, device=dev)
            check(L.up3d_group_max_scatter(fl, Gt, K, C4, ptr(dt), ptr(arg4), ptr(dz4), stream_ptr()), 1)
            side = SideStream(dev)                      # weight-gradient GEMMs overlap the dX chain
            dy3 = dz4 @ W4c
            gW4 = side.run(_wgrad, dz4, y3)
            # ---- BN2 + ReLU backward (+ group sums for the global half)
            n2 = (Gt + gpc - 1) // gpc
            part = torch.empty((n2, 2, C3), dtype=torch.float32, device=dev)
            check(L.up3d_gbn_bwd_reduce(fl, Gt, K, C3, gpc, ptr(dy3), ptr(zl), ptr(gpart), ptr(b3), ptr(stats2), ptr(part),
                                        stream_ptr()), 1)
            sums2, sums2g = _bn_backward_sums(part, C3, wd2)
            dz3 = torch.empty((R, C3), dtype=act, device=dev)
            dgp = torch.empty((Gt, C3), dtype=act, device=dev)
            check(L.up3d_gbn_bwd_apply(fl, Gt, K, C3, gpc, ptr(dy3), ptr(zl), ptr(gpart), ptr(b3), ptr(stats2), ptr(sums2g),
                                       float(R * wd2), ptr(dz3), ptr(dgp), stream_ptr()), 1)
            # ---- layer 3 (local + global halves)
            W3g, W3l = W3c[:, :C2], W3c[:, C2:]
            dfl = dz3 @ W3l
            dfg = dgp @ W3g
            gW3 = side.run(lambda a, b, c, d: torch.cat([_wgrad(a, b), _wgrad(c, d)], dim=1), dgp, fg, dz3, f)
            # ---- group max backward + layer 2
            gb2p = torch.empty(((Gt + gpc - 1) // gpc, 1, C2), dtype=torch.float32, device=dev)
            df = torch.empty((R, C2), dtype=act, device=dev)
            check(L.up3d_group_combine(fl, Gt, K, C2, gpc, ptr(dfl), ptr(dfg), ptr(arg2), ptr(df), ptr(gb2p), stream_ptr()), 1)
            gb2 = _reduce(gb2p, C2)[0]
            dy1 = df @ W2c
            gW2 = side.run(_wgrad, df, y1)
            # ---- BN1 + ReLU + layer 1 backward
            n1 = min((R + 127) // 128, 2 * _NUM_SMS)
            part1 = torch.empty((n1, 2, C1), dtype=torch.float32, device=dev)
            check(L.up3d_pn_conv1_bwd(fl, 0, R, GK, ptr(nb), ptr(w1), ptr(b1), ptr(stats1), ptr(dy1), None, 0.0, ptr(part1),
                                      n1, None, None, stream_ptr()), 1)
            sums1, sums1g = _bn_backward_sums(part1, C1, wd1)
            g1p = torch.empty((n1, 2, 2 * C1), dtype=torch.float32, device=dev)
            check(L.up3d_pn_conv1_bwd(fl, 1, R, GK, ptr(nb), ptr(w1), ptr(b1), ptr(stats1), ptr(dy1), ptr(sums1g),
                                      float(R * wd1), ptr(g1p), n1, None, None, stream_ptr()), 1)
            g1 = _reduce(g1p, 2 * C1).view(4, C1)              # rows: dW[:,0], dW[:,1], dW[:,2], db
            gW1, gb1 = g1[:3].t().contiguous(), g1[3]
            gb3 = torch.zeros(C3, dtype=torch.float32, device=dev)       # bias in front of a train-mode BatchNorm
            side.join()
        return (None, None, gW1.view(C1, 3, 1), gb1, sums1[1], sums1[0], gW2.view(C2, C1, 1), gb2, gW3.view(C3, 2 * C2, 1),
                gb3, sums2[1], sums2[0], gW4.view(C4, C3, 1), gb4)


def supports(encoder) -> bool:
    fc, sc = encoder.first_conv, encoder.second_conv
    try:
        ok = (fc[0].in_channels == 3 and fc[0].out_channels == 128 and fc[0].kernel_size == (1,)
              and isinstance(fc[1], torch.nn.modules.batchnorm._BatchNorm) and fc[1].affine and fc[1].track_running_stats
              and isinstance(sc[1], torch.nn.modules.batchnorm._BatchNorm) and sc[1].affine and sc[1].track_running_stats
              and sc[0].in_channels == 2 * fc[3].out_channels and fc[3].in_channels == 128
              and all(c.bias is not None for c in (fc[0], fc[3], sc[0], sc[3]))
              and all(c.out_channels % 8 == 0 for c in (fc[3], sc[0], sc[3])))
    except (AttributeError, IndexError):
        return False
    return bool(ok)


def _compute_copy(conv, act):
    w = conv.weight.detach()
    if act == torch.float32:
        return w.reshape(w.shape[0], w.shape[1]), conv.bias.detach()
    w16 = getattr(conv, "_w16", None)
    if w16 is None:
        return w.reshape(w.shape[0], w.shape[1]).to(act), conv.bias.detach().to(act)
    return w16.view(w.shape[0], w.shape[1]), conv._b16


def run_mini_pointnet(encoder, neighborhood):
    """encoder: backbone.Encoder (train mode); neighborhood (B,3,G,K) fp32 CUDA -> tokens (B,G,C)."""
    B, _, G, K = neighborhood.shape
    act = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else torch.float32
    if act not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"fused mini-PointNet: unsupported autocast dtype {act}")
    fc, sc = encoder.first_conv, encoder.second_conv
    W2c, b2c = _compute_copy(fc[3], act)
    W3c, _ = _compute_copy(sc[0], act)
    W4c, b4c = _compute_copy(sc[3], act)
    meta = _PNMeta(B, G, K, act, fc[1], sc[1], (W2c, b2c, W3c, W4c, b4c))
    return MiniPointNetFn.apply(neighborhood, meta, fc[0].weight, fc[0].bias, fc[1].weight, fc[1].bias, fc[3].weight,
                                fc[3].bias, sc[0].weight, sc[0].bias, sc[1].weight, sc[1].bias, sc[3].weight, sc[3].bias)
