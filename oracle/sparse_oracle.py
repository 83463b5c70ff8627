"""Dense-grid oracle of the sparse convolutions (TEST INFRASTRUCTURE ONLY; never imported by unipre3d_b200/).

The reference takes these layers from the un-vendored spconv v2 package (not in requirements.txt, absent from
/root/reference and from this image; call sites pointcept/models/sparse_unet/spconv_unet_v1m1_base.py:57-83, 153-160,
209-216, 247-253) and holds no test or golden vector for them: PARITY UNPINNED against spconv itself.  The oracle restates
spconv's published semantics on a densified grid with torch's own conv3d:
  * SubMConv3d(k):       conv3d(padding = k // 2) evaluated ONLY at the input's active voxels
  * SparseConv3d(2, 2):  conv3d(stride 2) evaluated at the voxels coord // 2 that have an active child
  * SparseInverseConv3d: out[i] = W[:, parity(i), :] @ in[parent(i)] at the original fine voxels
weights in spconv v2's (C_out, k, k, k, C_in) layout.  Operands are rounded to bf16 first (as the kernels do), products
and sums are fp64, so differences against the CUDA path are accumulation order only.
"""
from __future__ import annotations

import torch


def _bf16(x):
    return x.detach().to(torch.bfloat16).double()


def _densify(feats, idx, shape):
    B = int(idx[:, 0].max()) + 1
    C = feats.shape[1]
    g = torch.zeros((B, C, *shape), dtype=torch.float64)
    g[idx[:, 0].long(), :, idx[:, 1].long(), idx[:, 2].long(), idx[:, 3].long()] = feats
    return g


def subm_conv(feats, idx, weight, bias=None, round_bf16=True):
    """feats (n,Ci), idx (n,4) int (b,c0,c1,c2), weight (Co,k,k,k,Ci) -> (n,Co) at the same voxels."""
    f = _bf16(feats) if round_bf16 else feats.double()
    w = _bf16(weight) if round_bf16 else weight.double()
    k = weight.shape[1]
    shape = [int(idx[:, d].max()) + 1 for d in (1, 2, 3)]
    g = _densify(f, idx.cpu(), shape)
    out = torch.nn.functional.conv3d(g, w.permute(0, 4, 1, 2, 3), padding=k // 2)
    i = idx.cpu().long()
    o = out[i[:, 0], :, i[:, 1], i[:, 2], i[:, 3]]
    return o if bias is None else o + bias.double()


def down_conv(feats, idx, weight, round_bf16=True):
    """kernel 2 stride 2 -> (coarse idx (m,4) sorted by (b,c0,c1,c2), out (m,Co))."""
    f = _bf16(feats) if round_bf16 else feats.double()
    w = _bf16(weight) if round_bf16 else weight.double()
    i = idx.cpu().long()
    shape = [int(i[:, d].max()) // 2 * 2 + 2 for d in (1, 2, 3)]
    g = _densify(f, i, shape)
    out = torch.nn.functional.conv3d(g, w.permute(0, 4, 1, 2, 3), stride=2)
    coarse = torch.unique(torch.cat([i[:, :1], i[:, 1:] // 2], 1), dim=0)          # lexicographically sorted rows
    return coarse, out[coarse[:, 0], :, coarse[:, 1], coarse[:, 2], coarse[:, 3]]


def inverse_conv(coarse_feats, coarse_idx, fine_idx, weight, round_bf16=True):
    """-> (n_fine, Co): out[i] = W[:, parity(i), :] @ in[parent(i)]."""
    f = _bf16(coarse_feats) if round_bf16 else coarse_feats.double()
    w = _bf16(weight) if round_bf16 else weight.double()
    ci, fi = coarse_idx.cpu().long(), fine_idx.cpu().long()
    lut = {tuple(r.tolist()): n for n, r in enumerate(ci)}
    out = torch.zeros((fi.shape[0], weight.shape[0]), dtype=torch.float64)
    for n, r in enumerate(fi):
        p = lut[(int(r[0]), int(r[1]) // 2, int(r[2]) // 2, int(r[3]) // 2)]
        out[n] = w[:, int(r[1]) & 1, int(r[2]) & 1, int(r[3]) & 1, :] @ f[p]
    return out
