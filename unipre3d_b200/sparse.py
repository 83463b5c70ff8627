"""Sparse voxel tensors and sparse 3-D convolutions on B200 (host side of csrc/sparse_conv.cu).

Mirrors the slice of the spconv v2 API the reference's scene-level backbones use
(/root/reference/pointcept/models/sparse_unet/spconv_unet_v1m1_base.py:25-105, 150-253;
point_transformer_v3m1_base.py:281-287): `SparseConvTensor(features, indices, spatial_shape, batch_size)` with
`.features / .indices / .replace_feature`, `SubMConv3d`, `SparseConv3d(kernel 2, stride 2)`, `SparseInverseConv3d`
with `indice_key` sharing, `SparseSequential`.  Parameters keep spconv v2's layout and names
(`weight` of shape (C_out, k, k, k, C_in), optional `bias`), so state dicts written by the reference load unchanged.

Differences that do not change results: voxels are kept sorted by (batch, c0, c1, c2) -- `SparseConvTensor` sorts its
input once and remembers the permutation (`.perm`, `.features_in_input_order()`); spconv's row order is an artefact of
its hash table.  Rulebooks are int32 (kernel_volume, n_out) tables cached per `indice_key`.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr


def pack_keys(indices: torch.Tensor) -> torch.Tensor:
    """(n,4) int (batch, c0, c1, c2) -> int64 keys batch<<48 | c0<<32 | c1<<16 | c2."""
    i = indices.long()
    if i.numel() and (int(i.min()) < 0 or int(i[:, 1:].max()) > 65535 or int(i[:, 0].max()) > 32767):
        raise ValueError("voxel coordinates must lie in [0, 65535] (batch index in [0, 32767])")
    return (i[:, 0] << 48) | (i[:, 1] << 32) | (i[:, 2] << 16) | i[:, 3]


class SparseConvTensor:
    """features (n, C) fp32, indices (n, 4) int32 (batch, c0, c1, c2), sorted by key."""

    def __init__(self, features, indices, spatial_shape=None, batch_size=None, _sorted=False, _rules=None, _perm=None,
                 keep_order=False):
        """keep_order=True: rows stay in the caller's order (PTv3 keeps `point.feat` and the sparse tensor row-aligned);
        the sorted keys are kept on the side for rulebook construction.  Only SubMConv3d accepts such tensors."""
        self.keep_order = bool(keep_order)
        if keep_order:
            keys, perm = torch.sort(pack_keys(indices))
            if keys.numel() > 1 and bool((keys[1:] == keys[:-1]).any()):
                raise ValueError("duplicate voxel coordinates")
            self.perm, self.keys = perm, keys
            self.sorted_indices = indices.int()[perm].contiguous()
        elif not _sorted:
            keys = pack_keys(indices)
            keys, perm = torch.sort(keys)
            if keys.numel() > 1 and bool((keys[1:] == keys[:-1]).any()):
                raise ValueError("duplicate voxel coordinates")
            self.perm = perm
            features, indices = features[perm], indices[perm]
            self.keys = keys
        else:
            self.perm = _perm
            i = indices.long()                       # produced internally: already validated and sorted
            self.keys = (i[:, 0] << 48) | (i[:, 1] << 32) | (i[:, 2] << 16) | i[:, 3]
        self.features = features
        self.indices = indices.int().contiguous()
        if not hasattr(self, "keep_order"):
            self.keep_order = False
        self.spatial_shape, self.batch_size = spatial_shape, batch_size
        self.rules: Dict[str, dict] = _rules if _rules is not None else {}

    def replace_feature(self, features):
        t = SparseConvTensor.__new__(SparseConvTensor)
        t.__dict__.update(self.__dict__)
        t.features = features
        return t

    def features_in_input_order(self):
        """Rows in the order the constructor received them (inverse of the sort)."""
        if self.perm is None or self.keep_order:
            return self.features
        out = torch.empty_like(self.features)
        out[self.perm] = self.features
        return out

    @property
    def n(self) -> int:
        return int(self.features.shape[0])


# ----------------------------------------------------------------------------------------------- rulebooks
def subm_rulebook(x: SparseConvTensor, k: int) -> torch.Tensor:
    require_cuda(x.features)
    n = x.n
    nbr = torch.empty((k ** 3, n), dtype=torch.int32, device=x.features.device)
    sidx = x.sorted_indices if x.keep_order else x.indices
    with torch.cuda.device(x.features.device):
        check(_lib.lib.up3d_sparse_subm_rulebook(n, k, ptr(x.keys), ptr(sidx), ptr(nbr), stream_ptr()), launches=1)
    if x.keep_order:
        # sorted positions -> caller's rows: values through perm (-1 stays -1), columns scattered to row perm[s]
        perm_ext = torch.cat([x.perm.int(), x.perm.new_full((1,), -1).int()])
        rows = perm_ext[nbr.long()]
        nbr = torch.empty_like(rows)
        nbr[:, x.perm] = rows
    return nbr


def downsample_rule(x: SparseConvTensor):
    """kernel 2, stride 2, no padding: coarse voxel = coord // 2; returns (coarse indices (m,4), parent (n), parity (n),
    nbr_down (8, m), nbr_up (8, n))."""
    idx = x.indices
    coarse = torch.cat([idx[:, :1], idx[:, 1:] >> 1], 1)
    ckeys = pack_keys(coarse)
    ukeys, parent = torch.unique(ckeys, sorted=True, return_inverse=True)
    m = int(ukeys.numel())
    parity = ((idx[:, 1] & 1) * 4 + (idx[:, 2] & 1) * 2 + (idx[:, 3] & 1)).long()        # (a*2 + b)*2 + c
    n = x.n
    rows = torch.arange(n, device=idx.device, dtype=torch.int32)
    nbr_down = torch.full((8, m), -1, dtype=torch.int32, device=idx.device)
    nbr_down[parity, parent] = rows
    nbr_up = torch.full((8, n), -1, dtype=torch.int32, device=idx.device)
    nbr_up[parity, rows.long()] = parent.int()
    cidx = torch.stack([(ukeys >> 48) & 0xFFFF, (ukeys >> 32) & 0xFFFF, (ukeys >> 16) & 0xFFFF, ukeys & 0xFFFF], 1).int()
    return cidx, nbr_down, nbr_up


# ----------------------------------------------------------------------------------------------- the op
def _pad16(c: int) -> int:
    return (c + 15) // 16 * 16


def _conv_raw(n_out, nbr, feats, w_kio_bf16):
    """feats (n_in, Ci) fp32 contiguous, w (KV, Ci, Co) bf16 contiguous, channel counts multiples of 16."""
    KV, Ci, Co = w_kio_bf16.shape
    out = torch.empty((n_out, Co), dtype=torch.float32, device=feats.device)
    with torch.cuda.device(feats.device):
        check(_lib.lib.up3d_sparse_conv(n_out, Ci, Co, KV, ptr(nbr), ptr(feats), ptr(w_kio_bf16), ptr(out), stream_ptr()),
              launches=1)
    return out


class SparseConvFn(torch.autograd.Function):
    """out[i] = sum_o in[nbr_fwd[o][i]] @ W[o];  W given as (KV, C_in, C_out) fp32 (a view of the module parameter)."""

    @staticmethod
    def forward(ctx, feats, w_kio, nbr_fwd, nbr_bwd, n_out):
        require_cuda(feats, w_kio)
        KV, Ci, Co = w_kio.shape
        Cip, Cop = _pad16(Ci), _pad16(Co)
        f = feats.float()
        if Cip != Ci:
            f = torch.nn.functional.pad(f, (0, Cip - Ci))
        f = f.contiguous()
        w = w_kio.detach()
        if Cip != Ci or Cop != Co:
            w = torch.nn.functional.pad(w, (0, Cop - Co, 0, Cip - Ci))
        w16 = w.to(torch.bfloat16).contiguous()
        out = _conv_raw(n_out, nbr_fwd, f, w16)
        ctx.save_for_backward(f, w16, nbr_fwd, nbr_bwd)
        ctx.dims = (KV, Ci, Co, Cip, Cop, int(feats.shape[0]))
        return out[:, :Co] if Cop != Co else out

    @staticmethod
    def backward(ctx, dout):
        f, w16, nbr_fwd, nbr_bwd = ctx.saved_tensors
        KV, Ci, Co, Cip, Cop, n_in = ctx.dims
        n_out = int(dout.shape[0])
        d = dout.float()
        if Cop != Co:
            d = torch.nn.functional.pad(d, (0, Cop - Co))
        d = d.contiguous()
        din = dw = None
        if ctx.needs_input_grad[0]:
            din = _conv_raw(n_in, nbr_bwd, d, w16.transpose(1, 2).contiguous())[:, :Ci]
        if ctx.needs_input_grad[1]:
            dw = torch.zeros((KV, Cip, Cop), dtype=torch.float32, device=d.device)
            with torch.cuda.device(d.device):
                check(_lib.lib.up3d_sparse_conv_wgrad(n_out, Cip, Cop, KV, ptr(nbr_fwd), ptr(f), ptr(d), ptr(dw), stream_ptr()),
                      launches=1)
            dw = dw[:, :Ci, :Co]
        return din, dw, None, None, None


# ----------------------------------------------------------------------------------------------- modules
class _SparseConvBase(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, bias=True, indice_key=None):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, int(kernel_size)
        self.indice_key = indice_key
        k = self.kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, k, k, k, in_channels))      # spconv v2 layout
        nn.init.kaiming_uniform_(self.weight.view(out_channels, -1), a=5 ** 0.5)
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def _w_kio(self):
        k = self.kernel_size
        return self.weight.permute(1, 2, 3, 4, 0).reshape(k ** 3, self.in_channels, self.out_channels)

    def _finish(self, out):
        return out if self.bias is None else out + self.bias


class SubMConv3d(_SparseConvBase):
    """spconv.SubMConv3d: outputs only at the input's active voxels (stride / padding arguments are ignored there too)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, indice_key=None):
        super().__init__(in_channels, out_channels, kernel_size, bias, indice_key)

    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        k = self.kernel_size
        key = f"subm{k}:{self.indice_key}" if self.indice_key is not None else None
        rule = x.rules.get(key) if key else None
        if rule is None or rule["n"] != x.n:
            nbr = subm_rulebook(x, k)
            rule = {"n": x.n, "fwd": nbr, "bwd": torch.flip(nbr, dims=[0]).contiguous()}     # offset -o reads the mirror
            if key:
                x.rules[key] = rule
        out = SparseConvFn.apply(x.features, self._w_kio(), rule["fwd"], rule["bwd"], x.n)
        return x.replace_feature(self._finish(out))


class SparseConv3d(_SparseConvBase):
    """spconv.SparseConv3d restricted to what the reference uses: kernel 2, stride 2, no padding."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True, indice_key=None):
        if int(kernel_size) != 2 or int(stride) != 2 or int(padding) != 0:
            raise NotImplementedError("SparseConv3d: only kernel_size=2, stride=2, padding=0 (the reference's down-sampling)")
        super().__init__(in_channels, out_channels, kernel_size, bias, indice_key)

    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        if x.keep_order:
            raise NotImplementedError("SparseConv3d needs a sorted SparseConvTensor (keep_order=False)")
        cidx, nbr_down, nbr_up = downsample_rule(x)
        out = SparseConvFn.apply(x.features, self._w_kio(), nbr_down, nbr_up, int(cidx.shape[0]))
        shape = None if x.spatial_shape is None else [(s + 1) // 2 for s in x.spatial_shape]
        y = SparseConvTensor(self._finish(out), cidx, shape, x.batch_size, _sorted=True, _rules=x.rules, _perm=None)
        if self.indice_key is not None:
            x.rules["down:" + self.indice_key] = {"fine": x, "nbr_down": nbr_down, "nbr_up": nbr_up}
        return y


class SparseInverseConv3d(_SparseConvBase):
    """spconv.SparseInverseConv3d: back to the voxel set the SparseConv3d with the same indice_key came from."""

    def __init__(self, in_channels, out_channels, kernel_size, bias=True, indice_key=None):
        super().__init__(in_channels, out_channels, kernel_size, bias, indice_key)

    def forward(self, x: SparseConvTensor) -> SparseConvTensor:
        rule = x.rules.get("down:" + str(self.indice_key))
        if rule is None:
            raise RuntimeError(f"SparseInverseConv3d: no SparseConv3d with indice_key={self.indice_key!r} ran before")
        fine = rule["fine"]
        out = SparseConvFn.apply(x.features, self._w_kio(), rule["nbr_up"], rule["nbr_down"], fine.n)
        return fine.replace_feature(self._finish(out))


class SparseSequential(nn.Sequential):
    """spconv.SparseSequential: dense layers (BatchNorm1d, ReLU, ...) act on `.features`."""

    def forward(self, x):
        for m in self:
            if isinstance(m, (_SparseConvBase, SparseSequential)) or getattr(m, "is_sparse_module", False):
                x = m(x)
            elif isinstance(x, SparseConvTensor):
                x = x.replace_feature(m(x.features))
            else:
                x = m(x)
        return x
