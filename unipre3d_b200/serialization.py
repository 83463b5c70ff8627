"""Point-cloud serialization with the reference's surface (/root/reference/pointcept/models/utils/serialization/default.py:
`encode(grid_coord, batch, depth, order)`; pointcept/models/utils/structure.py:47-107 `Point.serialization`): the first
stage of the scene-level PTv3 backbone (SURVEY.md §8a row P1 -- the rest of that row is not built yet).

Keys come from one integer kernel (`up3d_zorder_keys`, csrc/serialize.cu) instead of the reference's 6 table gathers + ORs
per order (`up3d_hilbert_keys` for the "hilbert" / "hilbert-trans" orders: Skilling's transpose in registers instead of the
reference's bit-plane tensors); the ordering is a library radix sort (`torch.sort`).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr


@torch.no_grad()
def encode(grid_coord: torch.Tensor, batch: Optional[torch.Tensor] = None, depth: int = 16, order: str = "z") -> torch.Tensor:
    """grid_coord (n,3) integer voxel coordinates, batch (n) or None -> (n,) int64 keys."""
    if order not in {"z", "z-trans", "hilbert", "hilbert-trans"}:
        raise AssertionError(order)
    require_cuda(grid_coord, batch)
    g = grid_coord.to(torch.int32).contiguous()
    b = None if batch is None else batch.to(torch.int64).contiguous()
    code = torch.empty(g.shape[0], dtype=torch.int64, device=g.device)
    with torch.cuda.device(g.device):
        fn = _lib.lib.up3d_hilbert_keys if order.startswith("hilbert") else _lib.lib.up3d_zorder_keys
        check(fn(g.shape[0], int(depth), int(order.endswith("-trans")), ptr(g), ptr(b), ptr(code), stream_ptr()), launches=1)
    return code


@torch.no_grad()
def serialization(grid_coord: torch.Tensor, batch: torch.Tensor, order: Sequence[str] = ("z", "z-trans"),
                  depth: Optional[int] = None, shuffle_orders: bool = False) -> Tuple[int, torch.Tensor, torch.Tensor, torch.Tensor]:
    """structure.py:47-107: -> (depth, serialized_code (k,n), serialized_order (k,n), serialized_inverse (k,n))."""
    if depth is None:
        depth = int(grid_coord.max()).bit_length()          # adaptive cube depth, as the reference (one D2H read)
    n_batches = int(batch.max()) + 1 if batch.numel() else 1
    assert depth * 3 + n_batches.bit_length() <= 63 and depth <= 16
    code = torch.stack([encode(grid_coord, batch, depth, o) for o in order])
    srt = torch.sort(code, dim=1, stable=True)
    order_idx = srt.indices
    inverse = torch.zeros_like(order_idx).scatter_(
        1, order_idx, torch.arange(code.shape[1], device=code.device).repeat(code.shape[0], 1))
    if shuffle_orders:
        perm = torch.randperm(code.shape[0])
        code, order_idx, inverse = code[perm], order_idx[perm], inverse[perm]
    return depth, code, order_idx, inverse
