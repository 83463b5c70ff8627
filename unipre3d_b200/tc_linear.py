"""Host wrapper of up3d_tc_linear (csrc/gemm_tc.cu): the backbone's nn.Linear GEMMs on tcgen05 tensor cores.

Arithmetic = nn.Linear of /root/reference/openpoints/models/backbone/transformer.py:22-33, 52-77 (y = x W^T + b, optional
erf-GELU) and the dX product of its backward (dx = dy W), bf16 operands with fp32 accumulation in TMEM.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr

B_KMAJOR, B_NMAJOR = 0, 1                   # B stored (N,K) [weight as nn.Linear keeps it] / (K,N)
EPI_NONE, EPI_GELU, EPI_GELU_BWD = 0, 1, 2


def supported(T: int, N: int, K: int) -> bool:
    return K % 8 == 0 and N % 64 == 0


def tc_linear(a, b, bias=None, *, b_major=B_KMAJOR, epilogue=EPI_NONE, aux_in=None, aux_out=None, out=None,
              out_dtype=torch.bfloat16, tile_n=0, b_dynamic=False):
    """a (T,K) bf16; b (N,K) [b_major 0] or (K,N) [b_major 1] bf16 -> out (T,N) bf16|fp32 (see include/up3d.h).
    b_dynamic: b is an activation produced on this stream (not a weight), so it must not be prefetched early."""
    require_cuda(a, b)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise RuntimeError("tc_linear: operands must be bfloat16")
    if not (a.is_contiguous() and b.is_contiguous()):
        raise RuntimeError("tc_linear: operands must be contiguous")
    T, K = a.shape
    N = b.shape[0] if b_major == B_KMAJOR else b.shape[1]
    if (b.shape[1] if b_major == B_KMAJOR else b.shape[0]) != K:
        raise RuntimeError(f"tc_linear: inner sizes differ ({tuple(a.shape)} x {tuple(b.shape)}, b_major={b_major})")
    if out is None:
        out = torch.empty((T, N), dtype=out_dtype, device=a.device)
    elif out.shape != (T, N) or not out.is_contiguous() or out.dtype not in (torch.bfloat16, torch.float32):
        raise RuntimeError("tc_linear: bad output tensor")
    if bias is not None and (bias.dtype != torch.bfloat16 or bias.numel() != N):
        raise RuntimeError("tc_linear: bias must be bfloat16 of N elements")
    if epilogue == EPI_GELU and aux_out is None:
        aux_out = torch.empty((T, N), dtype=torch.bfloat16, device=a.device)
    for t in (aux_in, aux_out):
        if t is not None and (t.shape != (T, N) or t.dtype != torch.bfloat16 or not t.is_contiguous()):
            raise RuntimeError("tc_linear: aux tensors must be contiguous (T,N) bfloat16")
    with torch.cuda.device(a.device):
        check(_lib.lib.up3d_tc_linear(T, N, K, ptr(a), ptr(b), b_major | (2 if b_dynamic else 0), ptr(bias), epilogue, ptr(aux_in), ptr(aux_out),
                                      ptr(out), 1 if out.dtype == torch.float32 else 0, tile_n, stream_ptr()), launches=1)
    return (out, aux_out) if epilogue == EPI_GELU else out
