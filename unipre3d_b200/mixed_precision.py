"""bf16 tensor-core GEMMs with fp32 master weights, without autocast's per-call weight casts.

`torch.autocast` re-casts every Linear weight/bias fp32->bf16 in each forward and casts every weight gradient
bf16->fp32 in each backward: ~14 tiny kernels per transformer block per step (224 of the ~1000 launches of the C2 step).
Here each selected nn.Linear keeps a persistent bf16 SHADOW of its parameters, refreshed for the whole model by ONE
multi-tensor copy after the optimizer step; the backward GEMM writes the weight gradient directly in fp32
(`torch.mm(..., out_dtype=float32)`), so the fp32 master gradient needs no cast either.  Arithmetic is the same as
autocast's (same bf16-rounded operands, fp32 accumulation); the weight gradient is rounded once less.
"""
from __future__ import annotations

import types
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


_ONES = {}


def _ones_row(n: int, device) -> torch.Tensor:
    key = (n, device)
    t = _ONES.get(key)
    if t is None:
        if torch.cuda.is_current_stream_capturing():
            return torch.ones((1, n), dtype=torch.bfloat16, device=device)
        t = _ONES[key] = torch.ones((1, n), dtype=torch.bfloat16, device=device)
    return t


class _ShadowLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, w16, b16):
        x16 = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
        ctx.save_for_backward(x16, w16)
        ctx.in_dtype, ctx.has_bias = x.dtype, bias is not None
        return F.linear(x16, w16, b16)

    @staticmethod
    def backward(ctx, gy):
        x16, w16 = ctx.saved_tensors
        gy = gy.contiguous() if gy.dtype == torch.bfloat16 else gy.to(torch.bfloat16)
        gy2 = gy.reshape(-1, gy.shape[-1])
        x2 = x16.reshape(-1, x16.shape[-1])
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = (gy2 @ w16).reshape(x16.shape)
            if ctx.in_dtype != torch.bfloat16:
                gx = gx.to(ctx.in_dtype)
        if ctx.needs_input_grad[1]:
            gw = torch.mm(gy2.t(), x2, out_dtype=torch.float32)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            # column sum as a (1 x T) @ (T x out) GEMM with fp32 accumulation/output: same arithmetic as
            # gy2.sum(0, dtype=float32) but ~4 us instead of torch's ~12 us strided bf16 reduction
            gb = torch.mm(_ones_row(gy2.shape[0], gy2.device), gy2, out_dtype=torch.float32).reshape(-1)
        return gx, gw, gb, None, None


def _shadow_forward(self, x):
    return _ShadowLinearFn.apply(x, self.weight, self.bias, self._w16, self._b16)


class ShadowWeights:
    """Patches every nn.Linear under `module` (instance-level forward) and owns the bf16 shadows."""

    def __init__(self, module: nn.Module):
        self.masters: List[torch.Tensor] = []
        self.shadows: List[torch.Tensor] = []
        for m in module.modules():
            if isinstance(m, nn.Linear) and m.weight.is_cuda:
                m._w16 = m.weight.detach().to(torch.bfloat16)
                m._b16 = None if m.bias is None else m.bias.detach().to(torch.bfloat16)
                m.forward = types.MethodType(_shadow_forward, m)
                self.masters.append(m.weight)
                self.shadows.append(m._w16)
                if m.bias is not None:
                    self.masters.append(m.bias)
                    self.shadows.append(m._b16)
        self._add_pointwise_convs(module)

    def _add_pointwise_convs(self, module: nn.Module) -> None:
        """bf16 shadows of the mini-PointNet's 1x1 convolutions (consumed as GEMM operands by fused_pointnet.py)."""
        from .backbone import Encoder
        for enc in module.modules():
            if not isinstance(enc, Encoder):
                continue
            for m in enc.modules():
                if isinstance(m, nn.Conv1d) and m.kernel_size == (1,) and m.weight.is_cuda and m.bias is not None:
                    m._w16 = m.weight.detach().to(torch.bfloat16)
                    m._b16 = m.bias.detach().to(torch.bfloat16)
                    self.masters += [m.weight, m.bias]
                    self.shadows += [m._w16, m._b16]

    @torch.no_grad()
    def refresh(self) -> None:
        if self.masters:
            torch._foreach_copy_(self.shadows, self.masters)
