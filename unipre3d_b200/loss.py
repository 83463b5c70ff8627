"""Image losses of the render-loss path (/root/reference/utils/loss_utils.py:17-45,
/root/reference/train_network.py:260-302)."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr


def l1_loss(network_output, gt):
    return torch.abs((network_output - gt)).mean()


def l2_loss(network_output, gt):
    return ((network_output - gt) ** 2).mean()


def focal_l2_loss_torch(network_output, gt, bg_color, non_bg_color_loss_rate, bg_color_loss_rate):
    """Plain-torch statement of loss_utils.py:23-45 (used by tests as the fp32 reference of the fused kernel)."""
    base_loss = (network_output - gt) ** 2
    is_bg = (torch.isclose(gt[:, 0], bg_color[0], atol=1e-6) & torch.isclose(gt[:, 1], bg_color[1], atol=1e-6)
             & torch.isclose(gt[:, 2], bg_color[2], atol=1e-6))
    w_fg = 2 * non_bg_color_loss_rate / (bg_color_loss_rate + non_bg_color_loss_rate)
    w_bg = 2 * bg_color_loss_rate / (bg_color_loss_rate + non_bg_color_loss_rate)
    weights = torch.where(is_bg, w_bg, w_fg).unsqueeze(1)
    return (base_loss * weights).mean()


class _FocalL2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rendered, gt, bg, non_bg_rate, bg_rate):
        require_cuda(rendered, gt, bg)
        rendered, gt, bg = rendered.contiguous().float(), gt.contiguous().float(), bg.contiguous().float()
        n, c, H, W = rendered.shape
        assert c == 3 and gt.shape == rendered.shape
        out = torch.empty(4, dtype=torch.float32, device=rendered.device)
        dL = torch.empty_like(rendered)
        with torch.cuda.device(rendered.device):
            check(_lib.lib.up3d_focal_l2_loss(n, H, W, ptr(rendered), ptr(gt), ptr(bg), float(non_bg_rate),
                                              float(bg_rate), ptr(out), ptr(dL), stream_ptr()), launches=2)
        ctx.save_for_backward(dL)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (dL,) = ctx.saved_tensors
        return dL * g, None, None, None, None


def focal_l2_loss(network_output, gt, bg_color, non_bg_color_loss_rate, bg_color_loss_rate):
    """focal_l2_loss(rendered (n,3,H,W), gt, bg (3,), 4, 1) -> scalar; value and gradient from one fused kernel."""
    return _FocalL2.apply(network_output, gt, bg_color, non_bg_color_loss_rate, bg_color_loss_rate)
