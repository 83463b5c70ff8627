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


def _gt_layout(gt):
    """(base tensor for the pointer, is_u8, views_per_object, object stride in elements) of a gt that is either
    (n,3,H,W) contiguous or a (B,V',3,H,W) slice of a contiguous (B,V,3,H,W) tensor along the view axis."""
    if gt.dim() == 4:
        g = gt.contiguous()
        return g, g, int(g.shape[0]) or 1, 0
    B, V, C, H, W = gt.shape
    st = gt.stride()
    if not (st[4] == 1 and st[3] == W and st[2] == H * W and st[1] == C * H * W):
        g = gt.contiguous()
        return g, g, V, V * C * H * W
    return gt, gt, V, int(st[0])


class _FocalL2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rendered, gt, bg, non_bg_rate, bg_rate):
        require_cuda(rendered, gt, bg)
        rendered, bg = rendered.contiguous().float(), bg.contiguous().float()
        n, c, H, W = rendered.shape
        if gt.dtype not in (torch.uint8, torch.float32):
            gt = gt.float()
        keep, g, vpo, ostride = _gt_layout(gt)
        assert c == 3 and gt.numel() == rendered.numel(), "gt and rendered images differ in size"
        out = torch.empty(4, dtype=torch.float32, device=rendered.device)
        dL = torch.empty_like(rendered)
        with torch.cuda.device(rendered.device):
            check(_lib.lib.up3d_focal_l2_loss_strided(n, H, W, ptr(rendered), ptr(g), int(g.dtype == torch.uint8), vpo, ostride,
                                                      ptr(bg), float(non_bg_rate), float(bg_rate), ptr(out), ptr(dL),
                                                      stream_ptr()), launches=2)
        ctx.save_for_backward(dL)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (dL,) = ctx.saved_tensors
        unit = _UNIT.get(dL.device)
        if unit is not None and g.data_ptr() == unit.data_ptr():
            return dL, None, None, None, None      # d(loss)/d(loss) = 1 handed in by backward_unit: no 25 MB multiply
        return dL * g, None, None, None, None


_UNIT = {}


def backward_unit(loss: torch.Tensor) -> None:
    """`loss.backward()` with the upstream gradient 1 passed as a cached constant tensor that `_FocalL2.backward`
    recognises by address: the (V,3,H,W) image gradient is then handed to the rasterizer as it left the loss kernel
    instead of being multiplied by one first (a 25 MB read + write pass in front of `blend_backward`)."""
    unit = _UNIT.get(loss.device)
    if unit is None:
        if loss.is_cuda and torch.cuda.is_current_stream_capturing():
            return loss.backward()
        unit = _UNIT[loss.device] = torch.ones((), dtype=torch.float32, device=loss.device)
    if loss.dtype != torch.float32 or loss.dim() != 0:
        return loss.backward()
    torch.autograd.backward(loss, grad_tensors=unit)


def focal_l2_loss(network_output, gt, bg_color, non_bg_color_loss_rate, bg_color_loss_rate):
    """focal_l2_loss(rendered (n,3,H,W), gt, bg (3,), 4, 1) -> scalar; value and gradient from one fused kernel.
    gt: (n,3,H,W) float32, or -- read in place, no copy -- a (B,V',3,H,W) view-axis slice of the batch's gt_images,
    float32 or uint8 (8-bit images are divided by 255 per read)."""
    return _FocalL2.apply(network_output, gt, bg_color, non_bg_color_loss_rate, bg_color_loss_rate)
