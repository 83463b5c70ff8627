"""CPU port of one pre-training step of the render-loss path (TEST INFRASTRUCTURE / reported baseline ONLY).

Used by bench.py's `cpu_baseline` leg and `--impl reference` arm (kind "port") and by smoke()/tests as a checker.
The reference itself has NO CPU path for this step: its tokenizer kernels are CUDA-only
(openpoints/models/layers/subsample.py:93-100, group.py:93,156,194 allocate torch.cuda tensors) and its rasterizer is
an external CUDA extension.  This port therefore combines
  * the CPU restatement of the point-op kernels (oracle/pointops_oracle.c),
  * the backbone / head nn.Modules run by torch on the host cores (fp32, all threads), and
  * the C/OpenMP restatement of the rasterizer forward + backward (oracle/raster_oracle.c), called once per
    (object, view) exactly like the reference's Python double loop (train_network.py:418-442),
  * the plain-torch focal-L2 loss, clip_grad_norm_(1.0) and AdamW (train_network.py:156,386).
"""
from __future__ import annotations

import numpy as np
import torch

from . import oracle_lib as ol


class OracleSubsampleGroup(torch.nn.Module):
    """SubsampleGroup (group_embed.py:39-57) on the CPU oracle kernels."""

    def __init__(self, num_groups, group_size, radius):
        super().__init__()
        self.num_groups, self.group_size, self.radius = num_groups, group_size, radius

    def forward(self, p, x=None):
        pn = p.detach().cpu().numpy().astype(np.float32)
        fidx = ol.fps(pn, self.num_groups)
        center = np.take_along_axis(pn, fidx.astype(np.int64)[..., None], 1)
        idx = ol.ball_query(self.radius, self.group_size, pn, center)
        grouped = ol.group(np.ascontiguousarray(pn.transpose(0, 2, 1)), idx)
        neigh = grouped - center.transpose(0, 2, 1)[..., None]
        return torch.from_numpy(neigh), torch.from_numpy(center)


class _OracleRender(torch.autograd.Function):
    """One view through the C oracle (forward + hand-derived backward)."""

    @staticmethod
    def forward(ctx, xyz, opacity, scaling, rotation, shs, view, proj, campos, bg, W, H, tanfov, deg):
        sc = ol.Scene(xyz.detach().numpy(), opacity.detach().numpy(), scaling.detach().numpy(),
                      rotation.detach().numpy(), view.numpy(), proj.numpy(), campos.numpy(), W, H, tanfov, tanfov,
                      shs=shs.detach().numpy(), sh_degree=deg, bg=bg.numpy())
        ctx.sc = sc
        ctx.shapes = (opacity.shape, shs.shape)
        return torch.from_numpy(ol.render(sc)["color"])

    @staticmethod
    def backward(ctx, g):
        r = ol.render(ctx.sc, g.contiguous().numpy())["grads"]
        t = torch.from_numpy
        return (t(r["means3D"]), t(r["opacities"]).reshape(ctx.shapes[0]), t(r["scales"]), t(r["rotations"]),
                t(r["shs"]).reshape(ctx.shapes[1]), None, None, None, None, None, None, None, None)


def build_cpu_model(cfg, state_dict=None):
    """The product's nn.Modules (plain torch layers) on the CPU with the tokenizer swapped for the oracle kernels."""
    from unipre3d_b200.gaussian_predictor import GaussianSplatPredictor
    model = GaussianSplatPredictor(cfg)
    if state_dict is not None:
        model.load_state_dict(state_dict)
    gd = model.point_network.encoder.group_divider
    model.point_network.encoder.group_divider = OracleSubsampleGroup(gd.num_groups, gd.group_size, gd.radius)
    return model


def focal_l2_torch(r, gt, bg, non_bg_rate, bg_rate):
    base = (r - gt) ** 2
    is_bg = (torch.isclose(gt[:, 0], bg[0], atol=1e-6) & torch.isclose(gt[:, 1], bg[1], atol=1e-6)
             & torch.isclose(gt[:, 2], bg[2], atol=1e-6))
    w = torch.where(is_bg, 2 * bg_rate / (bg_rate + non_bg_rate), 2 * non_bg_rate / (bg_rate + non_bg_rate))
    return (base * w.unsqueeze(1)).mean()


def forward_loss(model, cfg, data):
    """Forward + loss of one batch dict on the CPU (differentiable).  Returns (loss, rendered, splats)."""
    import math
    ni = int(cfg.data.input_images)
    splats = model(data["point_cloud"], data["gt_images"][:, :ni], data["view_to_world_transforms"][:, :ni])
    R = int(cfg.data.training_resolution)
    tanfov = math.tan(cfg.data.fov * math.pi / 360)
    bg = torch.tensor([1.0, 1.0, 1.0] if cfg.data.white_background else [0.0, 0.0, 0.0])
    imgs, gts = [], []
    for b in range(data["gt_images"].shape[0]):
        shs = torch.cat([splats["features_dc"][b], splats["features_rest"][b]], 1)
        for r in range(ni, data["gt_images"].shape[1]):
            imgs.append(_OracleRender.apply(splats["xyz"][b], splats["opacity"][b], splats["scaling"][b],
                                            splats["rotation"][b], shs, data["world_view_transforms"][b, r],
                                            data["full_proj_transforms"][b, r], data["camera_centers"][b, r], bg,
                                            R, R, tanfov, int(cfg.model.max_sh_degree)))
            gts.append(data["gt_images"][b, r])
    rendered = torch.stack(imgs)
    loss = focal_l2_torch(rendered, torch.stack(gts), bg, cfg.opt.non_bg_color_loss_rate, cfg.opt.bg_color_loss_rate)
    return loss, rendered, splats


class CpuStepper:
    def __init__(self, cfg, state_dict=None):
        self.cfg = cfg
        self.model = build_cpu_model(cfg, state_dict)
        self.opt = torch.optim.AdamW([p for p in self.model.parameters() if p.requires_grad], lr=cfg.opt.base_lr,
                                     eps=1e-15, betas=tuple(cfg.opt.betas))

    def step(self, data) -> float:
        self.model.train()
        loss, _, _ = forward_loss(self.model, self.cfg, data)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm=1.0)
        self.opt.step()
        self.opt.zero_grad()
        return float(loss)
