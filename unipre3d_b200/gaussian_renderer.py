"""`render_predicted` with the reference's signature and return dict
(/root/reference/gaussian_renderer/__init__.py:13-104), plus `render_batch_predicted`, the batched form the
B200 trainer uses: all (object, view) pairs of a step in one launch set instead of the Python double loop of
/root/reference/train_network.py:418-442.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Union

import numpy as np
import torch

from .camera import focal2fov
from .diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from .rasterizer import RasterLayout, rasterize_batch


def _image_size(cfg):
    if hasattr(cfg.data, "training_resolution"):
        return int(cfg.data.training_resolution), int(cfg.data.training_resolution)
    return cfg.data.training_height, cfg.data.training_width


def _tanfov(cfg, focals_pixels=None):
    if focals_pixels is None:
        t = math.tan(cfg.data.fov * np.pi / 360)
        return t, t
    # NB: the reference passes the fov itself (not fov/2) to tan here (gaussian_renderer/__init__.py:39-40);
    # every call site passes focals_pixels=None, so this branch is kept only for signature parity.
    return (math.tan(focal2fov(focals_pixels[0].item(), cfg.data.training_resolution)),
            math.tan(focal2fov(focals_pixels[1].item(), cfg.data.training_resolution)))


def render_predicted(pc: dict, world_view_transform, full_proj_transform, camera_center, bg_color: torch.Tensor, cfg,
                     scaling_modifier=1.0, override_color=None, focals_pixels=None):
    """Render one view of one Gaussian set.  Background tensor (bg_color) must be on GPU!"""
    screenspace_points = torch.zeros_like(pc["xyz"], dtype=pc["xyz"].dtype, requires_grad=True, device="cuda") + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    tanfovx, tanfovy = _tanfov(cfg, focals_pixels)
    image_height, image_width = _image_size(cfg)
    raster_settings = GaussianRasterizationSettings(
        image_height=image_height, image_width=image_width, tanfovx=tanfovx, tanfovy=tanfovy, bg=bg_color,
        scale_modifier=scaling_modifier, viewmatrix=world_view_transform, projmatrix=full_proj_transform,
        sh_degree=cfg.model.max_sh_degree, campos=camera_center, prefiltered=False, debug=False, antialiasing=True)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)
    shs, colors_precomp = None, None
    if override_color is None:
        if "features_rest" in pc.keys():
            shs = torch.cat([pc["features_dc"], pc["features_rest"]], dim=1).contiguous()
        else:
            shs = pc["features_dc"]
    else:
        colors_precomp = override_color
    rendered_image, radii, _ = rasterizer(means3D=pc["xyz"], means2D=screenspace_points, shs=shs,
                                          colors_precomp=colors_precomp, opacities=pc["opacity"],
                                          scales=pc["scaling"], rotations=pc["rotation"], cov3D_precomp=None)
    return {"render": rendered_image, "viewspace_points": screenspace_points, "visibility_filter": radii > 0,
            "radii": radii}


class _LazyVisibility:
    """`[radii > 0 for each object]`, evaluated on first use (the training step never reads it: 8 launches saved)."""

    def __init__(self, radii_list):
        self._radii, self._vis = radii_list, None

    def _get(self):
        if self._vis is None:
            self._vis = [r > 0 for r in self._radii]
        return self._vis

    def __getitem__(self, i):
        return self._get()[i]

    def __len__(self):
        return len(self._radii)

    def __iter__(self):
        return iter(self._get())


def render_batch_predicted(pc: Dict[str, Union[torch.Tensor, List[torch.Tensor]]], world_view_transforms,
                           full_proj_transforms, camera_centers, bg_color: torch.Tensor, cfg,
                           scaling_modifier=1.0, view_slice: slice = slice(None)):
    """All objects x views at once.

    pc: the dict `GaussianSplatPredictor.forward` returns -- tensors (B,P,...) at object level or lists of
    per-scene tensors (P_b,...) at scene level.  world_view_transforms / full_proj_transforms (B,V,4,4),
    camera_centers (B,V,3).  `view_slice` selects which of the V views are rendered (the trainer renders
    views [input_images:], train_network.py:427).  Returns {"render": (B,V',3,H,W), "radii": list per
    object of (V',P_b) int32, "visibility_filter": same as bool}.
    """
    wv = world_view_transforms[:, view_slice]
    fp = full_proj_transforms[:, view_slice]
    cc = camera_centers[:, view_slice]
    B, V = wv.shape[0], wv.shape[1]
    is_list = isinstance(pc["xyz"], (list, tuple))
    if is_list:
        sizes = [int(x.shape[0]) for x in pc["xyz"]]
        cat = lambda k: torch.cat([x.reshape(x.shape[0], *x.shape[1:]) for x in pc[k]], 0)
        xyz, op, sc, rot = cat("xyz"), cat("opacity"), cat("scaling"), cat("rotation")
        shs = torch.cat([torch.cat([d, r], 1) for d, r in zip(pc["features_dc"], pc["features_rest"])], 0) \
            if "features_rest" in pc else cat("features_dc")
    else:
        P = int(pc["xyz"].shape[1])
        sizes = [P] * B
        xyz, op = pc["xyz"].reshape(B * P, 3), pc["opacity"].reshape(B * P)
        sc, rot = pc["scaling"].reshape(B * P, 3), pc["rotation"].reshape(B * P, 4)
        if "shs" in pc:           # already concatenated by the fused splat head (gaussian_predictor._fused_head)
            shs = pc["shs"].reshape(B * P, -1, 3)
        elif "features_rest" in pc:
            shs = torch.cat([pc["features_dc"], pc["features_rest"]], dim=2).reshape(B * P, -1, 3)
        else:
            shs = pc["features_dc"].reshape(B * P, -1, 3)
    tanfovx, tanfovy = _tanfov(cfg)
    H, W = _image_size(cfg)
    dev = xyz.device
    color, radii, _ = rasterize_batch(
        xyz, op, sc, rot, wv.reshape(B * V, 4, 4).to(dev), fp.reshape(B * V, 4, 4).to(dev), cc.reshape(B * V, 3).to(dev),
        bg_color, set_sizes=sizes, views_per_set=[V] * B, image_height=H, image_width=W, tanfovx=tanfovx,
        tanfovy=tanfovy, sh_degree=cfg.model.max_sh_degree, shs=shs.contiguous(), scale_modifier=scaling_modifier,
        antialiasing=True, invdepth=False)
    lay = RasterLayout.get(sizes, [V] * B, dev)
    rs = lay.view_rec_start_host
    radii_list = [radii[rs[b * V]: rs[(b + 1) * V]].view(V, sizes[b]) for b in range(B)]
    return {"render": color.view(B, V, 3, H, W), "radii": radii_list,
            "visibility_filter": _LazyVisibility(radii_list)}
