"""Drop-in for the external `diff_gaussian_rasterization` module the reference imports at
/root/reference/gaussian_renderer/__init__.py:8 (API generation with `antialiasing=` and a 3-tuple return,
lines 45-61 and 89-97).  Same names, field order, argument checks and return convention
(SURVEY.md §8b, Appendix A.10); the work is done by the batched sm_100a kernels with n_views = 1.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn as nn

from .rasterizer import rasterize_batch


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    antialiasing: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        if cov3D_precomp is not None:
            raise NotImplementedError("unipre3d_b200: cov3D_precomp is not on the reference's path "
                                      "(gaussian_renderer/__init__.py:66-72 always passes scales/rotations)")
        P = means3D.shape[0]
        color, radii, invdepth = rasterize_batch(
            means3D, opacities, scales, rotations, rs.viewmatrix.reshape(1, 4, 4), rs.projmatrix.reshape(1, 4, 4),
            rs.campos.reshape(1, 3), rs.bg, set_sizes=[P], views_per_set=[1], image_height=rs.image_height,
            image_width=rs.image_width, tanfovx=rs.tanfovx, tanfovy=rs.tanfovy, sh_degree=rs.sh_degree, shs=shs,
            colors_precomp=colors_precomp, means2D=means2D, scale_modifier=rs.scale_modifier,
            antialiasing=rs.antialiasing)
        return color[0], radii, invdepth[0]
