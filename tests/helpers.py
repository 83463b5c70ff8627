"""Shared seeded input builders for the test-suite (numpy, CPU)."""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from unipre3d_b200 import camera as cam  # noqa: E402

FOV_DEG = 49.13434264120263  # /root/reference/configs/transformer_pretraining.yaml:11


def make_camera(az=30.0, el=20.0, dist=1.75, fov_deg=FOV_DEG, znear=0.5, zfar=2.0):
    proj = cam.get_projection_matrix(znear, zfar, math.radians(fov_deg), math.radians(fov_deg))
    R, t = cam.look_at_pose(az, el, dist)
    v = cam.make_view(R, t, proj)
    tanfov = math.tan(fov_deg * math.pi / 360)
    return dict(view=v["world_view_transform"].numpy().astype(np.float32),
                proj=v["full_proj_transform"].numpy().astype(np.float32),
                campos=v["camera_center"].numpy().astype(np.float32), tanfovx=tanfov, tanfovy=tanfov)


def make_gaussians(P, seed=0, regime="reference", sh_coeffs=4, spread=0.5):
    """regime 'reference': N(0,1) head outputs through the reference activations
    (model/gaussian_predictor.py:249-254,298-328) -> sigma >= e^-1 (dense regime).
    regime 'small': small splats (sort-light / blend-heavy), unit quaternions."""
    rng = np.random.default_rng(seed)
    centers = rng.normal(size=(P, 3))
    centers = centers / np.linalg.norm(centers, axis=1, keepdims=True) * rng.uniform(0, 1, (P, 1)) ** (1 / 3) * spread
    raw = rng.normal(size=(P, 23)).astype(np.float32)
    if regime == "reference":
        xyz = np.tanh(raw[:, 0:3]) * 0.1 + centers
        opacity = 1 / (1 + np.exp(-raw[:, 3]))
        scales = np.exp(np.clip(raw[:, 4:7], -1, 20))
        rot = raw[:, 7:11] / np.maximum(np.linalg.norm(raw[:, 7:11], axis=0, keepdims=True), 1e-6)  # over points!
    elif regime == "small":
        xyz = centers
        opacity = rng.uniform(0.05, 1.0, P)
        scales = np.exp(rng.normal(-3.5, 0.5, (P, 3)))
        rot = raw[:, 7:11] / np.linalg.norm(raw[:, 7:11], axis=1, keepdims=True)
    elif regime == "mid":
        xyz = centers
        opacity = rng.uniform(0.05, 1.0, P)
        scales = np.exp(rng.normal(-2.0, 0.7, (P, 3)))
        rot = raw[:, 7:11]
    else:
        raise ValueError(regime)
    shs = raw[:, 11:11 + 3 * min(sh_coeffs, 4)].reshape(P, -1, 3)
    if sh_coeffs > 4:
        shs = np.concatenate([shs, rng.normal(size=(P, sh_coeffs - 4, 3)) * 0.3], 1)
    shs = shs[:, :sh_coeffs]
    return dict(means3D=xyz.astype(np.float32), opacities=opacity.astype(np.float32), scales=scales.astype(np.float32),
                rotations=rot.astype(np.float32), shs=np.ascontiguousarray(shs.astype(np.float32)))


def oracle_scene(g, c, W, H, sh_degree=1, bg=(0, 0, 0), colors_precomp=None, antialiasing=True, scale_modifier=1.0):
    from oracle import oracle_lib as ol
    return ol.Scene(g["means3D"], g["opacities"], g["scales"], g["rotations"], c["view"], c["proj"], c["campos"], W, H,
                    c["tanfovx"], c["tanfovy"], shs=None if colors_precomp is not None else g["shs"],
                    colors_precomp=colors_precomp, sh_degree=sh_degree, bg=bg, antialiasing=antialiasing,
                    scale_modifier=scale_modifier)
