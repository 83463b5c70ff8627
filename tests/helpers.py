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


from unipre3d_b200.synthetic import make_camera, make_gaussians  # noqa: E402,F401  (shared with bench.py)


def oracle_scene(g, c, W, H, sh_degree=1, bg=(0, 0, 0), colors_precomp=None, antialiasing=True, scale_modifier=1.0):
    from oracle import oracle_lib as ol
    return ol.Scene(g["means3D"], g["opacities"], g["scales"], g["rotations"], c["view"], c["proj"], c["campos"], W, H,
                    c["tanfovx"], c["tanfovy"], shs=None if colors_precomp is not None else g["shs"],
                    colors_precomp=colors_precomp, sh_degree=sh_degree, bg=bg, antialiasing=antialiasing,
                    scale_modifier=scale_modifier)
