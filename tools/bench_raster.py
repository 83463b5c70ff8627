"""Raster-only micro-benchmark (CUDA events on the launching stream): forward and backward of the batched
rasterizer on synthetic reference-regime Gaussians.  Usage: python tools/bench_raster.py [P] [B] [V] [R]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import make_camera, make_gaussians  # noqa: E402
from unipre3d_b200.rasterizer import rasterize_batch  # noqa: E402


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    V = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    R = int(sys.argv[4]) if len(sys.argv) > 4 else 256
    regime = sys.argv[5] if len(sys.argv) > 5 else "reference"
    dev = "cuda"
    gs = [make_gaussians(P, seed=100 + i, regime=regime) for i in range(B)]
    cat = {k: torch.tensor(np.concatenate([g[k] for g in gs], 0), device=dev).requires_grad_(True) for k in gs[0]}
    cams = [make_camera(az=360.0 * i / (B * V), el=5 + 2.0 * i) for i in range(B * V)]
    vm = torch.tensor(np.stack([c["view"] for c in cams]), device=dev)
    pm = torch.tensor(np.stack([c["proj"] for c in cams]), device=dev)
    cp = torch.tensor(np.stack([c["campos"] for c in cams]), device=dev)
    bg = torch.zeros(3, device=dev)
    kw = dict(set_sizes=[P] * B, views_per_set=[V] * B, image_height=R, image_width=R, tanfovx=cams[0]["tanfovx"],
              tanfovy=cams[0]["tanfovy"], sh_degree=1, shs=cat["shs"], invdepth=False)
    w = torch.randn(B * V, 3, R, R, device=dev)

    def step():
        color, radii, _ = rasterize_batch(cat["means3D"], cat["opacities"], cat["scales"], cat["rotations"], vm, pm, cp, bg, **kw)
        return color

    for _ in range(3):
        c = step(); c.backward(w)
    torch.cuda.synchronize()
    n = 20
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(n):
        e[0].record(); c = step(); e[1].record(); c.backward(w); e[2].record()
        torch.cuda.synchronize()
        tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
    with torch.no_grad():
        c = step()
    print(f"P={P} B={B} V={V} R={R} regime={regime}: fwd {tf / n:.3f} ms  bwd {tb / n:.3f} ms  "
          f"-> {B * V / ((tf + tb) / n) * 1e3:.0f} views/s (raster fwd+bwd only)  mean final colour {float(c.mean()):.4f}")


if __name__ == "__main__":
    main()
