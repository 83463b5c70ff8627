"""Synthetic ShapeNet-shaped batches (SURVEY.md §8d "synthetic inputs"): the dict the reference's ShapeNet
loader emits (/root/reference/dataset/shapenet.py:630-661, 530-535) without any dataset on disk.

Per sample: V = data.input_images + opt.imgs_per_obj views (shapenet.py:609-612), cameras on a sphere of radius
1.75 looking at the origin (shapenet.py:674-764), projection from znear/zfar/fov, all matrices transposed
(row-vector convention, shapenet.py:303-316); cloud = N points in a ball, mean-centred, max radius 0.5
(shapenet.py:480-492); GT images = U[0,1] inside a random disc on the exact background colour elsewhere
(so focal_l2 exercises both weights).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from . import camera as cam

CAMERA_DISTANCE = 1.75


def camera_pool(cfg, num: int = 24):
    """`num` poses: azimuth sweep, first half elevation 0-20 deg, second half 20-90 deg (shapenet.py:747-764)."""
    fov = math.radians(cfg.data.fov)
    proj = cam.get_projection_matrix(cfg.data.znear, cfg.data.zfar, fov, fov)
    half = num // 2
    az = np.linspace(-180, 180, half)
    poses = list(zip(az, np.linspace(0, 20, half))) + list(zip(az, np.linspace(20, 88, num - half)))
    return [cam.make_view(*cam.look_at_pose(a, e, CAMERA_DISTANCE), proj) for a, e in poses]


def make_cloud(n_points: int, rng: np.random.Generator, in_channels: int = 3) -> np.ndarray:
    d = rng.normal(size=(n_points, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p = d * rng.uniform(0, 1, (n_points, 1)) ** (1.0 / 3.0)
    p = p - p.mean(0, keepdims=True)
    p = p / np.linalg.norm(p, axis=1).max() * 0.5
    if in_channels == 4:  # height channel (shapenet.py:424-429)
        p = np.concatenate([p, p[:, 2:3] - p[:, 2:3].min()], 1)
    return p.astype(np.float32)


def make_batch(cfg, batch_size: int, n_points: int, seed: int = 0, pin: bool = False,
               image_dtype: str = "float32") -> Dict[str, torch.Tensor]:
    """image_dtype "float32": the dict the reference loader emits (images already divided by 255 on the host,
    shapenet.py:640-661).  "uint8": the same images as decoded from the dataset's 8-bit PNGs -- a quarter of the
    host->device bytes; Trainer divides by 255 on the device."""
    if image_dtype not in ("float32", "uint8"):
        raise ValueError(f"image_dtype {image_dtype!r}")
    rng = np.random.default_rng(seed)
    V = int(cfg.data.input_images) + int(cfg.opt.imgs_per_obj)
    R = int(cfg.data.training_resolution)
    pool = camera_pool(cfg)
    bgc = 1.0 if cfg.data.white_background else 0.0
    keys = ("world_view_transform", "view_to_world_transform", "full_proj_transform", "camera_center")
    out = {k + "s": [] for k in keys}
    clouds, gts = [], []
    yy, xx = np.mgrid[0:R, 0:R]
    for _ in range(batch_size):
        sel = rng.permutation(len(pool))[:V]
        for k in keys:
            out[k + "s"].append(torch.stack([pool[i][k] for i in sel]))
        clouds.append(make_cloud(n_points, rng, int(cfg.model.in_channels)))
        cx, cy, rad = rng.uniform(0.35 * R, 0.65 * R), rng.uniform(0.35 * R, 0.65 * R), rng.uniform(0.2 * R, 0.4 * R)
        disc = ((xx - cx) ** 2 + (yy - cy) ** 2) <= rad ** 2
        if image_dtype == "uint8":
            img = rng.integers(0, 256, (V, 3, R, R), dtype=np.uint8)
            gts.append(np.where(disc[None, None], img, np.uint8(255 * bgc)))
        else:
            img = rng.uniform(0, 1, (V, 3, R, R)).astype(np.float32)
            gts.append(np.where(disc[None, None], img, np.float32(bgc)))
    batch = {k: torch.stack(v).float() for k, v in out.items()}
    batch["gt_images"] = torch.from_numpy(np.stack(gts))
    batch["point_cloud"] = {"pos": torch.from_numpy(np.stack(clouds))}
    if pin and torch.cuda.is_available():
        batch = {k: ({kk: vv.pin_memory() for kk, vv in v.items()} if isinstance(v, dict) else v.pin_memory())
                 for k, v in batch.items()}
    return batch


def make_scene_batch(cfg, n_scenes: int, n_points: int, seed: int = 0, grid_size: float = 0.02,
                     room=(8.0, 6.0, 3.0)) -> Dict[str, torch.Tensor]:
    """ScanNet-shaped synthetic scenes: per scene `n_points` surface points of a room-sized box
    (floor, walls, a few boxes), voxelised at `grid_size` (first point per voxel kept, as GridSample's test mode),
    cameras inside the room looking around; the dict pointcept's collate emits ("coord", "grid_coord", "feat", "offset")
    plus the camera matrices / images of the object-level loader (row-vector, transposed matrices)."""
    rng = np.random.default_rng(seed)
    V = int(cfg.data.input_images) + int(cfg.opt.imgs_per_obj)
    H, W = (int(cfg.data.training_height), int(cfg.data.training_width)) if hasattr(cfg.data, "training_height") else \
        (int(cfg.data.training_resolution),) * 2
    fovx = math.radians(cfg.data.fov)
    fovy = 2 * math.atan(math.tan(fovx / 2) * H / W)
    proj_t = cam.get_projection_matrix(cfg.data.znear, cfg.data.zfar, fovx, fovy).transpose(0, 1)       # row-vector form
    rx, ry, rz = room
    coords, grids, feats, offsets, gts, v2w_all = [], [], [], [], [], []
    total = 0
    bgc = 1.0 if cfg.data.white_background else 0.0
    for _ in range(n_scenes):
        # surfaces: floor (40 %), four walls (40 %), three boxes (20 %)
        n_f, n_w = int(0.4 * n_points), int(0.4 * n_points)
        pts = [np.stack([rng.uniform(0, rx, n_f), rng.uniform(0, ry, n_f), np.zeros(n_f)], 1)]
        wall = rng.integers(0, 4, n_w)
        u, h = rng.uniform(0, 1, n_w), rng.uniform(0, rz, n_w)
        wx = np.where(wall == 0, 0.0, np.where(wall == 1, rx, u * rx))
        wy = np.where(wall == 2, 0.0, np.where(wall == 3, ry, u * ry))
        pts.append(np.stack([wx, wy, h], 1))
        n_b = n_points - n_f - n_w
        centre = rng.uniform([1, 1, 0.3], [rx - 1, ry - 1, 0.9], (3, 3))[rng.integers(0, 3, n_b)]
        pts.append(centre + rng.uniform(-0.4, 0.4, (n_b, 3)) * np.array([1, 1, 0.7]))
        p = np.concatenate(pts, 0).astype(np.float32)
        g = np.floor(p / grid_size).astype(np.int64)
        g -= g.min(0, keepdims=True)
        _, first = np.unique((g[:, 0] * 4096 + g[:, 1]) * 4096 + g[:, 2], return_index=True)
        first.sort()
        p, g = p[first], g[first]
        color = rng.uniform(0, 1, (p.shape[0], 3)).astype(np.float32)
        coords.append(p)
        grids.append(g.astype(np.int32))
        feats.append(np.concatenate([color, p / np.array(room, np.float32)], 1).astype(np.float32))     # 6 input channels
        total += p.shape[0]
        offsets.append(total)
        views = []
        for _v in range(V):
            R, t = cam.look_at_pose(rng.uniform(-180, 180), rng.uniform(-5, 25), 1.0)     # orientation only
            v2w = torch.tensor(cam.get_view2world(R, t)).transpose(0, 1).float().clone()
            v2w[3, :3] = torch.tensor(rng.uniform([2, 2, 1.2], [rx - 2, ry - 2, 1.8]), dtype=torch.float32)   # eye in the room
            views.append(v2w)
        v2w_all.append(torch.stack(views))
        gts.append(np.where(rng.uniform(0, 1, (V, 1, H, W)) < 0.7, rng.uniform(0, 1, (V, 3, H, W)), bgc).astype(np.float32))
    v2w = torch.stack(v2w_all)
    w2v = torch.linalg.inv(v2w)
    batch = {"view_to_world_transforms": v2w, "world_view_transforms": w2v, "full_proj_transforms": w2v @ proj_t,
             "camera_centers": v2w[:, :, 3, :3].clone(), "gt_images": torch.from_numpy(np.stack(gts)),
             "point_cloud": {"coord": torch.from_numpy(np.concatenate(coords)),
                             "grid_coord": torch.from_numpy(np.concatenate(grids)),
                             "feat": torch.from_numpy(np.concatenate(feats)),
                             "offset": torch.tensor(offsets, dtype=torch.int64)}}
    return batch


def batch_nbytes(batch) -> int:
    n = 0
    for v in batch.values():
        if isinstance(v, dict):
            n += sum(t.numel() * t.element_size() for t in v.values())
        else:
            n += v.numel() * v.element_size()
    return n


# ---------------------------------------------------------------------------------------------------------------------
# Raster-only inputs: seeded Gaussians + one camera, shared by the parity tests and bench.py's raster leg.
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
