"""Camera matrices in the reference's conventions.

Behavioural restatement of /root/reference/utils/graphics_utils.py:38-84 (getWorld2View2,
getView2World, getProjectionMatrix) and of how the ShapeNet loader assembles the per-view
transforms (/root/reference/dataset/shapenet.py:303-316): every 4x4 handed to the model or the
rasterizer is the TRANSPOSE of the column-vector matrix (row-vector convention), and
full_proj = world_view @ projection^T.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def get_world2view2(R: np.ndarray, t: np.ndarray, translate=(0.0, 0.0, 0.0), scale: float = 1.0) -> np.ndarray:
    """graphics_utils.py:38-50.  R is camera-to-world rotation, t the world-to-camera translation."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    C2W[:3, 3] = (C2W[:3, 3] + np.asarray(translate)) * scale
    return np.float32(np.linalg.inv(C2W))


def get_view2world(R: np.ndarray, t: np.ndarray, translate=(0.0, 0.0, 0.0), scale: float = 1.0) -> np.ndarray:
    """graphics_utils.py:52-62."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    C2W[:3, 3] = (C2W[:3, 3] + np.asarray(translate)) * scale
    return np.float32(C2W)


def get_projection_matrix(znear: float, zfar: float, fovX: float, fovY: float) -> torch.Tensor:
    """graphics_utils.py:64-84 (fov in radians)."""
    tanY, tanX = math.tan(fovY / 2), math.tan(fovX / 2)
    top, right = tanY * znear, tanX * znear
    bottom, left = -top, -right
    P = torch.zeros(4, 4)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def fov2focal(fov: float, pixels: float) -> float:
    return pixels / (2 * math.tan(fov / 2))


def focal2fov(focal: float, pixels: float) -> float:
    return 2 * math.atan(pixels / (2 * focal))


def look_at_pose(azimuth_deg: float, elevation_deg: float, distance: float):
    """Camera on a sphere looking at the origin (the geometry of shapenet.py:674-745: distance 1.75,
    azimuth sweep, elevation 0..90 deg).  Returns (R camera-to-world, t world-to-camera)."""
    az, el = math.radians(azimuth_deg), math.radians(elevation_deg)
    eye = distance * np.array([math.cos(el) * math.sin(az), math.cos(el) * math.cos(az), math.sin(el)])
    fwd = -eye / np.linalg.norm(eye)                   # camera +z looks at the origin
    up = np.array([0.0, 0.0, 1.0])
    if abs(np.dot(fwd, up)) > 0.999:
        up = np.array([0.0, 1.0, 0.0])
    right = np.cross(up, fwd)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd], axis=1)           # columns = camera axes in world coords
    t = -R.T @ eye
    return R, t


def make_view(R: np.ndarray, t: np.ndarray, projection: torch.Tensor):
    """shapenet.py:303-316 -> dict of the four per-view tensors the trainer consumes."""
    wv = torch.tensor(get_world2view2(R, t)).transpose(0, 1)
    vw = torch.tensor(get_view2world(R, t)).transpose(0, 1)
    full = wv.unsqueeze(0).bmm(projection.transpose(0, 1).unsqueeze(0)).squeeze(0)
    center = wv.inverse()[3, :3]
    return {"world_view_transform": wv, "view_to_world_transform": vw, "full_proj_transform": full,
            "camera_center": center}
