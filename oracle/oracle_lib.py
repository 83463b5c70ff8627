"""ctypes front-end of the CPU oracles (oracle/raster_oracle.c, oracle/pointops_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package `unipre3d_b200`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libup3d_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("raster_oracle.c", "pointops_oracle.c", "Makefile")]
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _SO


class _Scene(C.Structure):
    _fields_ = [("P", C.c_int), ("M", C.c_int), ("D", C.c_int), ("W", C.c_int), ("H", C.c_int),
                ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
                ("antialiasing", C.c_int),
                ("bg", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p),
                ("opacities", C.c_void_p), ("scales", C.c_void_p), ("rotations", C.c_void_p),
                ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p)]


class _Geom(C.Structure):
    _fields_ = [("depth", C.c_void_p), ("xy", C.c_void_p), ("conic_opacity", C.c_void_p), ("rgb", C.c_void_p),
                ("radii", C.c_void_p), ("rect", C.c_void_p), ("tiles_touched", C.c_void_p), ("clamped", C.c_void_p),
                ("cov3D", C.c_void_p)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.up3d_oracle_num_rendered.restype = C.c_int64
        _lib.up3d_oracle_render.restype = C.c_int64
        _lib.up3d_oracle_sizeof_scene.restype = C.c_size_t
        _lib.up3d_oracle_sizeof_geom.restype = C.c_size_t
        assert _lib.up3d_oracle_sizeof_scene() == C.sizeof(_Scene)
        assert _lib.up3d_oracle_sizeof_geom() == C.sizeof(_Geom)
    return _lib


def num_threads() -> int:
    return int(lib().up3d_oracle_num_threads())


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Scene:
    """One view of one Gaussian set, with the GaussianRasterizationSettings fields
    (gaussian_renderer/__init__.py:45-59)."""

    def __init__(self, means3D, opacities, scales, rotations, viewmatrix, projmatrix, campos, W, H, tanfovx, tanfovy,
                 shs=None, colors_precomp=None, sh_degree=0, bg=(0, 0, 0), scale_modifier=1.0, antialiasing=True):
        self.means3D = _f32(means3D).reshape(-1, 3)
        self.P = self.means3D.shape[0]
        self.opacities = _f32(opacities).reshape(-1)
        self.scales = _f32(scales).reshape(-1, 3)
        self.rotations = _f32(rotations).reshape(-1, 4)
        self.shs = None if shs is None else _f32(shs).reshape(self.P, -1, 3)
        self.colors = None if colors_precomp is None else _f32(colors_precomp).reshape(-1, 3)
        assert (self.shs is None) != (self.colors is None)
        self.M = 0 if self.shs is None else self.shs.shape[1]
        self.D = int(sh_degree)
        self.view = _f32(viewmatrix).reshape(16)
        self.proj = _f32(projmatrix).reshape(16)
        self.campos = _f32(campos).reshape(3)
        self.bg = _f32(bg).reshape(3)
        self.W, self.H = int(W), int(H)
        self.tanfovx, self.tanfovy = float(tanfovx), float(tanfovy)
        self.c = _Scene(self.P, self.M, self.D, self.W, self.H, self.tanfovx, self.tanfovy, float(scale_modifier),
                        int(bool(antialiasing)), _p(self.bg), _p(self.means3D), _p(self.shs), _p(self.colors),
                        _p(self.opacities), _p(self.scales), _p(self.rotations), _p(self.view), _p(self.proj),
                        _p(self.campos))

    @property
    def grid(self):
        return (self.W + 15) // 16, (self.H + 15) // 16


class Geom:
    def __init__(self, P):
        n = max(P, 1)
        self.depth = np.zeros(n, np.float32)
        self.xy = np.zeros((n, 2), np.float32)
        self.conic_opacity = np.zeros((n, 4), np.float32)
        self.rgb = np.zeros((n, 3), np.float32)
        self.radii = np.zeros(n, np.int32)
        self.rect = np.zeros((n, 4), np.int32)
        self.tiles_touched = np.zeros(n, np.int32)
        self.clamped = np.zeros((n, 3), np.uint8)
        self.cov3D = np.zeros((n, 6), np.float32)
        self.c = _Geom(*[_p(getattr(self, f)) for f, _ in _Geom._fields_])


def preprocess(sc: Scene) -> Geom:
    g = Geom(sc.P)
    lib().up3d_oracle_preprocess(C.byref(sc.c), C.byref(g.c))
    return g


def bin_tiles(sc: Scene, g: Geom):
    """-> (keys u64 [L], point_list u32 [L], ranges u32 [tiles,2]) : the reference's global sort (A.4)."""
    L = int(lib().up3d_oracle_num_rendered(sc.P, _p(g.tiles_touched)))
    keys = np.zeros(max(L, 1), np.uint64)
    vals = np.zeros(max(L, 1), np.uint32)
    gx, gy = sc.grid
    ranges = np.zeros((gx * gy, 2), np.uint32)
    lib().up3d_oracle_bin(sc.P, sc.W, sc.H, C.byref(g.c), _p(keys), _p(vals), _p(ranges))
    return keys[:L], vals[:L], ranges


def blend_forward(sc: Scene, g: Geom, point_list, ranges):
    color = np.zeros((3, sc.H, sc.W), np.float32)
    final_T = np.zeros((sc.H, sc.W), np.float32)
    n_contrib = np.zeros((sc.H, sc.W), np.uint32)
    pl = np.ascontiguousarray(point_list if len(point_list) else np.zeros(1, np.uint32))
    lib().up3d_oracle_blend_forward(sc.W, sc.H, _p(sc.bg), C.byref(g.c), _p(pl), _p(ranges), _p(color), _p(final_T),
                                    _p(n_contrib))
    return color, final_T, n_contrib


def blend_backward(sc: Scene, g: Geom, point_list, ranges, final_T, n_contrib, dL_dpix):
    P = max(sc.P, 1)
    d2 = np.zeros((P, 2), np.float32)
    dc = np.zeros((P, 3), np.float32)
    do = np.zeros(P, np.float32)
    dr = np.zeros((P, 3), np.float32)
    dL = _f32(dL_dpix).reshape(3, sc.H, sc.W)
    pl = np.ascontiguousarray(point_list if len(point_list) else np.zeros(1, np.uint32))
    lib().up3d_oracle_blend_backward(sc.P, sc.W, sc.H, _p(sc.bg), C.byref(g.c), _p(pl), _p(ranges), _p(final_T),
                                     _p(n_contrib), _p(dL), _p(d2), _p(dc), _p(do), _p(dr))
    return d2, dc, do, dr


def preprocess_backward(sc: Scene, g: Geom, d2, dc, do, dr):
    P = max(sc.P, 1)
    out = dict(means3D=np.zeros((P, 3), np.float32), means2D=np.zeros((P, 3), np.float32),
               shs=np.zeros((P, max(sc.M, 1), 3), np.float32), colors=np.zeros((P, 3), np.float32),
               opacities=np.zeros(P, np.float32), scales=np.zeros((P, 3), np.float32),
               rotations=np.zeros((P, 4), np.float32))
    lib().up3d_oracle_preprocess_backward(C.byref(sc.c), C.byref(g.c), _p(d2), _p(dc), _p(do), _p(dr),
                                          _p(out["means3D"]), _p(out["means2D"]), _p(out["shs"]), _p(out["colors"]),
                                          _p(out["opacities"]), _p(out["scales"]), _p(out["rotations"]))
    return {k: v[:sc.P] for k, v in out.items()}


def render(sc: Scene, dL_dpix=None):
    """Whole pipeline of one view. -> dict(color, radii, num_rendered[, grads])."""
    P = max(sc.P, 1)
    color = np.zeros((3, sc.H, sc.W), np.float32)
    radii = np.zeros(P, np.int32)
    res = {}
    if dL_dpix is None:
        L = lib().up3d_oracle_render(C.byref(sc.c), _p(color), _p(radii), None, None, None, None, None, None, None, None)
    else:
        dL = _f32(dL_dpix).reshape(3, sc.H, sc.W)
        g = dict(means3D=np.zeros((P, 3), np.float32), means2D=np.zeros((P, 3), np.float32),
                 shs=np.zeros((P, max(sc.M, 1), 3), np.float32), colors=np.zeros((P, 3), np.float32),
                 opacities=np.zeros(P, np.float32), scales=np.zeros((P, 3), np.float32),
                 rotations=np.zeros((P, 4), np.float32))
        L = lib().up3d_oracle_render(C.byref(sc.c), _p(color), _p(radii), _p(dL), _p(g["means3D"]), _p(g["means2D"]),
                                     _p(g["shs"]), _p(g["colors"]), _p(g["opacities"]), _p(g["scales"]),
                                     _p(g["rotations"]))
        res["grads"] = {k: v[:sc.P] for k, v in g.items()}
    res.update(color=color, radii=radii[:sc.P], num_rendered=int(L))
    return res


# ----------------------------------------------------------------------------- point ops
def fps_block_size(n: int) -> int:
    return int(lib().up3d_oracle_fps_block_size(int(n)))


def fps(xyz, npoint: int):
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    temp = np.full((B, N), 1e10, np.float32)
    out = np.zeros((B, npoint), np.int32)
    lib().up3d_oracle_fps(B, N, int(npoint), _p(xyz), _p(temp), _p(out))
    return out


def ball_query(radius: float, nsample: int, xyz, new_xyz):
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), np.int32)
    lib().up3d_oracle_ball_query(B, N, M, C.c_float(radius), int(nsample), _p(new_xyz), _p(xyz), _p(idx))
    return idx


def group(points, idx):
    points = _f32(points)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, Cc, N = points.shape
    _, M, K = idx.shape
    out = np.zeros((B, Cc, M, K), np.float32)
    lib().up3d_oracle_group(B, Cc, N, M, K, _p(points), _p(idx), _p(out))
    return out


def group_grad(grad_out, idx, N: int):
    grad_out = _f32(grad_out)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, Cc, M, K = grad_out.shape
    out = np.zeros((B, Cc, N), np.float32)
    lib().up3d_oracle_group_grad(B, Cc, int(N), M, K, _p(grad_out), _p(idx), _p(out))
    return out


def gather(points, idx):
    points = _f32(points)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    B, Cc, N = points.shape
    M = idx.shape[1]
    out = np.zeros((B, Cc, M), np.float32)
    lib().up3d_oracle_gather(B, Cc, N, M, _p(points), _p(idx), _p(out))
    return out


def eval_sh(deg: int, sh, pos, campos):
    """A.5 colour (after +0.5 and clamp) for points `pos` seen from `campos`; sh (n,M,3). -> (rgb (n,3), clamped (n,3))"""
    sh, pos, campos = _f32(sh), _f32(pos), _f32(campos)
    n, M = sh.shape[0], sh.shape[1]
    rgb = np.zeros((n, 3), np.float32)
    cl = np.zeros((n, 3), np.uint8)
    lib().up3d_oracle_eval_sh(int(deg), int(M), int(n), _p(sh), _p(pos), _p(campos), _p(rgb), _p(cl))
    return rgb, cl
