"""Batched differentiable Gaussian rasterizer: torch.autograd front-end of the C ABI.

One call renders every (object, view) pair of a training step -- the B*V `render_predicted` calls of
/root/reference/train_network.py:418-442 collapsed into one launch set -- and one backward call produces
the gradients that the reference obtains from B*V `_RasterizeGaussians.backward` nodes summed by autograd.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import RasterDesc, check, ptr, require_cuda, stream_ptr


class RasterLayout:
    """Index arrays + descriptor for a batch: `set_sizes[s]` Gaussians in set s, `views_per_set[s]` views of it.
    Cached per (sizes, views, device) so steady-state steps do not rebuild or re-upload anything."""
    _cache: dict = {}

    def __init__(self, set_sizes: Sequence[int], views_per_set: Sequence[int], device: torch.device):
        assert len(set_sizes) == len(views_per_set)
        self.set_sizes = [int(x) for x in set_sizes]
        self.views_per_set = [int(x) for x in views_per_set]
        self.n_sets = len(self.set_sizes)
        self.n_views = sum(self.views_per_set)
        self.n_gaussians = sum(self.set_sizes)
        set_offsets, set_view_start, view_set, view_rec_start = [0], [0], [], [0]
        for s, (p, v) in enumerate(zip(self.set_sizes, self.views_per_set)):
            set_offsets.append(set_offsets[-1] + p)
            set_view_start.append(set_view_start[-1] + v)
            for _ in range(v):
                view_set.append(s)
                view_rec_start.append(view_rec_start[-1] + p)
        self.n_records = view_rec_start[-1]
        self.max_set_size = max(self.set_sizes) if self.set_sizes else 0
        if self.n_records >= 2 ** 31 or self.n_gaussians >= 2 ** 31:
            raise RuntimeError("unipre3d_b200: more than 2^31 (view, Gaussian) records in one batch")
        i32 = dict(dtype=torch.int32, device=device)
        self.set_offsets = torch.tensor(set_offsets, **i32)
        self.set_view_start = torch.tensor(set_view_start, **i32)
        self.view_set = torch.tensor(view_set if view_set else [0], **i32)
        self.view_rec_start = torch.tensor(view_rec_start, **i32)
        self.view_rec_start_host = view_rec_start
        self.device = device

    @classmethod
    def get(cls, set_sizes, views_per_set, device) -> "RasterLayout":
        key = (tuple(int(x) for x in set_sizes), tuple(int(x) for x in views_per_set), str(device))
        lay = cls._cache.get(key)
        if lay is None:
            if len(cls._cache) > 64:
                cls._cache.clear()
            lay = cls._cache[key] = cls(set_sizes, views_per_set, device)
        return lay

    def desc(self, width, height, sh_degree, sh_coeffs, tanfovx, tanfovy, scale_modifier=1.0, antialiasing=True):
        return RasterDesc(self.n_sets, self.n_views, self.n_gaussians, self.n_records, self.max_set_size,
                          int(width), int(height), int(sh_degree), int(sh_coeffs), int(bool(antialiasing)),
                          float(tanfovx), float(tanfovy), float(scale_modifier),
                          self.set_offsets.data_ptr(), self.set_view_start.data_ptr(), self.view_set.data_ptr(),
                          self.view_rec_start.data_ptr())


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


class _RasterizeBatch(torch.autograd.Function):
    """Autograd node.  Gradient order mirrors upstream's _RasterizeGaussians.backward
    (SURVEY.md Appendix A.10): means3D, means2D, sh, colors_precomp, opacities, scales, rotations."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, viewmats, projmats, campos,
                bg, layout: RasterLayout, cfg: dict):
        require_cuda(means3D, sh, colors_precomp, opacities, scales, rotations, viewmats, projmats, campos, bg)
        dev = means3D.device
        means3D, sh, colors_precomp = _f32c(means3D), _f32c(sh), _f32c(colors_precomp)
        opacities, scales, rotations = _f32c(opacities), _f32c(scales), _f32c(rotations)
        viewmats, projmats, campos, bg = _f32c(viewmats), _f32c(projmats), _f32c(campos), _f32c(bg)
        M = 0 if sh is None else int(sh.shape[-2])
        W, H = cfg["image_width"], cfg["image_height"]
        d = layout.desc(W, H, cfg["sh_degree"], M, cfg["tanfovx"], cfg["tanfovy"], cfg["scale_modifier"],
                        cfg["antialiasing"])
        V, R = layout.n_views, layout.n_records
        if means3D.numel() != layout.n_gaussians * 3:
            raise RuntimeError(f"means3D has {means3D.numel() // 3} Gaussians, layout expects {layout.n_gaussians}")
        if viewmats.numel() != V * 16 or projmats.numel() != V * 16 or campos.numel() != V * 3:
            raise RuntimeError("viewmats/projmats/campos do not match the number of views in the layout")
        color = torch.empty((V, 3, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((R,), dtype=torch.int32, device=dev)
        invdepth = torch.empty((V, 1, H, W), dtype=torch.float32, device=dev) if cfg.get("invdepth", True) else None
        state = torch.empty((int(_lib.lib.up3d_raster_state_bytes(C.byref(d))),), dtype=torch.uint8, device=dev)
        scratch = torch.empty((int(_lib.lib.up3d_raster_scratch_bytes(C.byref(d))),), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib.up3d_raster_forward(C.byref(d), ptr(means3D), ptr(sh), ptr(colors_precomp), ptr(opacities),
                                               ptr(scales), ptr(rotations), ptr(viewmats), ptr(projmats), ptr(campos),
                                               ptr(bg), ptr(color), ptr(radii), ptr(invdepth), ptr(state),
                                               ptr(scratch), stream_ptr()), launches=3)
        ctx.layout, ctx.cfg, ctx.M = layout, cfg, M
        ctx.want_means2D = means2D is not None and means2D.requires_grad
        ctx.save_for_backward(means3D, sh, colors_precomp, opacities, scales, rotations, viewmats, projmats, campos, bg,
                              state)
        ctx.shapes = (None if sh is None else sh.shape, opacities.shape)
        ctx.mark_non_differentiable(radii)
        if invdepth is None:
            invdepth = torch.zeros((V, 1, H, W), dtype=torch.float32, device=dev)
        ctx.mark_non_differentiable(invdepth)
        return color, radii, invdepth

    @staticmethod
    def backward(ctx, grad_color, _grad_radii, _grad_invdepth):
        (means3D, sh, colors_precomp, opacities, scales, rotations, viewmats, projmats, campos, bg,
         state) = ctx.saved_tensors
        layout, cfg, M = ctx.layout, ctx.cfg, ctx.M
        dev = means3D.device
        d = layout.desc(cfg["image_width"], cfg["image_height"], cfg["sh_degree"], M, cfg["tanfovx"], cfg["tanfovy"],
                        cfg["scale_modifier"], cfg["antialiasing"])
        grad_color = _f32c(grad_color)
        G = layout.n_gaussians
        dmeans3D = torch.empty((G, 3), dtype=torch.float32, device=dev)
        dmeans2D = torch.empty((layout.n_records, 3), dtype=torch.float32, device=dev) if ctx.want_means2D else None
        dsh = torch.empty((G, M, 3), dtype=torch.float32, device=dev) if sh is not None else None
        dcol = torch.empty((G, 3), dtype=torch.float32, device=dev) if colors_precomp is not None else None
        dop = torch.empty((G,), dtype=torch.float32, device=dev)
        dsc = torch.empty((G, 3), dtype=torch.float32, device=dev)
        drot = torch.empty((G, 4), dtype=torch.float32, device=dev)
        scratch = torch.empty((int(_lib.lib.up3d_raster_scratch_bytes(C.byref(d))),), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib.up3d_raster_backward(C.byref(d), ptr(means3D), ptr(sh), ptr(colors_precomp), ptr(opacities),
                                                ptr(scales), ptr(rotations), ptr(viewmats), ptr(projmats), ptr(campos),
                                                ptr(bg), ptr(grad_color), ptr(state), ptr(scratch), ptr(dmeans3D),
                                                ptr(dmeans2D), ptr(dsh), ptr(dcol), ptr(dop), ptr(dsc), ptr(drot),
                                                stream_ptr()), launches=2)
        sh_shape, op_shape = ctx.shapes
        return (dmeans3D.view_as(means3D), dmeans2D, None if dsh is None else dsh.view(sh_shape), dcol,
                dop.view(op_shape), dsc.view_as(scales), drot.view_as(rotations), None, None, None, None, None, None)


def rasterize_batch(means3D, opacities, scales, rotations, viewmats, projmats, campos, bg, *, set_sizes,
                    views_per_set, image_height, image_width, tanfovx, tanfovy, sh_degree, shs=None,
                    colors_precomp=None, means2D=None, scale_modifier=1.0, antialiasing=True, invdepth=True):
    """Render all views of all Gaussian sets.

    means3D (G,3) [concatenated sets], shs (G,M,3) | colors_precomp (G,3), opacities (G,)|(G,1), scales (G,3),
    rotations (G,4); viewmats/projmats (V,4,4) row-vector convention, campos (V,3), bg (3,).
    Returns (color (V,3,H,W), radii (R,) int32, invdepth (V,1,H,W)); view v's radii are
    radii[layout.view_rec_start[v] : layout.view_rec_start[v+1]].
    """
    if (shs is None) == (colors_precomp is None):
        raise Exception("Please provide excatly one of either SHs or precomputed colors!")
    layout = RasterLayout.get(set_sizes, views_per_set, means3D.device)
    cfg = dict(image_height=int(image_height), image_width=int(image_width), tanfovx=float(tanfovx),
               tanfovy=float(tanfovy), sh_degree=int(sh_degree), scale_modifier=float(scale_modifier),
               antialiasing=bool(antialiasing), invdepth=bool(invdepth))
    return _RasterizeBatch.apply(means3D, means2D, shs, colors_precomp, opacities, scales, rotations, viewmats,
                                 projmats, campos, bg, layout, cfg)


# ----------------------------------------------------------------------------------------------- debug accessors
def debug_forward_state(means3D, opacities, scales, rotations, viewmats, projmats, campos, bg, *, set_sizes,
                        views_per_set, image_height, image_width, tanfovx, tanfovy, sh_degree, shs=None,
                        colors_precomp=None, scale_modifier=1.0, antialiasing=True, tile_lists=True):
    """Runs the forward and returns everything the parity tests compare with the oracle: per-record geometry,
    depth-sorted ids, per-tile lists as the reference's global (tile|depth) sort would give them, final_T, n_contrib."""
    with torch.no_grad():
        require_cuda(means3D)
        dev = means3D.device
        layout = RasterLayout.get(set_sizes, views_per_set, dev)
        M = 0 if shs is None else int(shs.shape[-2])
        d = layout.desc(image_width, image_height, sh_degree, M, tanfovx, tanfovy, scale_modifier, antialiasing)
        V, R, H, W = layout.n_views, layout.n_records, int(image_height), int(image_width)
        t = [_f32c(x) for x in (means3D, shs, colors_precomp, opacities, scales, rotations, viewmats, projmats, campos, bg)]
        color = torch.empty((V, 3, H, W), dtype=torch.float32, device=dev)
        radii = torch.empty((R,), dtype=torch.int32, device=dev)
        invdepth = torch.empty((V, 1, H, W), dtype=torch.float32, device=dev)
        state = torch.empty((int(_lib.lib.up3d_raster_state_bytes(C.byref(d))),), dtype=torch.uint8, device=dev)
        scratch = torch.empty((int(_lib.lib.up3d_raster_scratch_bytes(C.byref(d))),), dtype=torch.uint8, device=dev)
        check(_lib.lib.up3d_raster_forward(C.byref(d), *[ptr(x) for x in t], ptr(color), ptr(radii), ptr(invdepth),
                                           ptr(state), ptr(scratch), stream_ptr()), launches=3)
        f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
        out = dict(color=color, radii=radii, invdepth=invdepth, layout=layout,
                   n_visible=torch.zeros(V, **i32), sorted_ids=torch.zeros(max(R, 1), **i32),
                   depths=torch.zeros(max(R, 1), **f32), xy=torch.zeros((max(R, 1), 2), **f32),
                   conic_opacity=torch.zeros((max(R, 1), 4), **f32), rgb=torch.zeros((max(R, 1), 3), **f32),
                   rects=torch.zeros((max(R, 1), 4), **i32), final_T=torch.zeros((V, H, W), **f32),
                   n_contrib=torch.zeros((V, H, W), **i32))
        check(_lib.lib.up3d_raster_debug_state(C.byref(d), ptr(state), ptr(out["n_visible"]), ptr(out["sorted_ids"]),
                                               ptr(out["depths"]), ptr(out["xy"]), ptr(out["conic_opacity"]),
                                               ptr(out["rgb"]), ptr(out["rects"]), ptr(out["final_T"]),
                                               ptr(out["n_contrib"]), stream_ptr()), launches=1)
        out.update(bin_mode=torch.zeros(V, **i32), bin_total=torch.zeros(V, **i32))
        check(_lib.lib.up3d_raster_debug_bins(C.byref(d), ptr(state), ptr(out["bin_mode"]), ptr(out["bin_total"]), stream_ptr()))
        if tile_lists:
            tiles = ((W + 15) // 16) * ((H + 15) // 16)
            counts = torch.zeros(V * tiles, **i32)
            check(_lib.lib.up3d_raster_debug_tile_lists(C.byref(d), ptr(state), ptr(counts), None, None, stream_ptr()),
                  launches=1)
            offsets = torch.zeros(V * tiles + 1, **i32)
            offsets[1:] = torch.cumsum(counts, 0)
            lists = torch.zeros(max(int(offsets[-1].item()), 1), **i32)
            check(_lib.lib.up3d_raster_debug_tile_lists(C.byref(d), ptr(state), ptr(counts), ptr(offsets), ptr(lists),
                                                        stream_ptr()), launches=1)
            out.update(tile_counts=counts.view(V, tiles), tile_offsets=offsets, tile_lists=lists)
        return out
