"""fp64 torch-autograd restatement of the rasterizer FORWARD.

TEST INFRASTRUCTURE ONLY (see oracle/raster_oracle.c header; same "parity unpinned" caveat).

Purpose: gradients of this differentiable forward are the ground truth for the hand-derived
backward passes (the C oracle's and the CUDA product's), exactly the route Appendix A.9
recommends:  (i) straight-through on min(0.99, alpha),  (ii) sub-gradient 0 where the colour
was clamped at 0,  (iii) the max(0.000025, .) / max(0.1, .) floors as written,  (iv) the
EWA clamp treated as upstream does (clamped t.x/t.z -> no gradient through that coordinate).

Tile lists (which Gaussians, in which order, per tile) are discrete; they are taken from the
integer-exact C oracle (oracle_lib.bin_tiles) so that this module only restates arithmetic.
Pure-torch, meant for small cases (P <= a few hundred, images <= 64x64).
"""
from __future__ import annotations

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
         1.445305721320277, -0.5900435899266435]


def eval_sh_rgb(deg, sh, dirs):
    """sh (P,M,3), dirs (P,3) unit -> (P,3) before the +0.5 / clamp.  Same polynomials as
    /root/reference/utils/sh_utils.py:57-112 (the in-tree restatement of upstream's SH)."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    r = SH_C0 * sh[:, 0]
    if deg > 0:
        r = r - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        r = (r + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 6]
             + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
    if deg > 2:
        r = (r + SH_C3[0] * y * (3 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10]
             + SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
             + SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14]
             + SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return r


def project(means3D, opacities, scales, rotations, shs, colors_precomp, view, proj, campos, W, H, tanfovx, tanfovy,
            sh_degree, scale_modifier=1.0, antialiasing=True):
    """A.1-A.3, A.5 in fp64. view/proj are the flat row-vector 4x4s (16,). Returns xy, conic, opac*hs, rgb, depth."""
    dt = means3D.dtype
    view = view.reshape(4, 4).to(dt)  # view[c, r] = flat[4c + r]
    proj = proj.reshape(4, 4).to(dt)
    P = means3D.shape[0]
    ones = torch.ones(P, 1, dtype=dt)
    ph = torch.cat([means3D, ones], 1) @ proj          # (P,4): row-vector convention
    t = (torch.cat([means3D, ones], 1) @ view)[:, :3]
    p_w = 1.0 / (ph[:, 3] + 1e-7)
    ndc = ph[:, :2] * p_w[:, None]
    xy = torch.stack([((ndc[:, 0] + 1.0) * W - 1.0) * 0.5, ((ndc[:, 1] + 1.0) * H - 1.0) * 0.5], 1)
    # A.2
    r, x, y, z = rotations.unbind(1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(P, 3, 3)
    A = R * (scale_modifier * scales)[:, None, :]
    Sigma = A @ A.transpose(1, 2)
    # A.3
    fx, fy = W / (2.0 * tanfovx), H / (2.0 * tanfovy)
    limx, limy = 1.3 * tanfovx, 1.3 * tanfovy
    tz = t[:, 2]
    txtz, tytz = t[:, 0] / tz, t[:, 1] / tz
    cx = (txtz < -limx) | (txtz > limx)
    cy = (tytz < -limy) | (tytz > limy)
    tx = torch.where(cx, (txtz.clamp(-limx, limx) * tz).detach(), t[:, 0])
    ty = torch.where(cy, (tytz.clamp(-limy, limy) * tz).detach(), t[:, 1])
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz), zero, fy / tz, -(fy * ty) / (tz * tz)], 1).reshape(P, 2, 3)
    Rwc = view[:3, :3].t()  # Rwc[r, c] = flat[4c + r]
    T = J @ Rwc
    cov = T @ Sigma @ T.transpose(1, 2)
    c00, c01, c11 = cov[:, 0, 0], cov[:, 1, 0], cov[:, 1, 1]
    det0 = c00 * c11 - c01 * c01
    a, b, c = c00 + 0.3, c01, c11 + 0.3
    det = a * c - b * b
    hs = torch.sqrt(torch.clamp_min(det0 / det, 0.000025)) if antialiasing else torch.ones_like(det)
    conic = torch.stack([c / det, -b / det, a / det], 1)
    if colors_precomp is not None:
        rgb = colors_precomp
    else:
        d = means3D - campos.to(dt)[None]
        d = d / d.norm(dim=1, keepdim=True)
        rgb = torch.clamp_min(eval_sh_rgb(sh_degree, shs, d) + 0.5, 0.0)
    return xy, conic, opacities.reshape(-1) * hs, rgb, tz


def render(means3D, opacities, scales, rotations, shs, colors_precomp, view, proj, campos, bg, W, H, tanfovx, tanfovy,
           sh_degree, point_list, ranges, scale_modifier=1.0, antialiasing=True):
    """Differentiable forward; point_list / ranges come from the integer-exact oracle. -> (3,H,W) fp64."""
    xy, conic, op, rgb, _ = project(means3D, opacities, scales, rotations, shs, colors_precomp, view, proj, campos,
                                    W, H, tanfovx, tanfovy, sh_degree, scale_modifier, antialiasing)
    dt = means3D.dtype
    out = torch.zeros(3, H, W, dtype=dt)
    gx = (W + 15) // 16
    bg = bg.to(dt)
    pl = torch.as_tensor(point_list.astype("int64"))
    for tile in range(ranges.shape[0]):
        r0, r1 = int(ranges[tile, 0]), int(ranges[tile, 1])
        x0, y0 = (tile % gx) * 16, (tile // gx) * 16
        x1, y1 = min(x0 + 16, W), min(y0 + 16, H)
        ys, xs = torch.meshgrid(torch.arange(y0, y1, dtype=dt), torch.arange(x0, x1, dtype=dt), indexing="ij")
        pix = torch.stack([xs.reshape(-1), ys.reshape(-1)], 1)      # (n,2)
        n = pix.shape[0]
        if r1 <= r0:
            out[:, y0:y1, x0:x1] = bg[:, None, None].expand(3, y1 - y0, x1 - x0)
            continue
        ids = pl[r0:r1]
        d = xy[ids][None, :, :] - pix[:, None, :]                   # (n,L,2)
        cn = conic[ids][None]
        power = -0.5 * (cn[..., 0] * d[..., 0] ** 2 + cn[..., 2] * d[..., 1] ** 2) - cn[..., 1] * d[..., 0] * d[..., 1]
        alpha_raw = op[ids][None] * torch.exp(power)
        alpha = alpha_raw + (torch.clamp_max(alpha_raw, 0.99) - alpha_raw).detach()   # straight-through (A.7)
        keep = (power <= 0) & (alpha.detach() >= 1.0 / 255.0)
        a_eff = torch.where(keep, alpha, torch.zeros_like(alpha))
        one_m = 1.0 - a_eff
        T_incl = torch.cumprod(one_m, dim=1)
        T_before = torch.cat([torch.ones(n, 1, dtype=dt), T_incl[:, :-1]], 1)
        # termination: first kept entry whose T' < 1e-4 stops the pixel (that entry is NOT composited).
        # NB the oracle/CUDA decide this in fp32; tests use inputs away from the 1e-4 threshold or
        # accept the (rare) pixels where fp32/fp64 disagree via a robust comparison.
        stop = keep & (T_incl.detach() < 0.0001)
        dead = torch.cumsum(stop.to(torch.int64), dim=1) > 0
        w = torch.where(dead, torch.zeros_like(a_eff), a_eff * T_before)        # (n,L)
        col = w @ rgb[ids]                                                        # (n,3)
        T_fin = torch.where(dead, torch.ones_like(one_m), one_m).prod(dim=1)
        img = col + T_fin[:, None] * bg[None]
        out[:, y0:y1, x0:x1] = img.t().reshape(3, y1 - y0, x1 - x0)
    return out
