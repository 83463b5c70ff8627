"""CPU (numpy) restatements of the small algorithms around the backbone that the product runs as fused CUDA kernels.

TEST INFRASTRUCTURE ONLY -- imported by tests/, never by the product package `unipre3d_b200`.

Pinned (tests/test_oracle_backbone_cpu.py) against golden vectors generated from the reference's own Python
(tests/golden/process_output.npz, feature_fusion.npz) and against PyTorch's CPU operators:

  splat_head          /root/reference/model/gaussian_predictor.py:279-328 (`_process_network_output`, object branch),
                      activations 249-254, SH concatenation of gaussian_renderer/__init__.py:66-69
  fusion_geometry     /root/reference/fusion/feat_fusion.py:23-56 (`project_points_to_image`) and 88-119 (inside test,
                      per-pixel nearest-depth test with the `pixel_y * H + pixel_x` cell hash)
  layer_norm_*        nn.LayerNorm as used by openpoints/models/backbone/transformer.py:104,108,318
  gelu / gelu_grad    nn.GELU (erf form), transformer.py:17
  clip_adamw_step     /root/reference/train_network.py:368-390 (skip on NaN/Inf, clip_grad_norm_(1.0)) + 156-159
                      (torch.optim.AdamW, eps 1e-15)
  merge_mean_m2       pairwise (Chan et al.) merge of shifted partial sums -- the BatchNorm statistics of the fused
                      mini-PointNet (nn.BatchNorm1d batch statistics, transformer.py:216,221)
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import numpy as np


# ------------------------------------------------------------------------------------------------ splat head
def splat_head(raw: np.ndarray, center: np.ndarray, offset_scale: float, isotropic: bool, max_sh_degree: int) -> Dict[str, np.ndarray]:
    """raw (B, C, P) in the reference's channel-major layout [xyz 3 | opacity 1 | scaling 3 | rotation 4 | dc 3 | rest],
    center (B, P, 3).  Returns the reference's output dict (+ "shs" = [dc || rest] (B,P,M,3))."""
    raw = raw.astype(np.float32)
    B, C, P = raw.shape
    xyz_raw, opacity, scaling, rotation, dc = raw[:, 0:3], raw[:, 3:4], raw[:, 4:7], raw[:, 7:11], raw[:, 11:14]
    pos = (np.tanh(xyz_raw) * np.float32(offset_scale)).transpose(0, 2, 1) + center[:, :, :3].astype(np.float32)
    if isotropic:
        scaling = np.repeat(scaling[:, :1], 3, axis=1)
    flat = lambda x: x.reshape(x.shape[0], x.shape[1], -1).transpose(0, 2, 1)
    # F.normalize(x, dim=-1) on the (B,4,P) tensor: each quaternion COMPONENT is normalised over the P points
    norm = np.maximum(np.sqrt((rotation.astype(np.float64) ** 2).sum(-1, keepdims=True)), 1e-6).astype(np.float32)
    out = {"xyz": pos,
           "opacity": flat(1.0 / (1.0 + np.exp(-opacity))),
           "scaling": flat(np.exp(np.clip(scaling, -1.0, 20.0))),
           "rotation": flat(rotation / norm),
           "features_dc": flat(dc)[:, :, None, :]}
    if max_sh_degree > 0:
        rest = flat(raw[:, 14:])
        out["features_rest"] = rest.reshape(B, P, -1, 3)
        out["shs"] = np.concatenate([out["features_dc"], out["features_rest"]], axis=2)
    else:
        out["features_rest"] = np.zeros((B, 0, 3), np.float32)
        out["shs"] = out["features_dc"]
    return {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in out.items()}


# ------------------------------------------------------------------------------------------------ fusion geometry
def fusion_geometry(center: np.ndarray, c2w: np.ndarray, intrinsic: np.ndarray, H: int, W: int):
    """center (B,N,3), c2w (B,4,4) row-vector (transposed) camera-to-world.  -> ix, iy (B,N) int64 (0 where outside),
    inside (B,N) bool, keep (B,N) bool (nearest point of its pixel cell), depth (B,N) fp32."""
    center = center.astype(np.float32)
    B, N, _ = center.shape
    hom = np.concatenate([center, np.ones((B, N, 1), np.float32)], axis=2)
    w2c = np.linalg.inv(c2w.transpose(0, 2, 1).astype(np.float32)).astype(np.float32)
    cam = np.einsum("bij,bnj->bni", w2c, hom).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        px = (cam[..., 0] * np.float32(intrinsic[0][0])) / cam[..., 2] + np.float32(intrinsic[0][2])
        py = (cam[..., 1] * np.float32(intrinsic[1][1])) / cam[..., 2] + np.float32(intrinsic[1][2])
    fx, fy = np.rint(px), np.rint(py)                      # torch.round: half to even
    depth = cam[..., 2]
    inside = (fx >= 0) & (fy >= 0) & (fx < H) & (fy < W) & (depth >= 0)
    ix = np.where(inside, fx, 0).astype(np.int64)
    iy = np.where(inside, fy, 0).astype(np.int64)
    keep = np.zeros((B, N), bool)
    for b in range(B):
        best: Dict[int, float] = {}
        for n in range(N):
            if inside[b, n]:
                cell = int(iy[b, n]) * H + int(ix[b, n])
                best[cell] = min(best.get(cell, math.inf), float(depth[b, n]))
        for n in range(N):
            if inside[b, n]:
                keep[b, n] = float(depth[b, n]) == best[int(iy[b, n]) * H + int(ix[b, n])]
    return ix, iy, inside, keep, depth


def fuse_features(x: np.ndarray, center: np.ndarray, feat: np.ndarray, c2w: np.ndarray, intrinsic: np.ndarray,
                  w: np.ndarray, b: np.ndarray) -> np.ndarray:
    """FeatureFusion.__call__ (feat_fusion.py:58-145) for dense image features feat (B,C,H,W) and a Linear+ReLU
    fusion MLP (w (C, 2C), b (C)); x (B, N+1, C) with a CLS token or (B, N, C)."""
    B, N = center.shape[:2]
    C, H, W = feat.shape[1:]
    ix, iy, _, keep, _ = fusion_geometry(center, c2w, intrinsic, H, W)
    mapped = np.zeros((B, N, C), np.float32)
    for bb in range(B):
        for n in range(N):
            if keep[bb, n]:
                mapped[bb, n] = feat[bb, :, ix[bb, n], iy[bb, n]]          # the reference's [b, :, pixel_x, pixel_y]
    if x.shape[1] > N:
        patch = np.concatenate([x[:, 1:], mapped], -1)
        cls = np.concatenate([x[:, :1], np.zeros((B, 1, C), np.float32)], -1)
        xx = np.concatenate([cls, patch], 1)
    else:
        xx = np.concatenate([x, mapped], -1)
    return np.maximum(xx.astype(np.float32) @ w.T.astype(np.float32) + b.astype(np.float32), 0.0)


# ------------------------------------------------------------------------------------------------ LayerNorm / GELU
def layer_norm_fwd(x, gamma, beta, eps=1e-5):
    x = x.astype(np.float64)
    mean = x.mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(x.var(-1, keepdims=True) + eps)
    return ((x - mean) * rstd * gamma + beta), mean[..., 0], rstd[..., 0]


def layer_norm_bwd(dy, x, gamma, eps=1e-5):
    """-> dx, dgamma, dbeta  (dx = rstd (dy*gamma - mean(dy*gamma) - xhat mean(dy*gamma*xhat)))."""
    x, dy = x.astype(np.float64), dy.astype(np.float64)
    mean = x.mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(x.var(-1, keepdims=True) + eps)
    xhat = (x - mean) * rstd
    g = dy * gamma
    dx = rstd * (g - g.mean(-1, keepdims=True) - xhat * (g * xhat).mean(-1, keepdims=True))
    red = tuple(range(x.ndim - 1))
    return dx, (dy * xhat).sum(red), dy.sum(red)


_erf = np.vectorize(math.erf, otypes=[np.float64])


def gelu(x):
    x = x.astype(np.float64)
    return 0.5 * x * (1.0 + _erf(x / math.sqrt(2.0)))


def gelu_grad(x):
    x = x.astype(np.float64)
    return 0.5 * (1.0 + _erf(x / math.sqrt(2.0))) + x * np.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


# ------------------------------------------------------------------------------------------------ optimizer
def clip_adamw_step(params: Sequence[np.ndarray], grads: Sequence[np.ndarray], exp_avg: Sequence[np.ndarray],
                    exp_avg_sq: Sequence[np.ndarray], step: int, lrs: Sequence[float], betas=(0.9, 0.999), eps=1e-15,
                    weight_decay=1e-2, max_norm=1.0, grad_scale=1.0):
    """One trainer step in place (fp64): returns (new step count, total norm, applied).  NaN/Inf anywhere -> nothing changes."""
    total = math.sqrt(sum(float(((np.asarray(g, np.float64) * grad_scale) ** 2).sum()) for g in grads))
    if not math.isfinite(total):
        return step, total, False
    coef = min(max_norm / (total + 1e-6), 1.0) * grad_scale
    t = step + 1
    b1, b2 = betas
    for p, g, m, v, lr in zip(params, grads, exp_avg, exp_avg_sq, lrs):
        gg = np.asarray(g, np.float64) * coef
        p *= 1.0 - lr * weight_decay
        m += (1.0 - b1) * (gg - m)
        v *= b2
        v += (1.0 - b2) * gg * gg
        p -= (lr / (1.0 - b1 ** t)) * (m / (np.sqrt(v) / math.sqrt(1.0 - b2 ** t) + eps))
    return t, total, True


# ------------------------------------------------------------------------------------------------ BatchNorm statistics
def merge_mean_m2(shifts, sums, sumsqs, counts):
    """Partials of rows r in part p: shift_p, sum (z - shift_p), sum (z - shift_p)^2, n_p  ->  (mean, M2 = sum (z-mean)^2)."""
    shifts, sums, sumsqs = (np.asarray(a, np.float64) for a in (shifts, sums, sumsqs))
    counts = np.asarray(counts, np.float64).reshape(-1, *([1] * (shifts.ndim - 1)))
    N = counts.sum()
    mean = (counts * shifts + sums).sum(0) / N
    d = shifts - mean
    m2 = (sumsqs + 2.0 * d * sums + counts * d * d).sum(0)
    return mean, m2
