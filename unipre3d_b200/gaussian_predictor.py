"""GaussianSplatPredictor / PointFeaturePredictor with the reference's class names, constructor, forward
signature, output dict and state-dict keys (/root/reference/model/gaussian_predictor.py:16-447,
/root/reference/model/point_predictor.py:18-134).

Scope: the object-level fusion path every shipped object config uses
(`opt.use_fusion: true`, `opt.level: object`) with the `transformer` backbone.  Other backbones raise
NotImplementedError naming the §8 row that would add them.

Quirks reproduced on purpose (results must match the reference, SURVEY.md §7 "Quirk parity"):
  * rotation_activation = F.normalize(x, dim=-1) is applied to the (B,4,P) tensor, i.e. it normalises over
    the P points, not over the 4 quaternion components (gaussian_predictor.py:254,318);
  * scaling = exp(clamp(x, -1, 20))  -> sigma >= e^-1 (252);
  * `_forward_basic` of the reference calls a method that does not exist (98-110); the evident intent
    (= `_process_network_output(..., is_scene_level=False)` on the fusion-less backbone) is implemented instead.

Image branch: the reference runs a frozen SD-VAE (model/image_predictor.py:56-81, weights not shipped) under
no_grad and feeds `decoder_block_3` (128 channels at image resolution) to the trainable `image_conv`.
`FrozenImageStem` is a weight-free stand-in with the same interface and output shape.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from .backbone import PointTransformerEncoder
from .fusion import LazyImageFeatures


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class StemFeatureField:
    """The stem's output f[n,c,y,x] = sin(proj[c,:] . image[n,:,y,x] + shift[c]) held analytically: `dense()`
    materialises the (n,128,R,R) tensor the reference's image network would return; the B200 step only needs GroupNorm
    group statistics (one kernel over the 3-channel image, csrc/stem.cu) and the field at the sampled pixels."""

    def __init__(self, image: torch.Tensor, proj: torch.Tensor, shift: torch.Tensor):
        self.image, self.proj, self.shift = image.float().contiguous(), proj, shift
        n, _, H, W = image.shape
        self.shape = (n, proj.shape[0], H, W)

    @torch.no_grad()
    def dense(self) -> torch.Tensor:
        with torch.autocast(self.image.device.type, enabled=False):       # the frozen stem is defined in fp32
            f = torch.einsum("oc,nchw->nohw", self.proj, self.image) + self.shift.view(1, -1, 1, 1)
            return torch.sin(f)

    @torch.no_grad()
    def group_sums(self, num_groups: int) -> torch.Tensor:
        """-> (n, G, 2) fp64 [sum f, sum f^2] over each group's channels x pixels (csrc/stem.cu)."""
        from . import _lib
        from ._lib import check, ptr, require_cuda, stream_ptr
        require_cuda(self.image, self.proj, self.shift)
        n, Cc, H, W = self.shape
        sums = torch.empty((n, num_groups, 2), dtype=torch.float64, device=self.image.device)
        with torch.cuda.device(self.image.device):
            check(_lib.lib.up3d_stem_group_stats(n, H, W, Cc, num_groups, ptr(self.image), ptr(self.proj.contiguous()),
                                                 ptr(self.shift.contiguous()), ptr(sums), stream_ptr()), launches=1)
        return sums

    @torch.no_grad()
    def group_stats(self, num_groups: int):
        """-> (mean, biased var), each (n, G), over each group's channels x pixels (torch.var_mean(unbiased=False))."""
        return self.stats_from_sums(self.group_sums(num_groups), num_groups)

    @torch.no_grad()
    def stats_from_sums(self, sums: torch.Tensor, num_groups: int):
        n, Cc, H, W = self.shape
        cnt = float((Cc // num_groups) * H * W)
        mean = sums[..., 0] / cnt
        var = (sums[..., 1] / cnt - mean * mean).clamp_min(0.0)
        return mean.float(), var.float()

    @torch.no_grad()
    def gather(self, bidx, ix, iy) -> torch.Tensor:
        """-> (B, N, C) = dense()[bidx, :, ix, iy]"""
        with torch.autocast(self.image.device.type, enabled=False):
            px = self.image[bidx, :, ix, iy]                                # (B,N,3)
            return torch.sin(px @ self.proj.t() + self.shift)


class FrozenImageStem(nn.Module):
    """Stand-in for ImageFeaturePredictor: frozen, gradient-free, returns {"decoder_block_3": (n,128,R,R)} -- as a
    dense tensor, or (`lazy=True`, CUDA) as the analytic StemFeatureField the sampled image_conv path consumes."""

    def __init__(self, cfg, out_channels):
        super().__init__()
        self.cfg, self.out_channels = cfg, out_channels
        self.encoder_config = {"block_out_channels": [128, 256, 512, 512]}
        g = torch.Generator().manual_seed(1234)
        self.register_buffer("proj", torch.randn(128, 3, generator=g) * 0.7, persistent=False)
        self.register_buffer("shift", torch.randn(128, generator=g) * 0.3, persistent=False)

    @torch.no_grad()
    def forward(self, x: torch.Tensor, lazy: bool = False) -> Dict[str, torch.Tensor]:
        field = StemFeatureField(x, self.proj, self.shift)
        return {"decoder_block_3": field if (lazy and x.is_cuda) else field.dense()}


class RandnImageFeatures(nn.Module):
    """`model.image_branch=randn`: the synthetic image features SURVEY.md §8d prescribes in place of the absent SD-VAE
    weights -- a fixed N(0,1) (n,128,R,R) `decoder_block_3` tensor, materialised in HBM, which `image_conv` then reads
    (GroupNorm statistics over the whole 268 MB tensor, 1x1 convolution at the sampled pixels)."""

    def __init__(self, cfg, out_channels):
        super().__init__()
        self.cfg, self.out_channels = cfg, out_channels
        self.encoder_config = {"block_out_channels": [128, 256, 512, 512]}
        self._feat = None

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        n, _, H, W = x.shape
        if self._feat is None or self._feat.shape != (n, 128, H, W) or self._feat.device != x.device:
            g = torch.Generator(device=x.device).manual_seed(4321)
            self._feat = torch.randn((n, 128, H, W), generator=g, device=x.device)
        return {"decoder_block_3": self._feat}


class SplatHeadFn(torch.autograd.Function):
    """`_process_network_output` (object level) + the SH concatenation of the renderer as one launch each way
    (csrc/head.cu).  raw (B,P,11+3M) fp32, center (B,P,3) -> xyz, opacity (B,P,1), scaling, rotation, shs (B,P,M,3)."""

    @staticmethod
    def forward(ctx, raw, center, M, offset_scale, isotropic):
        from . import _lib
        from ._lib import check, ptr, require_cuda, stream_ptr
        require_cuda(raw, center)
        raw, center = raw.contiguous().float(), center.contiguous().float()
        B, P, Cr = raw.shape
        if Cr != 11 + 3 * M:
            raise RuntimeError(f"splat head: {Cr} channels do not match max_sh_degree (expected {11 + 3 * M})")
        dev = raw.device
        e = lambda *sh: torch.empty(sh, dtype=torch.float32, device=dev)
        xyz, op, sc, rot, shs, rn = e(B, P, 3), e(B, P, 1), e(B, P, 3), e(B, P, 4), e(B, P, M, 3), e(B, 4)
        with torch.cuda.device(dev):
            check(_lib.lib.up3d_splat_head_fwd(B, P, M, ptr(raw), ptr(center), float(offset_scale), int(bool(isotropic)),
                                               ptr(xyz), ptr(op), ptr(sc), ptr(rot), ptr(shs), ptr(rn), stream_ptr()), 1)
        ctx.save_for_backward(raw, rn)
        ctx.cfg = (M, float(offset_scale), int(bool(isotropic)))
        return xyz, op, sc, rot, shs

    @staticmethod
    def backward(ctx, d_xyz, d_op, d_sc, d_rot, d_shs):
        from . import _lib
        from ._lib import check, ptr, stream_ptr
        raw, rn = ctx.saved_tensors
        M, offset_scale, iso = ctx.cfg
        B, P, _ = raw.shape
        c = lambda t: None if t is None else t.contiguous().float()
        d_xyz, d_op, d_sc, d_rot, d_shs = c(d_xyz), c(d_op), c(d_sc), c(d_rot), c(d_shs)
        d_raw = torch.empty_like(raw)
        with torch.cuda.device(raw.device):
            check(_lib.lib.up3d_splat_head_bwd(B, P, M, ptr(raw), ptr(rn), offset_scale, iso, ptr(d_xyz), ptr(d_op), ptr(d_sc),
                                               ptr(d_rot), ptr(d_shs), ptr(d_raw), stream_ptr()), 1)
        return d_raw, None, None, None, None


class PointFeaturePredictor(nn.Module):
    """point_predictor.py:18-134: backbone + `final` MLP (-> 23 channels)."""

    def __init__(self, cfg, out_channels, pretrained_path=None):
        super().__init__()
        self.cfg, self.out_channels = cfg, out_channels
        model_type = cfg.model.backbone_type.lower()
        if model_type == "transformer":
            self.encoder = PointTransformerEncoder(in_channels=3, num_groups=128, encoder_dims=384, depth=16,
                                                   use_fusion=bool(getattr(cfg.opt, "use_fusion", True)))
            self.final = nn.Sequential(nn.Linear(384, 128), nn.ReLU(), nn.Linear(128, 23))
        elif model_type == "pointmlp":
            from .pointmlp import pointMLP
            self.encoder = pointMLP(cfg=cfg)
            self.final = nn.Sequential(nn.Linear(128, 64), nn.ReLU(), nn.Linear(64, 23))
        elif model_type == "sparseunet":
            from .sparse_unet import SpUNetBase
            self.encoder = SpUNetBase(in_channels=6, num_classes=64, cfg=cfg)          # point_predictor.py:64-67
            self.final = nn.Sequential(nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, self._head_width(out_channels)))
        elif model_type == "ptv3":
            from .ptv3 import PointTransformerV3
            self.encoder = PointTransformerV3(in_channels=6, cfg=cfg)                   # point_predictor.py:68-69
            self.final = nn.Sequential(nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, self._head_width(out_channels)))
        else:
            raise NotImplementedError(
                f"backbone_type={model_type!r}: 'transformer' (SURVEY.md §8a rows B1-B5), 'pointmlp' (row B6), 'sparseunet' "
                "(row P2) and 'ptv3' (row P1) are built; pcm/mamba3d are out of scope")
        if pretrained_path is not None:
            info = self.load_state_dict(torch.load(pretrained_path), strict=False)
            print(f"Loaded pretrained weights from {pretrained_path}")
            print(f"Missing keys: {info.missing_keys}")
            print(f"Unexpected keys: {info.unexpected_keys}")

    @staticmethod
    def _head_width(split_dimensions) -> int:
        """The reference hard-codes 23 output channels (SH degree 1, point_predictor.py:77-80); other degrees -- the SH
        degree 3 stress configuration of BASELINE.json configs[4] -- need sum(split_dimensions)."""
        return int(sum(split_dimensions))

    def forward(self, x):
        x, center = self.encoder(x, None, None, None, None)
        return self.final(x).permute(0, 2, 1), center

    def forward_feat_fusion(self, x, image_features, c2w_projection_matrix, fusion_mlps, intrinsic):
        x, center = self.encoder.forward(x, image_features, c2w_projection_matrix, fusion_mlps, intrinsic)
        return self.final(x).permute(0, 2, 1), center

    def forward_point_fusion(self, x, image_features, unprojected_coords, fusion_mlps):
        """point_predictor.py:117-134 (scene level): -> (per-point head outputs (n, 23), indices (n, 4) = (batch, grid
        coord)), both in the order of the INPUT points (the sparse tensor keeps voxels sorted internally)."""
        out = self.encoder.forward(x, image_features, unprojected_coords, fusion_mlps)
        if self.cfg.model.backbone_type.lower() == "ptv3":                              # point_predictor.py:128-132
            st = out.sparse_conv_feat
            return self.final(st.features), st.indices
        feats = self.final(out.features)
        if out.perm is not None:
            unsorted = torch.empty_like(feats)
            unsorted[out.perm] = feats
            idx = torch.empty_like(out.indices)
            idx[out.perm] = out.indices
            return unsorted, idx
        return feats, out.indices


class GaussianSplatPredictor(nn.Module):
    MODEL_CONFIGS = {
        "pointmlp": {"feature_dim": 128, "fusion_dim": 128, "final_dim": 128},
        "transformer": {"feature_dim": 384, "fusion_dim": 384, "final_dim": 384},
        "sparseunet": {"feature_dim": 128, "fusion_dim": 32, "final_dim": 32},
        "ptv3": {"feature_dim": 32, "fusion_dim": 32, "final_dim": 32},
    }

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.use_fusion = hasattr(cfg.opt, "use_fusion") and cfg.opt.use_fusion
        self.split_dimensions = [3, 1, 3, 4, 3]
        if cfg.model.max_sh_degree != 0:
            self.split_dimensions.append(((cfg.model.max_sh_degree + 1) ** 2 - 1) * 3)
        pretrained = getattr(cfg.opt, "pretrained_ckpt", None)
        if cfg.opt.level != "object":
            if cfg.opt.level != "scene":
                raise ValueError("Invalid optimization level")
            if self.use_fusion:
                raise NotImplementedError(
                    "opt.level='scene' with opt.use_fusion=true needs PointFusion (fusion/point_fusion.py:36-131: CPU/numpy "
                    "GridSample voxeliser), which is not built; run the scene backbones with opt.use_fusion=false")
            self.point_network = PointFeaturePredictor(cfg, self.split_dimensions, pretrained_path=pretrained)
        elif self.use_fusion:
            # image branch: "stem" (default) = weight-free analytic stand-in; "sdvae" = the reference's frozen
            # Stable-Diffusion VAE architecture (image_predictor.py; random-init unless model.vae_weights names the
            # reference's weights/diffusion_pytorch_model.bin)
            self.image_branch = str(getattr(cfg.model, "image_branch", "stem")).lower()
            if self.image_branch == "sdvae":
                from .image_predictor import ImageFeaturePredictor
                self.image_network = ImageFeaturePredictor(cfg, [128], pretrained_path=getattr(cfg.model, "vae_weights", None))
            elif self.image_branch == "stem":
                self.image_network = FrozenImageStem(cfg, [128])
            elif self.image_branch == "randn":
                self.image_network = RandnImageFeatures(cfg, [128])
            else:
                raise ValueError(f"model.image_branch={self.image_branch!r} (expected 'stem', 'randn' or 'sdvae')")
            self.point_network = PointFeaturePredictor(cfg, self.split_dimensions, pretrained_path=pretrained)
            mc = self.MODEL_CONFIGS[cfg.model.backbone_type]
            in_dim = self.image_network.encoder_config["block_out_channels"][0]
            self.image_conv = nn.Sequential(nn.GroupNorm(32, in_dim, eps=1e-06),
                                            nn.Conv2d(in_dim, mc["feature_dim"], kernel_size=1))
            self.fusion_mlps = nn.Sequential(nn.Linear(mc["feature_dim"] + mc["fusion_dim"], mc["fusion_dim"]),
                                             nn.ReLU())
            self.intrinsic = self._get_camera_intrinsics()
        else:
            self.point_network = PointFeaturePredictor(cfg, self.split_dimensions, pretrained_path=pretrained)
        if cfg.model.max_sh_degree > 0:
            v_to_sh = torch.tensor([[0, 0, -1], [-1, 0, 0], [0, 1, 0]], dtype=torch.float32)
            self.register_buffer("sh_to_v_transform", v_to_sh.transpose(0, 1).unsqueeze(0))
            self.register_buffer("v_to_sh_transform", v_to_sh.unsqueeze(0))

    # -- activations (gaussian_predictor.py:249-254)
    pos_act = staticmethod(torch.tanh)
    opacity_activation = staticmethod(torch.sigmoid)

    @staticmethod
    def scaling_activation(x):
        return torch.exp(torch.clamp(x, -1, 20))

    @staticmethod
    def rotation_activation(x):
        return torch.nn.functional.normalize(x, dim=-1, eps=1e-6)

    def _get_camera_intrinsics(self):
        fov, res = self.cfg.data.fov, self.cfg.data.training_resolution
        K = np.zeros((3, 4))
        K[2, 2] = 1
        focal = (res / 2.0) / math.tan(math.radians(fov / 2.0))
        K[0, 0] = K[1, 1] = focal
        K[0, 2] = K[1, 2] = res / 2.0
        return K

    def forward(self, point_cloud, image: Optional[torch.Tensor] = None,
                source_cameras_view_to_world: Optional[torch.Tensor] = None,
                unprojected_coords: Optional[torch.Tensor] = None, links: Optional[torch.Tensor] = None):
        if self.use_fusion:
            return self._forward_fusion(point_cloud, image, source_cameras_view_to_world, unprojected_coords)
        return self._forward_basic(point_cloud, source_cameras_view_to_world)

    def _forward_basic(self, point_cloud, source_cameras_view_to_world):
        if self.cfg.opt.level == "scene":
            # gaussian_predictor.py:157-167 without the image fusion: backbone -> final -> per-scene lists
            point_features, indices = self.point_network.forward_point_fusion(point_cloud, None, None, None)
            return self._process_network_output_scene(point_features.split(self.split_dimensions, dim=1),
                                                      point_cloud["coord"], indices)
        point_output, center = self.point_network(point_cloud)
        out = self._process_network_output(point_output.split(self.split_dimensions, dim=1), center)
        return {k: v.contiguous() for k, v in out.items()}

    def _forward_fusion(self, point_cloud, image, source_cameras_view_to_world=None, unprojected_coords=None):
        B, N_views = image.shape[0], image.shape[1]
        image = image.reshape(B * N_views, *image.shape[2:])
        if getattr(self.cfg.model, "dense_image_features", False):
            image_output = self.image_network.forward(image)
            image_features = self.image_conv.forward(image_output["decoder_block_3"])     # reference dataflow
        elif self.image_branch in ("sdvae", "randn"):
            # dense (n,128,R,R) decoder features from the frozen VAE; image_conv is still evaluated only at the sampled pixels
            image_features = LazyImageFeatures(self.image_network.forward(image)["decoder_block_3"], self.image_conv)
        else:
            image_output = self.image_network.forward(image, lazy=True)
            image_features = LazyImageFeatures(image_output["decoder_block_3"], self.image_conv)
            image_features.prefetch_stats(source_cameras_view_to_world)
        point_features, center = self.point_network.forward_feat_fusion(
            point_cloud, image_features, source_cameras_view_to_world, self.fusion_mlps, self.intrinsic)
        if point_features.is_cuda and not getattr(self, "force_module_path", False):
            out = self._fused_head(point_features, center)
        else:
            out = self._process_network_output(point_features.split(self.split_dimensions, dim=1), center)
        out = {k: v.reshape(B, N_views * v.shape[1], *v.shape[2:]) for k, v in out.items()}   # _multi_view_union
        # (features_dc / features_rest stay views of "shs" on the fused path: the renderer consumes "shs")
        return {k: (v if ("shs" in out and k.startswith("features_")) else v.contiguous()) for k, v in out.items()}

    def _fused_head(self, point_features, center) -> Dict[str, torch.Tensor]:
        """CUDA path of `_process_network_output`: point_features (B,C,P) is the permuted view of the head's (B,P,C)
        rows; one kernel writes the reference's output dict plus "shs" = [features_dc || features_rest] (B,P,M,3), the
        tensor render_predicted concatenates per call (gaussian_renderer/__init__.py:66-69)."""
        M = (int(self.cfg.model.max_sh_degree) + 1) ** 2
        raw = point_features.permute(0, 2, 1)
        xyz, op, sc, rot, shs = SplatHeadFn.apply(raw, center[:, :, :3], M, float(self.cfg.model.offset_scale),
                                                  bool(self.cfg.model.isotropic))
        out = {"xyz": xyz, "opacity": op, "scaling": sc, "rotation": rot, "features_dc": shs[:, :, :1],
               "features_rest": shs[:, :, 1:], "shs": shs}
        if M == 1:
            out["features_rest"] = torch.zeros((shs.shape[0], 0, 3), dtype=shs.dtype, device=shs.device)
        return out

    @staticmethod
    def _flatten_vector(x):
        return x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)

    def _process_network_output_scene(self, network_output: List[torch.Tensor], center, indices):
        """Scene branch of gaussian_predictor.py:298-364: inputs are (n, C) rows of ALL scenes, output values are lists
        with one (n_b, ...) tensor per scene; quaternions are normalised per point here (the (n, 4) layout makes
        F.normalize(dim=-1) the usual unit-quaternion normalisation, unlike the object branch)."""
        xyz_raw, opacity, scaling, rotation, features_dc = network_output[:5]
        pos = self.pos_act(xyz_raw) * self.cfg.model.offset_scale + center[:, :3]
        if self.cfg.model.isotropic:
            scaling = scaling[:, :1].expand(-1, 3)
        batch = indices[:, 0].long()
        counts = torch.bincount(batch).tolist()              # one host read per step (the reference reads .max().item())
        keys = ("xyz", "opacity", "scaling", "rotation", "features_dc", "features_rest")
        out = {k: [None] * len(counts) for k in keys}
        order = torch.argsort(batch, stable=True)            # scenes are contiguous in the loaders; stable = no-op then
        M1 = (self.cfg.model.max_sh_degree + 1) ** 2 - 1
        vals = {"xyz": pos, "opacity": self.opacity_activation(opacity), "scaling": self.scaling_activation(scaling),
                "rotation": self.rotation_activation(rotation), "features_dc": features_dc.unsqueeze(1)}
        if self.cfg.model.max_sh_degree > 0:
            vals["features_rest"] = network_output[5].reshape(network_output[5].shape[0], -1, 3)
        else:
            vals["features_rest"] = torch.zeros((pos.shape[0], M1, 3), dtype=pos.dtype, device=pos.device)
        start = 0
        for b, c in enumerate(counts):
            sel = order[start:start + c]
            for k in keys:
                out[k][b] = vals[k][sel]
            start += c
        return out

    def _process_network_output(self, network_output: List[torch.Tensor], center) -> Dict[str, torch.Tensor]:
        """Object branch of gaussian_predictor.py:279-328."""
        xyz_raw, opacity, scaling, rotation, features_dc = network_output[:5]
        pos = self.pos_act(xyz_raw) * self.cfg.model.offset_scale
        pos = pos.permute(0, 2, 1) + center[:, :, :3]
        if self.cfg.model.isotropic:
            scaling = scaling[:, :1].expand(-1, 3, -1)
        out = {
            "xyz": pos,
            "opacity": self._flatten_vector(self.opacity_activation(opacity)),
            "scaling": self._flatten_vector(self.scaling_activation(scaling)),
            "rotation": self._flatten_vector(self.rotation_activation(rotation)),
            "features_dc": self._flatten_vector(features_dc).unsqueeze(2),
        }
        if self.cfg.model.max_sh_degree > 0:
            rest = self._flatten_vector(network_output[5])
            out["features_rest"] = rest.reshape(*rest.shape[:2], -1, 3)
        else:
            dc = out["features_dc"]
            out["features_rest"] = torch.zeros((dc.shape[0], (self.cfg.model.max_sh_degree + 1) ** 2 - 1, 3),
                                               dtype=dc.dtype, device=dc.device)
        return out
