"""Reference arm of bench.py (`--impl reference`): the REFERENCE's own backbone modules on the host cores.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing under unipre3d_b200/ imports this file, and this file imports nothing
from unipre3d_b200/ at run time (the GPU box has no /root/reference either: everything it needs is staged by `stage()`).

What runs (SURVEY.md §8d "CPU baseline"):
  * backbone = the reference's `PointTransformerEncoder` (openpoints/models/backbone/transformer.py:246-327: mini-PointNet
    tokenizer, 16 pre-LN blocks, final norm) and `FeatureFusion` (fusion/feat_fusion.py:58-145), the UNMODIFIED files,
    copied by `stage()` from /root/reference into oracle/_ref/pyref/ (git-ignored, travels to the GPU box like the .so
    files) and imported with the same stubs tests/golden/make_golden.py uses (timm DropPath / trunc_normal_, the
    registry decorator, SubsampleGroup);
  * `SubsampleGroup` (FPS + ball query + grouping) = the CPU restatement of the reference's CUDA kernels
    (oracle/pointops_oracle.c) -- the reference has no CPU path for them (openpoints/models/layers/subsample.py:93-100);
  * image branch = `image_conv` of model/gaussian_predictor.py:210-215 (GroupNorm(32,128) + Conv2d(128,384,1)) evaluated
    densely on N(0,1) (B,128,R,R) decoder features, as the reference does (synthetic image features in place
    of the absent SD-VAE weights);
  * head = `final` of model/point_predictor.py:77-80 + the activations of model/gaussian_predictor.py:249-254, 298-328
    restated below (that module cannot be imported: diffusers / Mamba extensions);
  * rasterizer = C/OpenMP restatement oracle/raster_oracle.c, one call per (object, view) as train_network.py:418-442;
  * loss / clip / AdamW = plain torch (utils/loss_utils.py:23-45, train_network.py:156-159, 386).
All 8 objects x 4 views of the benchmarked batch are processed every step, fp32, all host threads.
"""
from __future__ import annotations

import importlib.util
import math
import os
import shutil
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
STAGE_DIR = os.path.join(HERE, "_ref", "pyref")
FILES = ["openpoints/models/backbone/transformer.py", "fusion/feat_fusion.py"]
BATCH_FILE = os.path.join(STAGE_DIR, "bench_batch.npz")


def stage(force: bool = False) -> bool:
    """Run where /root/reference exists (this container, from __graft_entry__.build()): copies the reference's files and
    writes the benchmarked synthetic batch as plain arrays.  Returns True when the staged tree is complete."""
    os.makedirs(STAGE_DIR, exist_ok=True)
    if os.path.isdir(REF_SRC):
        for f in FILES:
            dst = os.path.join(STAGE_DIR, f.replace("/", "__"))
            if force or not os.path.exists(dst):
                shutil.copyfile(os.path.join(REF_SRC, f), dst)
    if force or not os.path.exists(BATCH_FILE):
        root = os.path.dirname(HERE)
        if root not in sys.path:
            sys.path.insert(0, root)
        from unipre3d_b200 import synthetic            # build-time only: the batch bench.py's own arm uses (seed 0)
        from unipre3d_b200.config import compose
        cfg = compose(overrides=["data.training_resolution=256", "opt.batch_size=8"])
        b = synthetic.make_batch(cfg, 8, 8192, seed=0, image_dtype="uint8")
        np.savez(BATCH_FILE, pos=b["point_cloud"]["pos"].numpy(), gt_images=b["gt_images"].numpy(),
                 **{k: b[k].numpy() for k in ("world_view_transforms", "view_to_world_transforms", "full_proj_transforms",
                                              "camera_centers")},
                 cfg=np.array([cfg.data.fov, cfg.data.input_images, cfg.opt.imgs_per_obj, cfg.model.max_sh_degree,
                               cfg.opt.base_lr, cfg.opt.betas[0], cfg.opt.betas[1], cfg.opt.non_bg_color_loss_rate,
                               cfg.opt.bg_color_loss_rate, cfg.model.offset_scale], np.float64))
    return available()


def available() -> bool:
    return all(os.path.exists(os.path.join(STAGE_DIR, f.replace("/", "__"))) for f in FILES) and os.path.exists(BATCH_FILE)


class _DropPath(torch.nn.Module):      # timm.models.layers.DropPath (scale_by_keep=True)
    def __init__(self, drop_prob=0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        return x * mask.div_(keep)


class _Registry:
    def register_module(self, *a, **k):
        return lambda cls: cls


class OracleSubsampleGroup(torch.nn.Module):
    """SubsampleGroup (openpoints/models/layers/group_embed.py:39-57) on the CPU restatement of the reference kernels."""

    def __init__(self, num_groups, group_size, subsample="fps", group="ballquery", radius=0.1, **kw):
        super().__init__()
        self.num_groups, self.group_size, self.radius = num_groups, group_size, radius

    def forward(self, p, x=None):
        from . import oracle_lib as ol
        pn = p.detach().cpu().numpy().astype(np.float32)
        fidx = ol.fps(pn, self.num_groups)
        center = np.take_along_axis(pn, fidx.astype(np.int64)[..., None], 1)
        idx = ol.ball_query(self.radius, self.group_size, pn, center)
        grouped = ol.group(np.ascontiguousarray(pn.transpose(0, 2, 1)), idx)
        neigh = grouped - center.transpose(0, 2, 1)[..., None]
        return torch.from_numpy(neigh), torch.from_numpy(center)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference_modules():
    """-> (transformer module, FeatureFusion class) of the staged reference files."""
    if not available():
        raise RuntimeError("oracle/_ref/pyref is not staged (run __graft_entry__.build() where /root/reference exists)")
    _stub("timm"); _stub("timm.models")
    _stub("timm.models.layers", DropPath=_DropPath, trunc_normal_=torch.nn.init.trunc_normal_)
    _stub("openpoints"); _stub("openpoints.models")
    _stub("openpoints.models.build", MODELS=_Registry())
    _stub("openpoints.models.layers", SubsampleGroup=OracleSubsampleGroup)
    ff = _load("ref_feat_fusion", os.path.join(STAGE_DIR, "fusion__feat_fusion.py"))
    _stub("fusion", FeatureFusion=ff.FeatureFusion)
    tr = _load("ref_transformer", os.path.join(STAGE_DIR, "openpoints__models__backbone__transformer.py"))
    return tr, ff.FeatureFusion


class _OracleRender(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, opacity, scaling, rotation, shs, view, proj, campos, bg, W, H, tanfov, deg):
        from . import oracle_lib as ol
        sc = ol.Scene(xyz.detach().numpy(), opacity.detach().numpy(), scaling.detach().numpy(), rotation.detach().numpy(),
                      view.numpy(), proj.numpy(), campos.numpy(), W, H, tanfov, tanfov, shs=shs.detach().numpy(),
                      sh_degree=deg, bg=bg.numpy())
        ctx.sc, ctx.shapes = sc, (opacity.shape, shs.shape)
        return torch.from_numpy(ol.render(sc)["color"])

    @staticmethod
    def backward(ctx, g):
        from . import oracle_lib as ol
        r = ol.render(ctx.sc, g.contiguous().numpy())["grads"]
        t = torch.from_numpy
        return (t(r["means3D"]), t(r["opacities"]).reshape(ctx.shapes[0]), t(r["scales"]), t(r["rotations"]),
                t(r["shs"]).reshape(ctx.shapes[1]), None, None, None, None, None, None, None, None)


class RefStepper:
    """One pre-training step of the transformer config on the CPU with the reference's backbone modules."""

    def __init__(self, n_objects: int = 8, seed: int = 0):
        tr, _ = load_reference_modules()
        z = np.load(BATCH_FILE)
        (self.fov, ni, nv, deg, lr, b1, b2, self.non_bg, self.bg_rate, self.offset_scale) = [float(v) for v in z["cfg"]]
        self.ni, self.nv, self.deg = int(ni), int(nv), int(deg)
        B = self.B = n_objects
        t = lambda k: torch.from_numpy(z[k][:B].copy())
        self.pos = t("pos")
        self.gt = t("gt_images").float().div_(255.0)                     # the reference's loader divides on the host
        self.wv, self.v2w, self.fp, self.cc = t("world_view_transforms"), t("view_to_world_transforms"), \
            t("full_proj_transforms"), t("camera_centers")
        self.R = int(self.gt.shape[-1])
        torch.manual_seed(seed)
        # model/point_predictor.py:60-63, 77-80; model/gaussian_predictor.py:196-227
        self.encoder = tr.PointTransformerEncoder(in_channels=3, num_groups=128, encoder_dims=384, depth=16)
        M = (self.deg + 1) ** 2
        self.final = torch.nn.Sequential(torch.nn.Linear(384, 128), torch.nn.ReLU(), torch.nn.Linear(128, 11 + 3 * M))
        self.image_conv = torch.nn.Sequential(torch.nn.GroupNorm(32, 128, eps=1e-6), torch.nn.Conv2d(128, 384, 1))
        self.fusion_mlps = torch.nn.Sequential(torch.nn.Linear(768, 384), torch.nn.ReLU())
        self.split = [3, 1, 3, 4, 3] + ([3 * (M - 1)] if M > 1 else [])
        self.modules = torch.nn.ModuleList([self.encoder, self.final, self.image_conv, self.fusion_mlps])
        # the frozen SD-VAE's decoder_block_3 output, replaced by N(0,1) features
        self.decoder_features = torch.randn(B * self.ni, 128, self.R, self.R)
        focal = (self.R / 2.0) / math.tan(math.radians(self.fov / 2.0))
        self.intrinsic = np.zeros((3, 4))
        self.intrinsic[0, 0] = self.intrinsic[1, 1] = focal
        self.intrinsic[0, 2] = self.intrinsic[1, 2] = self.R / 2.0
        self.intrinsic[2, 2] = 1
        self.opt = torch.optim.AdamW([p for p in self.modules.parameters() if p.requires_grad], lr=lr, eps=1e-15,
                                     betas=(b1, b2))

    def n_views(self) -> int:
        return self.B * self.nv

    def _activations(self, out, center):
        """model/gaussian_predictor.py:298-328 (object level) with 249-254."""
        raw = out.split(self.split, dim=1)
        xyz_raw, opacity, scaling, rotation, dc = raw[:5]
        xyz = (torch.tanh(xyz_raw) * self.offset_scale).permute(0, 2, 1) + center
        res = {"xyz": xyz, "opacity": torch.sigmoid(opacity).permute(0, 2, 1),
               "scaling": torch.exp(torch.clamp(scaling, -1, 20)).permute(0, 2, 1),
               # F.normalize(dim=-1) on (B,4,P): over the POINT axis, as the reference does
               "rotation": torch.nn.functional.normalize(rotation, dim=-1, eps=1e-6).permute(0, 2, 1),
               "features_dc": dc.permute(0, 2, 1).unsqueeze(2)}
        B, P = xyz.shape[:2]
        res["features_rest"] = (raw[5].permute(0, 2, 1).reshape(B, P, -1, 3) if len(raw) > 5
                                else torch.zeros(B, P, 3, 3))
        return res

    def step(self) -> float:
        self.modules.train()
        image_features = self.image_conv(self.decoder_features)
        x, center = self.encoder.forward(self.pos, image_features, self.v2w[:, :self.ni], self.fusion_mlps, self.intrinsic)
        out = self.final(x).permute(0, 2, 1)
        sp = self._activations(out, center)
        tanfov = math.tan(self.fov * math.pi / 360)
        bg = torch.zeros(3)
        imgs, gts = [], []
        for b in range(self.B):
            shs = torch.cat([sp["features_dc"][b], sp["features_rest"][b]], 1)
            for r in range(self.ni, self.ni + self.nv):
                imgs.append(_OracleRender.apply(sp["xyz"][b], sp["opacity"][b], sp["scaling"][b], sp["rotation"][b], shs,
                                                self.wv[b, r], self.fp[b, r], self.cc[b, r], bg, self.R, self.R, tanfov,
                                                self.deg))
                gts.append(self.gt[b, r])
        r, gt = torch.stack(imgs), torch.stack(gts)
        is_bg = (torch.isclose(gt[:, 0], bg[0], atol=1e-6) & torch.isclose(gt[:, 1], bg[1], atol=1e-6)
                 & torch.isclose(gt[:, 2], bg[2], atol=1e-6))
        w = torch.where(is_bg, 2 * self.bg_rate / (self.bg_rate + self.non_bg), 2 * self.non_bg / (self.bg_rate + self.non_bg))
        loss = (((r - gt) ** 2) * w.unsqueeze(1)).mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.modules.parameters(), max_norm=1.0)
        self.opt.step()
        self.opt.zero_grad()
        return float(loss.detach())
