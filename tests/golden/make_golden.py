"""Generates tests/golden/*.npz by IMPORTING THE REFERENCE'S OWN PYTHON from /root/reference (run in the authoring
container only; the fixtures travel, /root/reference does not).

    python tests/golden/make_golden.py

Third-party modules the reference imports but that are absent here (timm, spconv, diffusers, easydict ...) are
replaced by minimal stubs; the CUDA-only tokenizer (SubsampleGroup) is stubbed by precomputed groups so the fixtures
pin the ARITHMETIC of the reference modules:
  transformer_encoder.npz : openpoints/models/backbone/transformer.py  PointTransformerEncoder (small config) fwd + grads
  transformer_encoder_w128.npz : the same module at width 128 / head_dim 64 (the widths the CUDA fused paths accept)
  feature_fusion.npz      : fusion/feat_fusion.py  FeatureFusion.__call__
  process_output.npz      : model/gaussian_predictor.py  _process_network_output / _init_activations (object level)
  utils.npz               : utils/loss_utils.py focal_l2_loss, utils/graphics_utils.py matrices, utils/sh_utils.py eval_sh
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class _DropPath(torch.nn.Module):  # timm.models.layers.DropPath (scale_by_keep=True)
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


class _FixedGroups(torch.nn.Module):
    """stands in for SubsampleGroup (CUDA-only in the reference): returns precomputed (neighborhood, center)"""
    neighborhood = None
    center = None

    def __init__(self, *a, **k):
        super().__init__()

    def forward(self, p, x=None):
        return _FixedGroups.neighborhood, _FixedGroups.center


def main():
    stub("timm"); stub("timm.models")
    stub("timm.models.layers", DropPath=_DropPath, trunc_normal_=torch.nn.init.trunc_normal_)
    stub("openpoints"); stub("openpoints.models")
    stub("openpoints.models.build", MODELS=_Registry())
    stub("openpoints.models.layers", SubsampleGroup=_FixedGroups)
    ff = load("ref_feat_fusion", os.path.join(REF, "fusion/feat_fusion.py"))
    stub("fusion", FeatureFusion=ff.FeatureFusion)
    tr = load("ref_transformer", os.path.join(REF, "openpoints/models/backbone/transformer.py"))

    # ---------------------------------------------------------------- transformer encoder (small config)
    torch.manual_seed(0)
    B, G, K, R = 2, 16, 8, 24
    cfgk = dict(num_groups=G, group_size=K, encoder_dims=64, trans_dim=48, depth=3, num_heads=6, drop_path_rate=0.1)
    enc = tr.PointTransformerEncoder(in_channels=3, **cfgk)
    with torch.no_grad():  # make BN affine / cls token non-trivial
        for m in enc.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
        enc.cls_token.normal_(0, 0.02)
    fusion_mlps = torch.nn.Sequential(torch.nn.Linear(96, 48), torch.nn.ReLU())
    neigh = torch.randn(B, 3, G, K) * 0.05
    center = torch.randn(B, G, 3) * 0.2
    _FixedGroups.neighborhood, _FixedGroups.center = neigh, center
    pts = torch.randn(B, 64, 3)
    img_feat = torch.randn(B, 48, R, R, requires_grad=True)
    from unipre3d_b200 import camera as cam
    import math
    fov = 49.13434264120263
    proj = cam.get_projection_matrix(0.5, 2.0, math.radians(fov), math.radians(fov))
    c2w = torch.stack([cam.make_view(*cam.look_at_pose(30.0 * i, 20.0, 1.75), proj)["view_to_world_transform"] for i in range(B)]).unsqueeze(1)
    K_in = np.zeros((3, 4)); focal = (R / 2.0) / math.tan(math.radians(fov / 2.0))
    K_in[0, 0] = K_in[1, 1] = focal; K_in[0, 2] = K_in[1, 2] = R / 2.0; K_in[2, 2] = 1
    enc.train()
    for m in enc.modules():
        if isinstance(m, _DropPath):
            m.drop_prob = 0.0
    enc.zero_grad()
    out_tr, _ = enc(pts, img_feat, c2w, fusion_mlps, K_in)   # train mode (batch-stat BN), no stochastic depth
    wsum = torch.randn_like(out_tr)
    (out_tr * wsum).sum().backward()
    enc.eval()                                               # eval AFTER the train pass: uses the saved running stats
    out_eval, c_eval = enc(pts, img_feat, c2w, fusion_mlps, K_in)
    sd = {k: v.detach().numpy() for k, v in enc.state_dict().items()}
    grads = {k: p.grad.numpy() for k, p in enc.named_parameters() if p.grad is not None}
    np.savez_compressed(os.path.join(OUT, "transformer_encoder.npz"),
                        cfg=np.array([G, K, 64, 48, 3, 6]), pts=pts.numpy(), neighborhood=neigh.numpy(), center=center.numpy(),
                        img_feat=img_feat.detach().numpy(), c2w=c2w.numpy(), intrinsic=K_in, wsum=wsum.numpy(),
                        out_eval=out_eval.detach().numpy(), out_train=out_tr.detach().numpy(),
                        grad_img_feat=img_feat.grad.numpy(),
                        fusion_w=fusion_mlps[0].weight.detach().numpy(), fusion_b=fusion_mlps[0].bias.detach().numpy(),
                        **{"sd." + k: v for k, v in sd.items()}, **{"grad." + k: v for k, v in grads.items()})

    # ---------------------------------------------------------------- FeatureFusion alone (incl. occlusion + out-of-image)
    torch.manual_seed(1)
    B, N, C, R = 3, 40, 8, 16
    x = torch.randn(B, N + 1, C)
    ctr = torch.randn(B, N, 3) * 0.35
    ctr[:, 5] = ctr[:, 4] * 1.02          # nearly colinear with the camera ray -> same pixel, depth test decides
    ctr[:, 7, :] = 5.0                    # outside the image
    feat = torch.randn(B, C, R, R)
    c2w = torch.stack([cam.make_view(*cam.look_at_pose(50.0 * i, 10.0 + 20 * i, 1.75), proj)["view_to_world_transform"] for i in range(B)]).unsqueeze(1)
    K2 = np.zeros((3, 4)); focal = (R / 2.0) / math.tan(math.radians(fov / 2.0))
    K2[0, 0] = K2[1, 1] = focal; K2[0, 2] = K2[1, 2] = R / 2.0; K2[2, 2] = 1
    mlp = torch.nn.Sequential(torch.nn.Linear(2 * C, C), torch.nn.ReLU())
    y = ff.FeatureFusion(mlp)(x, ctr, feat, c2w, K2)
    np.savez_compressed(os.path.join(OUT, "feature_fusion.npz"), x=x.numpy(), center=ctr.numpy(), feat=feat.numpy(),
                        c2w=c2w.numpy(), intrinsic=K2, w=mlp[0].weight.detach().numpy(), b=mlp[0].bias.detach().numpy(),
                        y=y.detach().numpy())

    # ---------------------------------------------------------------- GaussianSplatPredictor._process_network_output
    stub("spconv"); stub("spconv.pytorch")
    stub("model"); stub("model.image_predictor", ImageFeaturePredictor=object)
    stub("model.point_predictor", PointFeaturePredictor=object)
    gp = load("ref_gaussian_predictor", os.path.join(REF, "model/gaussian_predictor.py"))
    from types import SimpleNamespace as NS
    res = {}
    for name, (deg, iso) in {"deg1": (1, False), "deg0_iso": (0, True)}.items():
        cfg = NS(model=NS(max_sh_degree=deg, isotropic=iso, offset_scale=0.7), opt=NS(use_fusion=True, level="object"))
        m = gp.GaussianSplatPredictor.__new__(gp.GaussianSplatPredictor)
        torch.nn.Module.__init__(m)
        m.cfg = cfg
        m._init_activations()
        split = m._get_network_params()
        torch.manual_seed(2)
        Bq, P = 2, 37
        raw = torch.randn(Bq, sum(split), P) * 1.5
        raw[:, 4:7, :3] = torch.tensor([-3.0, 25.0, 0.0]).view(1, 3, 1)    # exercise clamp(-1, 20)
        ctr = torch.randn(Bq, P, 3)
        out = m._process_network_output(raw.split(split, dim=1), ctr, None, is_scene_level=False)
        out = m._make_contiguous(m._multi_view_union(out, Bq, 1))
        res.update({f"{name}.raw": raw.numpy(), f"{name}.center": ctr.numpy()})
        res.update({f"{name}.out.{k}": v.numpy() for k, v in out.items()})
    np.savez_compressed(os.path.join(OUT, "process_output.npz"), **res)

    # ---------------------------------------------------------------- utils: loss, camera matrices, SH
    lu = load("ref_loss_utils", os.path.join(REF, "utils/loss_utils.py"))
    gu = load("ref_graphics_utils", os.path.join(REF, "utils/graphics_utils.py"))
    su = load("ref_sh_utils", os.path.join(REF, "utils/sh_utils.py"))
    torch.manual_seed(3)
    r = torch.rand(4, 3, 12, 10)
    gt = torch.rand(4, 3, 12, 10)
    gt[:, :, :5] = 0.0
    gt[0, 0, 0, 0] = 5e-7      # still "close" to black (atol 1e-6)
    gt[1, 1, 1, 1] = 2e-6      # not close
    gt_w = gt.clone(); gt_w[:, :, :5] = 1.0
    l_black = lu.focal_l2_loss(r, gt, [0.0, 0.0, 0.0], 4, 1)
    l_white = lu.focal_l2_loss(r, gt_w, [1.0, 1.0, 1.0], 4, 1)
    Rm, t = cam.look_at_pose(33.0, 21.0, 1.75)
    w2v = gu.getWorld2View2(Rm, t); v2w = gu.getView2World(Rm, t)
    pm = gu.getProjectionMatrix(0.5, 2.0, math.radians(fov), math.radians(fov)).numpy()
    sh = torch.randn(50, 3, 16); dirs = torch.nn.functional.normalize(torch.randn(50, 3), dim=1)
    shv = {f"sh_deg{d}": su.eval_sh(d, sh, dirs).numpy() for d in range(4)}
    np.savez_compressed(os.path.join(OUT, "utils.npz"), r=r.numpy(), gt=gt.numpy(), gt_w=gt_w.numpy(),
                        l_black=float(l_black), l_white=float(l_white), R=Rm, t=t, w2v=w2v, v2w=v2w, proj=pm,
                        sh=sh.numpy(), dirs=dirs.numpy(), **shv)
    # ---------------------------------------------------------------- pointMLP encoder/decoder (small config, in_channels=4)
    from oracle import oracle_lib as ol

    def ref_fps(xyz, npoint):
        """what the reference's CUDA kernel does with a (B,N,C) buffer: it indexes it as (B,N,3)
        (sampling_gpu.cu:117-131) -- replayed by the CPU oracle on the same reinterpretation"""
        B, N, C = xyz.shape
        flat = xyz.detach().contiguous().reshape(-1).numpy()
        view = np.lib.stride_tricks.as_strided(flat, (B, N, 3), (N * 3 * 4, 12, 4)).copy()
        return torch.from_numpy(ol.fps(view, npoint))

    layers = sys.modules["openpoints.models.layers"]
    for n in ("random_sample", "LocalAggregation", "create_convblock2d", "three_interpolate", "three_nn",
              "gather_operation", "create_linearblock", "create_convblock1d", "create_grouper", "fps"):
        setattr(layers, n, None)
    layers.furthest_point_sample = ref_fps
    layers.__path__ = []
    stub("openpoints.models.layers.group", QueryAndGroup=None)
    pkg = stub("openpoints.models.backbone"); pkg.__path__ = []
    sys.modules["openpoints.models"].__path__ = []
    sys.modules["openpoints"].__path__ = []
    pm_mod = load("openpoints.models.backbone.pointmlp", os.path.join(REF, "openpoints/models/backbone/pointmlp.py"))
    torch.manual_seed(4)
    B, N, R = 2, 64, 16
    kw = dict(in_channels=4, embed_dim=8, groups=1, res_expansion=1.0, activation="relu", bias=False, use_xyz=False,
              normalize="anchor", dim_expansion=[2, 2, 2, 2], pre_blocks=[2, 2, 2, 2], pos_blocks=[2, 2, 2, 2],
              k_neighbors=[4, 4, 4, 4], reducers=[2, 2, 2, 2], de_blocks=[2, 2, 2, 2], de_dims=[64, 32, 16, 16])
    enc = pm_mod.PointMLPEncoder(**kw)
    with torch.no_grad():
        for m in enc.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
        for g in enc.local_grouper_list:
            g.affine_alpha.uniform_(0.5, 1.5); g.affine_beta.normal_(0, 0.2)
    xyz = torch.randn(B, N, 3) * 0.25
    p4 = torch.cat([xyz, xyz[:, :, 2:3] - xyz[:, :, 2:3].min(1, keepdim=True)[0]], -1)
    img = torch.randn(B, 16, R, R, requires_grad=True)
    fusion_mlps = torch.nn.Sequential(torch.nn.Linear(32, 16), torch.nn.ReLU())
    c2w = torch.stack([cam.make_view(*cam.look_at_pose(40.0 * i, 25.0, 1.75), proj)["view_to_world_transform"] for i in range(B)]).unsqueeze(1)
    K3 = np.zeros((3, 4)); focal = (R / 2.0) / math.tan(math.radians(fov / 2.0))
    K3[0, 0] = K3[1, 1] = focal; K3[0, 2] = K3[1, 2] = R / 2.0; K3[2, 2] = 1
    enc.train()
    out, p_out = enc({"pos": p4}, img, c2w, fusion_mlps, K3)
    wsum = torch.randn_like(out)
    (out * wsum).sum().backward()
    fps_stage1 = ref_fps(p4, N // 2)
    enc.eval()
    out_eval, _ = enc({"pos": p4}, img, c2w, fusion_mlps, K3)
    np.savez_compressed(os.path.join(OUT, "pointmlp_encoder.npz"), p4=p4.numpy(), img=img.detach().numpy(), c2w=c2w.numpy(),
                        intrinsic=K3, wsum=wsum.numpy(), out_train=out.detach().numpy(), out_eval=out_eval.detach().numpy(),
                        p_out=p_out.numpy(), grad_img=img.grad.numpy(), fps_stage1=fps_stage1.numpy(),
                        fusion_w=fusion_mlps[0].weight.detach().numpy(), fusion_b=fusion_mlps[0].bias.detach().numpy(),
                        **{"sd." + k: v.detach().numpy() for k, v in enc.state_dict().items()},
                        **{"grad." + k: q.grad.numpy() for k, q in enc.named_parameters() if q.grad is not None})
    print("wrote", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


def blocks_only():
    """transformer_blocks_w128.npz : openpoints/models/backbone/transformer.py TransformerEncoder (the Block stack alone,
    width 128 so that the CUDA fused stack -- width a multiple of 128 -- is pinned against the reference module)."""
    stub("timm"); stub("timm.models")
    stub("timm.models.layers", DropPath=_DropPath, trunc_normal_=torch.nn.init.trunc_normal_)
    stub("openpoints"); stub("openpoints.models")
    stub("openpoints.models.build", MODELS=_Registry())
    stub("openpoints.models.layers", SubsampleGroup=_FixedGroups)
    stub("fusion", FeatureFusion=object)
    tr = load("ref_transformer", os.path.join(REF, "openpoints/models/backbone/transformer.py"))
    torch.manual_seed(11)
    B, L, C, depth, heads = 3, 17, 128, 2, 4
    enc = tr.TransformerEncoder(embed_dim=C, depth=depth, num_heads=heads, drop_path_rate=[0.0, 0.1])
    with torch.no_grad():
        for q in enc.parameters():
            if q.ndim == 1:
                q.add_(torch.randn_like(q) * 0.2)     # non-trivial LayerNorm affine parameters and biases
    enc.train()
    for m in enc.modules():
        if isinstance(m, _DropPath):
            m.drop_prob = 0.0
    x = torch.randn(B, L, C, requires_grad=True)
    pos = (torch.randn(B, L, C) * 0.3).requires_grad_(True)
    out = enc(x, pos, None, None, None, None, None)
    wsum = torch.randn_like(out)
    (out * wsum).sum().backward()
    np.savez_compressed(os.path.join(OUT, "transformer_blocks_w128.npz"), cfg=np.array([B, L, C, depth, heads]),
                        x=x.detach().numpy(), pos=pos.detach().numpy(), wsum=wsum.numpy(), out=out.detach().numpy(),
                        grad_x=x.grad.numpy(), grad_pos=pos.grad.numpy(),
                        **{"sd." + k: v.detach().numpy() for k, v in enc.state_dict().items()},
                        **{"grad." + k: q.grad.numpy() for k, q in enc.named_parameters()})
    print("wrote transformer_blocks_w128.npz")


def encoder_w128():
    """transformer_encoder_w128.npz : the WHOLE PointTransformerEncoder (transformer.py:246-327) at a width the CUDA fused
    paths accept (trans_dim 128 = one LayerNorm vector group, head_dim 64 = the own attention kernel, 32 x 16-point
    groups through the fused mini-PointNet), train-mode forward + all gradients -- pins the composition of mini-PointNet,
    reduce_dim, cls/pos embedding, fused Block stack, FeatureFusion and the final norm on the GPU."""
    stub("timm"); stub("timm.models")
    stub("timm.models.layers", DropPath=_DropPath, trunc_normal_=torch.nn.init.trunc_normal_)
    stub("openpoints"); stub("openpoints.models")
    stub("openpoints.models.build", MODELS=_Registry())
    stub("openpoints.models.layers", SubsampleGroup=_FixedGroups)
    ff = load("ref_feat_fusion", os.path.join(REF, "fusion/feat_fusion.py"))
    stub("fusion", FeatureFusion=ff.FeatureFusion)
    tr = load("ref_transformer", os.path.join(REF, "openpoints/models/backbone/transformer.py"))
    torch.manual_seed(7)
    B, G, K, R, ed, td, depth, heads = 2, 32, 16, 24, 128, 128, 2, 2
    enc = tr.PointTransformerEncoder(in_channels=3, num_groups=G, group_size=K, encoder_dims=ed, trans_dim=td, depth=depth,
                                     num_heads=heads, drop_path_rate=0.1)
    with torch.no_grad():
        for m in enc.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
        enc.cls_token.normal_(0, 0.02)
    fusion_mlps = torch.nn.Sequential(torch.nn.Linear(2 * td, td), torch.nn.ReLU())
    neigh = torch.randn(B, 3, G, K) * 0.05
    center = torch.randn(B, G, 3) * 0.2
    _FixedGroups.neighborhood, _FixedGroups.center = neigh, center
    pts = torch.randn(B, 64, 3)
    img_feat = torch.randn(B, td, R, R, requires_grad=True)
    from unipre3d_b200 import camera as cam
    import math
    fov = 49.13434264120263
    proj = cam.get_projection_matrix(0.5, 2.0, math.radians(fov), math.radians(fov))
    c2w = torch.stack([cam.make_view(*cam.look_at_pose(30.0 * i, 20.0, 1.75), proj)["view_to_world_transform"] for i in range(B)]).unsqueeze(1)
    K_in = np.zeros((3, 4)); focal = (R / 2.0) / math.tan(math.radians(fov / 2.0))
    K_in[0, 0] = K_in[1, 1] = focal; K_in[0, 2] = K_in[1, 2] = R / 2.0; K_in[2, 2] = 1
    enc.train()
    for m in enc.modules():
        if isinstance(m, _DropPath):
            m.drop_prob = 0.0
    out_tr, _ = enc(pts, img_feat, c2w, fusion_mlps, K_in)
    wsum = torch.randn_like(out_tr)
    (out_tr * wsum).sum().backward()
    f16 = lambda a: a        # fp32 fixture: tolerances in the test are stated against exact reference values
    np.savez_compressed(os.path.join(OUT, "transformer_encoder_w128.npz"),
                        cfg=np.array([G, K, ed, td, depth, heads]), pts=pts.numpy(), neighborhood=neigh.numpy(),
                        center=center.numpy(), img_feat=img_feat.detach().numpy(), c2w=c2w.numpy(), intrinsic=K_in,
                        wsum=wsum.numpy(), out_train=out_tr.detach().numpy(), grad_img_feat=img_feat.grad.numpy(),
                        fusion_w=fusion_mlps[0].weight.detach().numpy(), fusion_b=fusion_mlps[0].bias.detach().numpy(),
                        grad_fusion_w=fusion_mlps[0].weight.grad.numpy(), grad_fusion_b=fusion_mlps[0].bias.grad.numpy(),
                        **{"sd." + k: f16(v.detach().numpy()) for k, v in enc.state_dict().items()},
                        **{"grad." + k: q.grad.numpy() for k, q in enc.named_parameters() if q.grad is not None})
    print("wrote transformer_encoder_w128.npz")


def serialization_fixture():
    """serialization.npz : pointcept/models/utils/serialization (encode: z / z-trans, with batch) on voxel coordinates --
    the reference's own functions, loaded straight from /root/reference."""
    pkg = stub("ref_serialization"); pkg.__path__ = [os.path.join(REF, "pointcept/models/utils/serialization")]
    z = load("ref_serialization.z_order", os.path.join(REF, "pointcept/models/utils/serialization/z_order.py"))
    h = load("ref_serialization.hilbert", os.path.join(REF, "pointcept/models/utils/serialization/hilbert.py"))
    d = load("ref_serialization.default", os.path.join(REF, "pointcept/models/utils/serialization/default.py"))
    g = torch.Generator().manual_seed(21)
    out = {}
    for name, n, depth, nb in (("small", 257, 7, 3), ("deep", 1000, 16, 5), ("bigbatch", 64, 10, 300)):
        hi = (1 << depth)
        coord = torch.randint(0, hi, (n, 3), generator=g, dtype=torch.int32)
        coord[0] = hi - 1                                      # all bits set
        coord[1] = 0
        batch = torch.sort(torch.randint(0, nb, (n,), generator=g))[0]
        out[f"{name}.coord"], out[f"{name}.batch"], out[f"{name}.depth"] = coord.numpy(), batch.numpy(), np.int64(depth)
        for o in ("z", "z-trans", "hilbert", "hilbert-trans"):
            out[f"{name}.code.{o}"] = d.encode(coord, batch, depth, order=o).numpy()
        out[f"{name}.code_nobatch.z"] = d.encode(coord, None, depth, order="z").numpy()
    np.savez_compressed(os.path.join(OUT, "serialization.npz"), **out)
    print("wrote serialization.npz")


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    if len(sys.argv) > 1 and sys.argv[1] == "serialization":
        serialization_fixture()
    elif len(sys.argv) > 1 and sys.argv[1] == "blocks":
        blocks_only()
    elif len(sys.argv) > 1 and sys.argv[1] == "encoder_w128":
        encoder_w128()
    else:
        main()
        blocks_only()
        encoder_w128()
        serialization_fixture()
