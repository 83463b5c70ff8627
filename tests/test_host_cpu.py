"""Host-side logic that needs no GPU: config surface, layouts, synthetic batches, EMA schedule, sharding,
state-dict naming, and the CPU port of the step."""
import math

import numpy as np
import pytest
import torch


def test_config_surface_and_overrides():
    from unipre3d_b200.config import compose
    cfg = compose()
    assert cfg.model.backbone_type == "transformer" and cfg.model.max_sh_degree == 1 and cfg.model.isotropic is False
    assert cfg.data.fov == pytest.approx(49.13434264120263) and cfg.data.training_resolution == 128
    assert cfg.data.znear == 0.5 and cfg.data.zfar == 2 and cfg.data.white_background is False
    assert cfg.opt.batch_size == 32 and cfg.opt.imgs_per_obj == 4 and cfg.opt.loss == "focal_l2"
    assert cfg.opt.ema.beta == 0.9999 and cfg.opt.betas == [0.9, 0.999] and cfg.general.multiple_gpu is False
    assert hasattr(cfg.data, "training_resolution") and not hasattr(cfg.data, "training_height")
    c2 = compose(overrides=["opt.batch_size=8", "general.device=[0,1]", "data.training_resolution=256"])
    assert c2.opt.batch_size == 8 and c2.general.multiple_gpu is True and c2.data.training_resolution == 256
    with pytest.raises(FileNotFoundError):
        compose("does_not_exist")


def test_raster_layout_indexing():
    from unipre3d_b200.rasterizer import RasterLayout
    lay = RasterLayout([5, 0, 7], [2, 1, 3], torch.device("cpu"))
    assert lay.n_sets == 3 and lay.n_views == 6 and lay.n_gaussians == 12 and lay.max_set_size == 7
    assert lay.set_offsets.tolist() == [0, 5, 5, 12]
    assert lay.set_view_start.tolist() == [0, 2, 3, 6]
    assert lay.view_set.tolist() == [0, 0, 1, 2, 2, 2]
    assert lay.view_rec_start.tolist() == [0, 5, 10, 10, 17, 24, 31] and lay.n_records == 31
    assert RasterLayout.get([4], [1], "cpu") is RasterLayout.get([4], [1], "cpu")


def test_synthetic_batch_matches_loader_contract():
    from unipre3d_b200 import synthetic
    from unipre3d_b200.config import compose
    cfg = compose(overrides=["data.training_resolution=32"])
    b = synthetic.make_batch(cfg, 3, 256, seed=0)
    V = 5
    assert b["gt_images"].shape == (3, V, 3, 32, 32) and b["point_cloud"]["pos"].shape == (3, 256, 3)
    for k in ("world_view_transforms", "view_to_world_transforms", "full_proj_transforms"):
        assert b[k].shape == (3, V, 4, 4)
    pos = b["point_cloud"]["pos"]
    assert torch.allclose(pos.mean(1), torch.zeros(3, 3), atol=1e-2) and float(pos.norm(dim=-1).max()) <= 0.5 + 1e-6
    wv, vw, cc = b["world_view_transforms"][0, 0], b["view_to_world_transforms"][0, 0], b["camera_centers"][0, 0]
    assert torch.allclose(wv @ vw, torch.eye(4), atol=1e-5)                       # row-vector convention inverses
    assert torch.allclose(cc, vw[3, :3], atol=1e-5) and abs(float(cc.norm()) - 1.75) < 1e-4
    p = torch.tensor([0.0, 0.0, 0.0, 1.0]) @ b["full_proj_transforms"][0, 0]      # origin projects to image centre
    assert abs(float(p[0] / p[3])) < 1e-5 and abs(float(p[1] / p[3])) < 1e-5
    gt = b["gt_images"]
    bgfrac = float((gt.sum(2) == 0).float().mean())
    assert 0.2 < bgfrac < 0.95


def test_ema_schedule_and_shard_batch():
    from unipre3d_b200.trainer import EMA, shard_batch
    lin = torch.nn.Linear(2, 2)
    ema = EMA(lin, beta=0.9999, update_every=10, update_after_step=100)
    w0 = lin.weight.detach().clone()
    ema.update()
    assert torch.equal(ema.ema_model.weight, w0)
    with torch.no_grad():
        lin.weight.add_(1.0)
    for _ in range(99):
        ema.update()                       # steps 1..99: copies at multiples of 10 (still <= update_after_step)
    assert torch.equal(ema.ema_model.weight, lin.weight)
    ema.step = 200
    assert ema.current_decay() == pytest.approx(1 - (1 + 99) ** (-2 / 3))
    with torch.no_grad():
        lin.weight.add_(1.0)
    ema.update()
    d = ema.current_decay.__func__(ema) if False else None
    assert not torch.equal(ema.ema_model.weight, lin.weight)
    data = {"a": torch.arange(8).view(8, 1), "point_cloud": {"pos": torch.arange(16).view(8, 2)}}
    s = shard_batch(data, 1, 4)
    assert s["a"].flatten().tolist() == [2, 3] and s["point_cloud"]["pos"].shape == (2, 2)
    with pytest.raises(ValueError):
        shard_batch(data, 0, 3)


def test_state_dict_keys_follow_reference_naming():
    from unipre3d_b200.config import compose
    from unipre3d_b200.gaussian_predictor import GaussianSplatPredictor
    m = GaussianSplatPredictor(compose())
    keys = set(m.state_dict().keys())
    for k in ["point_network.encoder.encoder.first_conv.0.weight", "point_network.encoder.encoder.second_conv.3.bias",
              "point_network.encoder.reduce_dim.weight", "point_network.encoder.cls_token", "point_network.encoder.cls_pos",
              "point_network.encoder.pos_embed.0.weight", "point_network.encoder.blocks.blocks.15.attn.qkv.weight",
              "point_network.encoder.blocks.blocks.0.mlp.fc2.bias", "point_network.encoder.norm.weight",
              "point_network.final.0.weight", "point_network.final.2.bias", "fusion_mlps.0.weight", "image_conv.0.weight",
              "image_conv.1.bias", "sh_to_v_transform", "v_to_sh_transform"]:
        assert k in keys, k
    assert "point_network.encoder.blocks.blocks.0.attn.qkv.bias" not in keys        # qkv_bias=False
    n_enc = sum(p.numel() for p in m.point_network.encoder.parameters())
    assert n_enc == 29066880                                                        # SURVEY.md §2.2 (measured on the reference)


def test_cpu_port_step_runs_and_learns():
    from oracle.cpu_step import CpuStepper
    from unipre3d_b200 import synthetic
    from unipre3d_b200.config import compose
    cfg = compose(overrides=["data.training_resolution=32", "opt.batch_size=2", "opt.imgs_per_obj=2"])
    torch.manual_seed(0)
    st = CpuStepper(cfg)
    for m in st.model.modules():
        if m.__class__.__name__ == "DropPath":
            m.drop_prob = 0.0
    d = synthetic.make_batch(cfg, 2, 512, seed=0)
    losses = [st.step(d) for _ in range(4)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]


def test_uint8_batches_and_in_place_target_layout():
    """Host-side pieces of the end-to-end path: 8-bit synthetic batches (a quarter of the H2D bytes, exact background
    value) and the layout detection that lets the fused loss read gt_images[:, input_images:] without a copy."""
    import torch
    from unipre3d_b200 import synthetic
    from unipre3d_b200.config import compose
    from unipre3d_b200.loss import _gt_layout
    cfg = compose(overrides=["data.training_resolution=32", "opt.batch_size=2"])
    b8 = synthetic.make_batch(cfg, 2, 128, seed=3, image_dtype="uint8")
    bf = synthetic.make_batch(cfg, 2, 128, seed=3)
    assert b8["gt_images"].dtype == torch.uint8 and bf["gt_images"].dtype == torch.float32
    assert b8["gt_images"].shape == bf["gt_images"].shape
    assert int(b8["gt_images"].min()) == 0                      # black background stays exactly 0
    assert synthetic.batch_nbytes(b8) < synthetic.batch_nbytes(bf)
    gt = b8["gt_images"]
    ni = int(cfg.data.input_images)
    view = gt[:, ni:]
    base, g, vpo, ostride = _gt_layout(view)
    assert g.data_ptr() == view.data_ptr() and vpo == gt.shape[1] - ni and ostride == gt.stride(0)   # in place
    flat = view.reshape(-1, *view.shape[2:])                    # the (n,3,H,W) form: contiguous copy, one "object"
    _, g2, vpo2, ostride2 = _gt_layout(flat)
    assert vpo2 == flat.shape[0] and ostride2 == 0
    weird = gt.permute(0, 1, 2, 4, 3)[:, ni:]                   # not a view-axis slice of a contiguous tensor -> copied
    _, g3, vpo3, ostride3 = _gt_layout(weird)
    assert g3.is_contiguous() and ostride3 == weird.shape[1] * weird.shape[2] * weird.shape[3] * weird.shape[4]


def test_fused_stack_parameter_order_and_support_predicate():
    """fused_encoder.stack_parameters (the order gradients come back in, and the order GradSync packs them) covers
    every Block parameter exactly once; supports() accepts the reference configuration and rejects what the kernels do
    not implement."""
    import torch
    from unipre3d_b200 import fused_encoder
    from unipre3d_b200.backbone import TransformerEncoder
    enc = TransformerEncoder(embed_dim=384, depth=3, num_heads=6, drop_path_rate=[0.0, 0.05, 0.1])
    ps = fused_encoder.stack_parameters(enc.blocks)
    assert len(ps) == 3 * fused_encoder.PARAMS_PER_BLOCK == len(list(enc.parameters()))
    assert {id(p) for p in ps} == {id(p) for p in enc.parameters()}
    assert fused_encoder.supports(enc.blocks)
    assert not fused_encoder.supports(TransformerEncoder(embed_dim=48, depth=1, num_heads=6).blocks)       # width % 128
    assert not fused_encoder.supports(TransformerEncoder(embed_dim=128, depth=1, num_heads=2, qkv_bias=True).blocks)
    assert not fused_encoder.supports(TransformerEncoder(embed_dim=128, depth=1, num_heads=2, drop_rate=0.1).blocks)


def test_lazy_visibility_filter_is_evaluated_on_demand():
    import torch
    from unipre3d_b200.gaussian_renderer import _LazyVisibility
    radii = [torch.tensor([[0, 3, 0], [1, 0, 2]]), torch.tensor([[5, 0, 0], [0, 0, 0]])]
    vis = _LazyVisibility(radii)
    assert vis._vis is None and len(vis) == 2
    assert torch.equal(vis[1], radii[1] > 0) and vis._vis is not None
    assert [v.dtype for v in vis] == [torch.bool, torch.bool]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port of the step on the host cores) prints ONE JSON line with the keys the
    driver reads; no GPU needed."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "views/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["n_gpus"] == 1 and d["steps"] >= 1 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    # "reference": the reference's own backbone modules staged under oracle/_ref/pyref; "port" when they are absent
    staged = os.path.exists(os.path.join(root, "oracle", "_ref", "pyref", "bench_batch.npz"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    # the reference arm must not map the product into its process
    probe = subprocess.run([sys.executable, "-c", "import sys; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','1'];"
                            "import runpy; runpy.run_path('bench.py', run_name='__main__');"
                            "bad=[m for m in sys.modules if m.startswith('unipre3d_b200')]; print('PRODUCT_MODULES', bad)"],
                           capture_output=True, text=True, timeout=600, cwd=root)
    if staged:
        assert "PRODUCT_MODULES []" in probe.stdout, probe.stdout[-500:] + probe.stderr[-500:]
    assert d["e2e"] == {"value": d["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_sd_vae_image_branch_architecture():
    """image_predictor.AutoencoderKL restates the sd-vae-ft-mse architecture the reference loads through diffusers
    (model/image_predictor.py:26-31): parameter count of the public checkpoint, diffusers' state-dict names, the four
    hooked decoder blocks with `decoder_block_3` = 128 channels at image resolution; frozen and deterministic."""
    import torch
    from unipre3d_b200.image_predictor import AutoencoderKL, ImageFeaturePredictor, count_parameters
    torch.manual_seed(0)
    ip = ImageFeaturePredictor(None, [128])
    assert count_parameters(ip.encoder) == 83_653_863
    sd = ip.encoder.state_dict()
    assert len(sd) == 248
    for k in ("encoder.down_blocks.1.resnets.0.conv_shortcut.weight", "encoder.down_blocks.2.downsamplers.0.conv.weight",
              "encoder.mid_block.attentions.0.to_q.weight", "encoder.mid_block.attentions.0.group_norm.bias",
              "decoder.up_blocks.2.upsamplers.0.conv.bias", "decoder.up_blocks.3.resnets.2.conv2.weight",
              "decoder.mid_block.attentions.0.to_out.0.weight", "quant_conv.weight", "post_quant_conv.bias"):
        assert k in sd, k
    assert "encoder.down_blocks.3.downsamplers.0.conv.weight" not in sd and "decoder.up_blocks.3.upsamplers.0.conv.weight" not in sd
    assert not any(p.requires_grad for p in ip.parameters())
    ip.train()
    assert not ip.encoder.training                                # the reference keeps the VAE in eval mode
    x = torch.rand(1, 3, 32, 32)
    a, b = ip(x), ip(x)
    assert {k: tuple(v.shape) for k, v in a.items()} == {"decoder_block_0": (1, 512, 8, 8), "decoder_block_1": (1, 512, 16, 16),
                                                         "decoder_block_2": (1, 256, 32, 32), "decoder_block_3": (1, 128, 32, 32)}
    assert all(torch.equal(a[k], b[k]) for k in a) and all(torch.isfinite(v).all() for v in a.values())
    # stopping after the last hooked block gives the same features as decoding to RGB
    feats = {}
    rgb = ip.encoder(x, collect=feats)
    assert rgb.shape == (1, 3, 32, 32) and torch.equal(feats["decoder_block_3"], a["decoder_block_3"])
    # a checkpoint in diffusers' layout loads strictly
    m2 = AutoencoderKL()
    m2.load_state_dict(sd, strict=True)


def test_predictor_selects_image_branch():
    import torch
    from unipre3d_b200.config import compose
    from unipre3d_b200.fusion import LazyImageFeatures
    from unipre3d_b200.gaussian_predictor import FrozenImageStem, GaussianSplatPredictor
    from unipre3d_b200.image_predictor import ImageFeaturePredictor
    assert isinstance(GaussianSplatPredictor(compose()).image_network, FrozenImageStem)
    m = GaussianSplatPredictor(compose(overrides=["model.image_branch=sdvae", "data.training_resolution=32"]))
    assert isinstance(m.image_network, ImageFeaturePredictor)
    trainable = [n for n, p in m.named_parameters() if p.requires_grad]
    assert trainable and not any(n.startswith("image_network.") for n in trainable)
    # the VAE's dense features feed the sampled image_conv evaluation
    feats = m.image_network(torch.rand(2, 3, 32, 32))["decoder_block_3"]
    lazy = LazyImageFeatures(feats, m.image_conv)
    bidx = torch.arange(2).unsqueeze(1).expand(2, 5)
    ix, iy = torch.randint(0, 32, (2, 5)), torch.randint(0, 32, (2, 5))
    assert torch.allclose(lazy.sample(bidx, ix, iy), lazy.dense()[bidx, :, ix, iy], atol=1e-4, rtol=1e-4)
    try:
        GaussianSplatPredictor(compose(overrides=["model.image_branch=nope"]))
        assert False
    except ValueError:
        pass


def test_checkpoint_round_trip_keeps_reference_keys(tmp_path):
    """ModelManager.save_checkpoint / load_checkpoint: the reference's checkpoint dict (train_network.py:200-210)."""
    import torch
    from unipre3d_b200.config import compose
    from unipre3d_b200.trainer import ModelManager
    cfg = compose(overrides=["data.training_resolution=32", "opt.batch_size=2"])
    torch.manual_seed(0)
    a = ModelManager(cfg, torch.device("cpu"))
    a.save_latest_checkpoint(123, 21.5, str(tmp_path))
    ck = torch.load(tmp_path / "model_latest.pth")
    ref_keys = {"iteration", "optimizer_state_dict", "model_state_dict", "best_PSNR"}      # train_network.py:200-210
    assert ref_keys <= set(ck) and set(ck) - ref_keys <= {"online_model_state_dict", "ema_state"}
    assert any(k.startswith("point_network.encoder.blocks.blocks.0.attn.qkv") for k in ck["model_state_dict"])
    torch.manual_seed(1)
    b = ModelManager(cfg, torch.device("cpu"))
    assert not torch.equal(b.model.fusion_mlps[0].weight, a.model.fusion_mlps[0].weight)
    meta = b.load_checkpoint(str(tmp_path / "model_latest.pth"))
    assert meta == {"iteration": 123, "best_PSNR": 21.5}
    for (ka, pa), (kb, pb) in zip(a.model.state_dict().items(), b.model.state_dict().items()):
        assert ka == kb and torch.equal(pa, pb), ka
    assert b._sched_step == 123                                       # StepLR phase follows the saved iteration
    if b.ema is not None:
        assert (b.ema.step, b.ema.initted) == (a.ema.step, a.ema.initted)
    # a checkpoint as the REFERENCE writes it: DDP "module." prefix, frozen AutoencoderKL under image_network.*, no extras
    ref_ck = {"iteration": 7, "best_PSNR": 1.0, "optimizer_state_dict": ck["optimizer_state_dict"],
              "model_state_dict": {**{"module." + k: v for k, v in ck["model_state_dict"].items()},
                                   "module.image_network.encoder.conv_in.weight": torch.zeros(2, 2)}}
    torch.save(ref_ck, tmp_path / "ref_style.pth")
    c = ModelManager(cfg, torch.device("cpu"))
    assert c.load_checkpoint(str(tmp_path / "ref_style.pth")) == {"iteration": 7, "best_PSNR": 1.0}
    assert c._sched_step == 7 and (c.ema is None or (c.ema.step == 7 and c.ema.initted))


def test_packed_batch_is_one_buffer_with_aligned_views():
    """trainer.PackedBatch: the flat pinned staging form of a batch dict (one H2D + one D2D per step)."""
    from unipre3d_b200 import synthetic
    from unipre3d_b200.config import compose
    from unipre3d_b200.trainer import PackedBatch
    cfg = compose(overrides=["data.training_resolution=32", "opt.batch_size=2"])
    b = synthetic.make_batch(cfg, 2, 128, seed=0, image_dtype="uint8")
    p = PackedBatch(b, pin=False)
    assert p.flat.dtype == torch.uint8 and p.flat.numel() == p.nbytes
    assert all(off % 256 == 0 for _, off, _, _, _ in p.layout)
    assert p.nbytes >= synthetic.batch_nbytes(b) and p.nbytes < synthetic.batch_nbytes(b) + 256 * len(p.layout)
    assert torch.equal(p.data["gt_images"], b["gt_images"]) and p.data["gt_images"].dtype == torch.uint8
    assert torch.equal(p.data["point_cloud"]["pos"], b["point_cloud"]["pos"])
    base = p.flat.data_ptr()
    assert all(base <= t.data_ptr() < base + p.nbytes for t in (p.data["gt_images"], p.data["camera_centers"]))
    mirror = p.views(p.flat.clone())                       # what the trainer builds on the device
    for k in ("world_view_transforms", "full_proj_transforms", "view_to_world_transforms", "camera_centers"):
        assert torch.equal(mirror[k], b[k])
    assert PackedBatch(b, pin=False).signature() == p.signature()


def test_scene_level_output_processing_matches_reference_loop():
    """GaussianSplatPredictor._process_network_output_scene == the mask loop of model/gaussian_predictor.py:331-364."""
    from types import SimpleNamespace as NS
    from unipre3d_b200.gaussian_predictor import GaussianSplatPredictor
    m = GaussianSplatPredictor.__new__(GaussianSplatPredictor)
    torch.nn.Module.__init__(m)
    m.cfg = NS(model=NS(max_sh_degree=1, isotropic=False, offset_scale=0.2))
    g = torch.Generator().manual_seed(0)
    n = 50
    raw = torch.randn(n, 23, generator=g)
    coord = torch.randn(n, 3, generator=g)
    batch = torch.sort(torch.randint(0, 3, (n,), generator=g))[0]
    indices = torch.cat([batch[:, None], torch.randint(0, 100, (n, 3), generator=g)], 1).int()
    out = m._process_network_output_scene(raw.split([3, 1, 3, 4, 3, 9], dim=1), coord, indices)
    xyz_raw, opacity, scaling, rotation, dc, rest = raw.split([3, 1, 3, 4, 3, 9], dim=1)
    pos = torch.tanh(xyz_raw) * 0.2 + coord
    for b in range(3):
        mask = indices[:, 0] == b
        assert torch.allclose(out["xyz"][b], pos[mask])
        assert torch.allclose(out["opacity"][b], torch.sigmoid(opacity[mask]))
        assert torch.allclose(out["scaling"][b], torch.exp(torch.clamp(scaling[mask], -1, 20)))
        assert torch.allclose(out["rotation"][b], torch.nn.functional.normalize(rotation[mask], dim=-1, eps=1e-6))
        assert torch.equal(out["features_dc"][b], dc[mask].unsqueeze(1))
        # _process_sh_features: (n, 9) -> reshape(n, 9, 1).permute(0, 2, 1).reshape(n, -1, 3)
        fr = rest[mask]
        ref = fr.reshape(fr.shape[0], fr.shape[1], -1).permute(0, 2, 1).reshape(fr.shape[0], -1, 3)
        assert torch.equal(out["features_rest"][b], ref) and ref.shape[1:] == (3, 3)


def test_scene_configs_compose_and_synthetic_scene_batch():
    from unipre3d_b200 import synthetic
    from unipre3d_b200.config import compose
    for name, bb in (("sparseunet_pretraining", "sparseunet"), ("ptv3_pretraining", "ptv3")):
        cfg = compose(name)
        assert cfg.model.backbone_type == bb and cfg.opt.level == "scene" and cfg.data.training_width == 160
        assert cfg.data.znear == 0.2 and cfg.data.zfar == 10 and cfg.data.white_background is True and cfg.opt.loss == "l2"
    cfg = compose("sparseunet_pretraining", overrides=["data.input_images=2", "opt.imgs_per_obj=2"])
    b = synthetic.make_scene_batch(cfg, 2, 5000, seed=1)
    pc = b["point_cloud"]
    assert pc["offset"].tolist()[-1] == pc["coord"].shape[0] == pc["grid_coord"].shape[0] == pc["feat"].shape[0]
    assert pc["feat"].shape[1] == 6 and int(pc["grid_coord"].min()) == 0
    assert b["gt_images"].shape == (2, 4, 3, 120, 160)
    key = (pc["grid_coord"][: pc["offset"][0]].long() * torch.tensor([1 << 40, 1 << 20, 1])).sum(1)
    assert key.unique().numel() == key.numel(), "one point per voxel"
    wv, vw = b["world_view_transforms"][1, 2], b["view_to_world_transforms"][1, 2]
    assert torch.allclose(wv @ vw, torch.eye(4), atol=1e-5)


def test_stack_grad_order_is_the_backward_slab_layout():
    """fused_encoder.stack_grad_order: the flat layout the data-parallel backward writes its gradients into -- four weight
    slabs stacked over the blocks, then per block the column-sum gradients (n1w n1b bproj n2w n2b b2 | b1)."""
    from unipre3d_b200 import fused_encoder as fe
    depth, P = 3, fe.PARAMS_PER_BLOCK
    order = fe.stack_grad_order(depth)
    assert sorted(order) == list(range(depth * P))
    assert order[:depth] == [2, 2 + P, 2 + 2 * P] and order[3 * depth:4 * depth] == [9, 9 + P, 9 + 2 * P]
    assert order[4 * depth:4 * depth + 7] == [0, 1, 4, 5, 6, 10, 8]


def test_clock_sampler_summary_and_disabled_rank():
    """bench.ClockSampler: parsing of the sampled lines (median SM clock, max clock, throttle reasons) and the disabled
    sampler of ranks > 0 (no thread, no child process, empty summary)."""
    import bench
    c = bench.ClockSampler(0, enabled=False)
    with c:
        pass
    assert c.proc is None and c.t is None
    assert c.summary() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": None}
    c.lines = ["0, 1965, 1965, , 0x0, Not Active, Not Active, Not Active, Not Active",
               "0, 1950, 1965, , 0x4, Not Active, Not Active, Not Active, Active",
               "0, 1935, 1965, , 0x0, Not Active, Not Active, Not Active, Not Active",
               "garbage"]
    c.source = "nvml"
    out = c.summary()
    assert out["sm_mhz"] == 1950.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["reasons"] == ["sw_power_cap"] and out["source"] == "nvml"
