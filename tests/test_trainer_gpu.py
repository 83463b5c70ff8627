"""The batched B200 step == the reference-shaped step (Python double loop of render_predicted + torch loss),
and the CUDA-graph step == the eager step."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg(res=64, bs=2):
    from unipre3d_b200.config import compose
    return compose(overrides=[f"data.training_resolution={res}", f"opt.batch_size={bs}", "opt.ema.use=false"])


def test_batched_step_equals_reference_shaped_loop():
    from unipre3d_b200 import synthetic
    from unipre3d_b200.gaussian_renderer import render_predicted
    from unipre3d_b200.loss import focal_l2_loss_torch
    from unipre3d_b200.trainer import Trainer, _to_device, prepare_model_inputs
    cfg = _cfg()
    tr = Trainer(cfg, use_cuda_graph=False)
    model = tr.model_manager.model
    data = _to_device(synthetic.make_batch(cfg, 2, 1024, seed=1), tr.device)
    # (a) batched path
    torch.manual_seed(7)
    loss_a = tr._forward_backward(data)
    grads_a = [p.grad.clone() for p in tr.params]
    for p in tr.params:
        p.grad = None
    # (b) reference-shaped loop: train_network.py:418-442 + loss_utils.focal_l2_loss
    torch.manual_seed(7)
    model.train()
    splats = model(**prepare_model_inputs(data, cfg, 2, tr.device))
    bg = tr.validation_manager.background
    imgs, gts = [], []
    for b in range(data["gt_images"].shape[0]):
        one = {k: v[b].contiguous() for k, v in splats.items() if len(v.shape) > 1}
        for r in range(cfg.data.input_images, data["gt_images"].shape[1]):
            imgs.append(render_predicted(one, data["world_view_transforms"][b, r], data["full_proj_transforms"][b, r],
                                         data["camera_centers"][b, r], bg, cfg, focals_pixels=None)["render"])
            gts.append(data["gt_images"][b, r])
    loss_b = focal_l2_loss_torch(torch.stack(imgs), torch.stack(gts), bg, cfg.opt.non_bg_color_loss_rate,
                                 cfg.opt.bg_color_loss_rate)
    loss_b.backward()
    assert abs(float(loss_a) - float(loss_b)) <= 1e-6 * max(1.0, abs(float(loss_b)))
    # conv biases feeding a BatchNorm have a mathematically zero gradient (pure rounding noise), hence the
    # global-scale floor in the denominator
    gmax = max(float(p.grad.abs().max()) for p in tr.params)
    worst = 0.0
    for ga, p in zip(grads_a, tr.params):
        scale = float(p.grad.abs().max()) + 1e-3 * gmax
        worst = max(worst, float((ga - p.grad).abs().max()) / scale)
    assert worst <= 2e-3, f"gradients of the batched step deviate from the per-view loop: rel {worst}"


def test_output_dict_shapes_and_quirks():
    from unipre3d_b200 import synthetic
    from unipre3d_b200.trainer import Trainer, _to_device, prepare_model_inputs
    cfg = _cfg()
    tr = Trainer(cfg)
    data = _to_device(synthetic.make_batch(cfg, 2, 2048, seed=2), tr.device)
    tr.model_manager.model.eval()
    with torch.no_grad():
        out = tr.model_manager.model(**prepare_model_inputs(data, cfg, 2, tr.device))
    assert out["xyz"].shape == (2, 128, 3) and out["opacity"].shape == (2, 128, 1)
    assert out["scaling"].shape == (2, 128, 3) and out["rotation"].shape == (2, 128, 4)
    assert out["features_dc"].shape == (2, 128, 1, 3) and out["features_rest"].shape == (2, 128, 3, 3)
    assert float(out["scaling"].min()) >= np.exp(-1) - 1e-6                    # exp(clamp(x,-1,20))
    # rotation is normalised over the POINT axis (gaussian_predictor.py:254,318)
    col_norm = out["rotation"].pow(2).sum(1).sqrt()
    assert torch.allclose(col_norm, torch.ones_like(col_norm), atol=1e-4)


def test_cuda_graph_step_matches_eager_and_learns():
    from unipre3d_b200 import synthetic
    from unipre3d_b200.trainer import Trainer
    cfg = _cfg(res=64, bs=2)
    data = synthetic.make_batch(cfg, 2, 1024, seed=3, pin=True)
    losses = {}
    for mode in (False, True):
        torch.manual_seed(0)
        tr = Trainer(cfg, use_cuda_graph=mode)
        for blk in tr.model_manager.model.modules():          # make the two runs deterministic: no DropPath
            if blk.__class__.__name__ == "DropPath":
                blk.drop_prob = 0.0
        losses[mode] = [tr.train_iteration(data) for _ in range(6)]
    # graph mode's warm-up steps are rolled back (parameters, BatchNorm buffers, moments, step counter), so the two
    # trajectories coincide step for step (float atomics in the raster backward leave ~1e-6 relative noise)
    assert np.isfinite(losses[True]).all() and np.isfinite(losses[False]).all()
    assert losses[False][-1] < losses[False][0], "eager loss does not decrease"
    for a, b in zip(losses[True], losses[False]):
        assert abs(a - b) <= 2e-3 * abs(b) + 1e-5, (losses[True], losses[False])
    assert abs(losses[True][0] - losses[False][0]) <= 1e-5 * abs(losses[False][0]) + 1e-7


@pytest.mark.parametrize("packed", [True, False])
def test_lookahead_grouping_equals_inline_grouping(packed):
    """Graph mode with `prefetch`: the FPS / ball-query grouping of batch i+1 is computed on a side stream next to step i
    and consumed by step i+1.  Same losses as eager steps that group in line, on a DIFFERENT batch every step (a stale or
    mismatched grouping would change the loss at once); then `replay_resident` (bench.py's `value` loop) on the last batch."""
    from unipre3d_b200 import synthetic
    from unipre3d_b200.trainer import Trainer
    cfg = _cfg(res=64, bs=2)
    batches = [synthetic.make_batch(cfg, 2, 1024, seed=20 + i, pin=True) for i in range(5)]
    losses = {}
    for mode in (False, True):
        torch.manual_seed(0)
        tr = Trainer(cfg, use_cuda_graph=mode)
        for blk in tr.model_manager.model.modules():
            if blk.__class__.__name__ == "DropPath":
                blk.drop_prob = 0.0
        if mode:
            bs = [tr.pack_batch(b) for b in batches] if packed else batches
            losses[mode] = [tr.train_iteration(bs[i], prefetch=bs[i + 1] if i + 1 < len(bs) else None)
                            for i in range(len(bs))]
            assert tr._la is not None and tr.lookahead_hits == len(bs) - 1
            assert tr.model_manager.model.point_network.encoder.group_divider.lookahead is None   # capture-only hook
            # resident replays: same batch again and again -> the loss keeps falling, grouping comes from the side stream
            hits = tr.lookahead_hits
            for _ in range(3):
                tr.replay_resident()
            torch.cuda.synchronize()
            assert tr.lookahead_hits == hits + 2 and float(tr._loss_buf) < losses[mode][-1]
        else:
            losses[mode] = [tr.train_iteration(b) for b in batches]
    assert abs(losses[True][0] - losses[False][0]) <= 1e-5 * abs(losses[False][0]) + 1e-7
    for a, b in zip(losses[True], losses[False]):
        assert abs(a - b) <= 2e-3 * abs(b) + 1e-5, (losses[True], losses[False])


def test_lazy_image_features_equal_dense_dataflow():
    """Sampling Conv1x1(GroupNorm(features)) at the projected pixels == indexing the dense tensor
    (the reference's dataflow, gaussian_predictor.py:139 + feat_fusion.py:121-131): values and parameter gradients."""
    from unipre3d_b200.fusion import LazyImageFeatures
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False          # the dense path is a cuDNN conv: compare in true fp32
    n, Cin, Cout, R, N = 3, 128, 384, 32, 50
    conv = torch.nn.Sequential(torch.nn.GroupNorm(32, Cin, eps=1e-6), torch.nn.Conv2d(Cin, Cout, 1)).cuda()
    with torch.no_grad():
        conv[0].weight.uniform_(0.5, 1.5); conv[0].bias.normal_()
    x = torch.randn(n, Cin, R, R, device="cuda")
    bidx = torch.arange(n, device="cuda").unsqueeze(1).expand(n, N)
    ix, iy = torch.randint(0, R, (n, N), device="cuda"), torch.randint(0, R, (n, N), device="cuda")
    w = torch.randn(n, N, Cout, device="cuda")
    a = LazyImageFeatures(x, conv).sample(bidx, ix, iy)
    ga = torch.autograd.grad((a * w).sum(), list(conv.parameters()))
    b = conv(x)[bidx, :, ix, iy]
    gb = torch.autograd.grad((b * w).sum(), list(conv.parameters()))
    assert torch.allclose(a, b, atol=2e-5, rtol=1e-5)
    for u, v in zip(ga, gb):
        assert torch.allclose(u, v, atol=1e-4 * float(v.abs().max()) + 1e-6, rtol=1e-4)


def test_shadow_bf16_weights_equal_autocast():
    """Persistent bf16 shadow weights (mixed_precision.py) == torch.autocast numerics, step by step."""
    from unipre3d_b200 import synthetic
    from unipre3d_b200.trainer import Trainer
    cfg = _cfg(res=64, bs=2)
    data = synthetic.make_batch(cfg, 2, 1024, seed=5, pin=True)
    out = {}
    for shadow in (False, True):
        torch.manual_seed(0)
        tr = Trainer(cfg, use_cuda_graph=False, autocast_dtype=torch.bfloat16, shadow_weights=shadow)
        for blk in tr.model_manager.model.modules():
            if blk.__class__.__name__ == "DropPath":
                blk.drop_prob = 0.0
        losses = [tr.train_iteration(data) for _ in range(3)]
        out[shadow] = (losses, [p.detach().clone() for p in tr.params])
    la, lb = out[False][0], out[True][0]
    assert abs(la[0] - lb[0]) <= 1e-3 * abs(la[0])              # same forward (bf16 rounding of the same operands)
    assert abs(la[-1] - lb[-1]) <= 2e-2 * abs(la[-1])
    assert lb[-1] < lb[0]


def test_uint8_host_images_equal_float_images():
    """8-bit host images divided by 255 on the device give the same step as the float images the reference's loader
    produces on the host (a quarter of the host->device bytes)."""
    from unipre3d_b200 import synthetic
    from unipre3d_b200.trainer import Trainer, _to_device
    cfg = _cfg(res=64, bs=2)
    b8 = synthetic.make_batch(cfg, 2, 1024, seed=9, image_dtype="uint8")
    assert b8["gt_images"].dtype == torch.uint8 and synthetic.batch_nbytes(b8) < 0.3 * synthetic.batch_nbytes(
        synthetic.make_batch(cfg, 2, 1024, seed=9))
    bf = dict(b8)
    bf["gt_images"] = b8["gt_images"].float() / 255.0
    losses = []
    for batch in (b8, bf):
        torch.manual_seed(0)
        tr = Trainer(cfg, use_cuda_graph=False)
        torch.manual_seed(1)
        losses.append(float(tr._forward_backward(_to_device(batch, tr.device))))
    assert abs(losses[0] - losses[1]) <= 1e-6 * abs(losses[1])
