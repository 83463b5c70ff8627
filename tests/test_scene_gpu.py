"""Scene-level step (BASELINE configs[4] shape, reduced): SpUNet on the sparse-convolution engine -> per-scene Gaussian
lists -> one batched render of all (scene, view) pairs with ragged set sizes -> L2 -> backward -> fused clip + AdamW."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_sparseunet_scene_step_runs_and_learns():
    from unipre3d_b200 import synthetic
    from unipre3d_b200.config import compose
    from unipre3d_b200.trainer import Trainer, _to_device, prepare_model_inputs
    cfg = compose("sparseunet_pretraining", overrides=["opt.use_fusion=false", "data.input_images=2", "opt.imgs_per_obj=2",
                                                        "opt.batch_size=2", "opt.ema.use=false"])
    tr = Trainer(cfg, use_cuda_graph=False)
    data = synthetic.make_scene_batch(cfg, 2, 6000, seed=0)
    dev = _to_device(data, tr.device)
    tr.model_manager.model.train()
    out = tr.model_manager.model(**prepare_model_inputs(dev, cfg, 2, tr.device))
    n = data["point_cloud"]["offset"].tolist()
    sizes = [n[0], n[1] - n[0]]
    assert isinstance(out["xyz"], list) and [int(x.shape[0]) for x in out["xyz"]] == sizes
    assert out["rotation"][0].shape == (sizes[0], 4) and out["features_rest"][1].shape == (sizes[1], 3, 3)
    assert torch.allclose(out["rotation"][0].norm(dim=-1), torch.ones(sizes[0], device=tr.device), atol=1e-4)
    # Gaussians sit within offset_scale of their source points (tanh * 0.2)
    assert float((out["xyz"][0] - dev["point_cloud"]["coord"][: sizes[0]]).abs().max()) <= 0.2 + 1e-5
    losses = [tr.train_iteration(data) for _ in range(5)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0], losses


def test_ptv3_scene_step_runs_and_learns():
    from unipre3d_b200 import synthetic
    from unipre3d_b200.config import compose
    from unipre3d_b200.trainer import Trainer
    cfg = compose("ptv3_pretraining", overrides=["opt.use_fusion=false", "data.input_images=2", "opt.imgs_per_obj=2",
                                                  "opt.batch_size=2", "opt.ema.use=false"])
    tr = Trainer(cfg, use_cuda_graph=False)
    for m in tr.model_manager.model.modules():
        if m.__class__.__name__ == "DropPath":
            m.drop_prob = 0.0
    data = synthetic.make_scene_batch(cfg, 2, 5000, seed=3)
    losses = [tr.train_iteration(data) for _ in range(5)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0], losses
