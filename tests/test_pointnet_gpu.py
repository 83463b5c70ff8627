"""Fused mini-PointNet (csrc/pointnet.cu + fused_pointnet.py) against the module path of backbone.Encoder -- itself
pinned against the reference's Encoder by tests/golden/transformer_encoder.npz -- on the same seeded inputs."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _encoder(C=384, seed=0):
    from unipre3d_b200.backbone import Encoder
    torch.manual_seed(seed)
    enc = Encoder(C).to(DEV)
    with torch.no_grad():
        for m in enc.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
    return enc


# K = 1: the max-pools are trivial, so both paths are the same smooth function and fp32 results must agree tightly.
# K > 1: an arg-max can flip between the two fp32 evaluation orders on a near-tie (the module path against its own
# fp64 evaluation shows the same ~2e-3 effect), so gradients get a looser max-norm bound plus a cosine check.
@pytest.mark.parametrize("B,G,K,tol", [(4, 64, 1, 2e-4), (2, 16, 8, 2e-2), (3, 37, 32, 2e-2), (8, 128, 32, 2e-2)])
def test_fused_mini_pointnet_equals_module_path_fp32(B, G, K, tol):
    from unipre3d_b200 import _lib, fused_pointnet
    enc_a = _encoder()
    enc_b = copy.deepcopy(enc_a)
    enc_b.force_module_path = True
    assert fused_pointnet.supports(enc_a)
    torch.manual_seed(1)
    nb = torch.randn(B, 3, G, K, device=DEV) * 0.05
    w = torch.randn(B, G, 384, device=DEV)
    enc_a.train(); enc_b.train()
    before = _lib.launch_count
    ta = enc_a.forward_grouped(nb)
    assert _lib.launch_count > before, "the fused CUDA path did not run"
    tb = enc_b.forward_grouped(nb)
    assert ta.shape == tb.shape == (B, G, 384)
    assert torch.allclose(ta, tb, atol=2e-4, rtol=1e-4), float((ta - tb).abs().max())
    (ta * w).sum().backward()
    (tb * w).sum().backward()
    gmax = max(float(p.grad.abs().max()) for p in enc_b.parameters())
    report = {}
    for (k, pa), (_, pb) in zip(enc_a.named_parameters(), enc_b.named_parameters()):
        if k in ("first_conv.0.bias", "first_conv.3.bias", "second_conv.0.bias"):
            # a bias that reaches a train-mode BatchNorm through linear maps only (conv1.bias -> BN1; conv2.bias -> max /
            # concat -> conv3 -> BN2; conv3.bias -> BN2) has a mathematically ZERO gradient: both paths hold rounding noise
            assert float(pa.grad.abs().max()) <= 1e-2 * gmax and float(pb.grad.abs().max()) <= 1e-2 * gmax, k
            continue
        scale = float(pb.grad.abs().max()) + 1e-4 * gmax
        err = float((pa.grad - pb.grad).abs().max()) / scale
        cos = float(torch.nn.functional.cosine_similarity(pa.grad.flatten(), pb.grad.flatten(), dim=0))
        report[k] = (err, cos)
    bad = {k: v for k, v in report.items() if v[0] > tol or v[1] < 0.9995}
    assert not bad, f"gradient mismatch (rel max err, cosine): {bad}  all: {report}"
    # running statistics (momentum 0.1, unbiased variance) and the batch counter
    for ma, mb in zip(enc_a.modules(), enc_b.modules()):
        if isinstance(ma, torch.nn.BatchNorm1d):
            assert torch.allclose(ma.running_mean, mb.running_mean, atol=1e-5, rtol=1e-4)
            assert torch.allclose(ma.running_var, mb.running_var, atol=1e-6, rtol=1e-4)
            assert int(ma.num_batches_tracked) == int(mb.num_batches_tracked) == 1


def test_fused_mini_pointnet_bf16_and_eval_dispatch():
    enc = _encoder()
    torch.manual_seed(2)
    nb = torch.randn(4, 3, 64, 32, device=DEV) * 0.05
    enc.train()
    ref = enc.forward_grouped(nb)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = enc.forward_grouped(nb)
    assert out.dtype == torch.bfloat16
    assert float((out.float() - ref).abs().max()) <= 0.06 * float(ref.abs().max())
    (out.float() ** 2).mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() and p.grad.dtype == torch.float32 for p in enc.parameters())
    # eval mode: the module path with running statistics
    enc.eval()
    a = enc.forward_grouped(nb)
    b = enc.forward(nb.permute(0, 2, 3, 1))
    assert torch.equal(a, b)


def test_group_max_first_occurrence_and_scatter():
    from unipre3d_b200 import _lib
    from unipre3d_b200._lib import check, ptr, stream_ptr
    Gt, K, C = 5, 32, 256
    torch.manual_seed(3)
    x = torch.randint(0, 4, (Gt * K, C), device=DEV).float()         # many ties
    out = torch.empty(Gt, C, device=DEV)
    arg = torch.empty(Gt, C, dtype=torch.int32, device=DEV)
    check(_lib.lib.up3d_group_max(0, Gt, K, C, ptr(x), ptr(out), ptr(arg), stream_ptr()), 1)
    x3 = x.view(Gt, K, C)
    assert torch.equal(out, x3.max(1)[0])
    first = (x3 == x3.max(1, keepdim=True)[0]).float().argmax(1)     # first maximal row
    assert torch.equal(arg.long(), first)
    d = torch.randn(Gt, C, device=DEV)
    dx = torch.empty(Gt * K, C, device=DEV)
    check(_lib.lib.up3d_group_max_scatter(0, Gt, K, C, ptr(d), ptr(arg), ptr(dx), stream_ptr()), 1)
    ref = torch.zeros(Gt, K, C, device=DEV).scatter_(1, first.unsqueeze(1), d.unsqueeze(1))
    assert torch.equal(dx.view(Gt, K, C), ref)


def test_group_combine_matches_torch():
    from unipre3d_b200 import _lib
    from unipre3d_b200._lib import check, ptr, stream_ptr
    Gt, K, C = 7, 32, 256
    torch.manual_seed(4)
    dl = torch.randn(Gt * K, C, device=DEV)
    dp = torch.randn(Gt, C, device=DEV)
    arg = torch.randint(0, K, (Gt, C), device=DEV, dtype=torch.int32)
    for gpc in (1, 2):
        dx = torch.empty_like(dl)
        part = torch.empty((Gt + gpc - 1) // gpc, 1, C, device=DEV)
        check(_lib.lib.up3d_group_combine(0, Gt, K, C, gpc, ptr(dl), ptr(dp), ptr(arg), ptr(dx), ptr(part), stream_ptr()), 1)
        cs = torch.empty(1, C, device=DEV)
        check(_lib.lib.up3d_bn_reduce_sums(part.shape[0], 1, C, ptr(part), ptr(cs), stream_ptr()), 1)
        cs = cs[0]
        ref = dl.view(Gt, K, C) + torch.zeros(Gt, K, C, device=DEV).scatter_(1, arg.long().unsqueeze(1), dp.unsqueeze(1))
        assert torch.allclose(dx.view(Gt, K, C), ref, atol=1e-6)
        assert torch.allclose(cs, ref.sum((0, 1)), atol=1e-3, rtol=1e-4)


@pytest.mark.parametrize("B,G", [(8, 128), (5, 100)])
def test_staged_group_tile_passes_are_bit_identical_to_direct_loads(B, G):
    """bf16 mini-PointNet, forward + backward: the shared-memory-ring variant of the group-tile kernels (cp.async.bulk,
    several groups per CTA, ragged last CTA) == the direct-load variant bit for bit -- outputs, every parameter gradient
    and the BatchNorm running statistics."""
    from unipre3d_b200 import _lib
    torch.manual_seed(5)
    nb = torch.randn(B, 3, G, 32, device=DEV) * 0.05
    w = torch.randn(B, G, 384, device=DEV)
    res = {}
    prev = _lib.lib.up3d_set_group_tile_staging(1)
    try:
        for staged in (1, 0):
            _lib.lib.up3d_set_group_tile_staging(staged)
            enc = _encoder(seed=7)
            enc.train()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = enc.forward_grouped(nb)
            (out.float() * w).sum().backward()
            res[staged] = ([out.detach().clone()] + [p.grad.clone() for p in enc.parameters()] +
                           [b.clone() for b in enc.buffers()])
    finally:
        _lib.lib.up3d_set_group_tile_staging(prev)
    assert len(res[0]) == len(res[1]) > 10
    for a, b in zip(res[1], res[0]):
        assert torch.equal(a, b)
