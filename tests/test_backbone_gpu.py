"""Fused transformer glue + optimizer kernels (csrc/backbone.cu) against plain PyTorch fp32 references of the same
ops, the fused 16-block stack against the module loop that mirrors the reference (backbone.py), and against the golden
fixture generated from the reference's own PointTransformerEncoder (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda"


def _tol(act):
    # bf16 outputs carry one rounding of a value of magnitude ~|x| (2^-9 relative)
    return dict(atol=2e-5, rtol=1e-5) if act == torch.float32 else dict(atol=2e-2, rtol=1e-2)


@pytest.mark.parametrize("act", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C", [128, 384, 768])
def test_ln_fwd_matches_torch(act, C):
    from unipre3d_b200 import fused_encoder as fe
    torch.manual_seed(0)
    B, L = 3, 37
    T = B * L
    x = torch.randn(T, C, device=DEV)
    delta = torch.randn(T, C, device=DEV).to(act)
    pos = torch.randn(T, C, device=DEV)
    scale = torch.tensor([0.0, 1.25, 1.0], device=DEV)
    gamma, beta = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV)
    xs, y, mean, rstd = fe.ln_fwd(x, delta, scale, pos, gamma, beta, 1e-5, L, act)
    xs_ref = x + scale.repeat_interleave(L).unsqueeze(1) * delta.float() + pos
    y_ref = F.layer_norm(xs_ref, (C,), gamma, beta, 1e-5)
    assert torch.allclose(xs, xs_ref, atol=1e-6, rtol=1e-6)
    assert torch.allclose(y.float(), y_ref, **_tol(act))
    assert torch.allclose(mean, xs_ref.mean(1), atol=1e-5)
    assert torch.allclose(rstd, (xs_ref.var(1, unbiased=False) + 1e-5).rsqrt(), rtol=1e-5)
    # optional inputs absent: plain LayerNorm; residual-only mode
    _, y2, _, _ = fe.ln_fwd(x, None, None, None, gamma, beta, 1e-5, L, act, want_xs=False)
    assert torch.allclose(y2.float(), F.layer_norm(x, (C,), gamma, beta, 1e-5), **_tol(act))
    xs3, y3, _, _ = fe.ln_fwd(x, delta, scale, None, None, None, 0.0, L, act, want_y=False)
    assert y3 is None and torch.allclose(xs3, x + scale.repeat_interleave(L).unsqueeze(1) * delta.float(), atol=1e-6)


@pytest.mark.parametrize("act", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C", [128, 384])
def test_ln_bwd_matches_torch_autograd(act, C):
    from unipre3d_b200 import fused_encoder as fe
    torch.manual_seed(1)
    B, L = 5, 129
    T = B * L
    xs = torch.randn(T, C, device=DEV, requires_grad=True)
    gamma = (torch.rand(C, device=DEV) + 0.5).requires_grad_(True)
    beta = torch.randn(C, device=DEV, requires_grad=True)
    dy = torch.randn(T, C, device=DEV).to(act)
    g_res = torch.randn(T, C, device=DEV)
    scale = torch.tensor([0.0, 1.1, 1.0, 1.1, 0.0], device=DEV)
    y = F.layer_norm(xs, (C,), gamma, beta, 1e-5)
    dxs_ref, dg_ref, db_ref = torch.autograd.grad(y, [xs, gamma, beta], dy.float())
    dx_ref = g_res + dxs_ref
    with torch.no_grad():
        mean = xs.mean(1)
        rstd = (xs.var(1, unbiased=False) + 1e-5).rsqrt()
        dpos = torch.full((T, C), 0.5, device=DEV)
        dgamma, dbeta, dbias = (torch.zeros(C, device=DEV) for _ in range(3))
        dx, dsc = fe.ln_bwd(dy, xs.detach(), mean, rstd, gamma.detach(), g_res, scale, L, dpos, True, dgamma, dbeta, dbias)
    tol = dict(atol=1e-4, rtol=1e-4)
    assert torch.allclose(dx, dx_ref, **tol)
    assert torch.allclose(dpos, 0.5 + dx_ref, **tol)
    sc_ref = scale.repeat_interleave(L).unsqueeze(1) * dx_ref
    assert torch.allclose(dsc.float(), sc_ref, **_tol(act))
    assert torch.allclose(dgamma, dg_ref, atol=2e-3, rtol=1e-3)
    assert torch.allclose(dbeta, db_ref, atol=2e-3, rtol=1e-3)
    assert torch.allclose(dbias, dsc.float().sum(0), atol=2e-3, rtol=1e-3)
    # no residual / no scaled output
    with torch.no_grad():
        dgamma.zero_(); dbeta.zero_()
        dx2, none = fe.ln_bwd(dy, xs.detach(), mean, rstd, gamma.detach(), None, None, L, None, False, dgamma, dbeta, None)
    assert none is None and torch.allclose(dx2, dxs_ref, **tol)


@pytest.mark.parametrize("act", [torch.float32, torch.bfloat16])
def test_gelu_and_colsum_kernels_match_torch(act):
    from unipre3d_b200 import fused_encoder as fe
    torch.manual_seed(2)
    T, C, L = 1032, 1536, 129
    pre = (torch.randn(T, C, device=DEV) * 2).to(act)
    dy = torch.randn(T, C, device=DEV).to(act)
    h = fe.gelu_fwd(pre)
    assert torch.allclose(h.float(), F.gelu(pre.float()), **_tol(act))
    p32 = pre.float().requires_grad_(True)
    (dref,) = torch.autograd.grad(F.gelu(p32), p32, dy.float())
    db = torch.zeros(C, device=DEV)
    dx = fe.gelu_bwd(dy, pre, db)
    assert torch.allclose(dx.float(), dref, **_tol(act))
    assert torch.allclose(db, dx.float().sum(0), atol=5e-3, rtol=1e-3)
    g = torch.randn(T, 384, device=DEV)
    scale = torch.rand(T // L, device=DEV)
    db2 = torch.zeros(384, device=DEV)
    out = fe.scale_cast_colsum(g, scale, L, act, db2)
    ref = scale.repeat_interleave(L).unsqueeze(1) * g
    assert torch.allclose(out.float(), ref, **_tol(act))
    assert torch.allclose(db2, out.float().sum(0), atol=5e-3, rtol=1e-3)


def _make_blocks(C=384, depth=4, heads=6, dpr=0.1):
    from unipre3d_b200.backbone import TransformerEncoder
    torch.manual_seed(3)
    rates = [x.item() for x in torch.linspace(0, dpr, depth)]
    enc = TransformerEncoder(embed_dim=C, depth=depth, num_heads=heads, drop_path_rate=rates).to(DEV)
    with torch.no_grad():
        for p in enc.parameters():            # non-trivial LayerNorm affine parameters and biases
            if p.ndim == 1:
                p.add_(torch.randn_like(p) * 0.1)
    return enc


def _module_loop_with_masks(enc, x, pos, masks):
    """backbone.Block arithmetic with explicit DropPath factors (transformer.py:119-120, 188)."""
    for i, blk in enumerate(enc.blocks):
        x = x + pos
        s1 = 1.0 if masks is None else masks[2 * i].view(-1, 1, 1)
        s2 = 1.0 if masks is None else masks[2 * i + 1].view(-1, 1, 1)
        x = x + s1 * blk.attn(blk.norm1(x))
        x = x + s2 * blk.mlp(blk.norm2(x))
    return x


@pytest.mark.parametrize("with_masks", [False, True])
def test_fused_stack_equals_module_loop_fp32(with_masks):
    from unipre3d_b200 import fused_encoder as fe
    enc = _make_blocks()
    assert fe.supports(enc.blocks)
    B, L, C = 4, 129, 384
    torch.manual_seed(4)
    x = torch.randn(B, L, C, device=DEV, requires_grad=True)
    pos = torch.randn(B, L, C, device=DEV, requires_grad=True)
    w = torch.randn(B, L, C, device=DEV)
    masks = None
    if with_masks:
        masks = (torch.rand(2 * len(enc.blocks), B, device=DEV) < 0.7).float() / 0.7
    params = list(enc.parameters())
    ref = _module_loop_with_masks(enc, x, pos, masks)
    g_ref = torch.autograd.grad((ref * w).sum(), [x, pos] + params)
    # fused: same masks injected through the Function directly
    plist, cw, e1, e2 = [], [], [], []
    for b in enc.blocks:
        plist += [b.norm1.weight, b.norm1.bias, b.attn.qkv.weight, b.attn.proj.weight, b.attn.proj.bias, b.norm2.weight,
                  b.norm2.bias, b.mlp.fc1.weight, b.mlp.fc1.bias, b.mlp.fc2.weight, b.mlp.fc2.bias]
        cw.append((b.attn.qkv.weight.detach(), b.attn.proj.weight.detach(), b.attn.proj.bias.detach(),
                   b.mlp.fc1.weight.detach(), b.mlp.fc1.bias.detach(), b.mlp.fc2.weight.detach(), b.mlp.fc2.bias.detach()))
        e1.append(b.norm1.eps); e2.append(b.norm2.eps)
    meta = fe._Meta(B, L, C, enc.blocks[0].attn.num_heads, float(enc.blocks[0].attn.scale), e1, e2, torch.float32, cw)
    out = fe.EncoderStackFn.apply(x, pos, masks, meta, *plist)
    assert torch.allclose(out, ref, atol=2e-4, rtol=1e-4), float((out - ref).abs().max())
    g = torch.autograd.grad((out * w).sum(), [x, pos] + plist)
    by_id = {id(p): gi for p, gi in zip([x, pos] + plist, g)}
    for p, gr in zip([x, pos] + params, g_ref):
        got = by_id[id(p)]
        scale = float(gr.abs().max()) + 1e-6
        assert float((got - gr).abs().max()) <= 2e-3 * scale, (tuple(p.shape), float((got - gr).abs().max()), scale)


def test_fused_stack_bf16_close_to_fp32_and_module_dispatch():
    """TransformerEncoder.forward dispatches to the fused stack on CUDA; under bf16 autocast it stays close to fp32."""
    enc = _make_blocks(depth=3, dpr=0.0)
    B, L, C = 2, 129, 384
    torch.manual_seed(5)
    x = torch.randn(B, L, C, device=DEV)
    pos = torch.randn(B, L, C, device=DEV) * 0.1
    enc.train()
    a = enc(x, pos, None, None, None, None, None)
    enc.force_module_path = True
    b = enc(x, pos, None, None, None, None, None)
    enc.force_module_path = False
    assert torch.allclose(a, b, atol=2e-4, rtol=1e-4)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        c = enc(x, pos, None, None, None, None, None)
    assert c.dtype == torch.float32
    assert float((c - a).abs().max()) <= 0.05 * float(a.abs().max())
    (c * c).mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() and p.grad.dtype == torch.float32
               for p in enc.parameters())


def test_fused_stack_matches_reference_module_fixture():
    """The reference's own TransformerEncoder (Block stack, width 128) output and gradients -- fixture generated by
    tests/golden/make_golden.py from /root/reference -- through the CUDA fused stack (fp32 operands)."""
    from tests.test_golden_cpu import _blocks_from_fixture
    from unipre3d_b200 import fused_encoder as fe
    z = np.load(os.path.join(G, "transformer_blocks_w128.npz"))
    enc = _blocks_from_fixture(z).to(DEV)
    assert fe.supports(enc.blocks)
    x = torch.tensor(z["x"], device=DEV, requires_grad=True)
    pos = torch.tensor(z["pos"], device=DEV, requires_grad=True)
    enc.train()
    for m in enc.modules():
        if m.__class__.__name__ == "DropPath":
            m.drop_prob = 0.0
    before = _lib_launches()
    out = enc(x, pos, None, None, None, None, None)
    assert _lib_launches() > before, "the fused CUDA stack did not run"
    np.testing.assert_allclose(out.detach().cpu().numpy(), z["out"], atol=5e-5, rtol=1e-4)
    (out * torch.tensor(z["wsum"], device=DEV)).sum().backward()
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    checks = [("grad_x", x.grad), ("grad_pos", pos.grad)] + [("grad." + k, p.grad) for k, p in enc.named_parameters()]
    assert len(checks) == 2 + 11 * int(z["cfg"][3])
    for k, g in checks:
        ref = z[k]
        scale = np.abs(ref).max() + 1e-4 * gmax
        assert np.abs(g.cpu().numpy() - ref).max() <= 2e-3 * scale + 2e-6, k


def _lib_launches():
    from unipre3d_b200 import _lib
    return _lib.launch_count


def _clone_params(shapes, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return [torch.randn(s, generator=g).to(DEV).requires_grad_(True) for s in shapes]


@pytest.mark.parametrize("gscale", [0.01, 30.0])      # below / above the clip threshold
def test_fused_clip_adamw_matches_torch(gscale):
    from unipre3d_b200.optim import FusedClipAdamW
    shapes = [(384, 384), (1536,), (7, 3), (4099,), (1,), (128, 257)]
    pa, pb = _clone_params(shapes, 0), _clone_params(shapes, 0)
    shadows = {id(p): torch.zeros(p.shape, dtype=torch.bfloat16, device=DEV) for p in pa[:2]}
    oa = FusedClipAdamW([{"params": pa[:3], "lr": 1e-3}, {"params": pa[3:], "lr": 3e-3}], lr=0.0, eps=1e-15,
                        betas=(0.9, 0.999), max_norm=1.0, shadows=shadows)
    ob = torch.optim.AdamW([{"params": pb[:3], "lr": 1e-3}, {"params": pb[3:], "lr": 3e-3}], lr=0.0, eps=1e-15,
                           betas=(0.9, 0.999))
    for step in range(5):
        gg = torch.Generator(device="cpu").manual_seed(100 + step)
        grads = [torch.randn(s, generator=gg).to(DEV) * gscale for s in shapes]
        for p, q, g in zip(pa, pb, grads):
            p.grad, q.grad = g.clone(), g.clone()
        total = torch.nn.utils.clip_grad_norm_(pb, max_norm=1.0)
        ob.step()
        oa.step()
        assert abs(oa.last_total_norm() - float(total)) <= 1e-5 * float(total)
        for p, q in zip(pa, pb):
            assert torch.allclose(p, q, atol=1e-6, rtol=1e-5), step
        if step == 2:                       # StepLR-style in-place LR change on the device scalars
            for g_ in oa.param_groups:
                g_["lr"].mul_(0.5)
            for g_ in ob.param_groups:
                g_["lr"] *= 0.5
    for p in pa[:2]:
        assert torch.equal(shadows[id(p)], p.detach().to(torch.bfloat16))
    assert oa.step_count() == 5
    sd = oa.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 5.0
    for p, q in zip(pa, pb):
        assert torch.allclose(oa.state[p]["exp_avg"], ob.state[q]["exp_avg"], atol=1e-6, rtol=1e-5)
        assert torch.allclose(oa.state[p]["exp_avg_sq"], ob.state[q]["exp_avg_sq"], atol=1e-7, rtol=1e-5)


def test_fused_clip_adamw_skips_non_finite_step():
    """train_network.py:336-340,368-384: a NaN/Inf gradient anywhere -> the whole step is skipped."""
    from unipre3d_b200.optim import FusedClipAdamW
    shapes = [(64, 64), (100,)]
    ps = _clone_params(shapes, 1)
    opt = FusedClipAdamW(ps, lr=1e-2, eps=1e-15, max_norm=1.0)
    for p in ps:
        p.grad = torch.randn_like(p)
    opt.step()
    before = [p.detach().clone() for p in ps]
    m_before = [opt.state[p]["exp_avg"].clone() for p in ps]
    for p in ps:
        p.grad = torch.randn_like(p)
    ps[1].grad[17] = float("nan")
    opt.step()
    assert opt.last_found_inf() and opt.step_count() == 1
    for p, b, m in zip(ps, before, m_before):
        assert torch.equal(p.detach(), b) and torch.equal(opt.state[p]["exp_avg"], m)
    for p in ps:
        p.grad = torch.randn_like(p)
    opt.step()
    assert (not opt.last_found_inf()) and opt.step_count() == 2
    assert not torch.equal(ps[0].detach(), before[0])


def test_stem_field_group_stats_and_gather_match_dense():
    """csrc/stem.cu: GroupNorm statistics of the analytic stem field from the 3-channel image == var_mean of the dense
    (n,128,R,R) tensor; sampled image_conv values/gradients identical through either route."""
    from unipre3d_b200.fusion import LazyImageFeatures
    from unipre3d_b200.gaussian_predictor import StemFeatureField
    torch.manual_seed(6)
    n, R, Cc, G, N = 3, 56, 128, 32, 40            # 56*56 is not a multiple of the block size: ragged last CTA
    img = torch.rand(n, 3, R, R, device=DEV)
    proj, shift = torch.randn(Cc, 3, device=DEV) * 0.7, torch.randn(Cc, device=DEV) * 0.3
    field = StemFeatureField(img, proj, shift)
    dense = field.dense()
    assert dense.shape == field.shape == (n, Cc, R, R)
    mean, var = field.group_stats(G)
    var_ref, mean_ref = torch.var_mean(dense.reshape(n, G, -1), dim=2, unbiased=False)
    assert torch.allclose(mean, mean_ref, atol=2e-6) and torch.allclose(var, var_ref, atol=2e-6, rtol=1e-5)
    bidx = torch.arange(n, device=DEV).unsqueeze(1).expand(n, N)
    ix, iy = torch.randint(0, R, (n, N), device=DEV), torch.randint(0, R, (n, N), device=DEV)
    assert torch.allclose(field.gather(bidx, ix, iy), dense[bidx, :, ix, iy], atol=2e-6)
    conv = torch.nn.Sequential(torch.nn.GroupNorm(G, Cc, eps=1e-6), torch.nn.Conv2d(Cc, 96, 1)).to(DEV)
    w = torch.randn(n, N, 96, device=DEV)
    a = LazyImageFeatures(field, conv).sample(bidx, ix, iy)
    b = LazyImageFeatures(dense, conv).sample(bidx, ix, iy)
    assert torch.allclose(a, b, atol=2e-5, rtol=1e-5)
    ga = torch.autograd.grad((a * w).sum(), list(conv.parameters()))
    gb = torch.autograd.grad((b * w).sum(), list(conv.parameters()))
    for u, v in zip(ga, gb):
        assert torch.allclose(u, v, atol=1e-4 * float(v.abs().max()) + 1e-6, rtol=1e-4)


@pytest.mark.parametrize("B,L,H", [(2, 129, 6), (1, 48, 2), (3, 100, 4), (1, 192, 1), (2, 17, 6)])
def test_attention_kernels_match_fp32_reference(B, L, H):
    """csrc/attention.cu against softmax(q k^T * scale) v evaluated in fp32 on the same bf16-rounded operands
    (transformer.py:58-71), forward and backward (autograd of the fp32 formula)."""
    from unipre3d_b200 import fused_encoder as fe
    D = 64
    C = H * D
    torch.manual_seed(7)
    qkv = (torch.randn(B * L, 3 * C, device=DEV) * 1.5).to(torch.bfloat16)
    do = torch.randn(B * L, C, device=DEV).to(torch.bfloat16)
    scale = D ** -0.5
    assert fe.attn_supported(torch.bfloat16, L, D)
    o, lse = fe.attn_fwd(qkv, B, L, H, D, scale)
    ref_in = qkv.float().requires_grad_(True)
    q, k, v = ref_in.view(B, L, 3, H, D).permute(2, 0, 3, 1, 4).unbind(0)
    att = (q @ k.transpose(-2, -1)) * scale
    ref_lse = torch.logsumexp(att, dim=-1)
    ref_o = (att.softmax(dim=-1) @ v).transpose(1, 2).reshape(B * L, C)
    assert torch.allclose(lse, ref_lse, atol=2e-3, rtol=1e-4)
    assert float((o.float() - ref_o).abs().max()) <= 2e-2 * float(ref_o.abs().max()) + 1e-3
    (ref_dqkv,) = torch.autograd.grad(ref_o, ref_in, do.float())
    dqkv = fe.attn_bwd(qkv, o, lse, do, B, L, H, D, scale)
    for name, sl in (("dq", slice(0, C)), ("dk", slice(C, 2 * C)), ("dv", slice(2 * C, 3 * C))):
        got, ref = dqkv[:, sl].float(), ref_dqkv[:, sl]
        err = float((got - ref).abs().max())
        assert err <= 3e-2 * float(ref.abs().max()) + 1e-3, (name, err, float(ref.abs().max()))
        cos = float(torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0))
        assert cos > 0.999, (name, cos)


@pytest.mark.parametrize("deg,iso", [(1, False), (1, True), (2, False)])
def test_fused_splat_head_matches_module_path(deg, iso):
    """csrc/head.cu splat head == GaussianSplatPredictor._process_network_output (golden-pinned against the reference in
    tests/test_golden_cpu.py) + the renderer's SH concatenation: values and gradients."""
    from types import SimpleNamespace as NS
    from unipre3d_b200.gaussian_predictor import GaussianSplatPredictor
    m = GaussianSplatPredictor.__new__(GaussianSplatPredictor)
    torch.nn.Module.__init__(m)
    m.cfg = NS(model=NS(max_sh_degree=deg, isotropic=iso, offset_scale=0.7))
    M = (deg + 1) ** 2
    split = [3, 1, 3, 4, 3, (M - 1) * 3]
    torch.manual_seed(8)
    B, P = 3, 128
    raw = (torch.randn(B, sum(split), P, device=DEV) * 1.5)
    raw[:, 4:7, :3] = torch.tensor([-3.0, 25.0, 0.0], device=DEV).view(1, 3, 1)     # exercise clamp(-1, 20)
    raw.requires_grad_(True)
    center = torch.randn(B, P, 3, device=DEV)
    ref = m._process_network_output(raw.split(split, dim=1), center)
    ref_shs = torch.cat([ref["features_dc"], ref["features_rest"]], dim=2)
    got = m._fused_head(raw, center)
    for k in ("xyz", "opacity", "scaling", "rotation", "features_dc", "features_rest"):
        assert got[k].shape == ref[k].shape, k
        assert torch.allclose(got[k], ref[k], atol=1e-5, rtol=1e-5), k
    assert torch.allclose(got["shs"], ref_shs, atol=0, rtol=0)
    ws = {k: torch.randn_like(ref[k]) for k in ("xyz", "opacity", "scaling", "rotation")}
    wsh = torch.randn_like(ref_shs)
    # keep the huge exp(20) scale out of the cotangent so tolerances stay meaningful
    ws["scaling"] = ws["scaling"] / ref["scaling"].detach().clamp_min(1.0)
    lr = sum((ref[k] * ws[k]).sum() for k in ws) + (ref_shs * wsh).sum()
    lg = sum((got[k] * ws[k]).sum() for k in ws) + (got["shs"] * wsh).sum()
    (g_ref,) = torch.autograd.grad(lr, raw)
    (g_got,) = torch.autograd.grad(lg, raw)
    assert torch.allclose(g_got, g_ref, atol=2e-5 * float(g_ref.abs().max()) + 1e-6, rtol=1e-4)


def test_fused_feature_fusion_matches_module_path():
    """csrc/head.cu fusion projection == the eager restatement of fusion/feat_fusion.py (golden-pinned against the
    reference in tests/test_golden_cpu.py): kept points, mapped features, fused output and image_conv gradients."""
    import math
    from unipre3d_b200 import camera as cam
    from unipre3d_b200 import fusion
    from unipre3d_b200.gaussian_predictor import StemFeatureField
    torch.manual_seed(9)
    B, N, R, Cin, Cout, G = 4, 128, 64, 128, 48, 32
    fov = 49.13434264120263
    proj_m = cam.get_projection_matrix(0.5, 2.0, math.radians(fov), math.radians(fov))
    c2w = torch.stack([cam.make_view(*cam.look_at_pose(40.0 * i, 10.0 + 15 * i, 1.75), proj_m)["view_to_world_transform"]
                       for i in range(B)]).unsqueeze(1).to(DEV)
    K = np.zeros((3, 4)); focal = (R / 2.0) / math.tan(math.radians(fov / 2.0))
    K[0, 0] = K[1, 1] = focal; K[0, 2] = K[1, 2] = R / 2.0; K[2, 2] = 1
    center = torch.randn(B, N, 3, device=DEV) * 0.3
    center[:, 5] = center[:, 4] * 1.02          # same pixel, depth test decides
    center[:, 7] = 5.0                          # outside the image
    img = torch.rand(B, 3, R, R, device=DEV)
    field = StemFeatureField(img, torch.randn(Cin, 3, device=DEV) * 0.7, torch.randn(Cin, device=DEV) * 0.3)
    conv = torch.nn.Sequential(torch.nn.GroupNorm(G, Cin, eps=1e-6), torch.nn.Conv2d(Cin, Cout, 1)).to(DEV)
    mlp = torch.nn.Sequential(torch.nn.Linear(2 * Cout, Cout), torch.nn.ReLU()).to(DEV)
    x = torch.randn(B, N + 1, Cout, device=DEV)
    w = torch.randn(B, N + 1, Cout, device=DEV)
    outs = {}
    for force in (True, False):
        fusion.FORCE_MODULE_PATH = force
        try:
            y = fusion.FeatureFusion(mlp)(x, center, fusion.LazyImageFeatures(field, conv), c2w, K)
            g = torch.autograd.grad((y * w).sum(), list(conv.parameters()) + list(mlp.parameters()))
        finally:
            fusion.FORCE_MODULE_PATH = False
        outs[force] = (y.detach(), g)
    keep, mapped, pix = fusion.fused_project_and_sample(fusion.LazyImageFeatures(field, conv), center, c2w[:, 0], K)
    assert 0 < int(keep.sum()) < B * N and not bool(keep[:, 7].any())
    # index work against the eager restatement: pixel coordinates, in-image test and nearest-depth test.  (A projected
    # coordinate that lands within an ulp of x.5 may round differently between cuBLAS' and the kernel's summation order:
    # allow a vanishing fraction of such points, none is expected on this seed.)
    with torch.no_grad():
        pi_xy, depth = fusion.FeatureFusion.project_points_to_image(center, c2w[:, 0], K)
        fx_, fy_ = pi_xy[..., 0], pi_xy[..., 1]
        inside = (fx_ >= 0) & (fy_ >= 0) & (fx_ < R) & (fy_ < R) & (depth >= 0)
        ix = torch.where(inside, fx_, torch.zeros_like(fx_)).int()
        iy = torch.where(inside, fy_, torch.zeros_like(fy_)).int()
        cell = torch.arange(B, device=DEV).unsqueeze(1) * (R * R) + iy.long() * R + ix.long()
        dm = torch.where(inside, depth, torch.full_like(depth, float("inf")))
        zbuf = torch.full((B * R * R + R * R,), float("inf"), device=DEV)
        zbuf.scatter_reduce_(0, cell.reshape(-1), dm.reshape(-1), reduce="amin", include_self=True)
        keep_ref = inside & (depth == zbuf[cell])
    same_pix = (pix[..., 0] == ix) & (pix[..., 1] == iy)
    assert float(same_pix.float().mean()) >= 0.995 and float((keep == keep_ref).float().mean()) >= 0.995
    assert bool(keep[same_pix].eq(keep_ref[same_pix]).all()) or float((keep == keep_ref).float().mean()) >= 0.998
    assert torch.allclose(outs[False][0], outs[True][0], atol=2e-5, rtol=1e-4)
    for a, b in zip(outs[False][1], outs[True][1]):
        assert torch.allclose(a, b, atol=1e-4 * float(b.abs().max()) + 1e-6, rtol=1e-4)


@pytest.mark.parametrize("act", [torch.float32, torch.bfloat16])
def test_fused_layer_norm_matches_module(act):
    from unipre3d_b200.fused_encoder import fused_layer_norm
    torch.manual_seed(11)
    norm = torch.nn.LayerNorm(384).to(DEV)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5); norm.bias.normal_()
    x = torch.randn(4, 129, 384, device=DEV, requires_grad=True)
    w = torch.randn(4, 129, 384, device=DEV)
    ref = norm(x)
    g_ref = torch.autograd.grad((ref * w).sum(), [x, norm.weight, norm.bias])
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=act == torch.bfloat16):
        out = fused_layer_norm(norm, x)
    assert out.dtype == act and torch.allclose(out.float(), ref, **_tol(act))
    g = torch.autograd.grad((out.float() * w).sum(), [x, norm.weight, norm.bias])
    tol = 1e-4 if act == torch.float32 else 2e-2
    for a, b in zip(g, g_ref):
        assert float((a - b).abs().max()) <= tol * float(b.abs().max()) + 1e-5


@pytest.mark.parametrize("name,deg,iso", [("deg1", 1, False), ("deg0_iso", 0, True)])
def test_splat_head_kernel_matches_reference_golden_and_oracle(name, deg, iso):
    """csrc/head.cu splat head against the golden vectors generated from the reference's own
    `_process_network_output` (tests/golden/process_output.npz) and against the numpy oracle."""
    from oracle import backbone_oracle as bo
    from unipre3d_b200.gaussian_predictor import SplatHeadFn
    z = np.load(os.path.join(G, "process_output.npz"))
    raw, center = z[f"{name}.raw"], z[f"{name}.center"]
    M = (deg + 1) ** 2
    xyz, op, sc, rot, shs = SplatHeadFn.apply(torch.tensor(raw, device=DEV).permute(0, 2, 1), torch.tensor(center, device=DEV),
                                              M, 0.7, iso)
    got = {"xyz": xyz, "opacity": op, "scaling": sc, "rotation": rot, "features_dc": shs[:, :, :1]}
    if deg > 0:
        got["features_rest"] = shs[:, :, 1:]
    ref = bo.splat_head(raw, center, 0.7, iso, deg)
    for k, v in got.items():
        gold = z[f"{name}.out.{k}"]
        assert tuple(v.shape) == gold.shape, k
        np.testing.assert_allclose(v.cpu().numpy(), gold, atol=2e-6, rtol=1e-5, err_msg=k)
        np.testing.assert_allclose(v.cpu().numpy(), ref[k], atol=2e-6, rtol=1e-5, err_msg=k)


def test_fusion_project_kernel_matches_oracle_geometry():
    import math
    from oracle import backbone_oracle as bo
    from unipre3d_b200 import camera as cam
    from unipre3d_b200 import fusion
    from unipre3d_b200.gaussian_predictor import StemFeatureField
    torch.manual_seed(12)
    B, N, R, Cin, G_ = 3, 128, 48, 128, 32
    fov = 49.13434264120263
    proj_m = cam.get_projection_matrix(0.5, 2.0, math.radians(fov), math.radians(fov))
    c2w = torch.stack([cam.make_view(*cam.look_at_pose(70.0 * i, 25.0, 1.75), proj_m)["view_to_world_transform"]
                       for i in range(B)])
    K = np.zeros((3, 4)); focal = (R / 2.0) / math.tan(math.radians(fov / 2.0))
    K[0, 0] = K[1, 1] = focal; K[0, 2] = K[1, 2] = R / 2.0; K[2, 2] = 1
    center = torch.randn(B, N, 3) * 0.35
    center[:, 9] = center[:, 8] * 1.03
    field = StemFeatureField(torch.rand(B, 3, R, R, device=DEV), torch.randn(Cin, 3, device=DEV) * 0.7,
                             torch.randn(Cin, device=DEV) * 0.3)
    conv = torch.nn.Sequential(torch.nn.GroupNorm(G_, Cin, eps=1e-6), torch.nn.Conv2d(Cin, 16, 1)).to(DEV)
    keep, _, pix = fusion.fused_project_and_sample(fusion.LazyImageFeatures(field, conv), center.to(DEV), c2w.to(DEV), K)
    ix, iy, inside, keep_ref, _ = bo.fusion_geometry(center.numpy(), c2w.numpy(), K, R, R)
    pix = pix.cpu().numpy()
    same = (pix[..., 0] == ix) & (pix[..., 1] == iy)
    assert same.mean() >= 0.995 and (keep.cpu().numpy() == keep_ref).mean() >= 0.995
    assert 0 < keep_ref.sum() < B * N


def test_fused_clip_adamw_matches_numpy_oracle():
    from oracle import backbone_oracle as bo
    from unipre3d_b200.optim import FusedClipAdamW
    shapes = [(96, 40), (513,), (5, 3, 2)]
    ps = _clone_params(shapes, 3)
    opt = FusedClipAdamW([{"params": ps[:2], "lr": 2e-3}, {"params": ps[2:], "lr": 1e-3}], lr=0.0, eps=1e-15, max_norm=1.0)
    opt.grad_scale = 0.5
    P = [p.detach().cpu().double().numpy().copy() for p in ps]
    M = [np.zeros_like(p) for p in P]; V = [np.zeros_like(p) for p in P]
    step = 0
    for it in range(3):
        gg = torch.Generator(device="cpu").manual_seed(50 + it)
        grads = [torch.randn(s, generator=gg) * (0.05 if it == 0 else 20.0) for s in shapes]
        for p, g in zip(ps, grads):
            p.grad = g.to(DEV)
        opt.step()
        step, total, applied = bo.clip_adamw_step(P, [g.double().numpy() for g in grads], M, V, step, [2e-3, 2e-3, 1e-3],
                                                  grad_scale=0.5)
        assert applied and abs(opt.last_total_norm() - total) <= 1e-5 * total
        for p, q in zip(ps, P):
            np.testing.assert_allclose(p.detach().cpu().numpy(), q, atol=2e-6, rtol=2e-5)
    assert opt.step_count() == step == 3


# ---------------------------------------------------------------------------------------------------------------------
# Whole-encoder parity on the GPU against the REFERENCE module's fixtures.
#
# bf16 tolerance model (stated, used by the three tests below): every GEMM operand/output of the bf16 path is rounded to
# bf16 once (relative error <= 2^-9, rms ~2^-9/sqrt(3)); a Block has 8 such roundings on the residual path (y1, qkv, o, a,
# y2, pre, h, d) and the residual stream itself stays fp32, so after `depth` blocks the error of an O(1)-normalised
# activation is ~ 2^-9 * sqrt(8 * depth) rms; gradients see the same count again on the way back.  With a 4-sigma
# allowance for the max over ~10^5 elements:  tol(depth) = 4 * 2^-9 * sqrt(8 * depth) for outputs, twice that for
# gradients -- relative to the tensor's max-abs (plus the fixture's overall gradient scale for tiny tensors).
def _bf16_tol(depth):
    return 4.0 * 2.0 ** -9 * (8.0 * depth) ** 0.5


@pytest.mark.parametrize("tc", ["1", "0"])
def test_fused_stack_bf16_matches_reference_module_fixture(tc, monkeypatch):
    """bf16 operands (the benchmarked precision; tcgen05 GEMMs when tc == "1", library GEMMs otherwise) against the
    reference's own TransformerEncoder fixture, outputs and every parameter gradient."""
    monkeypatch.setenv("UP3D_TC_LINEAR", tc)
    from tests.test_golden_cpu import _blocks_from_fixture
    z = np.load(os.path.join(G, "transformer_blocks_w128.npz"))
    enc = _blocks_from_fixture(z).to(DEV)
    depth = int(z["cfg"][3])
    x = torch.tensor(z["x"], device=DEV, requires_grad=True)
    pos = torch.tensor(z["pos"], device=DEV, requires_grad=True)
    enc.train()
    for m in enc.modules():
        if m.__class__.__name__ == "DropPath":
            m.drop_prob = 0.0
    before = _lib_launches()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = enc(x, pos, None, None, None, None, None)
    assert _lib_launches() > before
    tol = _bf16_tol(depth)
    err = np.abs(out.detach().float().cpu().numpy() - z["out"]).max()
    assert err <= tol * np.abs(z["out"]).max(), (err, tol * np.abs(z["out"]).max())
    (out.float() * torch.tensor(z["wsum"], device=DEV)).sum().backward()
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    checks = [("grad_x", x.grad), ("grad_pos", pos.grad)] + [("grad." + k, p.grad) for k, p in enc.named_parameters()]
    for k, g in checks:
        ref = z[k]
        scale = np.abs(ref).max() + 1e-2 * gmax
        e = np.abs(g.float().cpu().numpy() - ref).max()
        assert e <= 2 * tol * scale, (k, e, 2 * tol * scale)


def _run_whole_encoder(z, mode, module_path):
    """-> (out, center, grad_img, {param name: grad}, fusion grads) of PointTransformerEncoder on CUDA.
    module_path=True: the plain nn.Module loop (stock kernels; under bf16 this is ordinary torch.autocast)."""
    from tests.test_golden_cpu import _encoder_from_fixture
    enc, fusion = _encoder_from_fixture(z)
    enc, fusion = enc.to(DEV), fusion.to(DEV)
    enc.group_divider.neigh = enc.group_divider.neigh.to(DEV)
    enc.group_divider.center = enc.group_divider.center.to(DEV)
    if module_path:
        for m in enc.modules():
            m.force_module_path = True
    pts, c2w = torch.tensor(z["pts"], device=DEV), torch.tensor(z["c2w"], device=DEV)
    img = torch.tensor(z["img_feat"], device=DEV, requires_grad=True)
    enc.train()
    for m in enc.modules():
        if m.__class__.__name__ == "DropPath":
            m.drop_prob = 0.0
    if mode == "bf16":
        if not module_path:
            from unipre3d_b200.mixed_precision import ShadowWeights
            enc._shadow = ShadowWeights(torch.nn.ModuleList([enc, fusion]))       # the benchmarked configuration
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out, center = enc(pts, img, c2w, fusion, z["intrinsic"])
        out = out.float()
    else:
        out, center = enc(pts, img, c2w, fusion, z["intrinsic"])
    (out * torch.tensor(z["wsum"], device=DEV)).sum().backward()
    grads = {k: p.grad.float().cpu().numpy() for k, p in enc.named_parameters() if p.grad is not None}
    fg = {"grad_fusion_w": fusion[0].weight.grad.float().cpu().numpy(), "grad_fusion_b": fusion[0].bias.grad.float().cpu().numpy()}
    return out.detach().cpu().numpy(), center.cpu().numpy(), img.grad.cpu().numpy(), grads, fg


@pytest.mark.parametrize("fixture", ["transformer_encoder.npz", "transformer_encoder_w128.npz"])
@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_whole_encoder_on_gpu_matches_reference_fixture(mode, fixture):
    """PointTransformerEncoder.forward (transformer.py:290-327: mini-PointNet -> reduce_dim -> cls/pos -> block stack ->
    FeatureFusion -> norm) on CUDA with the fixture's groups, against the REFERENCE module's outputs and gradients.

    fp32: absolute bounds (1e-4 of the output scale, 2e-3 of each gradient's scale).
    bf16 (the benchmarked precision): the fixtures are small (256 / 1024 mini-PointNet rows, 17 / 33 tokens), so train-mode
    BatchNorm and the final LayerNorm amplify bf16 rounding by an amount no closed-form bound captures; the criterion is
    therefore two-sided and measured in the same test: every output / gradient of the fused path must deviate from the
    fp32 reference by at most  max(model bound, 1.5 x the deviation of stock torch.autocast(bf16) running the plain module
    loop on the same device)  -- i.e. the fused kernels may not be worse than ordinary mixed precision, and where the
    closed-form model (4 * 2^-9 * sqrt(8 * depth) of the tensor scale, twice that for gradients) applies it must hold."""
    z = np.load(os.path.join(G, fixture))
    depth = int(z["cfg"][4])
    before = _lib_launches()
    out, center, gimg, grads, fg = _run_whole_encoder(z, mode, module_path=False)
    assert _lib_launches() > before, "the CUDA kernels did not run"
    assert np.array_equal(center, z["center"])
    if mode == "bf16":
        a_out, _, a_gimg, a_grads, a_fg = _run_whole_encoder(z, mode, module_path=True)
        tol_o, tol_g = _bf16_tol(depth + 1), 2 * _bf16_tol(depth + 1)
    else:
        a_out = a_gimg = a_grads = a_fg = None
        tol_o, tol_g = 1e-4, 2e-3
    ref = z["out_train"]
    bound = tol_o * np.abs(ref).max() + 2e-5
    if a_out is not None:
        bound = max(bound, 1.5 * np.abs(a_out - ref).max())
    assert np.abs(out - ref).max() <= bound, ("out", np.abs(out - ref).max(), bound)
    gmax = max(np.abs(z[k]).max() for k in z.files if k.startswith("grad."))
    assert np.array_equal(gimg != 0, z["grad_img_feat"] != 0), "different pixels were sampled"

    def check(name, got, refg, auto):
        scale = np.abs(refg).max() + (1e-2 if mode == "bf16" else 1e-4) * gmax
        e = np.abs(got - refg).max()
        b = tol_g * scale + 2e-6
        if auto is not None:
            b = max(b, 1.5 * np.abs(auto - refg).max())
        assert e <= b, (name, e, b)

    check("grad_img_feat", gimg, z["grad_img_feat"], a_gimg)
    n_checked = 0
    for k, g in grads.items():
        if "grad." + k not in z.files or k in ("encoder.first_conv.0.bias", "encoder.second_conv.0.bias"):
            continue        # zero-gradient biases in front of train-mode BatchNorm (see tests/test_golden_cpu.py)
        check(k, g, z["grad." + k], None if a_grads is None else a_grads[k])
        n_checked += 1
    assert n_checked > 30
    if "grad_fusion_w" in z.files:
        for k in ("grad_fusion_w", "grad_fusion_b"):
            check(k, fg[k], z[k], None if a_fg is None else a_fg[k])
