"""Sparse 3-D convolutions (csrc/sparse_conv.cu through the C ABI) against the dense-grid oracle (oracle/sparse_oracle.py:
torch conv3d on the densified voxel grid, same bf16-rounded operands, fp64 sums) -- forward, input gradients and weight
gradients of SubMConv3d (k = 1, 3, 5), SparseConv3d (k 2, s 2) and SparseInverseConv3d, including empty neighbourhoods,
several batches, channel counts that need padding, and voxel counts that are not multiples of the 64-row CTA tile.

Tolerance: identical bf16 operands, fp32 accumulation on the device -> |err| <= 1e-5 * sum|a||w| + 1e-6 forward; gradients
see one extra bf16 rounding of dout / activations inside the kernels (2^-8 relative on each product term)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _voxels(n, grid, batches, seed):
    g = torch.Generator().manual_seed(seed)
    cells = torch.randperm(batches * grid ** 3, generator=g)[:n]
    b = cells // grid ** 3
    r = cells % grid ** 3
    return torch.stack([b, r // grid ** 2, (r // grid) % grid, r % grid], 1).int()


def _check_close(got, ref, scale, rel, name):
    err = (got.double().cpu() - ref).abs()
    bound = rel * scale + 1e-6
    assert bool((err <= bound).all()), (name, float(err.max()), float(bound.max() if torch.is_tensor(bound) else bound))


@pytest.mark.parametrize("n,grid,B,Ci,Co,k", [(300, 7, 2, 6, 32, 5), (1000, 12, 2, 32, 32, 3), (129, 6, 1, 64, 96, 3),
                                              (2000, 16, 3, 96, 64, 1), (70, 9, 1, 16, 256, 3), (5000, 24, 2, 128, 128, 3)])
def test_subm_conv_forward_backward(n, grid, B, Ci, Co, k):
    from oracle import sparse_oracle as so
    from unipre3d_b200.sparse import SparseConvTensor, SubMConv3d
    idx = _voxels(n, grid, B, seed=n)
    g = torch.Generator().manual_seed(n + 1)
    feats = torch.randn(n, Ci, generator=g)
    conv = SubMConv3d(Ci, Co, k, bias=True, indice_key="a").cuda()
    with torch.no_grad():
        conv.bias.normal_(0, 0.1)
    x = SparseConvTensor(feats.cuda().requires_grad_(True), idx.cuda())
    xf = x.features
    y = conv(x)
    sidx = x.indices.cpu()
    sfe = xf.detach().cpu()
    ref = so.subm_conv(sfe, sidx, conv.weight.detach().cpu(), conv.bias.detach().cpu())
    mag = so.subm_conv(sfe.abs(), sidx, conv.weight.detach().cpu().abs(), conv.bias.detach().cpu().abs())
    _check_close(y.features, ref, mag, 1e-5, "forward")
    # gradients through the dense oracle by autograd (fp64, un-rounded dout; kernels round dout / inputs to bf16)
    w = torch.randn(n, Co, generator=g)
    (y.features * w.cuda()).sum().backward()
    fr = sfe.double().requires_grad_(True)
    wr = conv.weight.detach().cpu().double().requires_grad_(True)
    dense = so.subm_conv(fr, sidx, wr, None, round_bf16=False)
    (dense * w.double()).sum().backward()
    gx = torch.autograd.grad((y.features * 0).sum(), [], allow_unused=True) if False else None
    # dX: gradient w.r.t. the SORTED features (x.features is the sorted view of the leaf)
    gin = torch.autograd.grad(conv(x).features, xf, w.cuda())[0]
    s = float(fr.grad.abs().max())
    assert float((gin.double().cpu() - fr.grad).abs().max()) <= 2e-2 * s + 1e-6
    sw = float(wr.grad.abs().max())
    assert float((conv.weight.grad.double().cpu() - wr.grad).abs().max()) <= 2e-2 * sw + 1e-6
    assert float((conv.bias.grad.cpu().double() - w.double().sum(0)).abs().max()) <= 1e-3 * float(w.abs().sum(0).max())


def test_down_and_inverse_conv_match_dense_oracle():
    from oracle import sparse_oracle as so
    from unipre3d_b200.sparse import SparseConv3d, SparseConvTensor, SparseInverseConv3d
    n, grid, B, Ci, Cm = 1500, 14, 2, 32, 64
    idx = _voxels(n, grid, B, seed=5)
    g = torch.Generator().manual_seed(6)
    feats = torch.randn(n, Ci, generator=g)
    down = SparseConv3d(Ci, Cm, 2, stride=2, bias=False, indice_key="s1").cuda()
    up = SparseInverseConv3d(Cm, Ci, 2, bias=False, indice_key="s1").cuda()
    x = SparseConvTensor(feats.cuda(), idx.cuda())
    xf = x.features.detach().requires_grad_(True)
    x = x.replace_feature(xf)
    y = down(x)
    cref_idx, cref = so.down_conv(xf.detach().cpu(), x.indices.cpu(), down.weight.detach().cpu())
    assert torch.equal(y.indices.cpu().long(), cref_idx), "coarse voxel set / order differs"
    mag = so.down_conv(xf.detach().cpu().abs(), x.indices.cpu(), down.weight.detach().cpu().abs())[1]
    _check_close(y.features, cref, mag, 1e-5, "down forward")
    z = up(y)
    assert torch.equal(z.indices, x.indices)
    zref = so.inverse_conv(y.features.detach().cpu(), y.indices.cpu(), x.indices.cpu(), up.weight.detach().cpu())
    mag = so.inverse_conv(y.features.detach().cpu().abs(), y.indices.cpu(), x.indices.cpu(), up.weight.detach().cpu().abs())
    _check_close(z.features, zref, mag, 1e-5, "inverse forward")
    # gradients of the pair through autograd vs the dense formulation (conv3d stride 2 / explicit inverse) in fp64
    w = torch.randn(n, Ci, generator=g)
    (z.features * w.cuda()).sum().backward()
    fr = xf.detach().cpu().double().requires_grad_(True)
    wd = down.weight.detach().cpu().double().requires_grad_(True)
    wu = up.weight.detach().cpu().double().requires_grad_(True)
    cidx, c = so.down_conv(fr, x.indices.cpu(), wd, round_bf16=False)
    fi = x.indices.cpu().long()
    par = {tuple(r.tolist()): k for k, r in enumerate(cidx)}
    rows = torch.tensor([par[(int(r[0]), int(r[1]) // 2, int(r[2]) // 2, int(r[3]) // 2)] for r in fi])
    wsel = wu[:, fi[:, 1] & 1, fi[:, 2] & 1, fi[:, 3] & 1, :]                       # (Co, n, Cm)
    zr = torch.einsum("onc,nc->no", wsel, c[rows])
    (zr * w.double()).sum().backward()
    for got, ref, name in ((xf.grad, fr.grad, "dX"), (down.weight.grad, wd.grad, "dW down"), (up.weight.grad, wu.grad, "dW up")):
        s = float(ref.abs().max())
        assert float((got.double().cpu() - ref).abs().max()) <= 3e-2 * s + 1e-6, name


def test_rulebook_is_cached_per_indice_key_and_empty_input():
    from unipre3d_b200.sparse import SparseConvTensor, SubMConv3d
    idx = _voxels(100, 6, 1, seed=1).cuda()
    x = SparseConvTensor(torch.randn(100, 16).cuda(), idx)
    a, b = SubMConv3d(16, 16, 3, indice_key="k").cuda(), SubMConv3d(16, 32, 3, indice_key="k").cuda()
    y = b(a(x))
    assert len([k for k in x.rules if k.startswith("subm3")]) == 1 and y.features.shape == (100, 32)
    with pytest.raises(ValueError, match="duplicate"):
        SparseConvTensor(torch.zeros(2, 16).cuda(), torch.zeros(2, 4, dtype=torch.int32).cuda())


# ---------------------------------------------------------------------------------------------------------------------
# Whole SpUNet (module mirror of pointcept/models/sparse_unet/spconv_unet_v1m1_base.py) on a ~5k-voxel scene against the
# same module tree executed with the dense-grid oracle convolutions on the CPU (fp64 sums, the same bf16 operand rounding
# at every convolution input).  A ~1e-7 accumulation difference in front of a bf16 rounding can flip that rounding
# (2^-9 relative), and the network stacks ~50 convolutions with train-mode BatchNorm between them: outputs must agree to
# 2e-2 of the output scale (measured ~3e-3).
def _oracle_run(mod, idx, feat, store):
    from oracle import sparse_oracle as so
    from unipre3d_b200 import sparse as sp
    from unipre3d_b200.sparse_unet import BasicBlock
    if isinstance(mod, BasicBlock):
        res_idx, res = idx, feat
        _, out = _oracle_run(mod.conv1, idx, feat, store)
        out = torch.relu(_oracle_run(mod.bn1, idx, out, store)[1])
        _, out = _oracle_run(mod.conv2, idx, out, store)
        out = _oracle_run(mod.bn2, idx, out, store)[1]
        return idx, torch.relu(out + _oracle_run(mod.proj, res_idx, res, store)[1])
    if isinstance(mod, sp.SparseSequential):
        for m in mod:
            idx, feat = _oracle_run(m, idx, feat, store)
        return idx, feat
    if isinstance(mod, sp.SubMConv3d):
        b = None if mod.bias is None else mod.bias.detach().cpu()
        return idx, so.subm_conv(feat, idx, mod.weight.detach().cpu(), b)
    if isinstance(mod, sp.SparseConv3d):
        cidx, out = so.down_conv(feat, idx, mod.weight.detach().cpu())
        store[mod.indice_key] = (idx, cidx)
        return cidx.int(), out
    if isinstance(mod, sp.SparseInverseConv3d):
        fine, cidx = store[mod.indice_key]
        return fine, so.inverse_conv(feat, cidx, fine, mod.weight.detach().cpu())
    if isinstance(mod, torch.nn.BatchNorm1d):                       # train mode: batch statistics
        f = feat.double()
        mu, var = f.mean(0), f.var(0, unbiased=False)
        return idx, (f - mu) / torch.sqrt(var + mod.eps) * mod.weight.detach().cpu().double() + mod.bias.detach().cpu().double()
    if isinstance(mod, torch.nn.ReLU):
        return idx, torch.relu(feat)
    if isinstance(mod, torch.nn.Identity):
        return idx, feat
    raise TypeError(type(mod))


def test_spunet_forward_matches_oracle_and_trains():
    from types import SimpleNamespace as NS
    from unipre3d_b200.sparse import pack_keys
    from unipre3d_b200.sparse_unet import SpUNetBase
    torch.manual_seed(0)
    n, grid = 5000, 40
    idx = _voxels(n, grid, 1, seed=11)
    # a surface-like occupancy (scenes are 2-D manifolds): keep voxels near two planes
    idx[:, 3] = (idx[:, 3] % 6) + (idx[:, 1] // 8)
    idx = torch.unique(idx, dim=0)
    n = idx.shape[0]
    feat = torch.randn(n, 6)
    net = SpUNetBase(6, 64, cfg=NS(opt=NS(use_fusion=False)), channels=(32, 64, 128, 256, 256, 128, 96, 96),
                     layers=(1, 1, 1, 1, 1, 1, 1, 1)).cuda().train()
    with torch.no_grad():
        for m in net.modules():                                     # non-trivial BatchNorm affine parameters
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.2)
    inp = {"grid_coord": idx[:, 1:].cuda(), "feat": feat.cuda(), "offset": torch.tensor([n]).cuda()}
    out = net(inp)
    assert out.features.shape == (n, 64) and torch.isfinite(out.features).all()
    # oracle on the sorted voxel list
    order = torch.argsort(pack_keys(idx))
    sidx, sfeat = idx[order], feat[order]
    assert torch.equal(out.indices.cpu(), sidx.int())
    store = {}
    i, f = _oracle_run(net.conv_input, sidx, sfeat, store)
    skips = [(i, f)]
    for s in range(net.num_stages):
        i, f = _oracle_run(net.down[s], i, f, store)
        i, f = _oracle_run(net.enc[s], i, f, store)
        skips.append((i, f))
    i, f = skips.pop(-1)
    for s in reversed(range(net.num_stages)):
        i, f = _oracle_run(net.up[s], i, f, store)
        si, sf = skips.pop(-1)
        assert torch.equal(torch.as_tensor(i).long(), torch.as_tensor(si).long())
        f = torch.cat([f, sf.double()], 1)
        i, f = _oracle_run(net.dec[s], i, f, store)
    _, ref = _oracle_run(net.final, i, f, store)
    err = float((out.features.double().cpu() - ref).abs().max())
    assert err <= 2e-2 * float(ref.abs().max()), (err, float(ref.abs().max()))
    # input order restored
    back = out.features_in_input_order()
    assert torch.equal(back[order.cuda()], out.features)
    # and the network trains: a few SGD steps on a fixed target reduce the loss
    target = torch.randn(n, 64).cuda()
    opt = torch.optim.SGD(net.parameters(), lr=0.05)
    losses = []
    for _ in range(4):
        opt.zero_grad()
        loss = ((net(inp).features - target) ** 2).mean()
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]
