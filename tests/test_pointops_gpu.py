"""Point ops through the C ABI vs (a) the CPU oracle and (b) the reference's OWN CUDA kernels
(oracle/_ref/pointnet2_batch_cuda.so, built unmodified from /root/reference by oracle/build_ref.py).
Integer outputs must be bit-exact, including tie-breaking."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_ext():
    so = os.path.join(ROOT, "oracle", "_ref", "pointnet2_batch_cuda.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("pointnet2_batch_cuda", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _cloud(B, N, seed, dup=False, grid=False):
    rng = np.random.default_rng(seed)
    if grid:  # lattice -> many exactly-equal distances (tie-break stress)
        side = int(round(N ** (1 / 3))) + 1
        pts = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), -1).reshape(-1, 3)[:N].astype(np.float32) * 0.125
        x = np.stack([pts[rng.permutation(N)] for _ in range(B)])
    else:
        x = rng.normal(size=(B, N, 3)).astype(np.float32)
        x = x / np.linalg.norm(x, axis=-1, keepdims=True) * rng.uniform(0, 1, (B, N, 1)).astype(np.float32) ** (1 / 3) * 0.5
    if dup:
        x[:, N // 2:] = x[:, : N - N // 2]
    return np.ascontiguousarray(x.astype(np.float32))


FPS_CASES = [(2, 1024, 128, False, False), (3, 8192, 128, False, False), (2, 1000, 100, True, False),
             (2, 4096, 512, False, True), (1, 777, 64, False, True), (2, 16384, 64, False, False),
             (1, 20000, 50, False, False), (2, 33, 33, False, False), (1, 2048, 1024, True, False)]


@pytest.mark.parametrize("B,N,M,dup,grid", FPS_CASES)
def test_fps_matches_oracle_and_reference(oracle, B, N, M, dup, grid):
    from unipre3d_b200.pointops import furthest_point_sample
    x = _cloud(B, N, seed=N + M, dup=dup, grid=grid)
    xt = torch.tensor(x, device="cuda")
    got = furthest_point_sample(xt, M).cpu().numpy()
    exp = oracle.fps(x, M)
    assert got.dtype == np.int32 and np.array_equal(got, exp)
    ext = _ref_ext()
    if ext is not None:
        out = torch.empty(B, M, dtype=torch.int32, device="cuda")
        temp = torch.full((B, N), 1e10, dtype=torch.float32, device="cuda")
        ext.furthest_point_sampling_wrapper(B, N, M, xt, temp, out)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), exp), "oracle disagrees with the reference's own kernel"


@pytest.mark.parametrize("B,N,M,K,r", [(2, 1024, 128, 32, 0.1), (3, 8192, 128, 32, 0.1), (2, 500, 77, 16, 0.05),
                                        (1, 4096, 256, 64, 0.3), (2, 300, 40, 8, 1e-4)])
def test_ball_query_group_match_oracle_and_reference(oracle, B, N, M, K, r):
    from unipre3d_b200.pointops import ball_query, furthest_point_sample, grouping_operation, subsample_group
    x = _cloud(B, N, seed=N + K)
    xt = torch.tensor(x, device="cuda")
    fidx = furthest_point_sample(xt, M)
    centers = torch.gather(xt, 1, fidx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    idx = ball_query(r, K, xt, centers)
    exp_idx = oracle.ball_query(r, K, x, centers.cpu().numpy())
    assert np.array_equal(idx.cpu().numpy(), exp_idx)
    feats = xt.transpose(1, 2).contiguous()
    grouped = grouping_operation(feats, idx)
    assert np.array_equal(grouped.cpu().numpy(), oracle.group(feats.cpu().numpy(), exp_idx))
    # fused path == composition (group_embed.py:39-57 / group.py:235-255)
    neigh, center, f2, idx2 = subsample_group(xt, M, K, r, return_idx=True)
    assert torch.equal(f2, fidx) and torch.equal(idx2, idx) and torch.equal(center, centers)
    assert torch.equal(neigh, grouped - centers.transpose(1, 2).unsqueeze(-1))
    ext = _ref_ext()
    if ext is not None:
        ridx = torch.zeros(B, M, K, dtype=torch.int32, device="cuda")
        ext.ball_query_wrapper(B, N, M, r, K, centers, xt, ridx)
        rout = torch.empty(B, 3, M, K, device="cuda")
        ext.group_points_wrapper(B, 3, N, M, K, feats, ridx, rout)
        torch.cuda.synchronize()
        assert torch.equal(ridx, idx) and torch.equal(rout, grouped)


def test_grouping_and_gather_gradients(oracle):
    from unipre3d_b200.pointops import gather_operation, grouping_operation
    B, C, N, M, K = 2, 5, 300, 40, 8
    rng = np.random.default_rng(0)
    f = torch.tensor(rng.normal(size=(B, C, N)).astype(np.float32), device="cuda", requires_grad=True)
    idx = torch.tensor(rng.integers(0, N, (B, M, K)).astype(np.int32), device="cuda")
    go = rng.normal(size=(B, C, M, K)).astype(np.float32)
    out = grouping_operation(f, idx)
    out.backward(torch.tensor(go, device="cuda"))
    exp = oracle.group_grad(go, idx.cpu().numpy(), N)
    np.testing.assert_allclose(f.grad.cpu().numpy(), exp, rtol=1e-5, atol=1e-5)
    f.grad = None
    gidx = torch.tensor(rng.integers(0, N, (B, M)).astype(np.int32), device="cuda")
    g = gather_operation(f, gidx)
    assert np.array_equal(g.detach().cpu().numpy(), oracle.gather(f.detach().cpu().numpy(), gidx.cpu().numpy()))
    # the reference's only in-tree pin for this path: gather_operation == torch.gather (subsample.py:163-191)
    assert torch.equal(g, torch.gather(f, 2, gidx.long().unsqueeze(1).expand(-1, C, -1)))
    g.sum().backward()
    exp = np.zeros((B, C, N), np.float32)
    for b in range(B):
        np.add.at(exp[b].T, gidx[b].cpu().numpy(), 1.0)
    np.testing.assert_allclose(f.grad.cpu().numpy(), exp, rtol=0, atol=1e-6)


def test_pointops_reject_cpu_tensors():
    from unipre3d_b200.pointops import furthest_point_sample
    with pytest.raises(RuntimeError, match="CUDA device"):
        furthest_point_sample(torch.zeros(1, 64, 3), 8)


def test_focal_l2_loss_kernel_matches_torch_formula():
    from unipre3d_b200.loss import focal_l2_loss, focal_l2_loss_torch
    rng = np.random.default_rng(0)
    for bgc in [(0.0, 0.0, 0.0), (1.0, 1.0, 1.0)]:
        gt = rng.uniform(0, 1, (5, 3, 40, 56)).astype(np.float32)
        mask = rng.uniform(size=(5, 1, 40, 56)) < 0.5
        gt = np.where(mask, np.asarray(bgc, np.float32)[None, :, None, None], gt)
        r = torch.tensor(rng.uniform(0, 1, gt.shape).astype(np.float32), device="cuda", requires_grad=True)
        g = torch.tensor(gt, device="cuda")
        bg = torch.tensor(bgc, device="cuda")
        l1 = focal_l2_loss(r, g, bg, 4, 1)
        l1.backward()
        g1 = r.grad.clone(); r.grad = None
        l2 = focal_l2_loss_torch(r, g, bg, 4, 1)
        l2.backward()
        assert abs(float(l1) - float(l2)) <= 1e-6 * max(1.0, abs(float(l2)))
        assert float((g1 - r.grad).abs().max()) <= 1e-9 + 1e-6 * float(r.grad.abs().max())


def test_focal_l2_reads_uint8_view_in_place():
    """The fused loss on a (B,V',3,H,W) view-axis slice of 8-bit images == the float formula on slice/255."""
    import torch
    from unipre3d_b200.loss import focal_l2_loss, focal_l2_loss_torch
    torch.manual_seed(0)
    B, V, H, W = 3, 5, 24, 20
    gt8 = torch.randint(0, 256, (B, V, 3, H, W), dtype=torch.uint8, device="cuda")
    gt8[:, :, :, :7] = 0                                   # background rows (black)
    r = torch.rand(B * (V - 1), 3, H, W, device="cuda", requires_grad=True)
    bg = torch.zeros(3, device="cuda")
    view = gt8[:, 1:]
    l1 = focal_l2_loss(r, view, bg, 4, 1)
    (g1,) = torch.autograd.grad(l1, r)
    gtf = (view.float() / 255.0).reshape(-1, 3, H, W)
    l2 = focal_l2_loss_torch(r, gtf, bg, 4, 1)
    (g2,) = torch.autograd.grad(l2, r)
    assert abs(float(l1) - float(l2)) <= 1e-6 * abs(float(l2))
    assert torch.allclose(g1, g2, atol=1e-9, rtol=1e-5)
    l3 = focal_l2_loss(r, (gt8.float() / 255.0)[:, 1:], bg, 4, 1)       # float view, also in place
    assert abs(float(l3) - float(l2)) <= 1e-6 * abs(float(l2))
