"""The oracle itself: hand-derived backward vs fp64 autograd of the forward restatement (Appendix A.9's route),
binning vs an independent numpy lexsort, point-op oracles vs brute-force numpy, and the closed-form statement of
the reference's FPS tie-break that the CUDA kernel implements."""
import numpy as np
import pytest
import torch

from tests.helpers import make_camera, make_gaussians, oracle_scene


@pytest.mark.parametrize("regime,P,W,H,deg,M", [("reference", 40, 48, 32, 1, 4), ("mid", 60, 40, 40, 3, 16),
                                                 ("small", 200, 32, 32, 2, 9), ("mid", 50, 32, 32, 0, 1)])
def test_c_oracle_backward_matches_fp64_autograd(oracle, regime, P, W, H, deg, M):
    from oracle import raster_ref as rr
    g = make_gaussians(P, seed=1, regime=regime, sh_coeffs=M)
    c = make_camera(az=40, el=25)
    bg = (0.2, 0.5, 0.7)
    sc = oracle_scene(g, c, W, H, sh_degree=deg, bg=bg)
    geo = oracle.preprocess(sc)
    keys, pl, ranges = oracle.bin_tiles(sc, geo)
    dL = np.random.default_rng(5).normal(size=(3, H, W)).astype(np.float32)
    res = oracle.render(sc, dL)
    t = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in g.items()}
    img = rr.render(t["means3D"], t["opacities"], t["scales"], t["rotations"], t["shs"], None,
                    torch.tensor(c["view"]).reshape(-1), torch.tensor(c["proj"]).reshape(-1), torch.tensor(c["campos"]),
                    torch.tensor(bg), W, H, c["tanfovx"], c["tanfovy"], deg, pl, ranges)
    assert np.abs(img.detach().numpy() - res["color"]).max() < 1e-5
    (img * torch.tensor(dL, dtype=torch.float64)).sum().backward()
    for k in ["means3D", "opacities", "scales", "rotations", "shs"]:
        a = t[k].grad.numpy()
        b = res["grads"][k].reshape(a.shape)
        assert np.abs(a - b).max() <= 5e-5 * np.abs(a).max() + 1e-7, k


def test_binning_equals_independent_lexsort(oracle):
    P, W, H = 700, 100, 70
    g = make_gaussians(P, seed=2, regime="small")
    g["means3D"][1::2] = g["means3D"][0::2]              # equal depths -> ties
    sc = oracle_scene(g, make_camera(az=12, el=33), W, H)
    geo = oracle.preprocess(sc)
    keys, pl, ranges = oracle.bin_tiles(sc, geo)
    gx = (W + 15) // 16
    tiles, ids = [], []
    for i in range(P):
        if geo.radii[i] > 0:
            x0, y0, x1, y1 = geo.rect[i]
            for y in range(y0, y1):
                for x in range(x0, x1):
                    tiles.append(y * gx + x); ids.append(i)
    tiles, ids = np.array(tiles), np.array(ids)
    order = np.lexsort((ids, geo.depth[ids].view(np.uint32), tiles))
    assert np.array_equal(pl, ids[order].astype(np.uint32))
    assert len(pl) == int(geo.tiles_touched.sum())
    for t in range(ranges.shape[0]):
        seg = tiles[order][ranges[t, 0]:ranges[t, 1]]
        assert (seg == t).all()
    assert (np.diff(keys.astype(np.uint64)) >= 0).all()


def test_empty_and_culled_inputs(oracle):
    g = make_gaussians(8, seed=3, regime="mid")
    g["means3D"][:] = [0, 0, -50]
    sc = oracle_scene(g, make_camera(az=0, el=0), 32, 32, bg=(0.1, 0.2, 0.3))
    r = oracle.render(sc, np.ones((3, 32, 32), np.float32))
    assert r["num_rendered"] == 0 or (r["radii"] >= 0).all()
    if (r["radii"] > 0).sum() == 0:
        assert np.allclose(r["color"], np.array([0.1, 0.2, 0.3], np.float32)[:, None, None])
        assert np.abs(r["grads"]["means3D"]).max() == 0


def _fps_numpy(x, m):
    idx = [0]
    d = np.full(x.shape[0], 1e10, np.float32)
    for _ in range(1, m):
        p = x[idx[-1]]
        dx, dy, dz = (x[:, 0] - p[0]).astype(np.float32), (x[:, 1] - p[1]).astype(np.float32), (x[:, 2] - p[2]).astype(np.float32)
        dist = (dz * dz + (dx * dx + dy * dy)).astype(np.float32)
        d = np.minimum(d, dist)
        idx.append(int(np.argmax(d)))
    return np.array(idx, np.int32)


def test_pointop_oracles_vs_numpy(oracle):
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 500, 3)).astype(np.float32)
    got = oracle.fps(x, 40)
    # without exact ties (random floats; fma vs separate rounding can differ only in ties) the plain argmax agrees
    for b in range(2):
        ref = _fps_numpy(x[b], 40)
        assert (got[b] == ref).mean() > 0.9 and got[b][0] == 0
    ctr = np.take_along_axis(x, got.astype(np.int64)[..., None], 1)
    idx = oracle.ball_query(0.6, 8, x, ctr)
    d2 = ((ctr[:, :, None] - x[:, None]) ** 2).sum(-1)
    for b in range(2):
        for m in range(40):
            hits = np.nonzero(d2[b, m] < 0.36 - 1e-5)[0]
            k = min(len(hits), 8)
            assert set(hits[:k]) <= set(idx[b, m]) or k == 0
            assert (idx[b, m][k:] == idx[b, m][0]).all() or len(np.nonzero(np.abs(d2[b, m] - 0.36) < 1e-4)[0]) > 0
    f = rng.normal(size=(2, 4, 500)).astype(np.float32)
    grp = oracle.group(f, idx)
    assert np.array_equal(grp, np.stack([f[b][:, idx[b]] for b in range(2)]))
    gg = oracle.group_grad(np.ones_like(grp), idx, 500)
    cnt = np.stack([np.bincount(idx[b].ravel(), minlength=500) for b in range(2)])
    assert np.allclose(gg, cnt[:, None, :].repeat(4, 1))
    # empty query ball -> zeros (group.py:194 zero-initialised buffer)
    far = np.full((2, 3, 3), 100.0, np.float32)
    assert (oracle.ball_query(0.1, 4, x, far) == 0).all()


def _bitrev(v, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (v & 1); v >>= 1
    return r


@pytest.mark.parametrize("n", [33, 64, 100, 777, 1024, 3000])
def test_fps_tie_break_closed_form(oracle, n):
    """The CUDA kernel uses: among equal distances, smallest (bitreverse_{log2 bs}(k mod bs), k div bs) wins, with
    bs = opt_n_threads(n).  Check that closed form against the literal replay of the reference's thread loop + tree
    reduction (the oracle) on clouds FULL of exact ties (integer lattice / duplicates)."""
    rng = np.random.default_rng(n)
    side = 5
    x = rng.integers(0, side, (1, n, 3)).astype(np.float32)      # many duplicates and equal distances
    m = min(n, 24)
    got = oracle.fps(x, m)[0]
    bs = oracle.fps_block_size(n)
    bits = bs.bit_length() - 1
    pts = x[0]
    d = np.full(n, 1e10, np.float32)
    idx = [0]
    ks = np.arange(n)
    prio = np.array([(_bitrev(int(k) % bs, bits) << 20) | (int(k) // bs) for k in ks])
    for _ in range(1, m):
        p = pts[idx[-1]]
        diff = pts - p
        dist = np.float32(diff[:, 2] * diff[:, 2]) + (np.float32(diff[:, 0] * diff[:, 0]) + np.float32(diff[:, 1] * diff[:, 1]))
        d = np.minimum(d, dist.astype(np.float32))
        best = d.max()
        cand = ks[d == best]
        idx.append(int(cand[np.argmin(prio[cand])]))
    assert np.array_equal(got, np.array(idx, np.int32))
