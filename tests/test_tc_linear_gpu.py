"""up3d_tc_linear (tcgen05 / TMEM / TMA GEMM, csrc/gemm_tc.cu) against a plain PyTorch fp32 reference of the same op
(nn.Linear of /root/reference/openpoints/models/backbone/transformer.py:22-33, 52-77 and its dX backward).

Tolerance: operands are the SAME bf16 values on both sides, products are exact in fp32 and both sides accumulate in
fp32, so the only differences are the summation order (|err| <= ~K * 2^-24 * sum|a||b|) and the single bf16 rounding
of the output (2^-9 relative) -- stated per test."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [  # (T, N, K): the transformer's Linear layers at 8 x 129 tokens, the mini-PointNet convs, ragged/small cases
    (1032, 1152, 384), (1032, 384, 384), (1032, 1536, 384), (1032, 384, 1536),
    (32768, 256, 128), (32768, 512, 512), (4096, 384, 512),
    (1, 64, 64), (127, 128, 72), (129, 192, 200), (300, 64, 8),
]


def _ops(T, N, K, b_major, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn((T, K), generator=g, device="cuda").bfloat16()
    w = (torch.randn((N, K), generator=g, device="cuda") / K ** 0.5).bfloat16()
    b = w if b_major == 0 else w.t().contiguous()
    bias = torch.randn(N, generator=g, device="cuda").bfloat16()
    return a, w, b, bias


@pytest.mark.parametrize("T,N,K", SHAPES)
@pytest.mark.parametrize("b_major", [0, 1])
def test_linear_matches_fp32_reference(T, N, K, b_major):
    from unipre3d_b200.tc_linear import tc_linear
    a, w, b, bias = _ops(T, N, K, b_major)
    ref = a.float() @ w.float().t() + bias.float()
    out32 = tc_linear(a, b, bias, b_major=b_major, out_dtype=torch.float32)
    scale = (a.float().abs() @ w.float().abs().t() + bias.float().abs())
    # fp32 accumulation-order bound
    assert ((out32 - ref).abs() <= 4e-6 * scale + 1e-6).all(), float((out32 - ref).abs().max())
    out16 = tc_linear(a, b, bias, b_major=b_major)
    assert out16.dtype == torch.bfloat16
    # one bf16 rounding on top (2^-8 covers the half-ulp plus the accumulation difference moving a rounding boundary)
    assert ((out16.float() - ref).abs() <= 2 ** -8 * ref.abs() + 4e-6 * scale + 1e-6).all()
    nobias = tc_linear(a, b, None, b_major=b_major, out_dtype=torch.float32)
    assert ((nobias - (ref - bias.float())).abs() <= 4e-6 * scale + 1e-6).all()


@pytest.mark.parametrize("tile_n", [32, 64, 96, 128])
def test_every_tile_width(tile_n):
    from unipre3d_b200.tc_linear import tc_linear
    T, N, K = 1032, 384, 384
    for b_major in (0, 1):
        if b_major == 1 and tile_n in (32, 96):
            continue
        a, w, b, bias = _ops(T, N, K, b_major, seed=tile_n)
        ref = a.float() @ w.float().t() + bias.float()
        out = tc_linear(a, b, bias, b_major=b_major, out_dtype=torch.float32, tile_n=tile_n)
        scale = a.float().abs() @ w.float().abs().t() + bias.float().abs()
        assert ((out - ref).abs() <= 4e-6 * scale + 1e-6).all()


@pytest.mark.parametrize("T,N,K", [(1032, 384, 1536), (1032, 384, 1152), (300, 128, 392), (32768, 256, 512)])
@pytest.mark.parametrize("ks", [2, 4])
@pytest.mark.parametrize("b_major", [0, 1])
def test_split_k_over_a_cluster(T, N, K, ks, b_major):
    """K range shared by a (1,1,ks) thread-block cluster, partial tiles reduced through distributed shared memory."""
    from unipre3d_b200.tc_linear import tc_linear
    a, w, b, bias = _ops(T, N, K, b_major, seed=ks)
    ref = a.float() @ w.float().t() + bias.float()
    scale = a.float().abs() @ w.float().abs().t() + bias.float().abs()
    for tn in (64, 128) if b_major else (32, 64, 96, 128):
        if N % tn:
            continue
        out = tc_linear(a, b, bias, b_major=b_major, out_dtype=torch.float32, tile_n=tn | (ks << 16))
        assert ((out - ref).abs() <= 4e-6 * scale + 1e-6).all(), (tn, float((out - ref).abs().max()))
        out16 = tc_linear(a, b, bias, b_major=b_major, tile_n=tn | (ks << 16))
        assert ((out16.float() - ref).abs() <= 2 ** -8 * ref.abs() + 4e-6 * scale + 1e-6).all()
    # the automatic choice (split-K for these deep-K shapes) is bitwise reproducible run to run
    o1 = tc_linear(a, b, bias, b_major=b_major, out_dtype=torch.float32)
    o2 = tc_linear(a, b, bias, b_major=b_major, out_dtype=torch.float32)
    assert torch.equal(o1, o2)
    assert ((o1 - ref).abs() <= 4e-6 * scale + 1e-6).all()


def test_gelu_epilogues():
    from unipre3d_b200.tc_linear import EPI_GELU, EPI_GELU_BWD, tc_linear
    T, N, K = 1032, 1536, 384
    a, w, b, bias = _ops(T, N, K, 0, seed=3)
    pre_ref = (a.float() @ w.float().t() + bias.float())
    h, pre = tc_linear(a, b, bias, epilogue=EPI_GELU)
    assert ((pre.float() - pre_ref).abs() <= 2 ** -8 * pre_ref.abs() + 1e-4).all()
    # GELU is evaluated on the bf16-rounded pre-activation the backward will read
    h_ref = torch.nn.functional.gelu(pre.float())
    assert ((h.float() - h_ref).abs() <= 2 ** -8 * h_ref.abs() + 1e-6).all()
    # backward: dpre = (dh = dd @ W2) * gelu'(pre), W2 (C, Hd) read as the (K,N) operand
    g = torch.Generator(device="cuda").manual_seed(5)
    dd = torch.randn((T, 384), generator=g, device="cuda").bfloat16()
    w2 = (torch.randn((384, N), generator=g, device="cuda") / 20).bfloat16()
    dpre = tc_linear(dd, w2, None, b_major=1, epilogue=EPI_GELU_BWD, aux_in=pre)
    x = pre.float().requires_grad_(True)
    torch.nn.functional.gelu(x).backward(dd.float() @ w2.float())
    ref = x.grad
    assert ((dpre.float() - ref).abs() <= 2 ** -7 * ref.abs() + 2e-4).all(), float((dpre.float() - ref).abs().max())


def test_argument_errors():
    from unipre3d_b200.tc_linear import tc_linear
    a = torch.zeros((8, 60), device="cuda").bfloat16()
    w = torch.zeros((64, 60), device="cuda").bfloat16()
    with pytest.raises(RuntimeError, match="multiple of 8"):
        tc_linear(a, w)
    with pytest.raises(RuntimeError, match="bfloat16"):
        tc_linear(a.float(), w)
    with pytest.raises(RuntimeError, match="CUDA device"):
        tc_linear(a.cpu(), w.cpu())
