"""A handful of eager up3d_tc_linear launches at the backbone's shapes, for `ncu --set full -k regex:gemm_kernel`:
fc1 + GELU epilogue, dX of fc2 + GELU-backward epilogue, qkv, proj, and the mini-PointNet's (32768 x 512) x (512 x 512)."""
import sys
import torch

sys.path.insert(0, ".")
from unipre3d_b200.tc_linear import B_NMAJOR, EPI_GELU, EPI_GELU_BWD, tc_linear  # noqa: E402

T, C, Hd = 1032, 384, 1536
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g, device="cuda").bfloat16()
y2, w1, b1 = r(T, C), r(Hd, C) * 0.05, r(Hd)
dd, w2 = r(T, C), r(C, Hd) * 0.05
wqkv, wproj = r(3 * C, C) * 0.05, r(C, C) * 0.05
big_a, big_w = r(32768, 512), r(512, 512) * 0.05
for _ in range(3):
    h, pre = tc_linear(y2, w1, b1, epilogue=EPI_GELU)
    dpre = tc_linear(dd, w2, None, b_major=B_NMAJOR, epilogue=EPI_GELU_BWD, aux_in=pre)
    qkv = tc_linear(y2, wqkv)
    a = tc_linear(y2, wproj, b1[:C].contiguous())
    big = tc_linear(big_a, big_w)
torch.cuda.synchronize()
print("ok", float(h.float().abs().mean()), float(dpre.float().abs().mean()))
