"""Micro-benchmark of the attention kernels at the transformer config (B=8, L=129, H=6, D=64) against the library
SDPA path (cuDNN flash fwd + autograd bwd).  CUDA events on the launching stream."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unipre3d_b200 import fused_encoder as fe  # noqa: E402

B, L, H, D = 8, 129, 6, 64
C = H * D
torch.manual_seed(0)
qkv = torch.randn(B * L, 3 * C, device="cuda").to(torch.bfloat16)
do = torch.randn(B * L, C, device="cuda").to(torch.bfloat16)
scale = D ** -0.5


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


o, lse = fe.attn_fwd(qkv, B, L, H, D, scale)
print(f"own fwd  {timeit(lambda: fe.attn_fwd(qkv, B, L, H, D, scale)):.2f} us")
print(f"own bwd  {timeit(lambda: fe.attn_bwd(qkv, o, lse, do, B, L, H, D, scale)):.2f} us")
ql = qkv.clone().requires_grad_(True)


def lib_fwd():
    q, k, v = ql.view(B, L, 3, H, D).permute(2, 0, 3, 1, 4).unbind(0)
    return torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=scale).transpose(1, 2).reshape(B * L, C)


ol = lib_fwd()
print(f"lib fwd  {timeit(lib_fwd):.2f} us")
print(f"lib bwd  {timeit(lambda: torch.autograd.grad(ol, ql, do, retain_graph=True)):.2f} us (incl. stack/copy)")
