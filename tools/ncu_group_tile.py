"""Workload for `ncu --set full -k regex:group_tile_kernel`: the fused mini-PointNet (bf16) forward + backward at the bench
step's sizes (8 objects x 128 groups x 32 neighbours), three times (profile the last pass: --launch-skip 16)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from unipre3d_b200.backbone import Encoder

torch.manual_seed(0)
enc = Encoder(384).cuda().train()
nb = torch.randn(8, 3, 128, 32, device="cuda") * 0.05
w = torch.randn(8, 128, 384, device="cuda")
for _ in range(3):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = enc.forward_grouped(nb)
    (out.float() * w).sum().backward()
    torch.cuda.synchronize()
print("ok")
