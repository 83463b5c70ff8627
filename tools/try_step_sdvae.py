"""Pre-training step with the SD-VAE image branch (model.image_branch=sdvae, random-init weights of that architecture):
a small eager step (finite loss) and the timed full-size step (8 objects, 8192 points, 256^2, CUDA graph, bf16)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unipre3d_b200 import synthetic  # noqa: E402
from unipre3d_b200.config import compose  # noqa: E402
from unipre3d_b200.trainer import Trainer  # noqa: E402

cfg = compose(overrides=["model.image_branch=sdvae", "data.training_resolution=64", "opt.batch_size=2", "opt.ema.use=false"])
tr = Trainer(cfg, use_cuda_graph=False, autocast_dtype=torch.bfloat16)
loss = [tr.train_iteration(synthetic.make_batch(cfg, 2, 1024, seed=0, pin=True, image_dtype="uint8")) for _ in range(3)]
print("small eager steps, loss:", [round(x, 5) for x in loss], flush=True)
assert all(x == x and abs(x) < 1e6 for x in loss)
del tr
torch.cuda.empty_cache()

for branch in ("sdvae", "stem"):
    cfg = compose(overrides=[f"model.image_branch={branch}", "data.training_resolution=256", "opt.batch_size=8", "opt.ema.use=false"])
    tr = Trainer(cfg, use_cuda_graph=True, autocast_dtype=torch.bfloat16)
    data = synthetic.make_batch(cfg, 8, 8192, seed=1, pin=True, image_dtype="uint8")
    for _ in range(5):
        tr.train_iteration(data)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        tr.replay_resident()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    print(f"image_branch={branch}: {ms:.3f} ms/step -> {32 / ms * 1e3:.0f} views/s (graph replays, resident inputs)", flush=True)
    del tr
    torch.cuda.empty_cache()
