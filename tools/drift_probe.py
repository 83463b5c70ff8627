"""Why does the step slow down as training on the synthetic batch proceeds?  Rasterizer kernel times (C-ABI event pairs)
and the whole step after 5 and after 85 optimisation steps on the same batch."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from unipre3d_b200 import _lib, synthetic
from unipre3d_b200.trainer import Trainer

NAMES = ["project", "depth_sort", "blend_forward", "grad_clear", "blend_backward", "geometry_backward"]


def raster_ms(tr, dev_batch, reps=3):
    ms = np.zeros((reps, 6))
    _lib.check(_lib.lib.up3d_raster_timing_enable(1))
    buf = (C.c_float * 6)()
    try:
        for r in range(reps):
            torch.cuda._sleep(100_000_000)
            tr._forward_backward(dev_batch)
            _lib.check(_lib.lib.up3d_raster_timing_read(buf))
            ms[r] = list(buf)
            for p in tr.params:
                p.grad = None
    finally:
        _lib.check(_lib.lib.up3d_raster_timing_enable(0))
    return dict(zip(NAMES, (1e3 * ms.mean(0)).round(1)))


def step_ms(tr, flush, K=10):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for a, b in ev:
        flush.zero_(); a.record(); tr.replay_resident(); b.record()
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in ev) / K


def main():
    dev = torch.device("cuda", 0)
    bench.select_config("transformer")
    cfg = bench.make_cfg(1)
    tr = Trainer(cfg, device=dev, use_cuda_graph=True, autocast_dtype=torch.bfloat16)
    pb = tr.pack_batch(synthetic.make_batch(cfg, bench.OBJECTS_PER_GPU, bench.N_POINTS, seed=0, pin=False, image_dtype="uint8"))
    for _ in range(5):
        tr.train_iteration(pb)
    tr.replay_resident()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    dev_batch = pb.views(pb.flat.to(dev))
    for stage in ("after ~16 steps", "after ~96 steps"):
        t = step_ms(tr, flush)
        print(stage, f"step {t:.4f} ms; raster kernels (us):", raster_ms(tr, dev_batch), flush=True)
        for _ in range(70):
            tr.replay_resident()
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
