"""Where does the end-to-end step lose time against the resident-input graph replay?  Variants of Trainer.train_iteration."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unipre3d_b200 import synthetic  # noqa: E402
from unipre3d_b200.config import compose  # noqa: E402
from unipre3d_b200.trainer import Trainer  # noqa: E402

cfg = compose(overrides=["data.training_resolution=256", "opt.batch_size=8"])
tr = Trainer(cfg, use_cuda_graph=True, autocast_dtype=torch.bfloat16)
batches = [synthetic.make_batch(cfg, 8, 8192, seed=i, pin=True, image_dtype="uint8") for i in range(4)]
for i in range(6):
    tr.train_iteration(batches[i % 4])
torch.cuda.synchronize()


def run(name, fn, n=40):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    print(f"{name:40s} {(time.perf_counter() - t0) / n * 1e3:.3f} ms/step (wall)", flush=True)


run("graph replay only", lambda i: tr.replay_resident())
run("lagged loss + prefetch", lambda i: tr.train_iteration(batches[i % 4], read_loss="lagged", prefetch=batches[(i + 1) % 4]))
run("no loss read + prefetch", lambda i: tr.train_iteration(batches[i % 4], read_loss=False, prefetch=batches[(i + 1) % 4]))
run("lagged loss, no prefetch", lambda i: tr.train_iteration(batches[i % 4], read_loss="lagged"))
run("blocking loss + prefetch", lambda i: tr.train_iteration(batches[i % 4], read_loss=True, prefetch=batches[(i + 1) % 4]))
t0 = time.perf_counter()
for i in range(40):
    tr._copy_into_static(tr._take_staged(batches[i % 4]))
torch.cuda.synchronize()
print(f"take_staged + copy_into_static alone     {(time.perf_counter() - t0) / 40 * 1e3:.3f} ms/step (wall)")

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timed(name, step_fn, n=20):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        flush.zero_()
        ev[i][0].record()
        step_fn(i)
        ev[i][1].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n * 1e3
    per = [a.elapsed_time(b) for a, b in ev]
    print(f"{name:40s} events {sum(per) / n:.3f} ms/step  wall {wall:.3f}  min {min(per):.3f} max {max(per):.3f}", flush=True)


timed("flush + graph replay", lambda i: tr.replay_resident())
timed("flush + e2e lagged/prefetch", lambda i: tr.train_iteration(batches[i % 4], read_loss="lagged", prefetch=batches[(i + 1) % 4]))
timed("flush + e2e no-loss/prefetch", lambda i: tr.train_iteration(batches[i % 4], read_loss=False, prefetch=batches[(i + 1) % 4]))
