"""Where does the end-to-end step (train_iteration with prefetch, pinned PackedBatch) spend time beyond the resident replay?
Per-step CUDA events around train_iteration, lookahead on / off, 3 x 20 steps each, L2 flush between steps as bench.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from unipre3d_b200 import synthetic
from unipre3d_b200.trainer import Trainer


def run(lookahead: bool, e2e_first: bool):
    dev = torch.device("cuda", 0)
    if lookahead:
        os.environ.pop("UP3D_NO_LOOKAHEAD", None)
    else:
        os.environ["UP3D_NO_LOOKAHEAD"] = "1"
    bench.select_config("transformer")
    cfg = bench.make_cfg(1)
    tr = Trainer(cfg, device=dev, use_cuda_graph=True, autocast_dtype=torch.bfloat16)
    pbs = [tr.pack_batch(synthetic.make_batch(cfg, bench.OBJECTS_PER_GPU, bench.N_POINTS, seed=i, pin=False, image_dtype="uint8"))
           for i in range(4)]
    for i in range(5):
        tr.train_iteration(pbs[i % 4])
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    K = 20
    E = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn):
        ev = [(E(), E()) for _ in range(K)]
        torch.cuda.synchronize()
        for i, (a, b) in enumerate(ev):
            flush.zero_()
            a.record()
            fn(i)
            b.record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev) / K

    def resident(i):
        tr.replay_resident()

    def e2e(i):
        tr.train_iteration(pbs[i % 4], read_loss="lagged", prefetch=pbs[(i + 1) % 4])

    order = [("e2e", e2e), ("res", resident)] if e2e_first else [("res", resident), ("e2e", e2e)]
    out = []
    for rep in range(2):
        for name, fn in order:
            if name == "res":
                tr.replay_resident()
            out.append(f"{name} {timed(fn):.4f}")
    print("lookahead" if tr._la is not None else "inline   ", "| ".join(out), flush=True)
    tr._graph = None


if __name__ == "__main__":
    for la in (True, False):
        for first in (False, True):
            run(la, first)
