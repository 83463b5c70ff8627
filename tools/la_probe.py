"""Why does the L2-flushed per-step timing of bench.py differ between lookahead grouping on / off?  Times replay_resident()
four ways (flush + per-step events; the same with a host sync per step; one event pair around the whole flushed loop; no
flush) and, per step, how long the main stream waits in _lookahead_take."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from unipre3d_b200 import synthetic
from unipre3d_b200.trainer import Trainer


def run(lookahead: bool, tc: bool = True, tc_all: bool = False):
    dev = torch.device("cuda", 0)
    os.environ["UP3D_TC_LINEAR"] = "1" if tc else "0"
    from unipre3d_b200 import fused_encoder as fe
    fe._TC_ALL = tc_all
    if lookahead:
        os.environ.pop("UP3D_NO_LOOKAHEAD", None)
    else:
        os.environ["UP3D_NO_LOOKAHEAD"] = "1"
    bench.select_config("transformer")
    cfg = bench.make_cfg(1)
    tr = Trainer(cfg, device=dev, use_cuda_graph=True, autocast_dtype=torch.bfloat16)
    pb = tr.pack_batch(synthetic.make_batch(cfg, bench.OBJECTS_PER_GPU, bench.N_POINTS, seed=0, pin=False, image_dtype="uint8"))
    for _ in range(5):
        tr.train_iteration(pb)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    K = 20
    E = lambda: torch.cuda.Event(enable_timing=True)
    out = []
    for rep in range(4):
        ev = [(E(), E()) for _ in range(K)]
        torch.cuda.synchronize()
        for a, b in ev:
            flush.zero_()
            a.record()
            tr.replay_resident()
            b.record()
        torch.cuda.synchronize()
        out.append(sum(a.elapsed_time(b) for a, b in ev) / K)
    tag = ("lookahead" if tr._la is not None else "inline   ") + (" tcgen05-all" if (tc and tc_all) else " tcgen05-fc1" if tc else " library-fc1")
    print(tag, "flush + per-step events, 4 x 20 steps:", " ".join(f"{x:.4f}" for x in out), "loss", float(tr._loss_buf), flush=True)
    tr._graph = None
    del tr


def main():
    for la, tc, ta in ((True, True, False), (True, True, False), (False, True, False)):
        run(la, tc, ta)


if __name__ == "__main__":
    main()
