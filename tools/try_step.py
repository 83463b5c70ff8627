import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unipre3d_b200 import synthetic
from unipre3d_b200.config import compose
from unipre3d_b200.trainer import Trainer
res, bs, npts = 256, 8, 8192
cfg = compose(overrides=[f"data.training_resolution={res}", f"opt.batch_size={bs}"])
data = synthetic.make_batch(cfg, bs, npts, seed=0, pin=True)
for mode, ac in ((False, None), (True, None), (True, torch.bfloat16)):
    tr = Trainer(cfg, use_cuda_graph=mode, autocast_dtype=ac)
    for _ in range(5):
        l = tr.train_iteration(data)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        l = tr.train_iteration(data)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"graph={mode} autocast={ac}: {ms:.3f} ms/step (wall {(time.perf_counter()-t0)/n*1e3:.3f}) -> {bs*4/ms*1e3:.0f} views/s loss {l:.5f}", flush=True)
    del tr
