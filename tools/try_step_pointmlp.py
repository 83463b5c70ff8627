import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unipre3d_b200 import synthetic
from unipre3d_b200.config import compose
from unipre3d_b200.trainer import Trainer
res, bs, npts = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 4, 8192
cfg = compose("pointmlp_pretraining", overrides=[f"data.training_resolution={res}", f"opt.batch_size={bs}"])
data = synthetic.make_batch(cfg, bs, npts, seed=0, pin=True)
for mode, ac in ((False, torch.bfloat16), (True, torch.bfloat16)):
    tr = Trainer(cfg, use_cuda_graph=mode, autocast_dtype=ac)
    for _ in range(4):
        l = tr.train_iteration(data)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        l = tr.train_iteration(data)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"pointmlp B={bs} graph={mode}: {ms:.3f} ms/step -> {bs*4/ms*1e3:.0f} views/s loss {l:.5f} mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
    if mode is False:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            tr.train_iteration(data); torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=25, max_name_column_width=60))
    del tr
