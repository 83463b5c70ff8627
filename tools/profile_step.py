"""One eager training step under the CUDA profiler range (for ncu --profile-from-start off) or torch.profiler."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unipre3d_b200 import synthetic
from unipre3d_b200.config import compose
from unipre3d_b200.trainer import Trainer
mode = sys.argv[1] if len(sys.argv) > 1 else "torch"
res, bs, npts = 256, 8, 8192
cfg = compose(overrides=[f"data.training_resolution={res}", f"opt.batch_size={bs}"])
data = synthetic.make_batch(cfg, bs, npts, seed=0, pin=True)
tr = Trainer(cfg, use_cuda_graph=False, autocast_dtype=torch.bfloat16 if "bf16" in sys.argv else None)
for _ in range(4):
    tr.train_iteration(data)
torch.cuda.synchronize()
if mode == "ncu":
    torch.cuda.profiler.start()
    tr.train_iteration(data)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            tr.train_iteration(data)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=int(os.environ.get('ROWS', '45')), max_name_column_width=70))
