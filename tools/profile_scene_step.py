"""Kernel breakdown (torch profiler, CUDA time) of one scene-level step: python tools/profile_scene_step.py [ptv3|sparseunet]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from unipre3d_b200 import synthetic  # noqa: E402
from unipre3d_b200.config import compose  # noqa: E402
from unipre3d_b200.trainer import Trainer  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "sparseunet"
sc = bench.SCENE_CONFIGS[name]
cfg = compose(sc["name"], overrides=["opt.use_fusion=false", "data.input_images=0", f"opt.imgs_per_obj={sc['views']}",
                                     "data.training_width=512", "data.training_height=512", "opt.batch_size=1",
                                     f"model.max_sh_degree={sc['sh']}", "opt.ema.use=false"])
tr = Trainer(cfg, use_cuda_graph=False)
data = synthetic.make_scene_batch(cfg, 1, sc["points"], seed=0)
for _ in range(3):
    tr.train_iteration(data)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.train_iteration(data)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=28, max_name_column_width=70))
