"""N-rank probe of the CUDA-graph step (prints stage markers; dumps Python stacks if it hangs)."""
import datetime, faulthandler, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
def log(*a):
    print(f"[r{rank} {time.strftime('%X')}]", *a, flush=True)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=60))
from unipre3d_b200 import synthetic
from unipre3d_b200.config import compose
from unipre3d_b200.trainer import Trainer
world = dist.get_world_size()
cfg = compose(overrides=["data.training_resolution=64", f"opt.batch_size={2*world}", "general.device=[" + ",".join(map(str, range(world))) + "]"])
data = synthetic.make_batch(cfg, 2, 1024, seed=rank, pin=True)
for graph in (False, True):
    faulthandler.dump_traceback_later(45, exit=True)
    tr = Trainer(cfg, use_cuda_graph=graph, autocast_dtype=torch.bfloat16)
    log("trainer built, graph =", graph)
    for i in range(6):
        l = tr.train_iteration(data)
        log("step", i, "loss", l)
    dist.barrier()
    faulthandler.cancel_dump_traceback_later()
    log("mode done", graph)
    del tr
log("done")
dist.destroy_process_group()
