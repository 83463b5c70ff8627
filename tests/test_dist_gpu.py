"""N = 2 on real GPUs (skipped below two devices): the data-parallel, CUDA-graph-captured step -- objects sharded over
ranks, SyncBatchNorm statistics merged across ranks inside the fused mini-PointNet, GradSync's overlapped NCCL
all-reduce, 1/world folded into the fused AdamW -- equals the single-GPU step on the global batch
(/root/reference/train_network.py:183-186 SyncBatchNorm + DistributedDataParallel semantics)."""
import os
import socket
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

STEPS = 3


def _cfg(bs):
    from unipre3d_b200.config import compose
    return compose(overrides=["data.training_resolution=64", f"opt.batch_size={bs}", "opt.ema.use=false"])


def _run_steps(tr, batches):
    """-> (losses, first moments after step 1).  AdamW(eps=1e-15) turns a gradient into ~lr*sign(g): parameters whose
    gradient is pure rounding noise (the conv biases in front of BatchNorm) move by +-lr at random, so the parameter
    values themselves are not comparable; Adam's first moment after step 1, (1-beta1) * clip * g, is linear in the
    all-reduced gradient and is."""
    for blk in tr.model_manager.model.modules():
        if blk.__class__.__name__ == "DropPath":
            blk.drop_prob = 0.0            # per-sample Bernoulli draws would differ between the two layouts
    losses = [tr.train_iteration(batches[0])]
    opt = tr.model_manager.optimizer
    m1 = [opt.state[p]["exp_avg"].detach().clone().cpu() for p in tr.params]
    bn = [b.detach().clone().cpu() for b in tr.model_manager.model.buffers() if b.dtype.is_floating_point]
    losses += [tr.train_iteration(b) for b in batches[1:]]
    return losses, m1, bn


def _worker(rank, world, port, out_dir, use_graph):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from unipre3d_b200 import synthetic
        from unipre3d_b200.trainer import Trainer, shard_batch
        cfg = _cfg(4)
        torch.manual_seed(0)
        # the graph-captured variant runs the benchmarked precision (bf16 GEMM operands, bench.py's configuration)
        tr = Trainer(cfg, use_cuda_graph=use_graph, autocast_dtype=torch.bfloat16 if use_graph else None)
        assert tr.world == 2 and tr.bs_per_gpu == 2
        batches = [shard_batch(synthetic.make_batch(cfg, 4, 1024, seed=10 + s), rank, world) for s in range(STEPS)]
        losses, m1, bn = _run_steps(tr, batches)
        gs = tr._grad_sync
        # the calibration step found the parameters downstream of the stack (their gradients leave with the early
        # all-reduce) and left the tokenizer in front of it for the tail
        assert gs.calibrated and len(gs._ready_idx) > 0 and len(gs._late_idx) > 0, (len(gs._ready_idx), len(gs._late_idx))
        t = torch.tensor(losses, device="cuda")
        dist.all_reduce(t)
        if rank == 0:
            torch.save({"m1": m1, "loss": (t / world).cpu(), "bn": bn},
                       os.path.join(out_dir, "dp.pt"))
        dist.barrier()
        torch.cuda.synchronize()
        if use_graph:
            # tearing the NCCL communicator down while CUDA graphs that captured its kernels are alive can hang (same
            # rule as bench.py): drop the graph and leave without the orderly teardown
            tr._graph = None
            os._exit(0)
    finally:
        if not use_graph:
            dist.destroy_process_group()


def _spawn(fn, args, nprocs, deadline_s=420):
    """mp.spawn with a deadline: a collective that never completes must fail this test, not hang the whole session."""
    import time
    import torch.multiprocessing as mp
    ctx = mp.spawn(fn, args=args, nprocs=nprocs, join=False)
    t0 = time.time()
    while not ctx.join(timeout=5):
        if time.time() - t0 > deadline_s:
            for p in ctx.processes:
                if p.is_alive():
                    p.kill()
            pytest.fail(f"data-parallel workers still running after {deadline_s} s")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("use_graph", [False, True])
def test_two_gpu_sharded_step_equals_single_gpu_global_batch(use_graph):
    from unipre3d_b200 import synthetic
    from unipre3d_b200.trainer import Trainer
    with tempfile.TemporaryDirectory() as td:
        _spawn(_worker, (2, _free_port(), td, use_graph), 2)
        got = torch.load(os.path.join(td, "dp.pt"))
    cfg = _cfg(4)
    torch.manual_seed(0)
    tr = Trainer(cfg, use_cuda_graph=False, autocast_dtype=torch.bfloat16 if use_graph else None)
    losses, m1, ref_bn = _run_steps(tr, [synthetic.make_batch(cfg, 4, 1024, seed=10 + s) for s in range(STEPS)])
    # mean of the per-rank losses == the global-batch loss (equal shard sizes, focal-L2 is a mean over pixels); steps 2..
    # see the noise-parameter drift described in _run_steps, hence the looser bound there
    # bf16 variant: identical roundings per object, but BatchNorm statistics / reductions are summed in another order,
    # which moves a few bf16 roundings (2^-9 each): 10x looser bounds
    k = 25.0 if use_graph else 1.0
    np.testing.assert_allclose(got["loss"].numpy()[0], np.float32(losses[0]), rtol=2e-3 if use_graph else 1e-5, atol=1e-7)
    np.testing.assert_allclose(got["loss"].numpy(), np.array(losses, np.float32), rtol=2e-2 if use_graph else 5e-3, atol=1e-6)
    gmax = max(float(m.abs().max()) for m in m1)
    worst = 0.0
    for a, b in zip(got["m1"], m1):
        worst = max(worst, float((a - b).abs().max()) / (float(b.abs().max()) + 1e-3 * gmax))
    # fp32: summation order only.  bf16: the two layouts merge BatchNorm statistics / reduce in another order, which moves
    # bf16 roundings; the bound is the gradient bound of the bf16 model in tests/test_backbone_gpu.py for the 16-block
    # stack, 2 * 4 * 2^-9 * sqrt(8 * 16) = 0.18 of the tensor scale (measured 0.03-0.06)
    g_tol = 2 * 4.0 * 2.0 ** -9 * (8.0 * 16) ** 0.5 if use_graph else 2e-3
    assert worst <= g_tol, f"all-reduced gradient deviates from the global-batch gradient: rel {worst}"
    # SyncBatchNorm running statistics after step 1 == BatchNorm statistics of the global batch
    assert len(ref_bn) == len(got["bn"]) and len(ref_bn) >= 4
    for a, b in zip(got["bn"], ref_bn):
        assert torch.allclose(a, b, rtol=k * 1e-4, atol=k * 1e-6)
