"""N > 1 host logic on CPU: two gloo ranks, objects sharded over ranks (shard_batch), gradients averaged by DDP
== the single-process gradient of the global batch.  Uses the CPU port (oracle tokenizer + oracle rasterizer), so it
exercises exactly the sharding / all-reduce plumbing the GPU trainer uses with NCCL."""
import os
import socket
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _cfg():
    from unipre3d_b200.config import compose
    return compose(overrides=["data.training_resolution=32", "opt.batch_size=2", "opt.imgs_per_obj=2"])


def _model(cfg):
    from oracle.cpu_step import build_cpu_model
    torch.manual_seed(0)
    m = build_cpu_model(cfg)
    m.train()
    for mod in m.modules():
        if mod.__class__.__name__ == "DropPath":
            mod.drop_prob = 0.0
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.eval()       # SyncBatchNorm has no CPU/gloo path; fixed statistics make shards == global batch
    return m


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.cpu_step import forward_loss
        from unipre3d_b200 import synthetic
        from unipre3d_b200.trainer import shard_batch
        cfg = _cfg()
        ddp = torch.nn.parallel.DistributedDataParallel(_model(cfg), find_unused_parameters=False)
        data = shard_batch(synthetic.make_batch(cfg, 2, 256, seed=0), rank, world)
        assert data["gt_images"].shape[0] == 1
        ni = int(cfg.data.input_images)
        # run through the DDP wrapper so its reducer hooks fire
        from oracle import cpu_step
        loss, _, _ = cpu_step.forward_loss(lambda *a: ddp(*a), cfg, data)
        loss.backward()
        grads = {k: p.grad.clone() for k, p in ddp.module.named_parameters() if p.grad is not None}
        losses = [torch.zeros(()) for _ in range(world)]
        dist.all_gather(losses, loss.detach())
        if rank == 0:
            torch.save({"grads": grads, "loss": float(torch.stack(losses).mean())}, os.path.join(out_dir, "ddp.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gloo_sharded_step_equals_global_batch():
    from oracle.cpu_step import forward_loss
    from unipre3d_b200 import synthetic
    cfg = _cfg()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker, args=(2, _free_port(), td), nprocs=2, join=True)
        got = torch.load(os.path.join(td, "ddp.pt"))
    model = _model(cfg)
    loss, _, _ = forward_loss(model, cfg, synthetic.make_batch(cfg, 2, 256, seed=0))
    loss.backward()
    assert abs(got["loss"] - float(loss)) <= 1e-6 * max(1.0, abs(float(loss)))
    gmax = max(float(p.grad.abs().max()) for p in model.parameters() if p.grad is not None)
    n = 0
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        d = float((got["grads"][k] - p.grad).abs().max())
        assert d <= 1e-4 * float(p.grad.abs().max()) + 1e-5 * gmax, (k, d)
        n += 1
    assert n > 100


def _sync_worker(rank, world, port, out_dir, overlap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from unipre3d_b200.grad_sync import GradSync
        torch.manual_seed(0)
        early = [torch.nn.Parameter(torch.zeros(s)) for s in ((7, 5), (11,), (3, 4, 2))]
        rest = [torch.nn.Parameter(torch.zeros(s)) for s in ((6,), (2, 9))]
        params = [rest[0]] + early + [rest[1]]                  # optimizer order differs from the flat-buffer order
        chunks = overlap == "chunks"
        direct = overlap == "direct"                            # the producer writes into the flat buffer (permuted layout)
        gs = GradSync(params, early, "cpu", overlap=bool(overlap), early_order=[2, 0, 1] if direct else None)
        if direct:
            assert [s[0] for s in gs._early_slices] == [24, 24 + 35, 0]
        for step in range(2):                                   # two steps: buffers and handles are reused
            g = torch.Generator().manual_seed(100 * step + rank)
            eg = [torch.randn(p.shape, generator=g) for p in early]
            rg = [torch.randn(p.shape, generator=g) for p in rest]
            if direct:
                views = gs._early_views()
                for v, t in zip(views, eg):
                    v.copy_(t)
                got = gs.early_hook(views)                      # nothing to pack: same memory
                assert all(a.data_ptr() == b.data_ptr() for a, b in zip(got, views))
            elif chunks:                                        # chunk-wise, last parameters first (backward order)
                gs.early_chunk_hook(2, eg[2:])
                gs.early_chunk_hook(0, eg[:2])
                got = eg
            else:
                got = gs.early_hook(eg)                         # fires in the middle of the "backward"
            assert direct or all(a is b for a, b in zip(got, eg))
            for p, t in zip(early, got):
                p.grad = t
            for p, t in zip(rest, rg):
                p.grad = t
            gs.finish()
            assert all(p.grad.data_ptr() >= gs.flat.data_ptr() for p in params)     # views of the flat buffer
        if rank == 0:
            torch.save([p.grad.clone() for p in early + rest], os.path.join(out_dir, f"sync_{overlap}.pt"))
    finally:
        dist.destroy_process_group()


def test_grad_sync_overlapped_all_reduce_sums_over_ranks():
    """grad_sync.GradSync (the N > 1 gradient exchange of the GPU trainer) over two gloo ranks: the early (in-backward)
    all-reduce on its own communicator + the tail all-reduce give the SUM over ranks, with and without the overlap."""
    world, shapes_e, shapes_r = 2, ((7, 5), (11,), (3, 4, 2)), ((6,), (2, 9))
    for overlap in (True, False, "chunks", "direct"):
        with tempfile.TemporaryDirectory() as td:
            mp.spawn(_sync_worker, args=(world, _free_port(), td, overlap), nprocs=world, join=True)
            got = torch.load(os.path.join(td, f"sync_{overlap}.pt"))
        want = None
        for rank in range(world):
            g = torch.Generator().manual_seed(100 * 1 + rank)   # the second step's gradients
            cur = [torch.randn(s, generator=g) for s in shapes_e] + [torch.randn(s, generator=g) for s in shapes_r]
            want = cur if want is None else [a + b for a, b in zip(want, cur)]
        for a, b in zip(got, want):
            assert torch.allclose(a, b, atol=1e-6)


def _ready_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from unipre3d_b200.grad_sync import GradSync
        early = [torch.nn.Parameter(torch.zeros(5, 3))]
        # rest: r0 complete before the hook (ready); r1 gets a second contribution after it (late although it has a
        # gradient at hook time); r2 only after the hook (late)
        rest = [torch.nn.Parameter(torch.zeros(s)) for s in ((4,), (2, 3), (7,))]
        gs = GradSync(rest[:1] + early + rest[1:], early, "cpu", overlap=True)
        out = []
        for step in range(3):
            g = torch.Generator().manual_seed(10 * step + rank)
            ge, g0, g1a, g1b, g2 = (torch.randn(s, generator=g) for s in ((5, 3), (4,), (2, 3), (2, 3), (7,)))
            rest[0].grad, rest[1].grad = g0.clone(), g1a.clone()
            gs.early_hook([ge])
            rest[1].grad = rest[1].grad + g1b
            rest[2].grad = g2.clone()
            early[0].grad = ge
            gs.finish()
            assert gs.calibrated and gs._ready_idx == [0] and gs._late_idx == [1, 2]
            out.append([p.grad.clone() for p in early + rest])
            for p in early + rest:
                p.grad = None
        if rank == 0:
            torch.save(out, os.path.join(out_dir, "ready.pt"))
    finally:
        dist.destroy_process_group()


def test_grad_sync_ready_parameters_ride_with_the_early_all_reduce():
    """Parameters whose gradients are final when the early hook fires are found by the calibration step and from then on
    all-reduced together with the early ones; a parameter that still accumulates afterwards stays late.  Sums are right
    on the calibration step and after it."""
    world = 2
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_ready_worker, args=(world, _free_port(), td), nprocs=world, join=True)
        got = torch.load(os.path.join(td, "ready.pt"))
    for step in range(3):
        want = None
        for rank in range(world):
            g = torch.Generator().manual_seed(10 * step + rank)
            ge, g0, g1a, g1b, g2 = (torch.randn(s, generator=g) for s in ((5, 3), (4,), (2, 3), (2, 3), (7,)))
            cur = [ge, g0, g1a + g1b, g2]
            want = cur if want is None else [a + b for a, b in zip(want, cur)]
        for a, b in zip(got[step], want):
            assert torch.allclose(a, b, atol=1e-6), step
