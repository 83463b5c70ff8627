"""Kernel timeline of the data-parallel step (run under torchrun, N ranks): CUPTI trace of graph replays of bench.py's
step; rank 0 writes gpurun_out/n{N}_timeline.csv (stream, start_us, dur_us, name) for ONE replay and prints the summary
profiles/README.md quotes: step span, busy time of the compute stream, idle gaps >= 5 us on it and what preceded them,
NCCL kernels (stream, start, duration)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    bench.bind_to_gpu_numa_node(local, world)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import datetime
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=180))
    from unipre3d_b200 import synthetic
    from unipre3d_b200.trainer import Trainer
    bench.select_config("transformer")
    cfg = bench.make_cfg(world)
    tr = Trainer(cfg, device=device, use_cuda_graph=True, autocast_dtype=torch.bfloat16)
    pb = tr.pack_batch(synthetic.make_batch(cfg, bench.OBJECTS_PER_GPU, bench.N_POINTS, seed=rank, pin=False, image_dtype="uint8"))
    for _ in range(5):
        tr.train_iteration(pb)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device) if os.environ.get("UP3D_PROFILE_FLUSH") else None
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(4):
            if flush is not None:
                flush.zero_()                      # bench.py's L2 flush between timed steps
            tr.replay_resident()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank == 0:
        evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "memcpy" not in e.name.lower()[:0]]
        rows = sorted(((e.time_range.start, e.time_range.end - e.time_range.start, getattr(e, "stream", -1) or -1, e.name) for e in evs))
        # one replay = the launches between the 2nd and 3rd occurrence of the first kernel name
        first = rows[0][3]
        starts = [i for i, r in enumerate(rows) if r[3] == first]
        per = len(rows) // 4
        a, b = 2 * per, 3 * per
        step = rows[a:b]
        t0 = step[0][0]
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/n{world}_timeline.csv", "w") as f:
            f.write("stream,start_us,dur_us,name\n")
            for s, d, st, n in step:
                f.write(f"{st},{s - t0:.2f},{d:.2f},\"{n[:120]}\"\n")
        span = max(s + d for s, d, _, _ in step) - t0
        by_stream = {}
        for s, d, st, n in step:
            by_stream.setdefault(st, []).append((s - t0, d, n))
        print(f"N={world}: {len(step)} kernels per replay, span {span:.1f} us")
        for st, ks in sorted(by_stream.items(), key=lambda kv: -sum(k[1] for k in kv[1])):
            print(f"  stream {st}: {len(ks)} kernels, busy {sum(k[1] for k in ks):.1f} us")
        main_st = max(by_stream.items(), key=lambda kv: len(kv[1]))[0]
        ks = by_stream[main_st]
        gaps = []
        for (s0, d0, n0), (s1, d1, n1) in zip(ks, ks[1:]):
            g = s1 - (s0 + d0)
            if g >= 5.0:
                gaps.append((g, s0 + d0, n0[:60], n1[:60]))
        print(f"  compute stream {main_st}: idle gaps >= 5 us: {len(gaps)}, total {sum(g[0] for g in gaps):.1f} us")
        for g in sorted(gaps, reverse=True)[:25]:
            print(f"    {g[0]:7.1f} us at {g[1]:8.1f}: after {g[2]} -> before {g[3]}")
        for s, d, st, n in step:
            if "nccl" in n.lower():
                print(f"  NCCL stream {st} start {s - t0:8.1f} dur {d:7.1f} {n[:80]}")
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tr._graph = None
    os._exit(0)


if __name__ == "__main__":
    main()
