#!/usr/bin/env python
"""bench.py -- views/sec of the UniPre3D render-loss pre-training step on B200.

Workload (BASELINE.json configs[1]): transformer_pretraining, 8 objects x 4 rendered views per GPU, 8192-point
clouds, 256x256 renders, synthetic ShapeNet-shaped batches, random-init weights.  One "step" = backbone forward ->
GaussianSplatPredictor -> 32 renders -> focal-L2 -> backward -> grad clip -> AdamW (-> NCCL all-reduce when N > 1).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched through torch.distributed.run)
  python bench.py --impl reference ...                       CPU port of the same step on the host cores (rank 0 only)

One JSON line on rank 0.  `value`: inputs resident in HBM.  `e2e`: same step through Trainer.train_iteration with
pinned HOST batches (H2D inside the timed region) and a D2H read of the loss every step.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OBJECTS_PER_GPU, N_POINTS, RES = 8, 8192, 256
CFG_NAME, GAUSSIANS_PER_OBJECT = "transformer_pretraining", 128
HBM_FALLBACK_GBS = 6650.0


def select_config(name: str) -> None:
    """--config transformer: BASELINE.json configs[1] (the headline, default).  --config pointmlp: configs[2] per GPU
    (pointmlp_pretraining, 4 objects x 4 views, 8192 pts -> 8192 Gaussians per object, 256x256)."""
    global OBJECTS_PER_GPU, CFG_NAME, GAUSSIANS_PER_OBJECT
    if name == "pointmlp":
        OBJECTS_PER_GPU, CFG_NAME, GAUSSIANS_PER_OBJECT = 4, "pointmlp_pretraining", N_POINTS


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).  In-process NVML
    (what nvidia-smi itself reads) from one background thread, every 100 ms; falls back to an `nvidia-smi -lms` child when
    NVML cannot be loaded.  `enabled=False` (ranks other than 0 of a multi-GPU run) samples nothing: eight `nvidia-smi`
    pollers on one box take driver locks often enough to stall the ranks' launches, and with the per-step all-reduce one
    stalled rank stalls all of them (N = 8 end-to-end step: 10 ms with a poller per rank)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, enabled: bool = True):
        self.idx, self.proc, self.lines, self.enabled = gpu_index, None, [], enabled
        self.source, self._stop, self.t = None, threading.Event(), None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip().isdigit()]
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(int(vis[self.idx]) if self.idx < len(vis) else self.idx)

    def _poll_nvml(self, nv, h):
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
                r = int(get_reasons(h))
                f = [str(self.idx), str(sm), str(mx), "", hex(r)] + ["Active" if r & bits[k] else "Not Active"
                                                                      for k in ("hw_slowdown", "hw_thermal_slowdown",
                                                                                "sw_thermal_slowdown", "sw_power_cap")]
                self.lines.append(", ".join(f))
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            nv, h = self._nvml_handle()
            self.source = "nvml"
            self.t = threading.Thread(target=self._poll_nvml, args=(nv, h), daemon=True)
            self.t.start()
            return self
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.source == "nvml":
            time.sleep(0.12)
            self._stop.set()
            self.t.join(timeout=2)
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "source": self.source}


def ncu_traffic(kernel_substr: str, pattern: str = "*_P128_ncu_raw.csv"):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel_substr` from the newest committed
    `ncu --set full` capture matching profiles/<pattern> (default: the transformer-config raster kernels); None if
    absent."""
    import csv
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    if not files:
        return None, None
    try:
        with open(files[-1]) as f:
            rows = list(csv.reader(f))
        hdr, units = rows[0], rows[1]
        ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = [float(r[ir].replace(",", "")) * scale.get(units[ir], 1.0) + float(r[iw].replace(",", "")) * scale.get(units[iw], 1.0)
                for r in rows[2:] if kernel_substr in r[ik]]
        return (float(np.mean(vals)) if vals else None), os.path.basename(files[-1])
    except Exception:
        return None, None


def make_cfg(n_gpus: int):
    from unipre3d_b200.config import compose
    ov = [f"data.training_resolution={RES}", f"opt.batch_size={OBJECTS_PER_GPU * n_gpus}"]
    if n_gpus > 1:
        ov.append("general.device=[" + ",".join(str(i) for i in range(n_gpus)) + "]")
    return compose(CFG_NAME, overrides=ov)


# ------------------------------------------------------------------------------------------------ reference arm
def run_cpu_port(args, as_reference: bool, budget_s: float = 25.0):
    """CPU port (oracle/) on a bounded sample of the same workload: 1 object (8192 pts, 4 views 256x256) per step."""
    from oracle import oracle_lib
    from oracle.cpu_step import CpuStepper
    from unipre3d_b200 import synthetic
    cfg = make_cfg(1)
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    stepper = CpuStepper(cfg)
    data = synthetic.make_batch(cfg, 1, N_POINTS, seed=0)
    views = int(cfg.opt.imgs_per_obj)
    steps, warm = (args.steps, args.warmup) if as_reference else (3, 1)
    t0 = time.perf_counter()
    for _ in range(warm):
        stepper.step(data)
    per = (time.perf_counter() - t0) / max(warm, 1)
    if as_reference:
        steps = max(1, min(steps, int(180.0 / max(per, 1e-3))))      # whole run within a few minutes
    else:
        steps = max(1, min(steps, int(budget_s / max(per, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        stepper.step(data)
    dt = (time.perf_counter() - t0) / steps
    cores = max(torch.get_num_threads(), oracle_lib.num_threads())
    return {"value": views / dt, "unit": "views/s", "cores": cores, "kind": "port",
            "sample": f"1 of {OBJECTS_PER_GPU} objects per step (8192 pts -> 128 Gaussians, {views} views 256x256), "
                      f"{steps} timed steps after {warm} warm-up, torch CPU backbone + C/OpenMP rasterizer oracle",
            "ms_per_step": dt * 1e3, "steps": steps, "warmup": warm}


def run_reference_modules(steps: int, warm: int, n_objects: int, budget_s: float):
    """The reference's OWN backbone modules (staged into oracle/_ref/pyref by build()) on the host cores, fp32, all
    threads; tokenizer kernels and rasterizer = the CPU restatements (the reference has no CPU path for either).
    No unipre3d_b200 import on this path."""
    from oracle import oracle_lib, ref_step
    oracle_lib.build()
    torch.set_num_threads(os.cpu_count() or 1)
    st = ref_step.RefStepper(n_objects=n_objects)
    t0 = time.perf_counter()
    for _ in range(max(warm, 1)):
        st.step()
    per = (time.perf_counter() - t0) / max(warm, 1)
    steps = max(1, min(steps, int(budget_s / max(per, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        st.step()
    dt = (time.perf_counter() - t0) / steps
    cores = max(torch.get_num_threads(), oracle_lib.num_threads())
    return {"value": st.n_views() / dt, "unit": "views/s", "cores": cores, "kind": "reference",
            "sample": f"{n_objects} of {OBJECTS_PER_GPU} objects x 4 views per step (8192 pts -> 128 Gaussians, 256x256), "
                      f"{steps} timed steps after {max(warm, 1)} warm-up; backbone = the reference's own "
                      "PointTransformerEncoder + FeatureFusion modules (oracle/_ref/pyref, unmodified files) + dense "
                      "image_conv on N(0,1) decoder features, torch CPU fp32; FPS/ball-query/group and the rasterizer = "
                      "C/OpenMP restatements (the reference has no CPU path for them)",
            "ms_per_step": dt * 1e3, "steps": steps, "warmup": max(warm, 1)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    same_config = False
    try:
        from oracle import ref_step
        if not ref_step.available():
            raise RuntimeError("oracle/_ref/pyref not staged")
        cb = run_reference_modules(args.steps, min(args.warmup, 2), OBJECTS_PER_GPU, budget_s=150.0)
        same_config = True
    except Exception as e:      # the staged reference files are absent: fall back to the self-contained CPU port
        print(f"[bench] reference modules unavailable ({e!r}); using the CPU port", file=sys.stderr)
        cb = run_cpu_port(args, as_reference=True)
    line = {"impl": "reference", "metric": "views/sec (8192 pts->256^2 render), full pre-training step",
            "value": cb["value"], "unit": "views/s", "n_gpus": args.gpus, "steps": cb["steps"], "warmup": cb["warmup"],
            "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "transformer_pretraining: 8 objects x 4 rendered views per GPU, 8192-pt clouds, "
                                   "256x256, 128 Gaussians/object, SH degree 1 (BASELINE.json configs[1])"
                                   + ("" if same_config else " (bounded sample: 1 object x 4 views per step)"),
                       "l2": "n/a (CPU)"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def raster_roofline(trainer, batch_dev, cfg, reps: int, peaks, peak_kind):
    """Per-kernel device time of the rasterizer on the step's own Gaussians (CUDA events recorded by the C ABI on the
    launching stream), and the HBM roofline of the dominant raster kernel."""
    from unipre3d_b200 import _lib
    from unipre3d_b200.trainer import prepare_model_inputs
    model = trainer.model_manager.model
    ms = np.zeros((reps, 6))
    _lib.check(_lib.lib.up3d_raster_timing_enable(1))
    buf = (C.c_float * 6)()
    try:
        for r in range(reps):
            # keep the GPU busy (~50 ms spin) while the host enqueues the whole eager step, so that no kernel waits for
            # its launch: the event pairs around the raster kernels then bracket device time only
            torch.cuda._sleep(100_000_000)
            trainer._forward_backward(batch_dev)          # eager forward + render + loss + backward (autocast as the step)
            _lib.check(_lib.lib.up3d_raster_timing_read(buf))
            ms[r] = list(buf)
            for p in trainer.params:
                p.grad = None
    finally:
        _lib.check(_lib.lib.up3d_raster_timing_enable(0))
    names = ["project", "depth_sort", "blend_forward", "grad_clear", "blend_backward", "geometry_backward"]
    mean = ms.mean(0)
    V = OBJECTS_PER_GPU * int(cfg.opt.imgs_per_obj)
    P, HW, M = GAUSSIANS_PER_OBJECT * int(cfg.data.input_images), RES * RES, 4     # Gaussians per object (x input views)
    # algorithmic bytes of blend_backward per launch (DESIGN.md "Kernels"): depth-ordered records read once per view
    # (48 B each), per-pixel final_T / n_contrib / dL_dcolor (20 B), per-record 9-float partials read-modify-write
    bwd_bytes = V * (P * 48 + HW * 20 + P * 9 * 4 * 2)
    dom = "blend_backward"
    t_dom = mean[names.index(dom)] * 1e-3
    achieved = bwd_bytes / t_dom / 1e9 if t_dom > 0 else None
    # traffic the reference's algorithm moves for the same views: P(240+36M) + 108 I + 40 WH per view
    I = P * (HW // 256)
    ref_bytes = V * (P * (240 + 36 * M) + 108 * I + 40 * HW)
    raster_ms = float(mean.sum())
    traffic, traffic_src = ncu_traffic("blend_backward_kernel")
    return {
        "roofline": {"bound": "hbm", "kernel": "up3d::blend_backward_kernel", "achieved": achieved,
                     "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": (achieved / peaks["hbm_gbs"]) if achieved else None,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_kind,
                     "note": "kernel is instruction-issue bound (ncu issue-active ~80%, DRAM ~3%): the reference "
                             "algorithm's 108*I bytes/view of sorted-list traffic do not exist here (DESIGN.md §4)",
                     "algorithmic_bytes_per_launch": bwd_bytes, "launch_ms": float(mean[names.index(dom)])},
        "raster_kernels_ms": {n: float(v) for n, v in zip(names, mean)},
        "raster_stage": {"ms_per_step": raster_ms, "views_per_s": V / (raster_ms * 1e-3) if raster_ms > 0 else None,
                         "reference_algorithm_bytes_per_step": ref_bytes,
                         "effective_GBs_vs_reference_traffic": ref_bytes / (raster_ms * 1e-3) / 1e9 if raster_ms > 0 else None},
    }


def raster_only_headline(device, reps: int, peaks, regime: str = "reference", B: int = 8, V: int = 4, P: int = 8192,
                         RES: int = 256):
    """Raster stage alone: by default at the headline Gaussian count (P = 8192 per object, dense reference regime, the
    pointMLP-true case of SURVEY §8d); `regime="small"` = the small-splat distribution SURVEY §8d asks to report
    separately (sort-light / blend-heavy; no HBM fraction is quoted for it)."""
    from unipre3d_b200 import _lib
    from unipre3d_b200.rasterizer import rasterize_batch
    from unipre3d_b200.synthetic import make_camera, make_gaussians
    gs = [make_gaussians(P, seed=100 + i, regime=regime) for i in range(B)]
    cat = {k: torch.tensor(np.concatenate([g[k] for g in gs], 0), device=device).requires_grad_(True) for k in gs[0]}
    cams = [make_camera(az=360.0 * i / (B * V), el=5 + 2.0 * i) for i in range(B * V)]
    vm = torch.tensor(np.stack([c["view"] for c in cams]), device=device)
    pm = torch.tensor(np.stack([c["proj"] for c in cams]), device=device)
    cp = torch.tensor(np.stack([c["campos"] for c in cams]), device=device)
    bg = torch.zeros(3, device=device)
    w = torch.randn(B * V, 3, RES, RES, device=device)
    kw = dict(set_sizes=[P] * B, views_per_set=[V] * B, image_height=RES, image_width=RES, tanfovx=cams[0]["tanfovx"],
              tanfovy=cams[0]["tanfovy"], sh_degree=1, shs=cat["shs"], invdepth=False)
    buf = (C.c_float * 6)()
    ms = np.zeros((reps, 6))
    _lib.check(_lib.lib.up3d_raster_timing_enable(1))
    try:
        for r in range(reps + 3):
            torch.cuda._sleep(20_000_000)                  # host runs ahead of the device (see raster_roofline)
            color, _, _ = rasterize_batch(cat["means3D"], cat["opacities"], cat["scales"], cat["rotations"], vm, pm, cp,
                                          bg, **kw)
            color.backward(w)
            _lib.check(_lib.lib.up3d_raster_timing_read(buf))
            if r >= 3:
                ms[r - 3] = list(buf)
    finally:
        _lib.check(_lib.lib.up3d_raster_timing_enable(0))
    tot = float(ms.mean(0).sum())
    if regime != "reference":
        from unipre3d_b200.rasterizer import debug_forward_state
        dbg = debug_forward_state(cat["means3D"].detach(), cat["opacities"].detach(), cat["scales"].detach(),
                                  cat["rotations"].detach(), vm, pm, cp, bg, set_sizes=[P] * B, views_per_set=[V] * B,
                                  image_height=RES, image_width=RES, tanfovx=cams[0]["tanfovx"], tanfovy=cams[0]["tanfovy"],
                                  sh_degree=1, shs=cat["shs"].detach(), tile_lists=True)
        I_total = int(dbg["tile_counts"].sum().item())
        return {"workload": f"{B} objects x {V} views, P={P} Gaussians/object (small-splat regime), {RES}x{RES}, fwd+bwd",
                "ms": tot, "views_per_s": B * V / (tot * 1e-3),
                "kernels_ms": {n: float(v) for n, v in zip(["project", "depth_sort(+bins)", "blend_forward", "grad_clear",
                                                             "blend_backward", "geometry_backward"], ms.mean(0))},
                "num_rendered_I": I_total, "instances_per_s": I_total / (tot * 1e-3),
                "binned_views": int(dbg["bin_mode"].sum().item()), "views": B * V,
                "bin_entries_per_visible_record": float(dbg["bin_total"].sum().item()) / max(1, int(dbg["n_visible"].sum().item())),
                "note": "blend-bound (exp / compositing arithmetic per (pixel, entry)); no HBM fraction is quoted for this regime"}
    I = P * (RES * RES // 256)
    ref_bytes = B * V * (P * (240 + 36 * 4) + 108 * I + 40 * RES * RES)
    return {"workload": f"{B} objects x {V} views, P=8192 Gaussians/object (reference scale regime), 256x256, fwd+bwd",
            "ms": tot, "views_per_s": B * V / (tot * 1e-3),
            "kernels_ms": {n: float(v) for n, v in zip(["project", "depth_sort", "blend_forward", "grad_clear",
                                                         "blend_backward", "geometry_backward"], ms.mean(0))},
            "reference_algorithm_bytes": ref_bytes,
            "effective_GBs_vs_reference_traffic": ref_bytes / (tot * 1e-3) / 1e9,
            "frac_of_hbm_peak_vs_reference_traffic": ref_bytes / (tot * 1e-3) / 1e9 / peaks["hbm_gbs"]}


def legacy_gpu_pointops(device, reps: int = 20):
    """the reference's OWN pointnet2_batch kernels (oracle/_ref/pointnet2_batch_cuda.so, compiled unmodified
    for sm_100 by oracle/build_ref.py) timed on this GPU next to the kernels that replace them, at the tokenizer's size
    (8 clouds x 8192 points -> 128 centres, 32 neighbours, radius 0.1).  CUDA events, L2-resident inputs (96 KB/cloud)."""
    import importlib.util
    from unipre3d_b200.pointops import furthest_point_sample, subsample_group
    so = os.path.join(ROOT, "oracle", "_ref", "pointnet2_batch_cuda.so")
    if not os.path.exists(so):
        return {"unavailable": "oracle/_ref/pointnet2_batch_cuda.so not built"}
    spec = importlib.util.spec_from_file_location("pointnet2_batch_cuda", so)
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    B, N, M, K, r = OBJECTS_PER_GPU, N_POINTS, 128, 32, 0.1
    g = torch.Generator(device="cpu").manual_seed(0)
    x = (torch.randn(B, N, 3, generator=g) * 0.2).to(device).contiguous()
    xt = x.transpose(1, 2).contiguous()

    def ev_time(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3

    idx = torch.empty(B, M, dtype=torch.int32, device=device)
    temp = torch.empty(B, N, dtype=torch.float32, device=device)
    nidx = torch.empty(B, M, K, dtype=torch.int32, device=device)
    grouped = torch.empty(B, 3, M, K, dtype=torch.float32, device=device)

    def ref_fps():
        temp.fill_(1e10)                              # subsample.py:93 (the reference re-creates it per call)
        ext.furthest_point_sampling_wrapper(B, N, M, x, temp, idx)

    ref_fps()
    center = torch.gather(x, 1, idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()

    def ref_group():
        ext.ball_query_wrapper(B, N, M, r, K, center, x, nidx)
        ext.group_points_wrapper(B, 3, N, M, K, xt, nidx, grouped)

    out = {"shape": f"B={B} N={N} centres={M} K={K} radius={r}",
           "reference_fps_us": ev_time(ref_fps), "up3d_fps_us": ev_time(lambda: furthest_point_sample(x, M)),
           "reference_ballquery_group_us": ev_time(ref_group),
           "up3d_subsample_group_us": ev_time(lambda: subsample_group(x, M, K, r)),
           "note": "up3d_subsample_group = FPS + ball query + grouping + centring in 2 launches (includes the FPS time); "
                   "the reference needs FPS + gather + ball query + grouping + subtraction"}
    # index-for-index agreement of the two implementations at this size
    out["fps_indices_equal"] = bool(torch.equal(furthest_point_sample(x, M), idx))
    return out


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(local_rank: int, world: int):
    """Pin this process (and, by first touch, its pinned staging buffers) to the host cores that are local to its GPU
    (/sys/bus/pci/devices/<gpu>/local_cpulist).  Without it the step's host->device copy may cross the socket
    interconnect: on this pool the same 8.66 MB copy took 0.3 ms or 4-5 ms depending on where the process happened to
    run.  Ranks that share a NUMA node split its cores.  UP3D_NO_NUMA_BIND=1 disables it."""
    if os.environ.get("UP3D_NO_NUMA_BIND"):
        return "disabled"
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = [v for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip().isdigit()]
        h = pynvml.nvmlDeviceGetHandleByIndex(int(vis[local_rank]) if local_rank < len(vis) else local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:                      # NVML prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            cpus = sorted(_parse_cpulist(f.read()))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if not allowed:
            return "no local cpus"
        if world > 1:                                        # ranks on the same node: disjoint slices of its cores
            per = max(1, len(allowed) // max(1, min(world, 4)))
            k = local_rank % max(1, len(allowed) // per)
            allowed = allowed[k * per:(k + 1) * per] or allowed
        os.sched_setaffinity(0, set(allowed))
        torch.set_num_threads(min(len(allowed), 8))
        return f"{len(allowed)} cores local to {bus}"
    except Exception as e:                                   # never fail the bench because of a missing sysfs file
        return f"unavailable ({type(e).__name__})"


def ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU port")
    binding = bind_to_gpu_numa_node(local, world)          # before the CUDA context and any pinned allocation
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=180))
    n_gpus = world
    from unipre3d_b200 import _lib, synthetic
    from unipre3d_b200.trainer import Trainer, _to_device
    cfg = make_cfg(n_gpus)
    use_graph = not args.no_graph
    autocast = None if args.fp32 else torch.bfloat16
    trainer = Trainer(cfg, device=device, use_cuda_graph=use_graph, autocast_dtype=autocast)
    n_batches = 4
    image_dtype = "float32" if args.float_images else "uint8"
    raw_batches = [synthetic.make_batch(cfg, OBJECTS_PER_GPU, N_POINTS, seed=1000 * rank + i, pin=False, image_dtype=image_dtype)
                   for i in range(n_batches)]
    # each batch lives in ONE pinned host buffer, as a loader writing into Trainer.pack_batch's views would leave it:
    # the step then moves its inputs with one host->device copy (+ one device->device copy into the graph's inputs)
    batches = [trainer.pack_batch(b) for b in raw_batches]
    h2d = batches[0].nbytes
    views_per_step = OBJECTS_PER_GPU * int(cfg.opt.imgs_per_obj) * n_gpus
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)     # > 126 MB L2
    peaks, peak_kind = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also captures the CUDA graph)
    for i in range(max(args.warmup, 3)):
        trainer.train_iteration(batches[i % n_batches])
    barrier()
    launches_before = _lib.launch_count

    def timed(step_fn):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for i in range(args.steps):
            flush.zero_()                      # L2 flush between timed iterations (outside the event pair)
            ev[i][0].record()
            step_fn(i)
            ev[i][1].record()
        barrier()
        total_ms = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([total_ms], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident steps ("value"): inputs already in HBM
    resident_flat = batches[0].flat.to(device)
    resident = batches[0].views(resident_flat)
    if use_graph:
        trainer._copy_into_static(resident_flat)
        torch.cuda.synchronize()

        def step_resident(i):
            # one replay of the captured step: takes its FPS / ball-query grouping from the previous replay's tail
            # branch and, beside its optimizer pass, groups the coordinates of the following step (the same work per
            # step as train_iteration(..., prefetch=next))
            trainer.replay_resident()
    else:
        def step_resident(i):
            trainer._step_body(resident)
    for _ in range(2):
        step_resident(0)                       # untimed: the first resident replay groups its batch in line
    with ClockSampler(local, enabled=(rank == 0)) as clk:
        ms_resident = timed(step_resident)
        # ---- end-to-end steps: pinned host batch -> H2D -> step -> D2H loss, through the public API
        losses = []
        # (the next step's pinned batch is prefetched on a copy stream while this step computes; every step still
        #  moves its full input batch host->device inside the timed region and reads its loss back)
        #  the loss of every step is read back device->host, consumed with a two-step lag as a logging loop would)
        ms_e2e = timed(lambda i: losses.append(trainer.train_iteration(
            batches[i % n_batches], read_loss="lagged", prefetch=batches[(i + 1) % n_batches])))
        losses = [x for x in losses if x is not None] + [trainer.last_loss()]
        # what this box's host->device path gives for one step's batch (median of 7 single copies from the pinned buffer):
        # boxes of the pool differ from ~53 GB/s to ~2 GB/s here, which is what moves e2e relative to `value`
        torch.cuda.synchronize()
        probe_dst, probe_ms = torch.empty_like(resident_flat), []
        for _ in range(7):
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            probe_dst.copy_(batches[0].flat, non_blocking=True)
            p1.record()
            torch.cuda.synchronize()
            probe_ms.append(p0.elapsed_time(p1))
        h2d_probe_gbs = h2d / (sorted(probe_ms)[3] * 1e-3) / 1e9
        del probe_dst
        # steady state without the artificial L2 flush (the step's own working set -- parameters, moments, gradients,
        # shadows: ~520 MB -- already exceeds the 126 MB L2 several times over); reported next to the flushed headline
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step_resident(i)
        e1.record()
        barrier()
        t_nf = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t_nf, op=dist.ReduceOp.MAX)
        ms_noflush = float(t_nf.item()) / args.steps
    clocks = clk.summary()
    # kernels of libunipre3d_b200 per step (FPS, subsample_group, project, depth_sort, blend_forward, focal_l2 x2,
    # blend_backward, geometry_backward), counted from the C-ABI calls of one eager step; a graph replay launches
    # the same kernel nodes
    before = _lib.launch_count
    trainer._step_body(resident)
    torch.cuda.synchronize()
    per_step_launches = _lib.launch_count - before
    gpu_launches = per_step_launches * args.steps * 2

    line = None
    extra = {}
    try:
        # eager (no graph) pass with event hooks for the per-kernel numbers; EVERY rank runs it (the model's
        # SyncBatchNorm layers are collectives when N > 1), rank 0 reports
        extra.update(raster_roofline(trainer, resident, cfg, max(3, min(args.steps, 10)), peaks, peak_kind))
    except Exception as e:  # never lose the headline because a side measurement failed
        extra["roofline_error"] = repr(e)
    try:
        # the optimizer pass: the step's largest HBM-bound kernel (one read-modify-write over param/grad/moments)
        trainer._forward_backward(resident)
        trainer._allreduce_grads()
        sq_ms, ap_ms, ap_bytes = trainer.model_manager.optimizer.time_passes(5)
        trainer.model_manager.optimizer.zero_grad(set_to_none=True)
        tr_adam, tr_src = ncu_traffic("adamw_kernel", "*_step_ncu_raw.csv")
        extra["roofline_hbm_kernel"] = {
            "bound": "hbm", "kernel": "up3d::adamw_kernel", "achieved": ap_bytes / (ap_ms * 1e-3) / 1e9,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ap_bytes / (ap_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "traffic": tr_adam, "traffic_source": tr_src, "peak_source": peak_kind,
            "algorithmic_bytes_per_launch": ap_bytes, "launch_ms": ap_ms,
            "note": "28 B per parameter (read param/grad/exp_avg/exp_avg_sq, write param/exp_avg/exp_avg_sq) + 2 B per "
                    "bf16-shadowed parameter; grad_sumsq_kernel (4 B/param) takes %.4f ms" % sq_ms}
    except Exception as e:
        extra["roofline_hbm_kernel_error"] = repr(e)
    if rank == 0 and n_gpus == 1 and CFG_NAME.startswith("transformer"):
        # two variants of the same step, each a fresh Trainer + CUDA graph, 10 resident-input replays (reported beside the
        # headline, never instead of it): (i) backbone GEMMs in fp32 = the reference's precision (no AMP anywhere in its
        # trainer); (ii) dense N(0,1) (B,128,R,R) image features in HBM instead of the analytic stem field
        for key, kw, ov in (("variant_fp32_backbone", dict(autocast_dtype=None), []),
                            ("variant_dense_randn_image_features", dict(autocast_dtype=autocast), ["model.image_branch=randn"])):
            try:
                from unipre3d_b200.config import compose as _compose
                cfg_v = _compose(CFG_NAME, overrides=[f"data.training_resolution={RES}", f"opt.batch_size={OBJECTS_PER_GPU}"] + ov)
                tv = Trainer(cfg_v, device=device, use_cuda_graph=True, **kw)
                pb = tv.pack_batch(raw_batches[0])
                for _ in range(3):
                    tv.train_iteration(pb)
                tv.replay_resident()
                torch.cuda.synchronize()
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
                for a_, b_ in evs:
                    flush.zero_()
                    a_.record(); tv.replay_resident(); b_.record()
                torch.cuda.synchronize()
                msv = sum(a_.elapsed_time(b_) for a_, b_ in evs) / len(evs)
                extra[key] = {"ms_per_step": msv, "views_per_s": views_per_step / (msv * 1e-3)}
                del tv
            except Exception as e:
                extra[key + "_error"] = repr(e)
    if rank == 0:
        try:
            extra["raster_only"] = raster_only_headline(device, 5, peaks)
            extra["raster_only_small_splat"] = raster_only_headline(device, 5, peaks, regime="small")
            # BASELINE configs[3]-shaped raster stress: one scene of 100k small Gaussians, 4 views of 512x512 (1024 tiles)
            extra["raster_only_scene_100k"] = raster_only_headline(device, 3, peaks, regime="small", B=1, V=4, P=100000, RES=512)
        except Exception as e:
            extra["raster_only_error"] = repr(e)
        cpu_baseline = None
        if n_gpus == 1 and not args.no_cpu_baseline and CFG_NAME.startswith("transformer"):
            try:
                from oracle import ref_step
                if ref_step.available():
                    cb = run_reference_modules(3, 1, 2, budget_s=20.0)      # bounded sample: 2 objects per step
                else:
                    cb = run_cpu_port(args, as_reference=False)
                cpu_baseline = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:
                cpu_baseline = {"error": repr(e)}
            try:
                extra["legacy_gpu_pointops"] = legacy_gpu_pointops(device)
            except Exception as e:
                extra["legacy_gpu_pointops_error"] = repr(e)
        value = views_per_step / (ms_resident / args.steps * 1e-3)
        e2e = views_per_step / (ms_e2e / args.steps * 1e-3)
        line = {"metric": "views/sec (8192 pts->256^2 render), full pre-training step", "value": value,
                "unit": "views/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_resident / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None,
                "dtype": "f32 rasterizer/point ops/loss/optimizer; backbone GEMMs " + ("f32" if args.fp32 else "bf16 (fp32 accumulate, fp32 master weights)"),
                "data": "synthetic",
                "config": {"workload": f"{CFG_NAME}: {OBJECTS_PER_GPU} objects x 4 rendered views per GPU, 8192-pt clouds, "
                                       f"256x256, {GAUSSIANS_PER_OBJECT} Gaussians/object, SH degree 1 (BASELINE.json "
                                       + ("configs[1])" if CFG_NAME.startswith("transformer") else "configs[2], per-GPU share)"),
                           "objects_per_gpu": OBJECTS_PER_GPU, "views_per_step": views_per_step, "parallelism": f"dp{n_gpus}",
                           "cuda_graph": bool(use_graph), "host_binding": binding,
                           "lookahead_grouping": ("every captured step groups (FPS + ball query) the following step's "
                                                  "points in a branch beside its optimizer pass (steps served so far: %d)"
                                                  % trainer.lookahead_hits
                                                  if trainer._la is not None else "off"),
                           "value_excludes": "host-side StepLR bookkeeping and the EMA update (every 10th step); e2e includes them",
                           "host_images": ("float32 (divided by 255 on the host, as the reference loader)" if args.float_images
                                           else "uint8 as decoded from the dataset's PNGs; /255 on the device"),
                           "l2": "256 MiB buffer written between timed iterations (L2 flush), per-step CUDA-event pairs"},
                "clocks": clocks,
                "e2e": {"value": e2e, "unit": "views/s", "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": h2d * n_gpus, "d2h_bytes_per_step": 4 * n_gpus,
                        "h2d_probe_GBs": h2d_probe_gbs, "last_loss": losses[-1] if losses else None},
                "gpu_launches": gpu_launches, "cpu_baseline": cpu_baseline,
                "steady_state_no_l2_flush": {"ms_per_step": ms_noflush, "views_per_s": views_per_step / (ms_noflush * 1e-3),
                                             "note": "same graph replays back to back, no flush between steps (not the "
                                                     "headline; the flushed number above is)"}}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        # tearing the NCCL communicator down while CUDA graphs that captured its kernels are alive can hang:
        # release the graph first, then skip the (optional) orderly teardown altogether
        trainer._graph = None
        os._exit(0)


SCENE_CONFIGS = {
    # BASELINE.json configs[3]: ptv3_pretraining (ScanNet-shaped scene), 100k points, 4 views 512x512 -- per-GPU share: 1 scene
    "ptv3": dict(name="ptv3_pretraining", points=100_000, views=4, sh=1),
    # configs[4]: sparseunet_pretraining scene, 200k points, 8 views 512x512, SH degree 3 (stress) -- per-GPU share: 1 scene
    "sparseunet": dict(name="sparseunet_pretraining", points=200_000, views=8, sh=3),
}


def ours_scene(args):
    """Scene-level configs (eager step: voxel counts are data dependent, no CUDA graph): one scene per GPU, backbone on
    the sparse-convolution engine -> per-scene Gaussian list -> all views in one binned raster call -> L2 -> backward ->
    fused clip + AdamW.  `opt.use_fusion=false` (PointFusion is not built)."""
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device")
    sc = SCENE_CONFIGS[args.config]
    binding = bind_to_gpu_numa_node(0, 1)
    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)
    from unipre3d_b200 import _lib, synthetic
    from unipre3d_b200.config import compose
    from unipre3d_b200.trainer import Trainer, _to_device
    cfg = compose(sc["name"], overrides=["opt.use_fusion=false", "data.input_images=0", f"opt.imgs_per_obj={sc['views']}",
                                         "data.training_width=512", "data.training_height=512", "opt.batch_size=1",
                                         f"model.max_sh_degree={sc['sh']}", "opt.ema.use=false"])
    autocast = None if args.fp32 else torch.bfloat16      # dense Linear / attention layers of PTv3 and the heads
    trainer = Trainer(cfg, device=device, use_cuda_graph=False, autocast_dtype=autocast)
    batches = [synthetic.make_scene_batch(cfg, 1, sc["points"], seed=i) for i in range(2)]
    n_vox = int(batches[0]["point_cloud"]["coord"].shape[0])
    for i in range(max(args.warmup, 3)):
        trainer.train_iteration(batches[i % 2])
    torch.cuda.synchronize()
    resident = _to_device(batches[0], device, non_blocking=False)
    steps = args.steps
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)

    def timed(fn):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        for i in range(steps):
            flush.zero_()
            ev[i][0].record()
            fn(i)
            ev[i][1].record()
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in ev) / steps

    before = _lib.launch_count
    with ClockSampler(0) as clk:
        ms = timed(lambda i: trainer._step_body(resident))
        ms_e2e = timed(lambda i: trainer.train_iteration(batches[i % 2]))
    launches = _lib.launch_count - before
    h2d = sum(t.numel() * t.element_size() for _, t in __import__("unipre3d_b200.trainer", fromlist=["_flat_items"])._flat_items(batches[0]))
    V = sc["views"]
    line = {"metric": "views/sec, full pre-training step (scene level)", "value": V / (ms * 1e-3), "unit": "views/s", "n_gpus": 1,
            "steps": steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32 features / rasterizer / optimizer; sparse-conv operands bf16 (fp32 accumulate); dense Linear / attention "
                     + ("f32" if args.fp32 else "bf16 autocast"),
            "data": "synthetic",
            "config": {"workload": f"{sc['name']} (opt.use_fusion=false): 1 scene per GPU, {sc['points']} surface points -> {n_vox} "
                                   f"voxels at 0.02 m = Gaussians, {V} views 512x512, SH degree {sc['sh']} (BASELINE.json "
                                   f"configs[{3 if args.config == 'ptv3' else 4}], per-GPU share)",
                       "cuda_graph": False, "host_binding": binding,
                       "l2": "256 MiB buffer written between timed iterations (L2 flush), per-step CUDA-event pairs"},
            "clocks": clk.summary(),
            "e2e": {"value": V / (ms_e2e * 1e-3), "unit": "views/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="transformer", choices=["transformer", "pointmlp", "ptv3", "sparseunet"],
                    help="transformer = BASELINE configs[1] (headline); pointmlp = configs[2] per-GPU share (4 objects); "
                         "ptv3 / sparseunet = configs[3] / [4] per-GPU share (1 scene, eager step)")
    ap.add_argument("--fp32", action="store_true", help="keep the backbone GEMMs in fp32 (reference precision)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--float-images", action="store_true", help="host batches carry float32 images (4x the H2D bytes)")
    args = ap.parse_args()
    if args.config in SCENE_CONFIGS and args.impl == "ours":
        return ours_scene(args)
    select_config(args.config)
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
