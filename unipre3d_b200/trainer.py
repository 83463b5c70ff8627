"""Pre-training step of the render-loss path with the reference trainer's shape
(/root/reference/train_network.py:136-220 ModelManager, 222-302 ValidationManager, 305-464 Trainer) on B200.

What changes relative to the reference loop (same math, fewer host round trips):
  * `render_validation_views` renders all B*V' (object, view) pairs in ONE batched launch set
    (gaussian_renderer.render_batch_predicted) instead of a Python double loop of render calls (418-442);
  * `focal_l2_loss` value + gradient come from one fused kernel;
  * `_check_and_clip_gradients` (368-390): the per-parameter `.any()` host syncs become one device-side
    total-norm; the "skip the step on NaN/Inf" rule is applied on the device through fused AdamW's `found_inf`
    hook, so the optimizer state is untouched on a bad step exactly as in the reference (336-340);
  * the whole step (forward, render, loss, backward, clip, AdamW) can be captured into ONE CUDA graph
    (`use_cuda_graph=True`), replayed per iteration with inputs copied into static buffers;
  * multi-GPU: one process per GPU, objects sharded over ranks (train_network.py:95-103 semantics), gradients
    all-reduced with NCCL (torch DDP buckets overlap the backward; SyncBatchNorm as train_network.py:183).
"""
from __future__ import annotations

import copy
import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from .gaussian_predictor import GaussianSplatPredictor
from .gaussian_renderer import render_batch_predicted
from .loss import focal_l2_loss, l1_loss, l2_loss, backward_unit


def _to_device(x, device, non_blocking=True):
    if isinstance(x, dict):
        return {k: _to_device(v, device, non_blocking) for k, v in x.items()}
    if torch.is_tensor(x):
        return x.to(device, non_blocking=non_blocking)
    return x


def prepare_model_inputs(data, cfg, bs_per_gpu, device):
    """utils/general_utils.py:251-293 (object level)."""
    point_cloud = data["point_cloud"] if "point_cloud" in data else data
    inputs = {"point_cloud": point_cloud,
              "source_cameras_view_to_world": data["view_to_world_transforms"][:, : cfg.data.input_images, ...],
              "image": None, "unprojected_coords": None}
    if cfg.opt.use_fusion:
        img = data["gt_images"][:, : cfg.data.input_images, ...]
        if img.dtype == torch.uint8:
            # 8-bit images as decoded from the dataset's PNGs: /255 on the device (the reference's loader divides on the
            # host, dataset/shapenet.py); only the source views are converted -- the loss reads the 8-bit targets in place
            img = img.to(device).float().div_(255.0)
        inputs["image"] = img
    return _to_device(inputs, device)


def shard_batch(data, rank: int, world: int):
    """Objects are the unit of data parallelism: rank r owns objects [r*B/world, (r+1)*B/world) of a
    global batch dict, with all of their views (the scene branch of the reference shards the same way through
    DistributedSampler, train_network.py:55-71,95-103)."""
    def cut(t):
        B = t.shape[0]
        if B % world != 0:
            raise ValueError(f"global batch of {B} objects does not divide over {world} ranks")
        per = B // world
        return t[rank * per:(rank + 1) * per]
    return {k: ({kk: cut(vv) for kk, vv in v.items()} if isinstance(v, dict) else cut(v)) for k, v in data.items()}


def _flat_items(data, prefix=""):
    for k in sorted(data.keys()):
        v = data[k]
        if isinstance(v, dict):
            yield from _flat_items(v, prefix + k + "/")
        elif torch.is_tensor(v):
            yield prefix + k, v


class PackedBatch:
    """A batch dict whose tensors are views into ONE pinned host buffer (256-byte aligned slots), so that a step moves its
    inputs with one host->device copy and one device->device copy instead of one pair per tensor.  `Trainer.pack_batch`
    builds it from an ordinary batch dict; a data loader can also fill `packed.data` in place (no extra host copy)."""

    def __init__(self, data, pin: bool = True):
        self.layout, off = [], 0
        for name, t in _flat_items(data):
            nbytes = t.numel() * t.element_size()
            self.layout.append((name, off, tuple(t.shape), t.dtype, nbytes))
            off += (nbytes + 255) // 256 * 256
        self.nbytes = off
        self.flat = torch.empty(off, dtype=torch.uint8)
        if pin and torch.cuda.is_available():
            self.flat = self.flat.pin_memory()
        self.data = self.views(self.flat)
        for name, t in _flat_items(data):
            self._get(self.data, name).copy_(t)

    def views(self, flat):
        """The batch dict as views into `flat` (a uint8 buffer of self.nbytes bytes on any device)."""
        out = {}
        for name, off, shape, dtype, nbytes in self.layout:
            v = flat[off:off + nbytes].view(dtype).view(shape)
            d, parts = out, name.split("/")
            for p_ in parts[:-1]:
                d = d.setdefault(p_, {})
            d[parts[-1]] = v
        return out

    @staticmethod
    def _get(d, name):
        for p_ in name.split("/"):
            d = d[p_]
        return d

    def signature(self):
        return tuple((n, s, str(dt)) for n, _, s, dt, _ in self.layout)


class EMA:
    """Exponential moving average of the model weights with ema_pytorch's default schedule (the reference
    wraps the model in ema_pytorch.EMA(beta, update_every, update_after_step), train_network.py:188-198):
    copy weights until `update_after_step`, then every `update_every` steps
    ema = lerp(ema, w, 1 - decay), decay = clamp(1 - (1 + t)^(-2/3), 0, beta), t = step - update_after_step - 1."""

    def __init__(self, model, beta=0.9999, update_every=10, update_after_step=100, inv_gamma=1.0, power=2.0 / 3.0):
        self.online = model
        self.ema_model = copy.deepcopy(model).requires_grad_(False).eval()
        self.beta, self.update_every, self.update_after_step = beta, update_every, update_after_step
        self.inv_gamma, self.power = inv_gamma, power
        self.step, self.initted = 0, False

    def _pairs(self):
        src = self.online.module if hasattr(self.online, "module") else self.online
        fl = lambda m: [p for p in list(m.parameters()) + list(m.buffers()) if p.dtype.is_floating_point]
        return fl(self.ema_model), fl(src)

    def current_decay(self) -> float:
        epoch = max(self.step - self.update_after_step - 1, 0)
        if epoch <= 0:
            return 0.0
        return min(max(1 - (1 + epoch / self.inv_gamma) ** -self.power, 0.0), self.beta)

    @torch.no_grad()
    def update(self):
        step = self.step
        self.step += 1
        if step % self.update_every != 0:
            return
        e, s = self._pairs()
        if step <= self.update_after_step or not self.initted:
            torch._foreach_copy_(e, s)
            self.initted = True
            return
        torch._foreach_lerp_(e, s, 1.0 - self.current_decay())


class ModelManager:
    """train_network.py:136-220."""

    def __init__(self, cfg, device, capturable: bool = False):
        self.cfg, self.device = cfg, device
        self.model = GaussianSplatPredictor(cfg).to(device)
        base_lr = cfg.opt.base_lr
        groups = [{"params": list(self.model.point_network.parameters()), "lr": base_lr}]
        if cfg.opt.use_fusion:
            groups += [{"params": list(self.model.fusion_mlps.parameters()), "lr": base_lr},
                       {"params": list(self.model.image_conv.parameters()), "lr": base_lr}]
        # train_network.py:156-159 AdamW(lr=0.0, eps=1e-15, betas) + 368-390 clip_grad_norm_(1.0)/skip-on-NaN, fused
        # (learning rates become device scalars at the first step, so a captured graph sees StepLR updates)
        from .optim import FusedClipAdamW
        self.optimizer = FusedClipAdamW(groups, lr=0.0, eps=1e-15, betas=tuple(cfg.opt.betas), max_norm=1.0)
        self.step_lr, self.lr_gamma, self._sched_step = cfg.opt.step_lr, cfg.opt.lr_gamma, 0
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        if self.world > 1:
            # train_network.py:183-186: SyncBatchNorm + data parallelism.  The gradient exchange is one NCCL
            # all-reduce of a flat fp32 buffer per step (Trainer._allreduce_grads) instead of DDP's bucket hooks, so
            # that the whole step -- collectives included -- can be captured into one CUDA graph.
            self.model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(self.model)
            with torch.no_grad():
                for t in list(self.model.parameters()) + list(self.model.buffers()):
                    c = t if t.is_contiguous() else t.contiguous()
                    dist.broadcast(c, src=0)
                    if c is not t:
                        t.copy_(c)
        self.ema = EMA(self.model, beta=cfg.opt.ema.beta, update_every=cfg.opt.ema.update_every,
                       update_after_step=cfg.opt.ema.update_after_step) if cfg.opt.ema.use else None

    @property
    def forward_model(self):
        return self.model

    def scheduler_step(self):
        """StepLR(step_size=opt.step_lr, gamma=opt.lr_gamma) (train_network.py:160-163), in place."""
        if self.step_lr == -1:
            return
        self._sched_step += 1
        if self._sched_step % self.step_lr == 0:
            for g in self.optimizer.param_groups:
                if torch.is_tensor(g["lr"]):
                    with torch.no_grad():
                        g["lr"].mul_(self.lr_gamma)
                else:
                    g["lr"] = g["lr"] * self.lr_gamma

    def save_checkpoint(self, iteration: int, best_psnr: float, save_path: str) -> None:
        """Same dict keys as train_network.py:200-210 (`model_state_dict` holds the EMA weights when EMA is on, as in
        the reference) plus what an exact resume needs and the reference drops: the online weights the Adam moments
        belong to and the EMA schedule position."""
        ckpt = {"iteration": iteration, "optimizer_state_dict": self.optimizer.state_dict(),
                "model_state_dict": (self.ema.ema_model.state_dict() if self.ema else self.model.state_dict()),
                "best_PSNR": best_psnr}
        if self.ema:
            ckpt["online_model_state_dict"] = self.model.state_dict()
            ckpt["ema_state"] = {"step": self.ema.step, "initted": self.ema.initted}
        torch.save(ckpt, save_path)

    def save_latest_checkpoint(self, iteration: int, best_psnr: float, save_dir: str) -> None:
        """train_network.py:212-215."""
        self.save_checkpoint(iteration, best_psnr, os.path.join(save_dir, "model_latest.pth"))

    def save_best_checkpoint(self, iteration: int, best_psnr: float, save_dir: str) -> None:
        """train_network.py:217-220."""
        self.save_checkpoint(iteration, best_psnr, os.path.join(save_dir, "model_best.pth"))

    def load_checkpoint(self, path: str, load_optimizer: bool = True) -> Dict[str, float]:
        """Resume from a checkpoint written by `save_checkpoint` (or by the reference's ModelManager: same keys).
        Frozen image-network weights absent from the file keep their current values."""
        ckpt = torch.load(path, map_location=self.device)

        def clean(sd):
            # a reference checkpoint may come from a DDP-wrapped model ("module." prefix) and always carries the frozen
            # AutoencoderKL under image_network.* (a submodule there; absent here unless model.image_branch=sdvae)
            sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
            own = set(self.model.state_dict().keys())
            return {k: v for k, v in sd.items() if k in own or not k.startswith("image_network.")}

        ema_sd = clean(ckpt["model_state_dict"])
        online_sd = clean(ckpt["online_model_state_dict"]) if "online_model_state_dict" in ckpt else ema_sd
        info = self.model.load_state_dict(online_sd, strict=False)
        bad = [k for k in info.missing_keys if not k.startswith("image_network.")] + list(info.unexpected_keys)
        if bad:
            raise RuntimeError(f"checkpoint does not match the model: {bad[:8]}")
        if self.ema:
            self.ema.ema_model.load_state_dict(ema_sd, strict=False)
            st = ckpt.get("ema_state")
            if st is not None:
                self.ema.step, self.ema.initted = int(st["step"]), bool(st["initted"])
            else:       # reference checkpoint: the EMA weights ARE the saved model; continue its schedule from `iteration`
                self.ema.step, self.ema.initted = int(ckpt.get("iteration", 0)), True
        if load_optimizer and "optimizer_state_dict" in ckpt:
            self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
            # StepLR phase: the decayed lr was just loaded, keep counting from the saved iteration (train_network.py:160-163)
            self._sched_step = int(ckpt.get("iteration", 0))
        for m in self.model.modules():          # bf16 weight shadows (mixed_precision.ShadowWeights) follow the masters
            w16 = getattr(m, "_w16", None)
            if w16 is not None:
                w16.copy_(m.weight.detach().reshape(w16.shape))
                if getattr(m, "_b16", None) is not None:
                    m._b16.copy_(m.bias.detach())
        return {"iteration": int(ckpt.get("iteration", 0)), "best_PSNR": float(ckpt.get("best_PSNR", 0.0))}


class ValidationManager:
    def __init__(self, cfg, device):
        self.cfg, self.device = cfg, device
        self.background = torch.tensor([1, 1, 1] if cfg.data.white_background else [0, 0, 0], dtype=torch.float32,
                                       device=device)

    def calculate_losses(self, rendered_images, gt_images, iteration: int) -> Dict[str, torch.Tensor]:
        """train_network.py:260-302.  LPIPS-VGG (after opt.start_lpips_after) needs pretrained VGG weights that
        are not available offline; it is outside the measured path."""
        losses = {}
        if self.cfg.opt.loss == "focal_l2":
            losses["l12_loss"] = focal_l2_loss(rendered_images, gt_images, self.background,
                                               self.cfg.opt.non_bg_color_loss_rate, self.cfg.opt.bg_color_loss_rate)
        else:
            losses["l12_loss"] = (l1_loss if self.cfg.opt.loss == "l1" else l2_loss)(rendered_images, gt_images)
        if self.cfg.opt.lambda_lpips != 0 and iteration > self.cfg.opt.start_lpips_after:
            raise NotImplementedError("LPIPS loss needs VGG weights (not shipped); set opt.lambda_lpips=0")
        losses["total_loss"] = losses["l12_loss"]
        return losses


class Trainer:
    """Step driver.  `train_iteration(data)` == reference 450-464 + 333-352 for one batch dict (host tensors)."""

    def __init__(self, cfg, device: Optional[torch.device] = None, use_cuda_graph: bool = False,
                 autocast_dtype: Optional[torch.dtype] = None, shadow_weights: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("unipre3d_b200.Trainer needs a CUDA device (there is no CPU fallback)")
        self.cfg = cfg
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        torch.manual_seed(cfg.general.random_seed)
        torch.cuda.manual_seed_all(cfg.general.random_seed)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.bs_per_gpu = cfg.opt.batch_size if self.world == 1 else cfg.opt.batch_size // self.world
        self.use_cuda_graph = use_cuda_graph
        self.autocast_dtype = autocast_dtype
        self.model_manager = ModelManager(cfg, self.device, capturable=use_cuda_graph)
        self.validation_manager = ValidationManager(cfg, self.device)
        self.params = [p for p in self.model_manager.model.parameters() if p.requires_grad]
        self._shadow = None
        if autocast_dtype == torch.bfloat16 and shadow_weights:
            from .mixed_precision import ShadowWeights
            self._shadow = ShadowWeights(self.model_manager.model)
            # the optimizer kernel rewrites the bf16 shadows together with the fp32 masters (no separate cast pass)
            self.model_manager.optimizer.set_shadows(
                {id(m): s for m, s in zip(self._shadow.masters, self._shadow.shadows)})
        self._grad_sync = None
        if self.world > 1:
            self._setup_grad_sync()
        else:
            from . import fused_encoder
            fused_encoder.GRAD_READY_HOOK = None      # a hook left by an earlier data-parallel Trainer of this process
            fused_encoder.GRAD_CHUNK_HOOK = None
            fused_encoder.GRAD_BUFFERS = None
        self.iteration = 0
        self._graph = None
        self._static: Optional[dict] = None
        if self._lpips_due(int(getattr(cfg.opt, "iterations", 0))):
            import warnings
            warnings.warn(f"opt.lambda_lpips={cfg.opt.lambda_lpips} switches the LPIPS-VGG term on after iteration "
                          f"{cfg.opt.start_lpips_after}; its pretrained weights are not shipped, so train_iteration "
                          f"raises there (set opt.lambda_lpips=0 to train without it)", stacklevel=2)
        self._loss_buf = torch.zeros((), dtype=torch.float32, device=self.device)
        self._loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
        # lagged loss reads: step i's loss is read back when step i + LOSS_LAG is enqueued, so the host may run LOSS_LAG
        # steps ahead of the device (data parallel: a rank whose host hiccups for a millisecond then no longer stalls
        # every rank through the per-step all-reduce)
        self.LOSS_LAG = 2
        self._loss_ring = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(self.LOSS_LAG + 1)]
        self._loss_ev = [None] * (self.LOSS_LAG + 1)
        # input pipelining: the next batch is copied host->device on a side stream while the current step runs
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._staged = None          # (key, device dict | flat device buffer, ready event)
        self._stage_flat = None      # device image of the PackedBatch being prefetched
        self._static_flat = None     # flat device buffer behind the captured graph's static inputs
        # lookahead grouping (graph mode, object-level transformer): the step graph groups (FPS + ball query) the NEXT
        # batch's points in a branch beside its optimizer pass and reads its own grouping from static tensors
        # (see _lookahead_setup)
        self._la = None
        self.lookahead_hits = 0      # steps whose grouping had been computed ahead (diagnostics / tests)

    def _lpips_due(self, iteration: int) -> bool:
        """train_network.py:229-231, 288-296: the LPIPS term joins the loss after opt.start_lpips_after."""
        return self.cfg.opt.lambda_lpips != 0 and iteration > self.cfg.opt.start_lpips_after

    def load_checkpoint(self, path: str, load_optimizer: bool = True) -> Dict[str, float]:
        """ModelManager.load_checkpoint + everything of the step that holds pointers into the replaced state: the
        captured graph (parameter / moment / pointer tables) is dropped and re-captured at the next iteration."""
        info = self.model_manager.load_checkpoint(path, load_optimizer=load_optimizer)
        self.iteration = info["iteration"]
        self._graph, self._static, self._staged, self._static_flat, self._la = None, None, None, None, None
        return info

    # ---------------------------------------------------------------------------------------------
    def render_validation_views(self, gaussian_splats, data) -> Tuple[torch.Tensor, torch.Tensor]:
        """train_network.py:392-448, batched: views [input_images:] of every object in one call."""
        bg = self.validation_manager.background
        sl = slice(int(self.cfg.data.input_images), None)
        out = render_batch_predicted(gaussian_splats, data["world_view_transforms"], data["full_proj_transforms"],
                                     data["camera_centers"], bg, self.cfg, view_slice=sl)
        rendered = out["render"]
        gt = data["gt_images"][:, sl]          # (B,V',3,H,W) view, float32 or uint8: the fused loss reads it in place
        if self.cfg.opt.loss != "focal_l2":
            gt = gt.reshape(-1, *gt.shape[2:])
            gt = gt.float().div(255.0) if gt.dtype == torch.uint8 else gt
        return rendered.reshape(-1, *rendered.shape[2:]), gt

    def _forward_backward(self, data, after_forward=None) -> torch.Tensor:
        """`after_forward`: called once the backbone's forward has been enqueued (lookahead grouping forks there)."""
        mm = self.model_manager
        model_inputs = prepare_model_inputs(data, self.cfg, self.bs_per_gpu, self.device)
        mm.forward_model.train()
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                splats = mm.forward_model(**model_inputs)
            splats = {k: ([t.float() for t in v] if isinstance(v, (list, tuple)) else v.float()) for k, v in splats.items()}
        else:
            splats = mm.forward_model(**model_inputs)
        if after_forward is not None:
            after_forward()
        rendered, gt = self.render_validation_views(splats, data)
        loss = self.validation_manager.calculate_losses(rendered, gt, self.iteration)["total_loss"]
        backward_unit(loss)
        return loss.detach()

    # ------------------------------------------------------------------------------------------ data parallelism
    def _setup_grad_sync(self) -> None:
        """grad_sync.GradSync over [transformer-stack parameters | everything else]: the stack's gradients are all-reduced
        from inside the backward (fused_encoder.GRAD_READY_HOOK), the rest after it; the 1/world factor is applied inside
        the optimizer kernels (FusedClipAdamW.grad_scale)."""
        from . import fused_encoder
        from .grad_sync import GradSync
        enc = getattr(getattr(self.model_manager.model, "point_network", None), "encoder", None)
        tb = getattr(getattr(enc, "blocks", None), "blocks", None)
        early = fused_encoder.stack_parameters(tb) if (tb is not None and fused_encoder.supports(tb)) else []
        direct = bool(early) and not os.environ.get("UP3D_NO_DIRECT_GRADS")
        order = fused_encoder.stack_grad_order(len(tb)) if direct else None
        self._grad_sync = GradSync(self.params, early, self.device, overlap=not os.environ.get("UP3D_NO_EARLY_SYNC"),
                                   early_order=order)
        self.model_manager.optimizer.grad_scale = 1.0 / self.world
        if direct:
            # the stack's backward writes its gradients straight into the exchange buffer (no 113 MB pack)
            gs_ = self._grad_sync

            def grad_buffers(depth, C, Hd):
                flat = gs_.flat_early
                sizes = [depth * 3 * C * C, depth * C * C, depth * Hd * C, depth * C * Hd, depth * (6 * C + Hd)]
                if depth != len(tb) or sum(sizes) != flat.numel():
                    return None
                a, b, c, d, e = torch.split(flat, sizes)
                return (a.view(depth, 3 * C, C), b.view(depth, C, C), c.view(depth, Hd, C), d.view(depth, C, Hd), e)
            fused_encoder.GRAD_BUFFERS = grad_buffers
        if self._grad_sync.overlap:
            fused_encoder.GRAD_READY_HOOK = self._grad_sync.early_hook
            if os.environ.get("UP3D_CHUNKED_SYNC"):
                # opt-in: blocks' gradients leave in chunks of GRAD_CHUNK_BLOCKS while the backward of earlier blocks still
                # runs.  Measured at N = 2: 3.32 ms/step vs 3.19 ms with the single in-backward all-reduce (16 extra
                # weight-gradient GEMM launches and 3 extra NCCL launches cost more than the shorter tail saves there)
                per = fused_encoder.PARAMS_PER_BLOCK
                gs = self._grad_sync
                fused_encoder.GRAD_CHUNK_HOOK = lambda first_block, grads: gs.early_chunk_hook(first_block * per, grads)

    def _allreduce_grads(self) -> None:
        """N > 1: SUM the gradients over ranks (the optimizer applies 1/world) and point every p.grad at its slice of
        the flat buffer."""
        if self.world == 1:
            return
        if self._grad_sync is None:
            self._setup_grad_sync()
        self._grad_sync.finish()

    def _clip_and_step(self) -> None:
        """368-390 + 343-344 in two launches (optim.FusedClipAdamW): total norm on the device; non-finite -> the
        update kernel leaves parameters, moments and the step counter untouched (the reference's `continue`); else
        gradients are scaled to max_norm 1.0 inside the AdamW pass, which also rewrites the bf16 weight shadows."""
        mm = self.model_manager
        mm.optimizer.step()
        mm.optimizer.zero_grad(set_to_none=True)   # grads are re-materialised at the same graph-pool addresses on replay

    def _step_body(self, data) -> torch.Tensor:
        mark = getattr(self.model_manager.optimizer, "mark_step_start", None)
        if mark is not None:
            mark()
        la = self._lookahead_graph_head()
        fork = None
        if la is not None:
            def fork():
                # grouping of the next batch as a branch that starts when the backbone's forward is done and runs beside
                # the rasterizer / loss / backward (measured placements: beside the tokenizer at the head of the step it
                # slowed those latency-bound kernels by as much as it saved; beside the optimizer pass it took 230 us
                # against the optimizer's 170 us)
                la["branch"].wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(la["branch"]):
                    la["module"].group(la["next_points"], out=la["next"])
        loss = self._forward_backward(data, after_forward=fork)
        self._allreduce_grads()
        self._clip_and_step()
        if la is not None:
            torch.cuda.current_stream().wait_stream(la["branch"])
        return loss

    @torch.no_grad()
    def _snapshot_state(self):
        m, opt = self.model_manager.model, self.model_manager.optimizer
        tens = list(m.parameters()) + list(m.buffers())
        had_state = bool(getattr(opt, "_built", False))
        moments = [(opt.state[p]["exp_avg"].clone(), opt.state[p]["exp_avg_sq"].clone())
                   for p in self.params] if had_state else None
        return {"tens": [t.detach().clone() for t in tens], "moments": moments,
                "opt_state": opt._state.clone() if had_state else None}

    @torch.no_grad()
    def _restore_state(self, snap) -> None:
        m, opt = self.model_manager.model, self.model_manager.optimizer
        for t, s in zip(list(m.parameters()) + list(m.buffers()), snap["tens"]):
            t.copy_(s)
        for p in self.params:
            st = opt.state.get(p)
            if not st or "exp_avg" not in st:
                continue
            if snap["moments"] is None:
                st["exp_avg"].zero_(); st["exp_avg_sq"].zero_()
        if snap["moments"] is not None:
            for p, (a, b) in zip(self.params, snap["moments"]):
                opt.state[p]["exp_avg"].copy_(a); opt.state[p]["exp_avg_sq"].copy_(b)
            opt._state.copy_(snap["opt_state"])
        elif getattr(opt, "_built", False):
            opt._state.zero_()
        for mod in m.modules():          # bf16 weight shadows follow the restored masters
            w16 = getattr(mod, "_w16", None)
            if w16 is not None:
                w16.copy_(mod.weight.detach().reshape(w16.shape))
                if getattr(mod, "_b16", None) is not None:
                    mod._b16.copy_(mod.bias.detach())

    # ---------------------------------------------------------------------------------------------
    def _copy_into_static(self, data) -> None:
        if self._la is not None and self._la["for"] is Trainer._RESIDENT:
            self._la["for"] = None                                      # the resident batch is being replaced
        if torch.is_tensor(data):                                       # flat device image of a PackedBatch
            if self._static_flat is None:
                raise RuntimeError("the step graph was captured from an unpacked batch; pass the same kind every step")
            self._static_flat.copy_(data, non_blocking=True)            # ONE device->device copy
            return

        def cp(dst, src):
            if isinstance(dst, dict):
                for k in dst:
                    cp(dst[k], src[k])
            else:
                dst.copy_(src, non_blocking=True)
        cp(self._static, data)

    # ------------------------------------------------------------------------------------------ lookahead grouping
    # FPS + ball-query grouping depends on the input coordinates alone and is a 140 us chain of latency-bound kernels on
    # 16 SMs at the head of every step.  In graph mode the captured step therefore (1) starts by taking its grouping from
    # the static tensors `next` (one small D2D copy into `cur`), and (2) carries a branch, forked after the backbone's
    # forward, that groups the points in the static buffer `next_points` into `next` for the step after it.
    # train_iteration(data, prefetch=nxt) feeds `next_points` with nxt's coordinates; when a step's grouping was not
    # prepared that way (first step, no prefetch) it is computed in line before the replay, as without the lookahead.
    _RESIDENT = object()

    @staticmethod
    def _points_of(batch):
        pc = batch["point_cloud"]
        pts = pc["pos"] if isinstance(pc, dict) else pc
        return pts[:, :, :3]

    def _lookahead_setup(self) -> None:
        """Off (None) for backbones without a SubsampleGroup tokenizer or with UP3D_NO_LOOKAHEAD=1."""
        from .backbone import SubsampleGroup
        enc = getattr(getattr(self.model_manager.model, "point_network", None), "encoder", None)
        gd = getattr(enc, "group_divider", None)
        if not isinstance(gd, SubsampleGroup) or os.environ.get("UP3D_NO_LOOKAHEAD"):
            self._la = None
            return
        pts = self._points_of(self._static)
        B, N = pts.shape[0], pts.shape[1]
        n_neigh, n_center = B * 3 * gd.num_groups * gd.group_size, B * gd.num_groups * 3

        def pair():
            flat = torch.empty(n_neigh + n_center, dtype=torch.float32, device=self.device)
            return flat, (flat[:n_neigh].view(B, 3, gd.num_groups, gd.group_size), flat[n_neigh:].view(B, gd.num_groups, 3))
        cur_flat, cur = pair()
        nxt_flat, nxt = pair()
        self._la = {"module": gd, "cur_flat": cur_flat, "cur": cur, "next_flat": nxt_flat, "next": nxt,
                    "next_points": pts.float().contiguous().clone(),          # (B,N,3) fp32, read by the graph's tail
                    "pts_stage": None, "pts_ready": None, "pts_free": None,  # host->device staging of the next points
                    "branch": torch.cuda.Stream(device=self.device), "h2d": torch.cuda.Stream(device=self.device),
                    "for": None}
        gd.group(self._la["next_points"], out=nxt)                            # the first replay takes it from `next`

    def _lookahead_feed(self, prefetch) -> None:
        """Host side, top of a step: start the H2D copy of the NEXT batch's coordinates into the staging buffer (its own
        stream; it waits only for the previous consumer of the staging buffer, so it runs while the previous step does)."""
        la = self._la
        host = self._points_of(prefetch.data if isinstance(prefetch, PackedBatch) else prefetch)
        if host.is_cuda:
            la["pts_stage"], la["pts_ready"] = host, None
            return
        if not host.is_contiguous():
            host = host.contiguous()
        if la["pts_stage"] is None or la["pts_stage"].shape != host.shape or la["pts_stage"].dtype != host.dtype or \
                not la["pts_stage"].is_contiguous():
            la["pts_stage"] = torch.empty(host.shape, dtype=host.dtype, device=self.device)
            la["pts_free"] = None
        st = la["h2d"]
        if la["pts_free"] is not None:
            st.wait_event(la["pts_free"])
        with torch.cuda.stream(st):
            la["pts_stage"].copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st)
        la["pts_ready"] = ev

    def _lookahead_before_replay(self, key, prefetch) -> None:
        """Current stream, after the step's inputs are in the static buffers: make `next` hold THIS batch's grouping (it
        does when the previous replay's tail was fed with it; else group in line) and `next_points` the next batch's
        coordinates for this replay's tail."""
        la = self._la
        cur = torch.cuda.current_stream()
        if la["for"] is not None and la["for"] is key:
            self.lookahead_hits += 1
        else:
            la["module"].group(self._points_of(self._static), out=la["next"])
        if prefetch is not None:
            if la["pts_ready"] is not None:
                cur.wait_event(la["pts_ready"])
            la["next_points"].copy_(la["pts_stage"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cur)
            la["pts_free"] = ev
        # without a prefetch the tail regroups whatever `next_points` holds; the result is not used (for = None)
        la["for"] = prefetch

    def _lookahead_graph_head(self):
        """Inside the captured step: cur <- next.  Returns the lookahead state when it is active for this capture."""
        la = self._la
        if la is None or la["module"].lookahead is None:
            return None
        la["cur_flat"].copy_(la["next_flat"], non_blocking=True)
        return la

    def replay_resident(self) -> None:
        """One step on the batch already resident in the graph's static inputs (bench.py's `value`): the replay takes its
        grouping from the previous replay's tail branch and regroups the same coordinates for the next one."""
        if self._graph is None:
            raise RuntimeError("replay_resident needs a captured step (run train_iteration once in graph mode)")
        la = self._la
        if la is not None:
            if la["for"] is Trainer._RESIDENT:
                self.lookahead_hits += 1
            else:
                pts = self._points_of(self._static)
                la["next_points"].copy_(pts, non_blocking=True)
                la["module"].group(pts, out=la["next"])
                la["for"] = Trainer._RESIDENT
        self._graph.replay()

    def pack_batch(self, data) -> PackedBatch:
        """Host batch dict -> PackedBatch (one flat pinned buffer): the form `train_iteration` moves with ONE H2D copy."""
        return PackedBatch(data)

    def _stage(self, data) -> None:
        """Start the H2D copy of `data` (pinned host dict or PackedBatch) on the copy stream."""
        if self._staged is not None and self._staged[0] is data:
            return
        self._copy_stream.wait_stream(torch.cuda.current_stream())   # staging buffers may still be read
        with torch.cuda.stream(self._copy_stream):
            if isinstance(data, PackedBatch):
                if self._stage_flat is None or self._stage_flat.numel() != data.nbytes:
                    self._stage_flat = torch.empty(data.nbytes, dtype=torch.uint8, device=self.device)
                self._stage_flat.copy_(data.flat, non_blocking=True)            # ONE host->device copy
                dev = self._stage_flat
            else:
                dev = _to_device(data, self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._staged = (data, dev, ev)

    def _take_staged(self, data):
        """Device copy of `data`: the prefetched one when it matches, else a fresh (stream-ordered) H2D copy."""
        if self._staged is not None and self._staged[0] is data:
            _, dev, ev = self._staged
            cur = torch.cuda.current_stream()
            cur.wait_event(ev)
            if torch.is_tensor(dev):                                    # flat staging buffer of a PackedBatch
                self._staged = None
                return dev
            for t in (list(dev["point_cloud"].values()) if isinstance(dev.get("point_cloud"), dict) else []) + \
                    [v for v in dev.values() if torch.is_tensor(v)]:
                t.record_stream(cur)     # allocated on the copy stream, consumed on the compute stream
            self._staged = None
            return dev
        if isinstance(data, PackedBatch):
            return data.flat.to(self.device, non_blocking=True)
        return _to_device(data, self.device)

    def train_iteration(self, data, read_loss: bool = True, prefetch=None):
        """One optimisation step on a host (ideally pinned) batch dict.  Returns the loss (float) when
        `read_loss` (a D2H read, as the reference's logging does) else the device scalar.  `prefetch`: the NEXT
        step's host batch; its H2D copy is overlapped with this step's compute."""
        self.iteration += 1
        mm = self.model_manager
        if self._lpips_due(self.iteration):
            # evaluated on the host EVERY iteration: a captured graph would otherwise keep replaying the LPIPS-free loss
            raise NotImplementedError("LPIPS loss needs VGG weights (not shipped); set opt.lambda_lpips=0")
        if not self.use_cuda_graph:
            dev = self._take_staged(data)
            if torch.is_tensor(dev):
                dev = data.views(dev.clone())       # eager mode: the staging buffer is rewritten by the next prefetch
            if prefetch is not None:
                self._stage(prefetch)
            loss = self._step_body(dev)
        else:
            if self._static is None:
                if isinstance(data, PackedBatch):
                    # static inputs = views into one flat device buffer (refreshed by a single copy per step)
                    self._static_flat = data.flat.to(self.device)
                    self._static = data.views(self._static_flat)
                    self._static_sig = data.signature()
                else:
                    self._static = _to_device(data, self.device, non_blocking=False)
                # warm-up on a side stream (allocator, cuBLAS handles, layout cache, lazy grads), then capture.  The
                # warm-up steps are real optimizer steps on the first batch: model, BatchNorm buffers, optimizer
                # moments and the step counter are snapshotted and put back, so iteration 1 starts from the same state
                # as in eager mode (only the RNG stream has advanced).
                self._lookahead_setup()
                gd = self._la["module"] if self._la is not None else None
                try:
                    if gd is not None:
                        gd.lookahead = self._la["cur"]      # only while warming up / capturing (see SubsampleGroup)
                    snap = self._snapshot_state()
                    s = torch.cuda.Stream()
                    s.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(s):
                        for _ in range(3):
                            self._step_body(self._static)
                    torch.cuda.current_stream().wait_stream(s)
                    torch.cuda.synchronize()
                    self._restore_state(snap)
                    self._graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
                        self._loss_buf.copy_(self._step_body(self._static))
                finally:
                    if gd is not None:
                        gd.lookahead = None
            if isinstance(data, PackedBatch) and getattr(self, "_static_sig", None) != data.signature():
                raise RuntimeError("PackedBatch layout differs from the one the step graph was captured with")
            if self._la is not None and prefetch is not None:
                self._lookahead_feed(prefetch)                  # the next batch's coordinates start moving first
            self._copy_into_static(self._take_staged(data))     # D2D when prefetched, H2D otherwise
            if self._la is not None:
                self._lookahead_before_replay(data, prefetch)
            if prefetch is not None:
                self._stage(prefetch)
            self._graph.replay()
            loss = self._loss_buf
        # NB (differs from the reference on a NaN/Inf step): the reference `continue`s before scheduler.step() and
        # ema.update() (train_network.py:336-352); here the skip decision lives on the device (found_inf inside the fused
        # optimizer, no host sync), so the StepLR counter and the EMA schedule advance on a skipped step as well.  The
        # parameters, moments and Adam step counter are untouched on such a step, as in the reference.
        mm.scheduler_step()
        if mm.ema:
            mm.ema.update()
        if read_loss == "lagged":
            # D2H read of THIS step's loss is enqueued now and consumed LOSS_LAG steps later (logging lags), so the host
            # never stalls the launch pipeline; returns the loss of step (this - LOSS_LAG) (None on the first calls)
            n = self.LOSS_LAG + 1
            slot = self.iteration % n
            old_slot = (self.iteration + 1) % n            # = (iteration - LOSS_LAG) mod n: the oldest read in flight
            prev = self._loss_ev[old_slot]
            out = None
            if prev is not None:
                prev.synchronize()
                out = float(self._loss_ring[old_slot])
                self._loss_ev[old_slot] = None
            self._loss_ring[slot].copy_(loss, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._loss_ev[slot] = ev
            return out
        if read_loss:
            self._loss_host.copy_(loss, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return float(self._loss_host)
        return loss

    def last_loss(self) -> float:
        """Blocks for the most recent step's lagged loss read."""
        slot = self.iteration % (self.LOSS_LAG + 1)
        if self._loss_ev[slot] is None:
            raise RuntimeError("no lagged loss read in flight")
        self._loss_ev[slot].synchronize()
        return float(self._loss_ring[slot])
