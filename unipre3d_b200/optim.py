"""Gradient clipping + AdamW of the reference trainer as two launches.

/root/reference/train_network.py:156-159 builds `torch.optim.AdamW(groups, lr=0.0, eps=1e-15, betas=cfg.opt.betas)`;
every iteration runs `_check_and_clip_gradients` (368-390: per-parameter NaN/Inf scans -> skip the step, else
`clip_grad_norm_(max_norm=1.0)`) and `optimizer.step()` (343-344).  Eagerly that is ~25 launches that move the 29 M
parameters' gradients three times (norm, scale, update) and their bf16 shadows once more; `up3d_adamw_step` does it
in one sum-of-squares pass and ONE read-modify-write pass (param, exp_avg, exp_avg_sq, grad -> + bf16 shadow), with
the clip coefficient, the skip-on-non-finite rule and the step counter all on the device (no host sync, graph-safe).

`FusedClipAdamW` subclasses torch.optim.AdamW so `state_dict()` / `load_state_dict()` keep the reference
checkpoint layout (train_network.py:200-210): per-parameter `step`, `exp_avg`, `exp_avg_sq`.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


class FusedClipAdamW(torch.optim.AdamW):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm: float = 1.0,
                 shadows: Optional[Dict[int, torch.Tensor]] = None):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.max_norm = float(max_norm)
        self.grad_scale = 1.0          # gradients are read as grad_scale * p.grad (1/world when p.grad holds the rank SUM)
        self._shadows = shadows or {}
        self._built = False
        self._ring, self._ring_ev, self._ring_i, self._frozen = [], [], 0, []
        self._last_ptrs = None

    # ------------------------------------------------------------------------------------------------------
    def set_shadows(self, shadows: Dict[int, torch.Tensor]) -> None:
        """{id(master parameter): bf16 tensor of the same shape} refreshed by the update kernel."""
        if self._built:
            raise RuntimeError("set_shadows must be called before the first step")
        self._shadows = dict(shadows)

    def _build(self) -> None:
        ps, group_of = [], []
        for gi, g in enumerate(self.param_groups):
            if g.get("amsgrad") or g.get("maximize"):
                raise NotImplementedError("FusedClipAdamW: amsgrad / maximize are not on the reference path")
            for p in g["params"]:
                if not p.requires_grad:
                    continue
                if p.grad is None:
                    raise RuntimeError("FusedClipAdamW: a trainable parameter has no gradient at the first step")
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedClipAdamW: parameters must be contiguous fp32 CUDA tensors (no CPU fallback)")
                ps.append(p)
                group_of.append(gi)
        if not ps:
            raise RuntimeError("FusedClipAdamW: no parameters")
        dev = ps[0].device
        self._params, self._dev = ps, dev
        g0 = self.param_groups[0]
        if any(tuple(g["betas"]) != tuple(g0["betas"]) or g["eps"] != g0["eps"] or g["weight_decay"] != g0["weight_decay"]
               for g in self.param_groups):
            raise NotImplementedError("FusedClipAdamW: betas/eps/weight_decay must be shared by all groups")
        # device scalars: [fp64 sum of squares | step, total_norm, clip_coef, found_inf | u32 ticket | pad]
        self._state = torch.zeros(8, dtype=torch.float32, device=dev)
        self._step_view = self._state[2]
        # learning rates live on the device (a captured graph must see StepLR updates): group["lr"] becomes a view
        lrs = torch.zeros(len(self.param_groups), dtype=torch.float32, device=dev)
        for gi, g in enumerate(self.param_groups):
            lrs[gi] = g["lr"].to(dev, torch.float32) if torch.is_tensor(g["lr"]) else float(g["lr"])
            g["lr"] = lrs[gi]
        self._lrs = lrs
        for p in ps:
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            elif "step" in st:
                self._step_view.fill_(float(st["step"]))       # resumed from a checkpoint
            st["step"] = self._step_view
        chunk = int(_lib.lib.up3d_adamw_chunk_elems())
        ct, cs = [], []
        for i, p in enumerate(ps):
            for s in range(0, p.numel(), chunk):
                ct.append(i)
                cs.append(s)
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)
        i64 = lambda v: torch.tensor(v, dtype=torch.int64, device=dev)
        self._chunk_tensor, self._chunk_start, self._n_chunks = i32(ct), i32(cs), len(ct)
        self._numel = i64([p.numel() for p in ps])
        self._group = i32(group_of)
        self._p_ptrs = i64([p.data_ptr() for p in ps])
        self._m_ptrs = i64([self.state[p]["exp_avg"].data_ptr() for p in ps])
        self._v_ptrs = i64([self.state[p]["exp_avg_sq"].data_ptr() for p in ps])
        sh = []
        for p in ps:
            s = self._shadows.get(id(p))
            if s is not None and not (s.dtype == torch.bfloat16 and s.is_contiguous() and s.numel() == p.numel()):
                raise RuntimeError("FusedClipAdamW: a shadow must be a contiguous bf16 tensor of its parameter's size")
            sh.append(0 if s is None else s.data_ptr())
        self._s_ptrs = i64(sh) if any(sh) else None
        self._g_ptrs = torch.zeros(len(ps), dtype=torch.int64, device=dev)
        self._capture_bufs = [torch.empty(len(ps), dtype=torch.int64).pin_memory() for _ in range(2)]
        # a fresh (all-zero) gradient-pointer table must be re-uploaded even when the gradients kept their addresses
        # (data-parallel mode: every p.grad is a fixed view of GradSync's flat buffer)
        self._last_ptrs = None
        self._ring, self._ring_ev, self._ring_i = [], [], 0
        self._side = torch.cuda.Stream(device=dev)      # carries the captured pointer-table upload (mark_step_start)
        self._built = True

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._built = False                    # moments were replaced: rebuild the pointer tables at the next step
        self._last_ptrs = None                 # ... and re-upload the gradient pointers into the new table

    def mark_step_start(self) -> None:
        """Called by the trainer at the top of a step that is being captured: the pointer-table upload of this step (a
        memcpy node whose content does not depend on anything the step computes) is then attached to this point of the
        graph instead of the end of the backward, which takes its ~10 us launch latency off the critical path."""
        self._start_ev = None
        if getattr(self, "_built", False) and torch.cuda.is_current_stream_capturing():
            self._start_ev = torch.cuda.Event()
            self._start_ev.record()

    def _upload_grad_ptrs(self) -> None:
        """The gradient tensors are re-allocated by autograd every step (at fixed graph-pool addresses under CUDA-graph
        capture), so their pointer table is refreshed from a pinned staging buffer on the launching stream."""
        ptrs = [p.grad.data_ptr() for p in self._params]
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing and ptrs == self._last_ptrs:
            return
        if capturing:
            # a captured memcpy node re-reads its pinned source on every replay: use a buffer that is never rewritten
            # (allocated in _build, i.e. during the eager warm-up steps -- no host allocation while capturing)
            if not self._capture_bufs:
                raise RuntimeError("FusedClipAdamW: run at least one eager step before capturing; at most 2 captures")
            buf = self._capture_bufs.pop()
            buf.copy_(torch.tensor(ptrs, dtype=torch.int64))
            self._frozen.append(buf)
            ev = getattr(self, "_start_ev", None)
            if ev is not None:
                cur = torch.cuda.current_stream()
                self._side.wait_event(ev)
                with torch.cuda.stream(self._side):
                    self._g_ptrs.copy_(buf, non_blocking=True)
                cur.wait_stream(self._side)
                self._start_ev = None
            else:
                self._g_ptrs.copy_(buf, non_blocking=True)
            self._last_ptrs = None
            return
        if len(self._ring) < 4:
            self._ring.append(torch.empty(len(ptrs), dtype=torch.int64).pin_memory())
            self._ring_ev.append(None)
        k = self._ring_i % len(self._ring)
        self._ring_i += 1
        if self._ring_ev[k] is not None:
            self._ring_ev[k].synchronize()     # the copy that last used this slot has finished
        self._ring[k].copy_(torch.tensor(ptrs, dtype=torch.int64))
        self._g_ptrs.copy_(self._ring[k], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._ring_ev[k] = ev
        self._last_ptrs = ptrs

    # ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        """clip_grad_norm_(max_norm) + skip-on-NaN/Inf + AdamW.  Returns None (no closure support)."""
        if closure is not None:
            raise NotImplementedError("FusedClipAdamW.step does not take a closure")
        if not self._built:
            self._build()
        for p in self._params:
            g = p.grad
            if g is None:
                raise RuntimeError("FusedClipAdamW: a parameter lost its gradient (the work list is static)")
            if not (g.dtype == torch.float32 and g.is_contiguous()):
                p.grad = g = g.float().contiguous()
        self._upload_grad_ptrs()
        g0 = self.param_groups[0]
        with torch.cuda.device(self._dev):
            check(_lib.lib.up3d_adamw_step(len(self._params), self._n_chunks, ptr(self._chunk_tensor), ptr(self._chunk_start),
                                           ptr(self._numel), ptr(self._p_ptrs), ptr(self._g_ptrs), ptr(self._m_ptrs),
                                           ptr(self._v_ptrs), ptr(self._s_ptrs), ptr(self._group), ptr(self._lrs),
                                           float(g0["betas"][0]), float(g0["betas"][1]), float(g0["eps"]),
                                           float(g0["weight_decay"]), self.max_norm, float(self.grad_scale),
                                           ptr(self._state), stream_ptr()),
                  launches=2)
        return None

    @torch.no_grad()
    def time_passes(self, reps: int = 5):
        """CUDA-event timing of the two launches on the current stream (bench.py roofline): -> (sumsq_ms, apply_ms,
        algorithmic bytes of the apply pass).  Every repetition is a REAL optimizer step on the current gradients."""
        if not self._built:
            self._build()
        self._upload_grad_ptrs()
        g0, L = self.param_groups[0], _lib.lib
        head = (len(self._params), self._n_chunks, ptr(self._chunk_tensor), ptr(self._chunk_start), ptr(self._numel))
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t_sq = t_ap = 0.0
        with torch.cuda.device(self._dev):
            for _ in range(reps):
                ev[0].record()
                check(L.up3d_grad_sumsq(*head, ptr(self._g_ptrs), float(self.grad_scale), ptr(self._state), stream_ptr()),
                      launches=1)
                ev[1].record()
                check(L.up3d_adamw_apply(*head, ptr(self._p_ptrs), ptr(self._g_ptrs), ptr(self._m_ptrs), ptr(self._v_ptrs),
                                         ptr(self._s_ptrs), ptr(self._group), ptr(self._lrs), float(g0["betas"][0]),
                                         float(g0["betas"][1]), float(g0["eps"]), float(g0["weight_decay"]), self.max_norm,
                                         float(self.grad_scale), ptr(self._state), stream_ptr()), launches=1)
                ev[2].record()
                torch.cuda.synchronize()
                t_sq += ev[0].elapsed_time(ev[1])
                t_ap += ev[1].elapsed_time(ev[2])
        n = sum(p.numel() for p in self._params)
        n_sh = sum(p.numel() for p in self._params if id(p) in self._shadows)
        # read param, grad, exp_avg, exp_avg_sq (16 B) + write param, exp_avg, exp_avg_sq (12 B) [+ 2 B bf16 shadow]
        return t_sq / reps, t_ap / reps, 28 * n + 2 * n_sh

    # device scalars of the most recent step (reading them is a D2H sync: diagnostics / tests only)
    def last_total_norm(self) -> float:
        return float(self._state[3])

    def last_found_inf(self) -> bool:
        return bool(self._state[5] != 0)

    def step_count(self) -> int:
        return int(self._state[2])
