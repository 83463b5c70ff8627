"""Gradient exchange of the data-parallel step (the reference wraps the model in DistributedDataParallel,
/root/reference/pointcept/engines/defaults.py:22-43, train_network.py:183-186: bucketed all-reduce + division by world).

`GradSync` keeps one flat fp32 buffer [early parameters | ready parameters | late parameters]:

  * `early_hook(grads)` -- called from INSIDE the backward (by the transformer stack's single backward node, whose
    gradients are ~97 % of all gradient bytes and complete ~0.35 ms before the backward ends): starts the all-reduce of
    [early | ready] on a second communicator (and, on CUDA, a second stream), so it overlaps the rest of the backward
    instead of following it.  The early gradients are either already in place (the producer wrote them into
    `flat_early`, see `early_order`) or packed here; the READY ones -- parameters of everything downstream of the stack,
    whose gradients are final when the stack's backward runs -- are packed on the communication stream, off the
    critical path;
  * `finish()` -- after the backward: packs and all-reduces the LATE gradients (the tokenizer in front of the stack: a
    few tensors), joins the early all-reduce and points every `p.grad` at its slice of the flat buffer.

Which of the non-early parameters are ready is measured, not assumed: on the first hooked backward outside a stream
capture their gradients are snapshotted at hook time and compared with the final ones in `finish()`; a parameter is
"ready" when the two are identical on every rank.  Until then all of them are treated as late.

Gradients are SUMMED; the 1/world of DDP's mean is applied by the optimizer (`FusedClipAdamW.grad_scale`), which saves a
pass over the 118 MB buffer.  Device-agnostic: the same code runs over gloo on CPU tensors (tests/test_dist_cpu.py).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, params: Sequence[torch.nn.Parameter], early_params: Sequence[torch.nn.Parameter], device,
                 overlap: bool = True, early_order: Optional[Sequence[int]] = None):
        """early_order: permutation of range(len(early_params)) = the order in which the early parameters lie in the
        flat buffer (default: as given).  A producer that knows this layout can write its gradients straight into
        `flat_early` (fused_encoder.GRAD_BUFFERS); early_hook then skips the pack for them."""
        self.world = dist.get_world_size()
        self.device = torch.device(device)
        early_ids = {id(p) for p in early_params}
        self.early: List[torch.nn.Parameter] = list(early_params)
        self.rest: List[torch.nn.Parameter] = [p for p in params if id(p) not in early_ids]
        self.n_early, n_rest = sum(p.numel() for p in self.early), sum(p.numel() for p in self.rest)
        self.flat = torch.zeros(self.n_early + n_rest, dtype=torch.float32, device=self.device)
        self.flat_early, self.flat_rest = self.flat[:self.n_early], self.flat[self.n_early:]
        order = list(early_order) if early_order is not None else list(range(len(self.early)))
        if sorted(order) != list(range(len(self.early))):
            raise ValueError("early_order must be a permutation of the early parameters")
        self._early_slices, off = [None] * len(self.early), 0
        for k in order:
            p = self.early[k]
            self._early_slices[k] = (off, off + p.numel(), tuple(p.shape))
            off += p.numel()
        # non-early parameters: [ready | late]; everything is late until the calibration step has run
        self._ready_idx: List[int] = []
        self._late_idx: List[int] = list(range(len(self.rest)))
        self._rest_views: List[torch.Tensor] = []
        self._n_ready = 0
        self._layout_rest()
        self.overlap = bool(overlap and self.early)
        self.calibrated = not self.overlap or not self.rest
        self._snap = None                          # calibration: {rest index: gradient at hook time}
        self._ready_sent = False                   # this backward's hook carried the ready gradients
        self._pending = None                       # "stream" (CUDA) or a dist.Work handle (CPU)
        self._chunks_seen = 0                      # early parameters whose chunk all-reduce was started this backward
        self._cpu_works = []
        self.is_cuda = self.device.type == "cuda"
        self._comm_stream = torch.cuda.Stream(device=self.device) if (self.is_cuda and self.overlap) else None
        # second communicator: the early all-reduce must not queue behind (or in front of) the SyncBatchNorm collectives
        # the rest of the backward issues on the default one
        self._pg = dist.new_group() if self.overlap else None

    # ------------------------------------------------------------------------------------------------ layout
    def _layout_rest(self) -> None:
        views, off = [None] * len(self.rest), self.n_early
        for i in self._ready_idx + self._late_idx:
            p = self.rest[i]
            views[i] = self.flat[off: off + p.numel()].view_as(p)
            off += p.numel()
        self._rest_views = views
        self._n_ready = sum(self.rest[i].numel() for i in self._ready_idx)

    def _early_views(self):
        return [self.flat[a:b].view(shape) for a, b, shape in self._early_slices]

    @staticmethod
    def _pack(views, grads) -> None:
        """Copy the gradients into their slices, except those the producer already wrote there."""
        todo = [(v, g) for v, g in zip(views, grads)
                if not (g.data_ptr() == v.data_ptr() and g.is_contiguous() and g.dtype == v.dtype)]
        # the multi-tensor copy kernel walks 64 K-element chunks with one CTA each: fine for the many small tensors, slow
        # (a handful of CTAs) for the few large ones, which get a full-grid copy of their own
        small = [(v, g) for v, g in todo if v.numel() < 32768]
        for v, g in todo:
            if v.numel() >= 32768:
                v.copy_(g)
        if small:
            torch._foreach_copy_([v for v, _ in small], [g for _, g in small])

    def _capturing(self) -> bool:
        return self.is_cuda and torch.cuda.is_current_stream_capturing()

    def _join(self) -> None:
        if self._pending is None:
            return
        if self.is_cuda:
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        elif self._pending == "works":
            for w in self._cpu_works:
                w.wait()
            self._cpu_works = []
        else:
            self._pending.wait()
        self._pending = None

    # ------------------------------------------------------------------------------------------------ backward hooks
    def early_hook(self, grads):
        """grads: the early parameters' gradients, in `early_params` order.  Returns what autograd should see."""
        if not self.overlap or len(grads) != len(self._early_slices):
            return grads
        self._join()                              # a previous backward that was never consumed (diagnostic passes)
        self._pack(self._early_views(), grads)
        ready = self._ready_idx if self.calibrated else []
        if not self.calibrated and not self._capturing():
            self._snap = {i: p.grad.detach().clone() for i, p in enumerate(self.rest) if p.grad is not None}
        if any(self.rest[i].grad is None for i in ready):
            raise RuntimeError("a parameter classified as ready has no gradient when the stack's backward runs")
        region = self.flat[:self.n_early + (self._n_ready if ready else 0)]
        self._ready_sent = bool(ready)
        if self.is_cuda:
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                if ready:                         # off the critical path: the backward goes on while these are packed
                    self._pack([self._rest_views[i] for i in ready], [self.rest[i].grad for i in ready])
                dist.all_reduce(region, op=dist.ReduceOp.SUM, group=self._pg)
            self._pending = "stream"
        else:
            if ready:
                self._pack([self._rest_views[i] for i in ready], [self.rest[i].grad for i in ready])
            self._pending = dist.all_reduce(region, op=dist.ReduceOp.SUM, group=self._pg, async_op=True)
        return grads                              # autograd adopts the originals (no copy); finish() re-points p.grad

    def early_chunk_hook(self, first_param: int, grads) -> None:
        """grads: gradients of early parameters [first_param, first_param + len(grads)) (early_params order), complete
        while the backward is still running: pack them and start the all-reduce of their slice of the flat buffer."""
        if not self.overlap:
            return
        if self._chunks_seen == 0:
            self._join()                          # a previous backward that was never consumed
        views = self._early_views()[first_param:first_param + len(grads)]
        torch._foreach_copy_(views, list(grads))
        a, b = self._early_slices[first_param][0], self._early_slices[first_param + len(grads) - 1][1]
        if self.is_cuda:
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self._pg)
            self._pending = "stream"
        else:
            w = dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.SUM, group=self._pg, async_op=True)
            self._cpu_works.append(w)
            self._pending = "works"
        self._chunks_seen += len(grads)

    # ------------------------------------------------------------------------------------------------ end of backward
    def finish(self) -> None:
        if any(p.grad is None for p in self.early + self.rest):
            raise RuntimeError("a trainable parameter received no gradient (data-parallel ranks would diverge)")
        if self._chunks_seen not in (0, len(self.early)):
            raise RuntimeError("chunked gradient exchange covered only part of the early parameters")
        ready_local = None
        if self._snap is not None:                # calibration step: which gradients were already final at hook time?
            ready_local = [int(i in self._snap and self._snap[i].shape == p.grad.shape
                               and bool(torch.equal(self._snap[i], p.grad))) for i, p in enumerate(self.rest)]
            self._snap = None
        if self._pending is None and self.early:  # the hook did not fire (module path / overlap off): reduce it now
            self._pack(self._early_views(), [p.grad for p in self.early])
            dist.all_reduce(self.flat_early, op=dist.ReduceOp.SUM)
            self._ready_sent = False
        self._chunks_seen = 0
        todo = self._late_idx if self._ready_sent else self._ready_idx + self._late_idx
        if todo:
            self._pack([self._rest_views[i] for i in todo], [self.rest[i].grad for i in todo])
            dist.all_reduce(self.flat[self.n_early + (self._n_ready if self._ready_sent else 0):], op=dist.ReduceOp.SUM)
        self._ready_sent = False
        self._join()
        for p, v in zip(self.early, self._early_views()):
            p.grad = v
        for p, v in zip(self.rest, self._rest_views):
            p.grad = v
        if ready_local is not None:
            # every rank must choose the same layout: ready = final at hook time on ALL ranks.  The views handed out
            # above keep this step's (old) layout of the same buffer; the new one applies from the next backward on.
            mask = torch.tensor(ready_local, dtype=torch.int32, device=self.device)
            dist.all_reduce(mask, op=dist.ReduceOp.MIN)
            mask = mask.tolist()
            self._ready_idx = [i for i, r in enumerate(mask) if r]
            self._late_idx = [i for i, r in enumerate(mask) if not r]
            self._layout_rest()                    # (the optimizer consumes this step's views before the next backward)
            self.calibrated = True
