"""The 16-Block transformer of the point backbone as ONE autograd node on B200.

Arithmetic = /root/reference/openpoints/models/backbone/transformer.py:89-121 (Block: x + drop_path(attn(norm1(x))),
x + drop_path(mlp(norm2(x)))) inside TransformerEncoder.forward's `x = block(x + pos)` loop (176-207), Attention
(36-77) and Mlp (10-33).  The eager module graph spends ~60 launches per Block and step on 1.5 MB tensors; here each
Block is 8 launches forward and ~17 backward:

    forward   ln_fwd (residual + DropPath scale + pos + LayerNorm, one pass) -> qkv GEMM -> attention -> proj GEMM(+bias)
              -> ln_fwd -> fc1 GEMM(+bias) -> gelu_fwd -> fc2 GEMM(+bias)
    backward  1 dX GEMM per Linear; the weight gradients of each Linear for ALL blocks are one batched GEMM at the end
              (their operands are written into (depth, T, .) stacks as they are produced), gelu_bwd (+ fc1 bias
              gradient), ln_bwd (+ residual add, LayerNorm parameter gradients, dpos accumulation, DropPath scale,
              cast and the bias gradient of the Linear in front -- all in the same pass), SDPA backward.

The hand-written passes are libunipre3d_b200's up3d_ln_fwd / up3d_ln_bwd / up3d_gelu_* / up3d_scale_cast_colsum
(csrc/backbone.cu) and up3d_attn_fwd / up3d_attn_bwd (csrc/attention.cu: bf16 operands, 1 launch each way, reading
the qkv GEMM output and writing the dqkv GEMM input in place); the dense GEMMs are library calls (cuBLASLt), as is
the attention of the fp32 reference-precision mode (torch SDPA).  The residual
stream, LayerNorm statistics and all parameter gradients are fp32; GEMM operands are fp32 (reference precision) or
bf16 (tensor cores, persistent bf16 weight shadows from mixed_precision.ShadowWeights).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F

import os

from . import _lib
from . import tc_linear as tcl
from ._lib import check, ptr, require_cuda, stream_ptr

_KEEP_CACHE = {}
SAVED_PER_BLOCK = 9     # xs, mean1, rstd1, x2, mean2, rstd2, pre, qkv, lse  (+ the stacked Y1, O, Y2, H once)
PARAMS_PER_BLOCK = 11   # norm1.w, norm1.b, qkv.w, proj.w, proj.b, norm2.w, norm2.b, fc1.w, fc1.b, fc2.w, fc2.b


def _flag(dtype) -> int:
    return 1 if dtype == torch.bfloat16 else 0


def ln_fwd(x, delta, scale, pos, gamma, beta, eps, L, act_dtype, want_xs=True, want_y=True, y_out=None):
    """xs = x + scale[b]*delta + pos ; y = LN(xs).  x (T,C) fp32 -> (xs fp32 | None, y act | None, mean, rstd)."""
    T, C = x.shape
    xs = torch.empty_like(x) if want_xs else None
    y = (y_out if y_out is not None else torch.empty((T, C), dtype=act_dtype, device=x.device)) if want_y else None
    mean = torch.empty(T, dtype=torch.float32, device=x.device) if want_y else None
    rstd = torch.empty(T, dtype=torch.float32, device=x.device) if want_y else None
    check(_lib.lib.up3d_ln_fwd(_flag(act_dtype), T, L, C, ptr(x), ptr(delta), ptr(scale), ptr(pos), ptr(gamma), ptr(beta),
                               float(eps), ptr(xs), ptr(y), ptr(mean), ptr(rstd), stream_ptr()), launches=1)
    return xs, y, mean, rstd


def ln_bwd(dy, xs, mean, rstd, gamma, g_res, scale, L, dpos, want_scaled, dgamma, dbeta, dbias, scaled_out=None):
    T, C = xs.shape
    dx = torch.empty_like(xs)
    dscaled = (scaled_out if scaled_out is not None else torch.empty((T, C), dtype=dy.dtype, device=xs.device)) \
        if want_scaled else None
    check(_lib.lib.up3d_ln_bwd(_flag(dy.dtype), T, L, C, ptr(dy), ptr(xs), ptr(mean), ptr(rstd), ptr(gamma), ptr(g_res),
                               ptr(scale), ptr(dx), ptr(dpos), ptr(dscaled), ptr(dgamma), ptr(dbeta),
                               ptr(dbias if want_scaled else None), stream_ptr()), launches=1)
    return dx, dscaled


def gelu_fwd(x, out=None):
    y = out if out is not None else torch.empty_like(x)
    check(_lib.lib.up3d_gelu_fwd(_flag(x.dtype), x.numel(), ptr(x), ptr(y), stream_ptr()), launches=1)
    return y


def gelu_bwd(dy, pre, dbias, out=None):
    T, C = pre.shape
    dx = out if out is not None else torch.empty_like(pre)
    check(_lib.lib.up3d_gelu_bwd(_flag(pre.dtype), T, C, ptr(dy), ptr(pre), ptr(dx), ptr(dbias), stream_ptr()), launches=1)
    return dx


def scale_cast_colsum(g, scale, L, act_dtype, dbias, out=None):
    T, C = g.shape
    out = out if out is not None else torch.empty((T, C), dtype=act_dtype, device=g.device)
    check(_lib.lib.up3d_scale_cast_colsum(_flag(act_dtype), T, L, C, ptr(g), ptr(scale), ptr(out), ptr(dbias),
                                          stream_ptr()), launches=1)
    return out


def attn_supported(act_dtype, L, D) -> bool:
    return act_dtype == torch.bfloat16 and D == 64 and L <= int(_lib.lib.up3d_attn_max_len())


def attn_fwd(qkv, B, L, H, D, scale, out=None):
    """qkv (B*L, 3*H*D) bf16 -> o (B*L, H*D) bf16, lse (B,H,L) fp32 (csrc/attention.cu)."""
    o = out if out is not None else torch.empty((B * L, H * D), dtype=qkv.dtype, device=qkv.device)
    lse = torch.empty((B, H, L), dtype=torch.float32, device=qkv.device)
    check(_lib.lib.up3d_attn_fwd(B, L, H, D, float(scale), ptr(qkv), ptr(o), ptr(lse), stream_ptr()), launches=1)
    return o, lse


def attn_bwd(qkv, o, lse, do, B, L, H, D, scale, out=None):
    dqkv = out if out is not None else torch.empty_like(qkv)
    check(_lib.lib.up3d_attn_bwd(B, L, H, D, float(scale), ptr(qkv), ptr(o), ptr(lse), ptr(do), ptr(dqkv), stream_ptr()),
          launches=1)
    return dqkv


def use_tc(act_dtype) -> bool:
    """bf16 GEMMs go to the hand-written tcgen05 kernel (csrc/gemm_tc.cu); UP3D_TC_LINEAR=0 keeps the library GEMMs."""
    return act_dtype == torch.bfloat16 and os.environ.get("UP3D_TC_LINEAR", "1") != "0"


def _wgrad(dy, x):
    """dW = dy^T @ x written in fp32 by the GEMM itself (no cast pass for the fp32 master gradient)."""
    if dy.dtype == torch.float32:
        return dy.t() @ x
    return torch.mm(dy.t(), x, out_dtype=torch.float32)


def _wgrad_batched(dy, x, out=None):
    """(depth, T, out) x (depth, T, in) -> (depth, out, in) fp32: the weight gradients of one Linear of ALL blocks as a
    single batched GEMM (16 launches of ~5 us become one).  `out`: where the GEMM writes (data parallel: the slab of the
    gradient-exchange buffer, GRAD_BUFFERS)."""
    if dy.dtype == torch.float32:
        return torch.bmm(dy.transpose(1, 2), x) if out is None else torch.bmm(dy.transpose(1, 2), x, out=out)
    if out is None:
        return torch.bmm(dy.transpose(1, 2), x, out_dtype=torch.float32)
    return torch.bmm(dy.transpose(1, 2), x, out_dtype=torch.float32, out=out)


def _linear_deep(x, w, b, tc):
    """y = x W^T + b for the deep-K Linear (fc2: K = 4C)."""
    if tc and _TC_ALL:
        return tcl.tc_linear(x, w, b)
    return F.linear(x, w, b)


def _dx_deep(dy, w, tc):
    """dx = dy W for the Linears whose OUTPUT is wide (qkv: 3C, fc1: 4C), i.e. a deep-K dX product."""
    if tc and _TC_ALL:
        return tcl.tc_linear(dy, w, None, b_major=tcl.B_NMAJOR)
    return dy @ w


# Which Linears run on the hand-written tcgen05 kernel (csrc/gemm_tc.cu).  Default: the two whose epilogue absorbs a
# whole elementwise pass (fc1 + bias + GELU forward; dX of fc2 + GELU backward) -- measured 4.9 us vs 4.0 + 4.5 us and
# 4.4 us vs 3.9 + 7.3 us per block against library GEMM + separate kernel.  The plain 1032-row GEMMs are L2->SM
# bandwidth-bound; there the library's 2-CTA multicast kernels are still 0.5-1 us faster per call (tools/bench_tc_linear.py),
# so they stay on the library unless UP3D_TC_ALL=1.
_TC_ALL = os.environ.get("UP3D_TC_ALL", "0") != "0"
_SIDE_STREAMS = {}
# Data-parallel hook (set by trainer.Trainer when world > 1): called at the end of the stack's backward with the list of
# its parameter gradients (order = run_encoder_stack's parameter order); returns the tensors autograd should see.  The
# trainer uses it to start the all-reduce of these ~97 % of all gradient bytes while the rest of the backward still runs.
GRAD_READY_HOOK = None
# Chunked variant (preferred when set): called from INSIDE the stack's backward every GRAD_CHUNK_BLOCKS blocks with
# (first_block, gradients of blocks [first_block, first_block + n) in stack_parameters order), as soon as those blocks'
# weight gradients exist -- the all-reduce of a chunk then overlaps the backward of the blocks in front of it instead of
# starting when the whole stack is done.
GRAD_CHUNK_HOOK = None
GRAD_CHUNK_BLOCKS = 4
# Data parallel, zero-copy variant of GRAD_READY_HOOK: callable(depth, C, Hd) -> (gWqkv (depth,3C,C), gWproj (depth,C,C),
# gW1 (depth,Hd,C), gW2 (depth,C,Hd), small (depth*(6C+Hd),)) fp32 slabs of the gradient-exchange buffer laid out as
# stack_grad_order() says, or None.  The batched weight-gradient GEMMs and the column-sum kernels then write there
# directly and the 113 MB pack in front of the all-reduce disappears.
GRAD_BUFFERS = None
# Opt-in (UP3D_WGRAD_CHUNK=n > 0): every n blocks (counting down) the four batched weight-gradient GEMMs of those blocks
# (+ fc1's stacked bias column-sum) are issued on the side stream, beside the dX chain of the blocks in front of them,
# instead of as 100 us of serial GEMMs after the chain.  Measured (profiles/r2_step_ab.txt): 2.61 / 2.63 / 2.60 ms per
# step for n = 4 / 2 / 8 against 2.58 ms for the serial batch at the end -- the GEMMs slow the latency-bound chain by
# more than they hide -- hence 0 (off) by default.
WGRAD_CHUNK = int(os.environ.get("UP3D_WGRAD_CHUNK", "0"))


def stack_grad_order(depth: int):
    """Permutation of stack_parameters() that describes the flat layout GRAD_BUFFERS hands out: the four weight slabs
    stacked over the blocks, then per block the column-sum gradients in the backward's `small` order."""
    P = PARAMS_PER_BLOCK
    order = []
    for k in (2, 3, 7, 9):                           # qkv.weight, proj.weight, fc1.weight, fc2.weight
        order += [i * P + k for i in range(depth)]
    for i in range(depth):                           # n1w n1b bproj n2w n2b b2 | b1
        order += [i * P + k for k in (0, 1, 4, 5, 6, 10, 8)]
    return order


SIDE_STREAM_INLINE = os.environ.get("UP3D_SIDE_INLINE", "0") != "0"


class SideStream:
    """Runs work that is OFF the backward critical path (the weight-gradient GEMMs: nothing downstream in the backward
    needs them) on a second stream, so their launch latency and tails overlap the dX chain -- at 1032 tokens every GEMM of
    this stack is latency-, not throughput-bound.  Fork/join with stream events (captured as parallel graph branches);
    inputs handed to the side stream are kept alive until the join so the allocator cannot recycle them early."""

    def __init__(self, device):
        self.main = torch.cuda.current_stream(device)
        key = (device.index if device.index is not None else torch.cuda.current_device())
        if key not in _SIDE_STREAMS:
            _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
        self.side = _SIDE_STREAMS[key]
        self.keep = []

    def run(self, fn, *tensors):
        if SIDE_STREAM_INLINE:                      # (A/B switch: run the work in line on the current stream)
            return fn(*tensors)
        self.side.wait_stream(self.main)
        with torch.cuda.stream(self.side):
            out = fn(*tensors)
        self.keep.extend(tensors)
        return out

    def join(self):
        if SIDE_STREAM_INLINE:
            self.keep.clear()
            return
        self.main.wait_stream(self.side)
        self.keep.clear()


class _Meta:
    """Non-tensor arguments of the stack (one object so that autograd sees a single opaque input)."""

    def __init__(self, B, L, C, heads, scale, eps1, eps2, act_dtype, compute_weights):
        self.B, self.L, self.C, self.heads, self.scale = B, L, C, heads, scale
        self.eps1, self.eps2, self.act_dtype = eps1, eps2, act_dtype
        self.compute_weights = compute_weights     # per block: (wqkv, wproj, bproj, w1, b1, w2, b2) in act dtype


class EncoderStackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pos, masks, meta: _Meta, *params):
        B, L, C, H = meta.B, meta.L, meta.C, meta.heads
        T, D, act = B * L, C // H, meta.act_dtype
        depth = len(params) // PARAMS_PER_BLOCK
        require_cuda(x, pos)
        xcur = x.reshape(T, C).contiguous().float()
        pos2 = pos.reshape(T, C).contiguous().float()
        pend, pend_scale = None, None
        saved, attn_nodes = [], []
        own_attn = attn_supported(act, L, D)
        Hd = meta.compute_weights[0][3].shape[0]
        tc = use_tc(act) and tcl.supported(T, C, C) and tcl.supported(T, Hd, C)
        with torch.cuda.device(x.device), torch.autocast("cuda", enabled=False):
            # the A-side operands of the four weight-gradient GEMMs, stacked over the blocks (one batched GEMM each in
            # the backward): LayerNorm outputs, attention outputs, GELU outputs
            e = lambda n: torch.empty((depth, T, n), dtype=act, device=x.device)
            Y1, O, Y2, HH = e(C), e(C), e(C), e(Hd)
            for i in range(depth):
                n1w, n1b, _, _, _, n2w, n2b, _, _, _, _ = params[i * PARAMS_PER_BLOCK:(i + 1) * PARAMS_PER_BLOCK]
                wqkv, wproj, bproj, w1, b1, w2, b2 = meta.compute_weights[i]
                s1 = masks[2 * i] if masks is not None else None
                s2 = masks[2 * i + 1] if masks is not None else None
                xs, y1, mu1, rs1 = ln_fwd(xcur, pend, pend_scale, pos2, n1w, n1b, meta.eps1[i], L, act, y_out=Y1[i])
                qkv = tcl.tc_linear(y1, wqkv) if (tc and _TC_ALL) else y1 @ wqkv.t()
                if own_attn:
                    o, lse = attn_fwd(qkv, B, L, H, D, meta.scale, out=O[i])
                    attn_nodes.append(None)
                else:                                   # fp32 operands (reference precision): library SDPA
                    with torch.enable_grad():
                        qkv_l = qkv.detach().requires_grad_(True)
                        q, k, v = qkv_l.view(B, L, 3, H, D).permute(2, 0, 3, 1, 4).unbind(0)
                        o4 = F.scaled_dot_product_attention(q, k, v, scale=meta.scale)
                        o_l = o4.transpose(1, 2).reshape(T, C)
                    o, lse = O[i].copy_(o_l.detach()), mu1.new_empty(0)
                    attn_nodes.append((qkv_l, o_l))
                a = tcl.tc_linear(o, wproj, bproj) if (tc and _TC_ALL) else F.linear(o, wproj, bproj)
                x2, y2, mu2, rs2 = ln_fwd(xs, a, s1, None, n2w, n2b, meta.eps2[i], L, act, y_out=Y2[i])
                if tc:      # fc1 + bias + GELU in the GEMM epilogue (pre = the GELU input the backward re-reads)
                    h, pre = tcl.tc_linear(y2, w1, b1, epilogue=tcl.EPI_GELU, out=HH[i])
                else:
                    pre = F.linear(y2, w1, b1)
                    h = gelu_fwd(pre, out=HH[i])
                d = _linear_deep(h, w2, b2, tc)
                saved += [xs, mu1, rs1, x2, mu2, rs2, pre, qkv, lse]
                xcur, pend, pend_scale = x2, d, s2
            out, _, _, _ = ln_fwd(xcur, pend, pend_scale, None, None, None, 0.0, L, act, want_y=False)
        saved += [Y1, O, Y2, HH]
        ctx.meta, ctx.depth, ctx.attn_nodes, ctx.own_attn = meta, depth, attn_nodes, own_attn
        ctx.has_masks, ctx.tc = masks is not None, tc
        ctx.save_for_backward(*(saved + list(params) + ([masks] if masks is not None else [])))
        return out.view(B, L, C)

    @staticmethod
    def backward(ctx, gout):
        meta, depth = ctx.meta, ctx.depth
        B, L, C = meta.B, meta.L, meta.C
        T, act = B * L, meta.act_dtype
        sv = ctx.saved_tensors
        n_act = SAVED_PER_BLOCK * depth + 4
        H, D = meta.heads, C // meta.heads
        Y1, O, Y2, HH = sv[n_act - 4:n_act]
        params = sv[n_act:n_act + PARAMS_PER_BLOCK * depth]
        masks = sv[-1] if ctx.has_masks else None
        Hd = meta.compute_weights[0][3].shape[0]
        dev, tc = gout.device, ctx.tc
        with torch.cuda.device(dev), torch.autocast("cuda", enabled=False):
            g = gout.reshape(T, C).contiguous().float()
            # all column-sum gradients of the stack in one zero-filled buffer (the kernels accumulate atomically)
            per = 6 * C + Hd      # n1w n1b bproj n2w n2b b2 (C each) + b1 (Hd)
            bufs = GRAD_BUFFERS(depth, C, Hd) if (GRAD_BUFFERS is not None and GRAD_CHUNK_HOOK is None) else None
            if bufs is not None:
                small = bufs[4].zero_()
            else:
                small = torch.zeros(depth * per, dtype=torch.float32, device=dev)
            dpos = torch.zeros((T, C), dtype=torch.float32, device=dev)
            grads: List[Optional[torch.Tensor]] = [None] * (PARAMS_PER_BLOCK * depth)

            def sm(i, k, n=C):
                o = i * per + k * C
                return small[o:o + n]

            # slots: 0 n1w, 1 n1b, 2 bproj, 3 n2w, 4 n2b, 5 b2, 6.. b1
            s2_last = masks[2 * depth - 1] if masks is not None else None
            # the dY operands of the four weight-gradient GEMMs, stacked over the blocks
            e = lambda n: torch.empty((depth, T, n), dtype=act, device=dev)
            DD, DPRE, DA, DQKV = e(C), e(Hd), e(C), e(3 * C)
            dd = scale_cast_colsum(g, s2_last, L, act, sm(depth - 1, 5), out=DD[depth - 1])
            chunked = GRAD_CHUNK_HOOK is not None and depth % GRAD_CHUNK_BLOCKS == 0
            wg = {}                                 # block index -> (gWqkv, gWproj, gW1, gW2) when produced chunk-wise
            CH = WGRAD_CHUNK if (WGRAD_CHUNK > 0 and not chunked and depth % max(WGRAD_CHUNK, 1) == 0
                                 and depth > WGRAD_CHUNK and g.is_cuda) else 0
            side = SideStream(dev) if CH else None

            def side_wgrads(sl):
                """(side stream) weight gradients of blocks `sl` into their slabs (exchange buffer or fresh tensors)."""
                oq, op, o1, o2 = (b[sl] for b in bufs[:4]) if bufs is not None else (None,) * 4
                c2, c1 = _wgrad_batched(DD[sl], HH[sl], o2), _wgrad_batched(DPRE[sl], Y2[sl], o1)
                cp, cq = _wgrad_batched(DA[sl], O[sl], op), _wgrad_batched(DQKV[sl], Y1[sl], oq)
                if tc:
                    small.view(depth, per)[sl, 6 * C:].copy_(DPRE[sl].sum(dim=1, dtype=torch.float32))
                return cq, cp, c1, c2

            def block_grads(i, gWqkv_i, gWproj_i, gW1_i, gW2_i):
                return [sm(i, 0), sm(i, 1), gWqkv_i, gWproj_i, sm(i, 2), sm(i, 3), sm(i, 4), gW1_i, sm(i, 6, Hd), gW2_i, sm(i, 5)]

            for i in range(depth - 1, -1, -1):
                xs, mu1, rs1, x2, mu2, rs2, pre, qkv, lse = sv[SAVED_PER_BLOCK * i:SAVED_PER_BLOCK * (i + 1)]
                n1w, _, _, _, _, n2w, _, _, _, _, _ = params[i * PARAMS_PER_BLOCK:(i + 1) * PARAMS_PER_BLOCK]
                wqkv, wproj, _, w1, _, w2, _ = meta.compute_weights[i]
                s1 = masks[2 * i] if masks is not None else None
                s2_prev = masks[2 * i - 1] if (masks is not None and i > 0) else None
                # ---- MLP branch
                if tc:      # dX of fc2 with the GELU backward in the epilogue; fc1's bias gradient = column sums, below
                    dpre = tcl.tc_linear(dd, w2, None, b_major=tcl.B_NMAJOR, epilogue=tcl.EPI_GELU_BWD, aux_in=pre, out=DPRE[i])
                else:
                    dh = dd @ w2
                    dpre = gelu_bwd(dh, pre, sm(i, 6, Hd), out=DPRE[i])
                dy2 = _dx_deep(dpre, w1, tc)
                dx2, da = ln_bwd(dy2, x2, mu2, rs2, n2w, g, s1, L, None, True, sm(i, 3), sm(i, 4), sm(i, 2),
                                 scaled_out=DA[i])
                # ---- attention branch
                do = tcl.tc_linear(da, wproj, None, b_major=tcl.B_NMAJOR) if (tc and _TC_ALL) else da @ wproj
                if ctx.own_attn:
                    dqkv = attn_bwd(qkv, O[i], lse, do, B, L, H, D, meta.scale, out=DQKV[i])
                else:
                    qkv_l, o_l = ctx.attn_nodes[i]
                    (dqkv,) = torch.autograd.grad(o_l, qkv_l, do)
                    dqkv = DQKV[i].copy_(dqkv)
                dy1 = _dx_deep(dqkv, wqkv, tc)
                g, dd = ln_bwd(dy1, xs, mu1, rs1, n1w, dx2, s2_prev, L, dpos, i > 0, sm(i, 0), sm(i, 1),
                               sm(i - 1, 5) if i > 0 else None, scaled_out=DD[i - 1] if i > 0 else None)
                if CH and i % CH == 0:
                    sl = slice(i, i + CH)
                    cq, cp, c1, c2 = side.run(lambda *keep, sl=sl: side_wgrads(sl), DD, HH, DPRE, Y2, DA, O, DQKV, Y1, small)
                    for k in range(CH):
                        wg[i + k] = (cq[k], cp[k], c1[k], c2[k])
                if chunked and i % GRAD_CHUNK_BLOCKS == 0:
                    # data parallel: the weight gradients of blocks [i, i + CH) now, so their all-reduce starts now
                    sl = slice(i, i + GRAD_CHUNK_BLOCKS)
                    c2, c1 = _wgrad_batched(DD[sl], HH[sl]), _wgrad_batched(DPRE[sl], Y2[sl])
                    cp, cq = _wgrad_batched(DA[sl], O[sl]), _wgrad_batched(DQKV[sl], Y1[sl])
                    if tc:
                        small.view(depth, per)[sl, 6 * C:].copy_(DPRE[sl].sum(dim=1, dtype=torch.float32))
                    chunk = []
                    for k in range(GRAD_CHUNK_BLOCKS):
                        wg[i + k] = (cq[k], cp[k], c1[k], c2[k])
                        chunk += block_grads(i + k, *wg[i + k])
                    GRAD_CHUNK_HOOK(i, chunk)
            if side is not None:
                side.join()                         # (only the chunk of the first blocks can still be in flight)
            elif not chunked:
                # ---- weight gradients of all blocks: four batched GEMMs (fp32 written by the GEMM)
                oq, op, o1, o2 = bufs[:4] if bufs is not None else (None,) * 4
                gW2, gW1 = _wgrad_batched(DD, HH, o2), _wgrad_batched(DPRE, Y2, o1)
                gWproj, gWqkv = _wgrad_batched(DA, O, op), _wgrad_batched(DQKV, Y1, oq)
                if tc:          # fc1 bias gradients of all blocks: one column-sum pass over the stacked dpre
                    small.view(depth, per)[:, 6 * C:].copy_(DPRE.sum(dim=1, dtype=torch.float32))
                for i in range(depth):
                    wg[i] = (gWqkv[i], gWproj[i], gW1[i], gW2[i])
            for i in range(depth):
                grads[i * PARAMS_PER_BLOCK:(i + 1) * PARAMS_PER_BLOCK] = block_grads(i, *wg[i])
            ctx.attn_nodes = None
            if GRAD_READY_HOOK is not None and not chunked:
                grads = GRAD_READY_HOOK(grads)
        gx = g.view(B, L, C) if ctx.needs_input_grad[0] else None
        gp = dpos.view(B, L, C) if ctx.needs_input_grad[1] else None
        return (gx, gp, None, None) + tuple(grads)


class FusedLayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the last dimension through up3d_ln_fwd / up3d_ln_bwd (1 launch each way instead of 2 + 3); used
    for the encoder's final `self.norm` (transformer.py:325)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, act_dtype):
        require_cuda(x)
        shape = x.shape
        C = shape[-1]
        x2 = x.reshape(-1, C).contiguous().float()
        with torch.cuda.device(x.device):
            _, y, mean, rstd = ln_fwd(x2, None, None, None, weight, bias, eps, x2.shape[0], act_dtype, want_xs=False)
        ctx.save_for_backward(x2, mean, rstd, weight)
        ctx.shape = shape
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd, weight = ctx.saved_tensors
        C = x2.shape[1]
        dy2 = dy.reshape(-1, C).contiguous()
        if dy2.dtype not in (torch.float32, torch.bfloat16):
            dy2 = dy2.float()
        dwb = torch.zeros(2 * C, dtype=torch.float32, device=x2.device)
        with torch.cuda.device(x2.device):
            dx, _ = ln_bwd(dy2, x2, mean, rstd, weight, None, None, x2.shape[0], None, False, dwb[:C], dwb[C:], None)
        return dx.view(ctx.shape), dwb[:C], dwb[C:], None, None


def fused_layer_norm(norm, x):
    """norm: nn.LayerNorm; x (..., C) on CUDA -> LayerNorm(x) in the autocast dtype (fp32 without autocast)."""
    act = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else torch.float32
    C = x.shape[-1]
    ok = (x.is_cuda and isinstance(norm, torch.nn.LayerNorm) and norm.elementwise_affine and norm.bias is not None
          and tuple(norm.normalized_shape) == (C,) and C % 128 == 0 and C // 128 in (1, 2, 3, 4, 6, 8)
          and act in (torch.float32, torch.bfloat16))
    if not ok:
        return norm(x)
    return FusedLayerNormFn.apply(x, norm.weight, norm.bias, float(norm.eps), act)


def supports(blocks) -> bool:
    """The fused stack covers the configuration the reference instantiates (qkv_bias=False, no dropout, erf-GELU,
    LayerNorm with affine parameters, width a multiple of 128)."""
    import torch.nn as nn
    for b in blocks:
        C = b.norm1.normalized_shape[0]
        if C % 128 != 0 or C // 128 not in (1, 2, 3, 4, 6, 8):
            return False
        if b.attn.qkv.bias is not None or b.attn.proj.bias is None or b.mlp.fc1.bias is None or b.mlp.fc2.bias is None:
            return False
        if b.attn.attn_drop.p != 0.0 or b.attn.proj_drop.p != 0.0 or b.mlp.drop.p != 0.0:
            return False
        if not isinstance(b.mlp.act, nn.GELU) or getattr(b.mlp.act, "approximate", "none") != "none":
            return False
        if not (isinstance(b.norm1, nn.LayerNorm) and b.norm1.elementwise_affine and b.norm1.bias is not None):
            return False
    return True


def _compute_copy(lin, act_dtype):
    """(weight, bias) of an nn.Linear in the GEMM operand dtype: the persistent bf16 shadows when they exist."""
    if act_dtype == torch.float32:
        return lin.weight.detach(), (None if lin.bias is None else lin.bias.detach())
    w16 = getattr(lin, "_w16", None)
    if w16 is None:
        w16 = lin.weight.detach().to(act_dtype)
        b16 = None if lin.bias is None else lin.bias.detach().to(act_dtype)
    else:
        b16 = lin._b16
    return w16, b16


def stack_parameters(blocks):
    """The stack's parameters in the order EncoderStackFn receives them (and GRAD_READY_HOOK their gradients)."""
    params = []
    for b in blocks:
        params += [b.norm1.weight, b.norm1.bias, b.attn.qkv.weight, b.attn.proj.weight, b.attn.proj.bias,
                   b.norm2.weight, b.norm2.bias, b.mlp.fc1.weight, b.mlp.fc1.bias, b.mlp.fc2.weight, b.mlp.fc2.bias]
    return params


def run_encoder_stack(blocks, x, pos, training: bool):
    """blocks: nn.ModuleList of backbone.Block; x, pos (B,L,C) on CUDA -> (B,L,C) fp32."""
    B, L, C = x.shape
    act = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else torch.float32
    if act not in (torch.float32, torch.bfloat16):
        raise RuntimeError(f"fused encoder stack: unsupported autocast dtype {act}")
    params, cw, eps1, eps2, keep = [], [], [], [], []
    for b in blocks:
        params += [b.norm1.weight, b.norm1.bias, b.attn.qkv.weight, b.attn.proj.weight, b.attn.proj.bias,
                   b.norm2.weight, b.norm2.bias, b.mlp.fc1.weight, b.mlp.fc1.bias, b.mlp.fc2.weight, b.mlp.fc2.bias]
        wqkv, _ = _compute_copy(b.attn.qkv, act)
        wproj, bproj = _compute_copy(b.attn.proj, act)
        w1, b1 = _compute_copy(b.mlp.fc1, act)
        w2, b2 = _compute_copy(b.mlp.fc2, act)
        cw.append((wqkv, wproj, bproj, w1, b1, w2, b2))
        eps1.append(b.norm1.eps)
        eps2.append(b.norm2.eps)
        p = float(getattr(b.drop_path, "drop_prob", 0.0))
        keep += [1.0 - p, 1.0 - p]
    masks = None
    if training and any(k < 1.0 for k in keep):
        # DropPath (timm semantics, scale_by_keep): one Bernoulli(keep)/keep factor per (branch, sample), all 2*depth
        # branches drawn at once.  The keep-probability column is cached on the device (no H2D copy per step).
        key = (x.device, tuple(keep))
        kt = _KEEP_CACHE.get(key)
        if kt is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("fused encoder stack: run one eager step before CUDA-graph capture")
            kt = _KEEP_CACHE[key] = torch.tensor(keep, dtype=torch.float32, device=x.device).unsqueeze(1)
        masks = (torch.rand(len(keep), B, device=x.device) < kt).float() / kt
    meta = _Meta(B, L, C, blocks[0].attn.num_heads, float(blocks[0].attn.scale), eps1, eps2, act, cw)
    return EncoderStackFn.apply(x, pos, masks, meta, *params)
