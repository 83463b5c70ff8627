"""fc1 (+bias +GELU) as the step sees it: a CUDA graph of 16 'blocks', each = [small producer kernel writing A] ->
fc1 variant -> [library fc2 GEMM], every block with its OWN weights, L2 flushed between replays.  Reports us per block
for several fc1 variants minus the filler-only graph -- the in-step cost, as opposed to tools/bench_tc_linear.py's
back-to-back launches on L2-resident operands."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from unipre3d_b200 import fused_encoder as fe  # noqa: E402
from unipre3d_b200 import tc_linear as tcl  # noqa: E402

T, C, H, D = 1032, 384, 1536, 16


def main():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    src = rn(T, C).bfloat16()
    a = torch.empty_like(src)
    W1 = [(rn(H, C) * 0.05).bfloat16() for _ in range(D)]
    B1 = [(rn(H) * 0.05).bfloat16() for _ in range(D)]
    W2 = [(rn(C, H) * 0.05).bfloat16() for _ in range(D)]
    B2 = [(rn(C) * 0.05).bfloat16() for _ in range(D)]
    HH = torch.empty((D, T, H), device=dev, dtype=torch.bfloat16)
    PRE = torch.empty((D, T, H), device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def block(i, variant):
        torch.add(src, 1.0, out=a)                                   # stands for ln_fwd: A is produced right before fc1
        if variant == "filler":
            h = HH[i]
        elif variant == "tc_gelu":
            h, _ = tcl.tc_linear(a, W1[i], B1[i], epilogue=tcl.EPI_GELU, out=HH[i], aux_out=PRE[i])
        elif variant == "tc_gelu_nobias":
            h, _ = tcl.tc_linear(a, W1[i], None, epilogue=tcl.EPI_GELU, out=HH[i], aux_out=PRE[i])
        elif variant == "tc_gelu_hotw":
            h, _ = tcl.tc_linear(a, W1[0], B1[0], epilogue=tcl.EPI_GELU, out=HH[i], aux_out=PRE[i])
        elif variant == "tc_gelu_hotout":
            h, _ = tcl.tc_linear(a, W1[i], B1[i], epilogue=tcl.EPI_GELU, out=HH[0], aux_out=PRE[0])
        elif variant == "tc_plain":
            h = tcl.tc_linear(a, W1[i], B1[i], out=HH[i])
        elif variant == "tc_plain+gelu":
            pre = tcl.tc_linear(a, W1[i], B1[i], out=PRE[i])
            h = fe.gelu_fwd(pre, out=HH[i])
        elif variant == "lib+gelu":
            pre = F.linear(a, W1[i], B1[i])
            h = fe.gelu_fwd(pre, out=HH[i])
        else:
            raise ValueError(variant)
        return F.linear(h, W2[i], B2[i])

    def graph_us(variant):
        for i in range(D):
            block(i, variant)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(D):
                block(i, variant)
        ts = []
        for _ in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2] * 1e3 / D

    base = graph_us("filler")
    print(f"filler (producer + library fc2) {base:7.2f} us per block")
    for v in ("tc_gelu", "tc_gelu_nobias", "tc_gelu_hotw", "tc_gelu_hotout", "tc_plain", "tc_plain+gelu", "lib+gelu"):
        t = graph_us(v)
        print(f"{v:<18s} {t:7.2f} us per block  -> fc1 part {t - base:6.2f} us", flush=True)


if __name__ == "__main__":
    main()
