"""Per-launch time of up3d_tc_linear vs the library GEMM (torch -> cuBLASLt) at the backbone's shapes, measured as the
step graph sees them: a CUDA graph of `reps` dependent launches (output of one is not reused; same stream, back to back)."""
import sys
import torch

sys.path.insert(0, ".")
from unipre3d_b200.tc_linear import tc_linear  # noqa: E402

SHAPES = [(1032, 1152, 384), (1032, 384, 384), (1032, 1536, 384), (1032, 384, 1536), (32768, 256, 128), (32768, 512, 512),
          (32768, 384, 512)]


def graph_time(fn, reps=50, iters=20):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * iters)


def main():
    print(f"{'T':>6} {'N':>5} {'K':>5} {'bm':>2} {'tile':>6} | {'tc us':>8} {'lib us':>8} | TFLOP/s tc")
    for T, N, K in SHAPES:
        a = torch.randn((T, K), device="cuda").bfloat16()
        w = torch.randn((N, K), device="cuda").bfloat16()
        wt = w.t().contiguous()
        bias = torch.randn(N, device="cuda").bfloat16()
        out = torch.empty((T, N), device="cuda", dtype=torch.bfloat16)
        lib0 = graph_time(lambda: torch.nn.functional.linear(a, w, bias))
        dy = torch.randn((T, N), device="cuda").bfloat16()
        lib1 = graph_time(lambda: dy @ w)
        flops = 2.0 * T * N * K
        for bm, tiles in ((0, (0, 32, 64, 96, 128)), (1, (0, 64, 128))):
            for tn in tiles:
                for ks in ((0,) if tn == 0 else (1, 2, 4)):
                    if bm == 0:
                        if tn and (N % tn or (ks > 1 and K // 64 // ks < 2)):
                            continue
                        t = graph_time(lambda: tc_linear(a, w, bias, out=out, tile_n=tn | (ks << 16)))
                    else:   # dx = dy @ w : (T,N) x (N,K): the (K_gemm = N, N_gemm = K) problem, B stored (K_gemm, N_gemm)
                        if tn and (K % tn or (ks > 1 and N // 64 // ks < 2)):
                            continue
                        o2 = torch.empty((T, K), device="cuda", dtype=torch.bfloat16)
                        t = graph_time(lambda: tc_linear(dy, w, None, b_major=1, out=o2, tile_n=tn | (ks << 16)))
                    tag = "auto" if tn == 0 else f"{tn}/{ks}"
                    print(f"{T:6d} {N:5d} {K:5d} {bm:2d} {tag:>6} | {t:8.2f} {(lib0 if bm == 0 else lib1):8.2f} | {flops / t * 1e-6:7.1f}")


if __name__ == "__main__":
    main()
