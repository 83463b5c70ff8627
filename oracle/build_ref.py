"""Build the reference's own point-op CUDA extension into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Compiles the UNMODIFIED sources where they lie under
/root/reference/openpoints/cpp/pointnet2_batch/src (nothing is copied into this repo) into
oracle/_ref/pointnet2_batch_cuda.so for sm_100.  The .so travels to the GPU box with the
snapshot (oracle/_ref is git-ignored, not gpurun-ignored) and is used by the `-m gpu`
parity tests as the *real reference* for FPS / ball query / grouping.
It is never imported by the product package.
"""
import glob
import os
import sys

REF_SRC = "/root/reference/openpoints/cpp/pointnet2_batch/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def build(verbose: bool = False) -> str | None:
    so = os.path.join(OUT, "pointnet2_batch_cuda.so")
    if not os.path.isdir(REF_SRC):
        return so if os.path.exists(so) else None
    srcs = sorted(glob.glob(os.path.join(REF_SRC, "*.cpp")) + glob.glob(os.path.join(REF_SRC, "*.cu")))
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load
    load(name="pointnet2_batch_cuda", sources=srcs, build_directory=OUT,
         extra_cflags=["-O2"], extra_cuda_cflags=["-O2"], verbose=verbose, is_python_module=False)
    return so if os.path.exists(so) else None


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
