"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol include/up3d.h declares,
argument validation works without a GPU, and the product package never routes through the oracle."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "up3d.h")).read()
    return sorted(set(re.findall(r"UP3D_API\s+[\w \*]+?\b(up3d_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from unipre3d_b200 import _lib
    syms = _declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(_lib.lib, s), f"libunipre3d_b200.so does not export {s}"
    assert sorted(_lib.EXPORTED) == syms
    assert _lib.lib.up3d_version() >= 100


def test_argument_validation_without_gpu():
    from unipre3d_b200 import _lib
    L = _lib.lib
    assert L.up3d_fps(1, 16, 32, None, None, None, None) != 0
    assert b"npoint" in L.up3d_last_error()
    assert L.up3d_ball_query(1, 0, 4, 0.1, 8, None, None, None, None) != 0
    d = _lib.RasterDesc(1, 1, 4, 4, 4, 0, 16, 1, 4, 1, 0.5, 0.5, 1.0, None, None, None, None)
    assert L.up3d_raster_forward(C.byref(d), *([None] * 16)) != 0
    assert b"image size" in L.up3d_last_error()
    d = _lib.RasterDesc(1, 1, 4, 4, 4, 64, 64, 5, 4, 1, 0.5, 0.5, 1.0, None, None, None, None)
    assert L.up3d_raster_forward(C.byref(d), *([None] * 16)) != 0
    assert b"sh_degree" in L.up3d_last_error()
    # empty work is a no-op success
    assert L.up3d_fps(0, 16, 4, None, None, None, None) == 0
    d0 = _lib.RasterDesc(0, 0, 0, 0, 0, 64, 64, 1, 4, 1, 0.5, 0.5, 1.0, None, None, None, None)
    assert L.up3d_raster_state_bytes(C.byref(d0)) > 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "unipre3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "oracle_lib" not in src, f"{f} references oracle_lib"


def test_cpu_tensors_fail_loudly():
    import torch
    from unipre3d_b200.rasterizer import rasterize_batch
    z = torch.zeros
    with pytest.raises(RuntimeError, match="CUDA device"):
        rasterize_batch(z(4, 3), z(4), z(4, 3), z(4, 4), z(1, 4, 4), z(1, 4, 4), z(1, 3), z(3), set_sizes=[4],
                        views_per_set=[1], image_height=16, image_width=16, tanfovx=0.5, tanfovy=0.5, sh_degree=0,
                        shs=z(4, 1, 3))
