#!/bin/bash
# ncu captures of the rasterizer kernels (run under gpurun, 1 GPU).  Outputs land in gpurun_out/.
set -x
P=${1:-8192}; TAG=${2:-r1}
ncu --set full --clock-control none --import-source on -k regex:'blend_backward_kernel|blend_forward_kernel|depth_sort_kernel' \
    -s 12 -c 3 -o gpurun_out/raster_${TAG}_P${P} -f python tools/bench_raster.py $P 8 4 256 > gpurun_out/ncu_raster_${TAG}_P${P}.log 2>&1
tail -2 gpurun_out/ncu_raster_${TAG}_P${P}.log
