#!/bin/bash
# 1-GPU validation: the -m gpu suite, smoke(), the default bench line and the ncu launch list of the bench command
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 400 2>&1 | tail -6 > gpurun_out/final_tests.log
cat gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/final_bench.err
python - <<PY
import json
for l in open("gpurun_out/final_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "clocks")})
        print(d["e2e"]); print(d["roofline"]); print(d.get("cpu_baseline"))
        print({k: d.get(k) for k in d if k.startswith("variant") or k.startswith("steady") or k.startswith("raster_only") or k.startswith("legacy")})
PY
if [ -n "$WITH_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2700 --launch-count 900 --csv \
    --log-file gpurun_out/r2c_bench_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
  echo "ncu rc=$?"; wc -l gpurun_out/r2c_bench_launches.csv
fi
