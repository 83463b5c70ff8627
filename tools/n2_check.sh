#!/bin/bash
# N = 2 validation on a 2-GPU box: data-parallel parity tests, the bench line, and the kernel timeline of one replay
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1000 python -m pytest tests/test_dist_gpu.py -x -q --timeout 480 2>&1 | tail -8 > gpurun_out/n2_tests.log
cat gpurun_out/n2_tests.log
for mode in ${N2_MODES:-default}; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_$mode.json 2> gpurun_out/n2_$mode.err
  echo "== $mode rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/n2_$mode.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"].get("h2d_probe_GBs"), d.get("steady_state_no_l2_flush", {}).get("ms_per_step"))
PY
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/profile_n2.py 2>&1 | grep -v "^W\|OMP_NUM\|^\*\*\*\|Warning\|_warn_once\|tr = Trainer" | tail -40
