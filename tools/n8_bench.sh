#!/bin/bash
# N-GPU bench line exactly as the driver launches it (N = number of GPUs of the box)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo "N=$N rc=$?"
python - <<PY
import json
for l in open("gpurun_out/n${N}_bench.json"):
    if l.startswith("{"):
        d = json.loads(l); print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"], d["clocks"])
PY
tail -3 gpurun_out/n${N}_bench.err
