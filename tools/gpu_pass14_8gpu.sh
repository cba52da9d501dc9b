#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/symm_gather_check.py 2>&1 | grep -E "gather|MISMATCH|Error" | tail -5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 bench.py --gpus 8 --steps 200 --warmup 10 --no-workloads --no-reference-gpu --no-e2e > gpurun_out/n8_root4.json 2> gpurun_out/n8_root4.err; echo "rc $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/n8_root4.json"))
print("N=8 root (4 copy streams): value %.3e ms %.4f without %s" % (d["value"], d["ms_per_step"], (d.get("without_gather") or {}).get("ms_per_step")))
PY
