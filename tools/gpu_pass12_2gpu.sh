#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for mode in root all; do
  echo "== bench --gpus 2 --gather auto --gather-to $mode"
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 100 --warmup 10 --no-workloads --no-reference-gpu --no-e2e --gather-to $mode > gpurun_out/n2_peer_$mode.json 2> gpurun_out/n2_peer_$mode.err; echo "rc $?"; tail -2 gpurun_out/n2_peer_$mode.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/n2_peer_$mode.json"))
    print("value %.3e ms %.4f gather %s reserve %s without %s" % (d["value"], d["ms_per_step"], d["config"].get("gather"), d["config"]["sms_reserved_for_gather"], (d.get("without_gather") or {}).get("ms_per_step")))
except Exception as e: print("no json", e)
PY
done
echo "== full default line at N=2 (as the driver runs it)"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_default.json 2> gpurun_out/n2_default.err; echo "rc $?"; tail -2 gpurun_out/n2_default.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/n2_default.json"))
print("value %.3e ms %.4f e2e %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_step")))
for k,v in (d.get("workloads") or {}).items(): print(k, v.get("ms_per_step"), v.get("error"))
PY
