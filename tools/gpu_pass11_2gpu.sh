#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== gather check"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/symm_gather_check.py 2>&1 | tail -6
for mode in root all; do
  echo "== bench --gpus 2 --gather auto --gather-to $mode"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 100 --warmup 10 --no-workloads --no-reference-gpu --no-e2e --gather-to $mode > gpurun_out/n2_peer_$mode.json 2> gpurun_out/n2_peer_$mode.err; echo "rc $?"; tail -2 gpurun_out/n2_peer_$mode.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/n2_peer_$mode.json"))
    print("value %.3e ms %.4f gather %s reserve %s without %s" % (d["value"], d["ms_per_step"], d["config"].get("gather"), d["config"]["sms_reserved_for_gather"], (d.get("without_gather") or {}).get("ms_per_step")))
except Exception as e: print("no json", e)
PY
done
echo "== pyramid + cfg4 (1 GPU part)"
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "cqt2010 or vqt or VQT or CQT2010 or cfg4 or sweep-cqt-2010 or gamma" 2>&1 | tail -4
timeout 200 python bench.py --workload cfg4 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu > gpurun_out/q_cfg4.json 2>> gpurun_out/q_err.txt; python -c "
import json; d=json.load(open('gpurun_out/q_cfg4.json')); print('cfg4 ms %.4f launches/step %.1f' % (d['ms_per_step'], d['gpu_launches']/d['steps']))"
NNAB_TALL=0 timeout 200 python bench.py --workload cfg4 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu > gpurun_out/q_cfg4_notall.json 2>> gpurun_out/q_err.txt; python -c "
import json; d=json.load(open('gpurun_out/q_cfg4_notall.json')); print('cfg4 (dense octaves) ms %.4f' % (d['ms_per_step']))"
tail -3 gpurun_out/q_err.txt
