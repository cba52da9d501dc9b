#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== cqt tests (tall kernel) + determinism"; timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "repeatable or cqt1992 or CQT1992 or cfg3 or sweep" 2>&1 | tail -4
echo "== multi-device (one process)"; timeout 300 python -m pytest tests/test_gpu_multidevice.py -m gpu -q --timeout 200 2>&1 | tail -4
echo "== symm gather check"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/symm_gather_check.py 2>&1 | tail -5
for g in auto nccl; do
  echo "== bench --gpus 2 --gather $g"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 100 --warmup 10 --no-workloads --no-reference-gpu --gather $g > gpurun_out/n2_$g.json 2> gpurun_out/n2_$g.err; echo "rc $?"; tail -2 gpurun_out/n2_$g.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/n2_$g.json"))
    print("value %.3e ms %.4f gather %s reserve %s without %s e2e %s" % (d["value"], d["ms_per_step"], d["config"].get("gather"), d["config"]["sms_reserved_for_gather"], (d.get("without_gather") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("ms_per_step")))
except Exception as e: print("no json", e)
PY
done
echo "== bench --gpus 2 --gather auto --gather-to root"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 100 --warmup 10 --no-workloads --no-reference-gpu --no-e2e --gather-to root > gpurun_out/n2_root.json 2> gpurun_out/n2_root.err; tail -1 gpurun_out/n2_root.err; python -c "
import json; d=json.load(open('gpurun_out/n2_root.json')); print('root: value %.3e ms %.4f' % (d['value'], d['ms_per_step']))"
