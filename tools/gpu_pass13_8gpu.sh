#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
run() { # n mode extra
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 2954$1 bench.py --gpus $1 --steps 100 --warmup 10 --no-workloads --no-reference-gpu --no-e2e --gather-to $2 $3 > gpurun_out/n$1_$2$4.json 2> gpurun_out/n$1_$2$4.err; echo "rc $?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/n$1_$2$4.json"))
    print("N=$1 $2 $3: value %.3e ms %.4f gather %s without %s clocks %s" % (d["value"], d["ms_per_step"], d["config"].get("gather"), (d.get("without_gather") or {}).get("ms_per_step"), d["clocks"].get("sm_mhz")))
except Exception as e: print("no json", e); print(open("gpurun_out/n$1_$2$4.err").read()[-600:])
PY
}
run 8 root "" ""
run 8 all "" ""
run 4 root "" ""
run 8 all "--gather nccl" "_nccl"
echo "== default driver line at N=8"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8_default.json 2> gpurun_out/n8_default.err; echo "rc $?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/n8_default.json"))
print("N=8 default: value %.3e ms %.4f e2e %s" % (d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("ms_per_step")))
for k,v in (d.get("workloads") or {}).items(): print(k, v.get("ms_per_step"), v.get("value"), v.get("error"))
PY
