#!/usr/bin/env bash
# cfg3 (CQT1992v2) check after a change of the tall-A kernel: parity, determinism, full size, bench line.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -k "1992 or cfg3 or cqt" 2>&1 | tail -5
Q="--no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu"
timeout 300 python bench.py --workload cfg3 --steps 50 --warmup 10 $Q > gpurun_out/q_cfg3.json 2> gpurun_out/q_cfg3.err || tail -5 gpurun_out/q_cfg3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/q_cfg3.json"))
print("cfg3 ms %.4f frac %.3f launches %s" % (d["ms_per_step"], d["roofline"]["frac"], d.get("gpu_launches")))
PY
