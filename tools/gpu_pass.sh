#!/usr/bin/env bash
# Full -m gpu suite, then the quick benches of every workload and (optionally) the full bench line.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
echo "== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -12
show() { python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); r = d["roofline"]
print("%s ms/step %.4f value %.3e frac %.3f tensor_pipe %s launch_ms %.4f share %.2f launches %d" % (
    sys.argv[2], d["ms_per_step"], d["value"], r["frac"] or 0, (r.get("tensor_pipe") or {}).get("frac"),
    r["avg_launch_ms"], r["share_of_step"] or 0, d["gpu_launches"]))
PY
}
for wl in stft2048 cfg2 cfg5 cfg3 cfg4; do
  timeout 200 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu \
      > gpurun_out/q_$wl.json 2>> gpurun_out/q_err.txt && show gpurun_out/q_$wl.json $wl
done
tail -5 gpurun_out/q_err.txt
