#!/usr/bin/env bash
# GPU pass over the block-partial STFT kernel: STFT-family parity, quick benches, then the full bench line.
set -u
mkdir -p gpurun_out
echo "== STFT-family parity"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --timeout 300 -x -k "stft or mel or mfcc or gammatone or cfg1 or cfg2 or cfg5" 2>&1 | tail -5
show() { python - "$1" "$2" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); r = d["roofline"]
print("%s ms/step %.4f value %.3e frac %.3f tensor_pipe %s launch_ms %.4f share %.2f launches %d" % (
    sys.argv[2], d["ms_per_step"], d["value"], r["frac"] or 0, (r.get("tensor_pipe") or {}).get("frac"),
    r["avg_launch_ms"], r["share_of_step"] or 0, d["gpu_launches"]))
PY
}
for wl in stft2048 cfg2 cfg5; do
  timeout 200 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu \
      > gpurun_out/blk_$wl.json 2>> gpurun_out/blk_err.txt && show gpurun_out/blk_$wl.json $wl
done
echo "== full bench line"
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc $?"; tail -3 gpurun_out/bench_full.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_full.json"))
print("value %.3e ms %.4f frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["frac"]))
print("e2e", d.get("e2e")); print("pcie", d.get("pcie"))
for k, v in (d.get("workloads") or {}).items():
    print(k, {kk: v.get(kk) for kk in ("ms_per_step", "value", "error")}, (v.get("roofline") or {}).get("frac"))
print("reference_gpu", d.get("reference_gpu")); print("cpu", d.get("cpu_baseline")); print("clocks", d.get("clocks"))
PY
tail -5 gpurun_out/blk_err.txt
