#!/usr/bin/env bash
# The driver's two bench arms with the final code + launch lists aligned to whole steps.
set -u
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "rc $?"; tail -2 gpurun_out/r02_final_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_final_bench_reference.json 2>/dev/null; echo "rc $?"
Q="--no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu"
for wl in cfg2 cfg3 cfg4 cfg5 stft2048; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_$wl.csv python bench.py --workload $wl --steps 4 --warmup 3 $Q > /dev/null 2>&1
done
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_final_bench.json"))
r = d["roofline"]
print("cfg2 value %.4e ms %.4f frac %.3f pipe %.3f | e2e %.4f ms (%.2f of link; pcie %s) | cpu %s" % (
    d["value"], d["ms_per_step"], r["frac"], r["tensor_pipe"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["frac_of_link"],
    d["pcie"], d.get("cpu_baseline", {}).get("value")))
for k, v in d["workloads"].items():
    print(k, "ms %.4f value %.3e frac %.3f" % (v["ms_per_step"], v["value"], v["roofline"]["frac"]))
print("reference_gpu", {k: v.get("ms_per_step") for k, v in d.get("reference_gpu", {}).items()})
print("clocks", d["clocks"])
r2 = json.load(open("gpurun_out/r02_final_bench_reference.json"))
print("reference arm", r2["value"], r2["cpu_baseline"]["cores"], r2["cpu_baseline"]["kind"])
PY
