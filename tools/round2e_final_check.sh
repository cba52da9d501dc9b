#!/usr/bin/env bash
# Final GPU pass of round 2 (1 GPU): the full -m gpu suite, smoke, and the default bench line on the final
# code (balanced tall-kernel schedule and dense filterbanks on the tensor cores both on by default).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
echo "== pytest -m gpu"
timeout 420 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/r02e_pytest_gpu.txt
echo "== smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r02e_smoke.txt
echo "== bench (default line)"
timeout 600 python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "rc $?"; tail -2 gpurun_out/r02e_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02e_bench.json"))
    r = d["roofline"]
    print("cfg2 value %.4e ms %.4f frac %.3f | e2e %.4f ms | launches %s" % (d["value"], d["ms_per_step"], r["frac"], d["e2e"]["ms_per_step"], d.get("gpu_launches")))
    for k, v in d["workloads"].items():
        print(k, v.get("error") or "ms %.4f value %.3e frac %.3f" % (v["ms_per_step"], v["value"], v["roofline"]["frac"]))
    print("clocks", d["clocks"])
except Exception as e:
    print("default bench unreadable:", e)
PY
