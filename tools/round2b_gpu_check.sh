#!/usr/bin/env bash
# Second GPU pass of round 2 (one short gpurun call, 1 GPU): the full -m gpu suite (the new CFP and
# balanced-schedule tests run last), the cfg3 A/B of the balanced tall-kernel schedule, smoke, and the
# default bench line.  Most important first: the call may be cut short by the remaining GPU budget.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
Q="--no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu"
echo "== pytest -m gpu"
timeout 420 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/r02b_pytest_gpu.txt
echo "== cfg3 static vs balanced"
NNAB_TALL_BALANCE=0 timeout 120 python bench.py --workload cfg3 --steps 30 --warmup 5 $Q > gpurun_out/r02b_bench_cfg3_static.json 2> gpurun_out/r02b_bench_cfg3_static.err; echo "rc $?"
NNAB_TALL_BALANCE=1 timeout 120 python bench.py --workload cfg3 --steps 30 --warmup 5 $Q > gpurun_out/r02b_bench_cfg3_balanced.json 2> gpurun_out/r02b_bench_cfg3_balanced.err; echo "rc $?"
python - <<'PY'
import json
for k in ("static", "balanced"):
    try:
        d = json.load(open(f"gpurun_out/r02b_bench_cfg3_{k}.json"))
        print(k, "ms %.4f value %.4e frac %.3f" % (d["ms_per_step"], d["value"], d["roofline"]["frac"]), d["clocks"])
    except Exception as e:
        print(k, "unreadable:", e)
PY
echo "== smoke"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r02b_smoke.txt
echo "== bench (default line)"
timeout 600 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "rc $?"; tail -2 gpurun_out/r02b_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02b_bench.json"))
    r = d["roofline"]
    print("cfg2 value %.4e ms %.4f frac %.3f | e2e %.4f ms" % (d["value"], d["ms_per_step"], r["frac"], d["e2e"]["ms_per_step"]))
    for k, v in d["workloads"].items():
        print(k, "ms %.4f value %.3e frac %.3f" % (v["ms_per_step"], v["value"], v["roofline"]["frac"]))
except Exception as e:
    print("default bench unreadable:", e)
PY
echo "== CFP timing"
timeout 200 python tools/bench_cfp.py > gpurun_out/r02b_bench_cfp.json 2> gpurun_out/r02b_bench_cfp.err; echo "rc $?"; cat gpurun_out/r02b_bench_cfp.json
