#!/usr/bin/env bash
# One-call evidence run for a round (under gpurun): GPU tests, smoke, every BASELINE workload,
# the CPU arm, an ncu launch list and one full capture of the dominant kernel.
# Outputs land in gpurun_out/ev_* ; summaries are copied into profiles/ afterwards.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/ev_smi.txt
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/ev_pytest.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ev_smoke.txt 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/ev_bench_cfg2.json 2> gpurun_out/ev_err.txt
for wl in cfg3 cfg4 cfg5 stft2048; do
  timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e \
      > gpurun_out/ev_bench_$wl.json 2>> gpurun_out/ev_err.txt
done
timeout 200 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/ev_bench_reference.json 2>> gpurun_out/ev_err.txt
B="--steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv \
    --log-file gpurun_out/ev_launches_cfg2.csv python bench.py $B > /dev/null 2>> gpurun_out/ev_err.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:framed_tc2 -s 3 -c 1 \
    -o gpurun_out/ev_prof_cfg2 python bench.py $B > /dev/null 2>> gpurun_out/ev_err.txt
cat gpurun_out/ev_pytest.txt gpurun_out/ev_smoke.txt | tail -12
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/ev_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline") or {}
        print(f.split("ev_bench_")[1], round(d["value"]), round(d["ms_per_step"], 4), r.get("frac"),
              (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
tail -3 gpurun_out/ev_err.txt
