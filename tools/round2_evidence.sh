#!/usr/bin/env bash
# Final evidence of round 2, one gpurun call (1 GPU): full -m gpu suite, smoke, the driver's bench lines,
# launch lists and ncu --set full captures of the dominant kernels.  Outputs under gpurun_out/ (r02_*).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.json
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6 | tee gpurun_out/r02_final_pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r02_final_smoke.txt
echo "== bench (default line)"; timeout 900 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err; echo "rc $?"; tail -2 gpurun_out/r02_final_bench.err
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_final_bench_reference.json 2>/dev/null; echo "rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_final_bench.json"))
r = d["roofline"]
print("cfg2 value %.4e ms %.4f frac %.3f pipe %.3f | e2e %.4f ms (%.2f of link) | cpu %s" % (
    d["value"], d["ms_per_step"], r["frac"], r["tensor_pipe"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["frac_of_link"],
    d.get("cpu_baseline", {}).get("value")))
for k, v in d["workloads"].items():
    print(k, "ms %.4f value %.3e frac %.3f" % (v["ms_per_step"], v["value"], v["roofline"]["frac"]))
print("reference_gpu", {k: v.get("ms_per_step") for k, v in d.get("reference_gpu", {}).items()})
print("clocks", d["clocks"])
r2 = json.load(open("gpurun_out/r02_final_bench_reference.json"))
print("reference arm", r2["value"], r2["cpu_baseline"]["cores"], r2["cpu_baseline"]["kind"])
PY
Q="--no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu"
echo "== launch lists"
for wl in cfg2 cfg3 cfg4 cfg5; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/r02_launches_$wl.csv python bench.py --workload $wl --steps 4 --warmup 3 $Q > /dev/null 2>&1
done
echo "== ncu --set full"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:framed_tcb -s 3 -c 1 -f -o gpurun_out/r02_cfg2_block python bench.py --workload cfg2 --steps 3 --warmup 3 $Q > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:framed_tcb -s 3 -c 1 -f -o gpurun_out/r02_stft2048_block python bench.py --workload stft2048 --steps 3 --warmup 3 $Q > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:framed_tc2t -s 3 -c 1 -f -o gpurun_out/r02_cfg3_tall python bench.py --workload cfg3 --steps 3 --warmup 3 $Q > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fir_tc_kernel -s 7 -c 1 -f -o gpurun_out/r02_cfg4_fir python bench.py --workload cfg4 --steps 3 --warmup 3 $Q > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:octave_tc_kernel -s 8 -c 1 -f -o gpurun_out/r02_cfg4_octave python bench.py --workload cfg4 --steps 3 --warmup 3 $Q > /dev/null 2>&1
echo "== MMA rate probe"
timeout 200 python tools/probe_mma_rate.py > gpurun_out/probe_mma_rate.log 2>&1; tail -3 gpurun_out/probe_mma_rate.log
ls -la gpurun_out/*.ncu-rep
