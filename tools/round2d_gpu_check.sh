#!/usr/bin/env bash
# Fourth GPU pass of round 2: small-batch latency figures, ncu captures of two deep octave CQT launches (cfg4).
set -u
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu"
echo "== latency"
timeout 200 python tools/bench_latency.py > gpurun_out/r02d_latency.json 2> gpurun_out/r02d_latency.err; echo "rc $?"; cat gpurun_out/r02d_latency.json; tail -3 gpurun_out/r02d_latency.err
echo "== ncu cfg4 deep octaves"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:octave_tc_kernel -s 10 -c 2 -f -o gpurun_out/r02d_cfg4_octave_deep python bench.py --workload cfg4 --steps 3 --warmup 3 $Q > /dev/null 2>&1; echo "rc $?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
