#!/usr/bin/env bash
# Last GPU pass of round 2: launch list and one ncu --set full capture of the dense-filterbank path (Gammatonegram).
set -u
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu"
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 30 --csv --log-file gpurun_out/r02f_launches_gammatone.csv python bench.py --workload gammatone --steps 4 --warmup 3 $Q > /dev/null 2>&1; echo "rc $?"
timeout 100 ncu --set full --clock-control none --import-source on -k regex:framed_tcb -s 3 -c 1 -f -o gpurun_out/r02f_gammatone_block python bench.py --workload gammatone --steps 3 --warmup 3 $Q > /dev/null 2>&1; echo "rc $?"
ls -la gpurun_out/r02f*
