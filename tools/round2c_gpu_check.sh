#!/usr/bin/env bash
# Third GPU pass of round 2 (one short gpurun call, 1 GPU): dense filterbanks on the tensor cores
# (NNAB_FB_PLANES), Gammatonegram timing, ncu capture + launch list of the balanced tall-A CQT kernel.
set -u
mkdir -p gpurun_out
Q="--no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu"
echo "== fb planes tests"
timeout 300 python -m pytest tests/test_zz_gpu_fb_planes.py -m gpu -q --timeout 200 -p no:cacheprovider 2>&1 | tail -25 | tee gpurun_out/r02c_pytest_fb_planes.txt
echo "== gammatone timing"
timeout 200 python tools/bench_gammatone.py > gpurun_out/r02c_bench_gammatone.json 2> gpurun_out/r02c_bench_gammatone.err; echo "rc $?"; cat gpurun_out/r02c_bench_gammatone.json; tail -3 gpurun_out/r02c_bench_gammatone.err
echo "== ncu cfg3 balanced"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:framed_tc2t -s 3 -c 1 -f -o gpurun_out/r02c_cfg3_tall_balanced python bench.py --workload cfg3 --steps 3 --warmup 3 $Q > /dev/null 2>&1; echo "rc $?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 24 --csv --log-file gpurun_out/r02c_launches_cfg3.csv python bench.py --workload cfg3 --steps 4 --warmup 3 $Q > /dev/null 2>&1; echo "rc $?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
