#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== probe"; timeout 120 python tools/probe_rowoffset.py 2>&1 | tail -4
echo "== cqt / determinism tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_determinism.py tests/test_gpu_fullsize.py -m gpu -q --timeout 300 -k "cqt1992 or sweep-cqt-1992 or CQT1992 or cfg3 or repeatable" 2>&1 | tail -4
for wl in cfg3; do
  timeout 200 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu > gpurun_out/q_$wl.json 2>> gpurun_out/q_err.txt
  python -c "
import json; d=json.load(open('gpurun_out/q_$wl.json')); r=d['roofline']; print('$wl ms %.4f frac %.3f pipe %.3f' % (d['ms_per_step'], r['frac'], r['tensor_pipe']['frac']))"
done
echo "== ncu stft2048 block kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:framed_tcb -s 3 -c 1 -o gpurun_out/r02_stft2048_block \
   python bench.py --workload stft2048 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu > gpurun_out/ncu_stft.log 2>&1; tail -2 gpurun_out/ncu_stft.log
echo "== ncu cfg2 block kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:framed_tcb -s 3 -c 1 -o gpurun_out/r02_cfg2_block \
   python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu > gpurun_out/ncu_cfg2.log 2>&1; tail -2 gpurun_out/ncu_cfg2.log
ls -la gpurun_out/*.ncu-rep
