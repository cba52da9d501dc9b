#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "mel or mfcc or MFCC or Mel or repeatable or cfg2 or cfg5 or gammatone or cqt1992 or cfg3" 2>&1 | tail -4
for wl in cfg2 cfg5 cfg3; do
  timeout 200 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-workloads --no-reference-gpu > gpurun_out/q_$wl.json 2>> gpurun_out/q_err.txt
  python -c "
import json; d=json.load(open('gpurun_out/q_$wl.json')); r=d['roofline']; print('$wl ms %.4f frac %.3f pipe %.3f launch %.4f share %.2f' % (d['ms_per_step'], r['frac'], r['tensor_pipe']['frac'], r['avg_launch_ms'], r['share_of_step']))"
done
tail -3 gpurun_out/q_err.txt
